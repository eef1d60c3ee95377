for ib in 32768 49152 65536 98304 131072; do
  CFM_B200_LIB=profiles/ab/tune.so CFM_GN_ITEM_BYTES=$ib python profiles/quick_perf.py 1024 ops_gn_$ib.txt > gpurun_out/qp_gn_$ib.log 2>&1
  echo "ITEM_BYTES=$ib: $(grep groupnorm gpurun_out/qp_gn_$ib.log | head -1)"
  grep -E "output_blocks.(9|10|8|6).0.in_layers.0 |out.0 |input_blocks.1.0.in_layers.0 " gpurun_out/ops_gn_$ib.txt
done
