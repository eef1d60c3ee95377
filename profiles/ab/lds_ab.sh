for v in base default base default; do
  if [ $v = default ]; then unset CFM_B200_LIB; else export CFM_B200_LIB=profiles/ab/$v.so; fi
  python profiles/ab_loop.py 1024 100 2
  python profiles/quick_perf.py 1024 ops_lds_$v.txt > gpurun_out/qp_lds_$v.log 2>&1
  grep -E "input_blocks.(1|4|5).0.(in_layers.2|out_layers.3)|input_blocks.4.1.attention|output_blocks.10.0.in_layers.2" gpurun_out/ops_lds_$v.txt | awk '{printf "%s %s | ", $1, $3} END {print ""}'
done
