for v in base default base default; do
  if [ $v = default ]; then unset CFM_B200_LIB; else export CFM_B200_LIB=profiles/ab/$v.so; fi
  python profiles/ab_loop.py 1024 100 2
  python profiles/quick_perf.py 1024 ops_nswap_$v.txt > gpurun_out/qp_nswap_$v.log 2>&1
  grep -E "input_blocks.(7|8).0|output_blocks.(3|5).(0|1)" gpurun_out/ops_nswap_$v.txt | awk '{printf "%s %s | ", $1, $3} END {print ""}'
done
