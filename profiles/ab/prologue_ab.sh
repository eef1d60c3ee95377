for v in base default base default; do
  if [ $v = default ]; then unset CFM_B200_LIB; else export CFM_B200_LIB=profiles/ab/$v.so; fi
  python profiles/ab_loop.py 1024 100 2
  python profiles/ab_loop.py 128 100 4
done
