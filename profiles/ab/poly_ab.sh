for v in poly0 default poly50; do
  if [ $v = default ]; then unset CFM_B200_LIB; else export CFM_B200_LIB=profiles/ab/$v.so; fi
  python profiles/perop_mnist.py 4096 > gpurun_out/perop_mnist_$v.txt 2>&1
  echo "$v mnist: $(head -1 gpurun_out/perop_mnist_$v.txt) attn: $(grep 'input_blocks.1.1.attention' gpurun_out/perop_mnist_$v.txt)"
  python profiles/ab_loop.py 1024 100 2
done
