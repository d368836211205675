cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r3k_gpu_tests.log 2>&1; tail -2 gpurun_out/r3k_gpu_tests.log
( for v in new pso1 psostat pso6; do
  if [ $v = new ]; then lib=hand-pose-estimation_b200/libhpe_b200.so; else lib=hand-pose-estimation_b200/build/variants/libhpe_b200_$v.so; fi
  echo $v; HPE_B200_LIB=$PWD/$lib python tools/pso_batch_probe.py 1184; HPE_B200_LIB=$PWD/$lib python tools/pso_batch_probe.py 5920
done ) > gpurun_out/r3k_pso.txt 2>&1
cat gpurun_out/r3k_pso.txt
