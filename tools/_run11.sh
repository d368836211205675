cd /root/repo
python -m pytest tests/test_gpu_swarm.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_sequences.py -x -q -m gpu > gpurun_out/r3j_tests.log 2>&1; tail -2 gpurun_out/r3j_tests.log
( for v in prev new pso1 psostat psostat1; do
  if [ $v = new ]; then lib=hand-pose-estimation_b200/libhpe_b200.so; else lib=hand-pose-estimation_b200/build/variants/libhpe_b200_$v.so; fi
  echo $v; HPE_B200_LIB=$PWD/$lib python tools/pso_batch_probe.py 1184; HPE_B200_LIB=$PWD/$lib python tools/pso_batch_probe.py 5920
done
for v in prev new; do
  if [ $v = new ]; then lib=hand-pose-estimation_b200/libhpe_b200.so; else lib=hand-pose-estimation_b200/build/variants/libhpe_b200_$v.so; fi
  echo $v; HPE_B200_LIB=$PWD/$lib python tools/track_probe.py 1000
done ) > gpurun_out/r3j_pso.txt 2>&1
cat gpurun_out/r3j_pso.txt
