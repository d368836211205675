cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r3l_gpu_tests.log 2>&1; tail -4 gpurun_out/r3l_gpu_tests.log
( for v in prev new; do
  if [ $v = new ]; then lib=hand-pose-estimation_b200/libhpe_b200.so; else lib=hand-pose-estimation_b200/build/variants/libhpe_b200_$v.so; fi
  echo $v; HPE_B200_LIB=$PWD/$lib python tools/pso_batch_probe.py 1184; HPE_B200_LIB=$PWD/$lib python tools/pso_batch_probe.py 5920
  HPE_B200_LIB=$PWD/$lib python tools/observe_probe.py 2368; HPE_B200_LIB=$PWD/$lib python tools/track_probe.py 1000
  HPE_B200_LIB=$PWD/$lib timeout 120 python bench.py --steps 300 --warmup 5 --no-extras 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'evals/s %.4g  ms %.4f  frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))"
done ) > gpurun_out/r3l_fold.txt 2>&1
cat gpurun_out/r3l_fold.txt
