cd /root/repo
python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_selftest.py -x -q -m gpu > gpurun_out/r3m_tests.log 2>&1; tail -2 gpurun_out/r3m_tests.log
python tools/observe_probe.py 2368; python tools/observe_probe.py 1184
