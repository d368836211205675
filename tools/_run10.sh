cd /root/repo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3_bench_reference_arm.json 2> gpurun_out/r3_bench_reference_arm.err
python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; tail -c 2500 gpurun_out/r3_bench_n1.json
