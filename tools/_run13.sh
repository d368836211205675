cd /root/repo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3_bench_reference_arm.json 2> gpurun_out/r3_bench_reference_arm.err
python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; tail -c 300 gpurun_out/r3_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r3_ncu_launches.log 2>&1; tail -1 gpurun_out/r3_ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:cost_kernel -s 3 -c 1 -o gpurun_out/r3_cost python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/r3_ncu_cost.log 2>&1; tail -1 gpurun_out/r3_ncu_cost.log
ncu --set full --clock-control none --import-source on -k regex:pso_kernel -c 1 -o gpurun_out/r3_pso python tools/pso_batch_probe.py 1184 > gpurun_out/r3_ncu_pso.log 2>&1; tail -1 gpurun_out/r3_ncu_pso.log
