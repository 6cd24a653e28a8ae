python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest_final.log; tail -4 gpurun_out/r2_gputest_final.log
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python tools/kernel_hash.py > gpurun_out/kernel_hash.json
ncu --set full --import-source on --clock-control none -k regex:rk4_ensemble_kernel -s 3 -c 1 -f -o gpurun_out/prof_bench_dpend_micro python bench.py --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_dpend.log 2>&1; tail -1 gpurun_out/ncu_dpend.log | head -c 300
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launch_list.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_list.log 2>&1; tail -1 gpurun_out/ncu_bench_list.log | head -c 300
