mkdir -p gpurun_out
V=$PWD/sopot_b200/lib/variants
{ python tools/time_dpend.py; SOPOT_B200_LIB=$V/libsopot_b200_s3_inth.so python tools/time_dpend.py; python tools/time_dpend.py; SOPOT_B200_LIB=$V/libsopot_b200_s3_inth.so python tools/time_dpend.py; } 2>&1 | tee gpurun_out/r2e_dpend_inth.txt
export SOPOT_B200_LIB=$V/libsopot_b200_s3_inth.so
python tools/kernel_hash.py $SOPOT_B200_LIB > gpurun_out/kernel_hash_inth.json
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dpend" > gpurun_out/r2e_gputest_dpend.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_gputest_dpend.log; tail -2 gpurun_out/r2e_gputest_dpend.log
timeout 100 ncu --set full --import-source on --clock-control none -k regex:rk4_ensemble_kernel -s 3 -c 1 -f -o gpurun_out/prof_bench_dpend_inth python bench.py --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_dpend_inth.log 2>&1; echo "ncu rc=$?"
