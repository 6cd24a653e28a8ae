# one GPU call: the third-form double-pendulum step (SB_DPEND_STEP=3, the tree's default build) -- tests, bench line, ncu capture, A/B timings
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_gputest.log; tail -4 gpurun_out/r2b_gputest.log
python bench.py --no-extras > gpurun_out/r2b_bench_step3_noextras.json 2> gpurun_out/r2b_bench_step3_noextras.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2b_bench_step3_noextras.json; echo
python tools/kernel_hash.py > gpurun_out/kernel_hash_step3.json
for i in 1 2; do
python tools/time_dpend.py
for t in step2 minb6 minb5 b64 b256 ipt2 r32; do SOPOT_B200_LIB=$PWD/sopot_b200/lib/variants/libsopot_b200_s3_$t.so python tools/time_dpend.py; done
done 2>&1 | tee gpurun_out/r2b_dpend_variants.txt
timeout 240 ncu --set full --import-source on --clock-control none -k regex:rk4_ensemble_kernel -s 3 -c 1 -f -o gpurun_out/prof_bench_dpend_step3 python bench.py --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_dpend_step3.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/ncu_dpend_step3.log | head -c 200; echo
timeout 200 python bench.py > gpurun_out/r2b_bench_step3_full.json 2> gpurun_out/r2b_bench_step3_full.err; echo "full bench rc=$?"
