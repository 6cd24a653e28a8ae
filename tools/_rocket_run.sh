python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rocket" > gpurun_out/r2_rocket_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_rocket_tests.log; tail -3 gpurun_out/r2_rocket_tests.log
(
echo "== stock (stage loop, 4 CTAs)"; python tools/time_cfg.py rocket
for v in rl3 rl5 rnoloop; do echo "== $v"; SOPOT_B200_LIB=sopot_b200/lib/variants/libsopot_b200_$v.so python tools/time_cfg.py rocket; done
) > gpurun_out/r2_rocket_timing.log 2>&1
cat gpurun_out/r2_rocket_timing.log
