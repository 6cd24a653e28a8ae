python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "unified or grid2d" > gpurun_out/r2_ugrid_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ugrid_tests.log; tail -4 gpurun_out/r2_ugrid_tests.log
(
echo "== march (default)"; python tools/time_cfg.py ugrid | grep -v pairs
echo "== gather"; SOPOT_B200_UGRID_GATHER=1 python tools/time_cfg.py ugrid | grep -v pairs
for rows in 16 64 128; do echo "== march rows $rows"; SOPOT_B200_UGRID_MARCH_ROWS=$rows python tools/time_cfg.py ugrid | grep -v pairs | grep "rk4\|euler"; done
echo "== grid2d"; python tools/time_cfg.py grid | head -1; python tools/time_cfg.py grid_strict | head -1
) > gpurun_out/r2_ugrid_timing.log 2>&1
cat gpurun_out/r2_ugrid_timing.log
