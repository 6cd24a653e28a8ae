python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grid2d" > gpurun_out/r2_grid_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_grid_tests.log; tail -3 gpurun_out/r2_grid_tests.log
SOPOT_B200_GRID_MARCH_STRICT=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grid2d" > gpurun_out/r2_grid_tests_ms.log 2>&1; echo "rc=$?" >> gpurun_out/r2_grid_tests_ms.log; tail -2 gpurun_out/r2_grid_tests_ms.log
(
echo "== default (contract: march, strict: tile with odd stride)"; python tools/time_cfg.py grid; python tools/time_cfg.py grid_strict
echo "== tile kernel for both"; SOPOT_B200_GRID_TILE_KERNEL=1 python tools/time_cfg.py grid
) > gpurun_out/r2_grid_timing.log 2>&1
cat gpurun_out/r2_grid_timing.log | grep -v per-stage
SOPOT_B200_GRID_TILE_KERNEL=1 ncu --set full --import-source on --clock-control none -k regex:grid_rk4_fused -c 1 -f -o gpurun_out/prof_tile_ws41 python tools/profile_driver.py grid 2 > gpurun_out/ncu_tile.log 2>&1; tail -1 gpurun_out/ncu_tile.log
