python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grid2d" > gpurun_out/r2_grid_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_grid_tests.log; tail -5 gpurun_out/r2_grid_tests.log
(
echo "== default (march both modes)"; python tools/time_cfg.py grid; python tools/time_cfg.py grid_strict
echo "== strict on the tile kernel"; SOPOT_B200_GRID_TILE_STRICT=1 python tools/time_cfg.py grid_strict
) > gpurun_out/r2_grid_timing.log 2>&1
cat gpurun_out/r2_grid_timing.log | grep -v per-stage
ncu --set full --import-source on --clock-control none -k regex:grid_rk4_march -c 1 -f -o gpurun_out/prof_march_strict2 python tools/profile_driver.py grid_strict 2 > gpurun_out/ncu_march_strict2.log 2>&1; tail -1 gpurun_out/ncu_march_strict2.log
