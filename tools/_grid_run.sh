for n in 1152 1280 1536 1792; do for cfg in grid grid_strict; do
echo "== $cfg $n"
SOPOT_B200_GRID_TILE_KERNEL=1 python tools/time_cfg.py $cfg $n 2>&1 | head -1 | cut -c1-70
SOPOT_B200_GRID_MARCH_KERNEL=1 python tools/time_cfg.py $cfg $n 2>&1 | head -1 | cut -c1-70
SOPOT_B200_GRID_MARCH_KERNEL=1 SOPOT_B200_MARCH_ROWS=32 python tools/time_cfg.py $cfg $n 2>&1 | head -1 | cut -c1-70
SOPOT_B200_GRID_MARCH_KERNEL=1 SOPOT_B200_MARCH_ROWS=48 python tools/time_cfg.py $cfg $n 2>&1 | head -1 | cut -c1-70
done; done
