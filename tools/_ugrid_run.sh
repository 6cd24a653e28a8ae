python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "unified or ugrid" 2>&1 | tail -8
for n in 512 1024; do python tools/time_cfg.py ugrid $n 2>&1 | grep -v "stage pairs" | cut -c1-70; done
