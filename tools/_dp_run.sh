for i in 1 2; do
python tools/time_dpend.py
SOPOT_B200_LIB=sopot_b200/lib/variants/libsopot_b200_dpold.so python tools/time_dpend.py
done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dpend" 2>&1 | tail -3
