run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 tools/bench_grid_multigpu.py 8192 20 2>/dev/null | grep '^{' ; }
N=${1:-8}
echo "== default"; run $N 29661
echo "== no overlap"; SOPOT_B200_NO_HALO_OVERLAP=1 run $N 29662
echo "== no overlap rows 205"; SOPOT_B200_NO_HALO_OVERLAP=1 SOPOT_B200_MARCH_ROWS=205 run $N 29663
echo "== no overlap rows 64"; SOPOT_B200_NO_HALO_OVERLAP=1 SOPOT_B200_MARCH_ROWS=64 run $N 29664
echo "== overlap rows 64"; SOPOT_B200_MARCH_ROWS=64 run $N 29665
echo "== overlap rows 205"; SOPOT_B200_MARCH_ROWS=205 run $N 29666
