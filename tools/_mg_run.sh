N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tools/bench_grid_multigpu.py 2048 40 2>/dev/null | grep '^{' ; }
echo "== overlap"; run 29661
echo "== no overlap"; SOPOT_B200_UGRID_NO_OVERLAP=1 run 29663
echo "== overlap"; run 29665
echo "== no overlap"; SOPOT_B200_UGRID_NO_OVERLAP=1 run 29667
