N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tools/bench_grid_multigpu.py 8192 40 2>/dev/null | grep '^{' ; }
echo "== p2p"; run 29661
