mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/r2d_bench_final.json 2> gpurun_out/r2d_bench_final.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2d_bench_final.json; echo
