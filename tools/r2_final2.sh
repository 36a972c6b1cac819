set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_gputest.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2g_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
timeout 300 python bench.py --workload c4 --steps 24 --warmup 8 --equil 64 --no-legs > gpurun_out/r2g_bench_c4.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_ref.json 2>/dev/null
