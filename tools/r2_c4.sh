python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest17.log 2>&1
python bench.py --workload c4 --steps 24 --warmup 8 --equil 64 --no-legs > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
