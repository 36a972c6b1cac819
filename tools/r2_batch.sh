python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest18.log 2>&1
python bench.py > gpurun_out/r2_bench_default2.json 2> gpurun_out/r2_bench_default2.err
