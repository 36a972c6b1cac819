set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest9.log 2>&1
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --equil 1 --no-cpu-baseline --no-legs > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cycle -s 36 -c 1 -f -o gpurun_out/r2_cycle_c2 python bench.py --steps 2 --no-cpu-baseline --no-legs > gpurun_out/r2_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rdf -s 4 -c 1 -f -o gpurun_out/r2_rdf_c5 python bench.py --workload c5 --steps 2 --rdf-samples 512 --no-cpu-baseline > gpurun_out/r2_ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cycle -s 16 -c 1 -f -o gpurun_out/r2_cycle_c3 python bench.py --workload c3 --equil 12 --steps 2 --no-cpu-baseline --no-legs > gpurun_out/r2_ncu_c3.log 2>&1
ls -la gpurun_out
