set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gputest.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2h_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_ref.json 2>/dev/null
timeout 300 python bench.py --workload c3 --steps 4 --no-legs > gpurun_out/r2h_bench_c3.json 2>/dev/null
NM_NO_HELPERS=1 timeout 300 python bench.py --workload c3 --steps 4 --no-legs --no-cpu-baseline > gpurun_out/r2h_bench_c3_nohelp.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2h_launches_c2.csv python bench.py --steps 2 --warmup 3 --equil 1 --no-cpu-baseline --no-legs > gpurun_out/r2h_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cycle -s 36 -c 1 -f -o gpurun_out/r2h_cycle_c2 python bench.py --steps 2 --no-cpu-baseline --no-legs > gpurun_out/r2h_ncu_c2.log 2>&1
