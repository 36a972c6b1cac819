for pf in -1 0 1 2; do
  NM_LIST_PF=$pf python bench.py --workload c3 --equil 12 --steps 3 --no-legs --no-cpu-baseline > gpurun_out/r2_pf_c3_$pf.json 2>/dev/null
  NM_LIST_PF=$pf python bench.py --steps 4 --no-legs --no-cpu-baseline > gpurun_out/r2_pf_c2_$pf.json 2>/dev/null
done
