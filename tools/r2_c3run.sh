# N2 acceptance: a C3-shaped run (32 x 32 replicas of 4000 atoms) with a few recorded cycles: wall time, peak host memory, output sizes
mkdir -p /tmp/c3run && cd /tmp/c3run
/usr/bin/time -v python $GRAFT_REPO_ROOT/scripts/lammps_remcmc.py -v -bm -ss 10 -pn 32 -tn 32 -sn 3 -sm 16 -nt 8 -n c3shape -dn > $GRAFT_REPO_ROOT/gpurun_out/r2_c3run.log 2> $GRAFT_REPO_ROOT/gpurun_out/r2_c3run.time
ls -la /tmp/c3run | head -30 >> $GRAFT_REPO_ROOT/gpurun_out/r2_c3run.log
du -sh /tmp/c3run >> $GRAFT_REPO_ROOT/gpurun_out/r2_c3run.log
