# driver-level check: scripts/lammps_remcmc.py on a 4 x 8 grid of 4000-atom replicas, with and without helper CTAs:
# the .thrm / .traj files must be identical
set -e
rm -rf /tmp/drvA /tmp/drvB; mkdir -p /tmp/drvA /tmp/drvB
(cd /tmp/drvA && python $GRAFT_REPO_ROOT/scripts/lammps_remcmc.py -v -n t -ss 10 -pn 4 -tn 8 -sn 3 -sm 16 -bm > log.txt 2>&1)
(cd /tmp/drvB && NM_NO_HELPERS=1 python $GRAFT_REPO_ROOT/scripts/lammps_remcmc.py -v -n t -ss 10 -pn 4 -tn 8 -sn 3 -sm 16 -bm > log.txt 2>&1)
ls -la /tmp/drvA | head -20
for f in $(cd /tmp/drvA && ls | grep -v log.txt); do cmp /tmp/drvA/$f /tmp/drvB/$f && echo "identical: $f"; done
tail -3 /tmp/drvA/log.txt
