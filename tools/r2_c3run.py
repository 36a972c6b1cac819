"""N2 acceptance (development aid): a C3-shaped driver run (32 x 32 replicas of 4000 atoms, a few recorded cycles) -- wall
time, peak host memory (sampled RSS of the process), output sizes"""
import os, resource, subprocess, sys, threading, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work = "/tmp/c3run"
os.makedirs(work, exist_ok=True)
cmd = [sys.executable, os.path.join(root, "scripts", "lammps_remcmc.py"), "-v", "-bm", "-ss", "10", "-pn", "32", "-tn", "32", "-sn", sys.argv[1] if len(sys.argv) > 1 else "3",
       "-sm", "16", "-nt", "8", "-n", "c3shape", "-dn"]
t0 = time.time()
p = subprocess.Popen(cmd, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
peak = [0]
def watch():
    while p.poll() is None:
        try:
            with open("/proc/%d/status" % p.pid) as fh:
                for line in fh:
                    if line.startswith("VmRSS"):
                        peak[0] = max(peak[0], int(line.split()[1]))
        except OSError:
            pass
        time.sleep(0.2)
th = threading.Thread(target=watch); th.start()
out = p.communicate()[0]; th.join()
print(out[-3000:])
print("exit %d  wall %.1f s  peak RSS %.2f GB" % (p.returncode, time.time() - t0, peak[0] / 1e6))
tot = 0
for f in sorted(os.listdir(work)):
    sz = os.path.getsize(os.path.join(work, f)); tot += sz
    print("%-40s %12d" % (f, sz))
print("total %.2f GB" % (tot / 1e9))
