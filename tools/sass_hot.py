"""instruction mix of the force loops (those with 4 MUFU.RCP per iteration) of nm::k_cycle<NTHR> in the built library"""
import re, subprocess, sys
from collections import Counter
nthr = sys.argv[1] if len(sys.argv) > 1 else "512"
out = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN2nm7k_cycleILi%sEEEvNS_3DevEx" % nthr, sys.argv[2] if len(sys.argv) > 2 else "neuralmelting_b200/libnm_b200.so"], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
print("kernel instructions:", len(ins))
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA (0x[0-9a-f]+)', t)
    if not m: continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr:
        body = ins[addr[tgt]:i + 1]
        if len(body) < 400 and sum('MUFU.RCP' in x for _, x in body) == 4:
            c = Counter((x.split()[1] if x.startswith('@') else x.split()[0]).split('.')[0] for _, x in body)
            f64 = c['DFMA'] + c['DMUL'] + c['DADD']
            print("%#x..%#x  %d instr, %d FP64, %d FP32, slots(2*FP64+rest) %d  %s" % (tgt, a, len(body), f64, c['FFMA'] + c['FMUL'] + c['FADD'], len(body) + f64, dict(c.most_common(12))))
            if "-v" in sys.argv:
                for aa, x in body: print("   %#x %s" % (aa, x))
