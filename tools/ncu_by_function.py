"""ncu report -> share of executed warp instructions / issue slots / stall samples per source function (development aid).
usage: python tools/ncu_by_function.py report.ncu-rep [source.cu]
Lines are attributed to the enclosing top-level function of the source file (found by a brace-depth scan)."""
import csv, io, re, subprocess, sys, collections

def functions(path):
    """line -> name of the enclosing function (depth-0 definitions, also inside one namespace)"""
    names, out, depth, cur, start_depth = {}, {}, 0, None, None
    pend = None
    src = open(path).read().splitlines()
    for no, l in enumerate(src, 1):
        code = re.sub(r'//.*', '', l)
        if cur is None:
            m = re.search(r'([A-Za-z_][A-Za-z_0-9]*)\s*\([^;]*$', code) if not code.strip().startswith('#') else None
            if m and m.group(1) not in ('if', 'for', 'while', 'switch', 'asm', 'namespace', '__launch_bounds__', 'NM_CTAS_PER_SM'):
                pend = m.group(1)
            if pend and '{' in code and not re.match(r'\s*namespace', code):
                cur, start_depth = pend, depth
                pend = None
        if cur: out[no] = cur
        depth += code.count('{') - code.count('}')
        if cur is not None and depth <= start_depth and '}' in code:
            cur = None
    return out

def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = csv.reader(io.StringIO(txt))
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    fmap, fpath, hdr, line = {}, None, None, None
    lineagg = collections.defaultdict(lambda: [0, 0, 0])
    for r in rows:
        if not r: continue
        if r[0] == "File Path":
            fpath = r[1]
            import os
            loc = os.path.join("neuralmelting_b200/csrc", os.path.basename(fpath))
            fmap = functions(loc) if os.path.exists(loc) else {}
            continue
        if r[0] == "Line No": hdr = r; continue
        if r[0] == "Function Name" or hdr is None: continue
        if r[0] != "":
            line = int(r[0]); continue
        ix = {n: i for i, n in enumerate(hdr)}
        sass = r[3]
        num = lambda v: int(v) if v not in ("", "-") else 0
        ex = num(r[ix["Instructions Executed"]])
        smp = num(r[ix["# Samples"]])
        op = sass.split()[1] if sass.strip().startswith('@') else sass.split()[0]
        f64 = op.split('.')[0] in ("DFMA", "DMUL", "DADD", "DSETP")
        key = (os.path.basename(fpath), fmap.get(line, "?"))
        a = agg[key]; a[0] += ex; a[1] += ex * (2 if f64 else 1); a[2] += smp
        for k in ("stall_short_sb", "stall_long_sb", "stall_barrier", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_lg", "stall_mio", "stall_dispatch", "stall_branch_resolving"):
            if k in ix: a[3][k] += num(r[ix[k]])
        la = lineagg[(os.path.basename(fpath), line)]; la[0] += ex; la[1] += ex * (2 if f64 else 1); la[2] += smp
    tot = [sum(a[i] for a in agg.values()) for i in range(3)]
    print("total warp instructions %.3e  issue slots (FP64 x2) %.3e  samples %d" % tuple(tot))
    print("%-18s %-28s %8s %8s %8s   top stalls" % ("file", "function", "instr%", "slots%", "sample%"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        if a[2] < 0.002 * tot[2] and a[0] < 0.002 * tot[0]: continue
        st = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(1, a[2])) for k, v in a[3].most_common(4))
        print("%-18s %-28s %8.2f %8.2f %8.2f   %s" % (key[0], key[1], 100.0 * a[0] / tot[0], 100.0 * a[1] / tot[1], 100.0 * a[2] / tot[2], st))
    if "-l" in sys.argv:
        print("hottest lines by samples:")
        for key, a in sorted(lineagg.items(), key=lambda kv: -kv[1][2])[:40]:
            print("  %s:%d  instr %.2f%%  samples %.2f%%" % (key[0], key[1], 100.0 * a[0] / tot[0], 100.0 * a[2] / tot[2]))

main()
