"""ncu report -> the metric summary kept under profiles/ (development aid).
usage: python tools/ncu_summary.py report.ncu-rep "header line" > profiles/NAME.txt"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__t_sector_hit_rate\.pct|launch__(block_size|grid_size|occupancy_limit_registers|"
                  r"occupancy_limit_shared_mem|registers_per_thread|shared_mem_per_block_dynamic)$|lts__t_sector_hit_rate\.pct|"
                  r"lts__t_bytes\.sum$|sm__cycles_(active\.avg|elapsed\.avg|elapsed\.max)$|sm__inst_executed_pipe_(alu|fma|fmaheavy|fp64|lsu|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_issued\.avg\.per_cycle_active|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__inst_executed_op_(local|shared|global)_(ld|st)\.sum$|smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"sm__sass_inst_executed_op_local_(ld|st)\.sum$|smsp__sass_inst_executed_op_(local|shared|global)_(ld|st)\.sum$|l1tex__t_bytes_pipe_lsu_mem_local_op_(ld|st)\.sum$)")

def main():
    rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    if header:
        print("# " + header)
    kn = names.index("Kernel Name") if "Kernel Name" in names else None
    if kn is not None:
        print("# kernel: %s   grid %s block %s" % (vals[kn], vals[names.index("Grid Size")] if "Grid Size" in names else "?", vals[names.index("Block Size")] if "Block Size" in names else "?"))
    for n, u, v in sorted(zip(names, units, vals)):
        if KEEP.match(n):
            print("%-95s %-16s %s" % (n, u, v))

main()
