"""Summarise an .ncu-rep (read on the CPU box): python scripts/ncu_summary.py <file.ncu-rep> [regex ...]
Prints the metrics that decide the roofline position of the hybrid-ODE kernels (issue, pipes, stalls, DRAM, occupancy)."""
import csv
import io
import re
import subprocess
import sys

DEFAULT = [
    r"^gpu__time_duration\.sum", r"launch__registers_per_thread", r"launch__occupancy_limit", r"launch__grid_size", r"launch__block_size",
    r"launch__occupancy_per", r"sm__warps_active\.avg\.pct_of_peak_sustained_active", r"smsp__issue_active\.avg\.pct",
    r"smsp__inst_executed\.sum$", r"sm__inst_executed_pipe_(fma|alu|xu|lsu|fmaheavy|fmalite|uniform|cbu|adu)\S*\.avg\.pct_of_peak_sustained_active",
    r"sm__pipe_(fma|alu|xu|fmaheavy|fmalite)\w*_cycles_active\.avg\.pct_of_peak_sustained_active",
    r"dram__bytes_(read|write)\.sum$", r"dram__throughput\.avg\.pct", r"smsp__warps_eligible\.avg\.per_cycle_active",
    r"smsp__average_warps?_issue_stalled_\w+_per_issue_active", r"smsp__average_warp_latency_issue_stalled", r"l1tex__t_sector_hit_rate",
    r"sm__throughput\.avg\.pct", r"sm__cycles_elapsed\.max", r"smsp__thread_inst_executed_per_inst_executed", r"sm__inst_executed\.avg\.per_cycle_(active|elapsed)",
    r"smsp__inst_executed\.avg\.per_cycle_active", r"local_(load|store)", r"lts__t_sector_hit_rate",
]


def main():
    path = sys.argv[1]
    pats = [re.compile(p) for p in (sys.argv[2:] or DEFAULT)]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("== {}  grid {} block {}".format(name[:150], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for i, h in enumerate(hdr):
            short = h.split(".TriageCompute.")[-1]
            if any(p.search(short) for p in pats):
                print("  {:<90} {:>18} {}".format(short, r[i], units[i]))


if __name__ == "__main__":
    main()
