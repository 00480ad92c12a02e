"""Markdown summary (key metrics + top stall reasons) of every launch in an ncu report.  usage: REPORT.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    print(f"## `{name[:110]}`\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for w in want:
        if w in h:
            print(f"| {w} | {r[h.index(w)]} | {units[h.index(w)]} |")
    st = []
    for i, c in enumerate(h):
        if "pcsamp_warps_issue_stalled" in c and "not_issued" not in c:
            try:
                st.append((float(r[i].replace(",", "")), c.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print("\nWarp-state samples: " + ", ".join(f"{n} {100 * v / tot:.0f} %" for v, n in sorted(st, reverse=True)[:7]) + "\n")
