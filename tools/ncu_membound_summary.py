"""One table of the memory-bound kernels in an ncu report: time, DRAM traffic, achieved GB/s against the measured HBM peak.
usage: REPORT.ncu-rep [peak_GBs]"""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.isfile(pk) else 6650.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]


def val(r, name):
    if name not in h:
        return float("nan")
    i = h.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6,
             "byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
    return v * scale


print(f"| kernel | time us | DRAM read MB | DRAM write MB | achieved GB/s | of {peak:.0f} GB/s | ncu dram % | grid x block |")
print("|---|---:|---:|---:|---:|---:|---:|---|")
for r in rows[2:]:
    name = r[h.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "")
    if "FillFunctor" in name or "elementwise" in name or "distribution" in name:
        continue  # the harness' own torch fills / randn
    us, rd, wr = val(r, "gpu__time_duration.sum"), val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us * 1e3
    pct = val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    grid = r[h.index("launch__grid_size")] + " x " + r[h.index("launch__block_size")]
    print(f"| `{name[:70]}` | {us:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} % | {pct:.0f} | {grid} |")
