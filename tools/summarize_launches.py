"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"])).replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total serialized device time: {tot / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches\n")
print("| kernel | launches | total us | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:70]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
