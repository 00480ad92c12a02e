"""Opcode histogram + hottest instructions from an ncu report's source page.
usage: python tools/ncu_ophist.py REPORT.ncu-rep LAUNCH_INDEX [N_TOP]"""
import csv
import subprocess
import sys
from collections import Counter

rep, idx = sys.argv[1], int(sys.argv[2])
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
h = rows[1]
seen, data = set(), []
for r in rows[2:]:
    if len(r) > 10 and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0])
        data.append(r)
iS, iI = h.index("# Samples"), h.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data)
ex, sm = Counter(), Counter()
for r in data:
    t = r[1].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ex[op] += int(r[iI])
    sm[op] += int(r[iS])
te = sum(ex.values())
print(f"total warp-instr {te}  samples {tot}")
for op, c in ex.most_common(22):
    print(f"{op:10s} exec {c:11d} {100*c/te:5.1f}%  samples {sm[op]:6d} {100*sm[op]/tot:5.1f}%")
print()
for r in sorted(data, key=lambda r: -int(r[iS]))[:ntop]:
    print(f"{int(r[iS]):6d} {int(r[iI]):9d}  {r[1][:100]}")
