"""Per-source-line stall samples from an ncu report (needs --import-source on and -lineinfo).
usage: python tools/ncu_srclines.py REPORT.ncu-rep LAUNCH_INDEX [N_TOP]"""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None
lines = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        h = r
        continue
    if h and len(r) == len(h) and r[0].isdigit():
        lines.append((fname, r))
iS, iI = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[iS]) for _, r in lines)
print(f"total samples {tot}")
for f, r in sorted(lines, key=lambda fr: -int(fr[1][iS]))[:ntop]:
    st = sorted(((int(r[i]), h[i][6:]) for i in stall_cols if int(r[i]) > 0), reverse=True)[:3]
    print(f"{int(r[iS]):6d} {100 * int(r[iS]) / tot:5.1f}% {int(r[iI]):10d}  {f}:{r[0]:>4s}  {r[1].strip()[:90]}   {st}")
