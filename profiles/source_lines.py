#!/usr/bin/env python
"""profiles/source_lines.py <report.ncu-rep> <kernel regex> [top] — executed warp instructions and stall samples per
CUDA source line of one kernel (from `ncu --page source --csv --print-source cuda,sass`; needs -lineinfo builds and
--import-source on captures)."""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, h, lines, seen = None, None, [], set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        h = r
        ie, isamp, ithr = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
        continue
    if h is None or len(r) <= ie or not r[0].isdigit() or r[2] != "-":   # source-line rows carry "-" as address
        continue
    key = (cur, r[0])
    if key in seen:      # a second launch of the same kernel
        continue
    seen.add(key)
    lines.append((int(r[ie]), int(r[isamp]), int(r[ithr]), cur, int(r[0]), r[1].strip()))
tot = sum(x[0] for x in lines) or 1
tots = sum(x[1] for x in lines) or 1
print(f"{rx}: {tot} warp instructions, {tots} stall samples")
for n, s, t, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}% inst {100 * s / tots:5.1f}% stall  thr/inst {t / max(n, 1):5.1f}  {f}:{ln:<4d} {src[:110]}")
