#!/usr/bin/env python
"""profiles/sass_hist.py <report.ncu-rep> <kernel regex> — executed-instruction histogram by SASS opcode
(from `ncu --page source --csv`), to see where a kernel's issue slots go."""
import collections
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ops, samp, tot, h = collections.Counter(), collections.Counter(), 0, None
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        if h is not None:
            break          # second launch of the same kernel: the first is enough
        h = r
        ie, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        continue
    if h is None or len(r) <= ie:
        continue
    parts = r[isrc].split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    v = int(r[ie])
    ops[op] += v
    samp[op] += int(r[isamp])
    tot += v
print(f"{rx}: {tot} warp instructions executed")
for op, v in ops.most_common(28):
    print(f"  {op:12s} {v:11d} {100 * v / tot:5.1f} %   stall samples {samp[op]}")
