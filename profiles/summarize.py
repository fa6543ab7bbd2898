#!/usr/bin/env python
"""profiles/summarize.py <tag> — turn what profiles/capture.sh brought back in gpurun_out/ into the
small text files that are committed under profiles/:

    <tag>_bench.json            the bench line of that call (NOT measured under a profiler)
    <tag>_launches.csv          ncu launch list of the same bench command (cold-cache, serialised: shares only)
    <tag>_launch_shares.txt     per-kernel launches / mean duration / share of the summed kernel time
    <tag>_step_sequence.txt     the kernels of one step in launch order
    <tag>_kernels.csv           selected counters of the `ncu --set full` capture of the dominant kernels
    <tag>_pipes_stalls.txt      busiest pipes and top stall reasons per captured kernel

Reads the .ncu-rep with `ncu -i ... --page raw --csv` (works without a GPU).
"""
from __future__ import annotations

import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(tag: str):
    src = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith("==")]
    with open(os.path.join(DST, f"{tag}_launches.csv"), "w") as f:
        f.writelines(lines)
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for x in rows:
        agg.setdefault(x["Kernel Name"].split("(")[0], []).append(float(x["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(DST, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {len(rows)} launches, "
                f"{tot / 1e3:.1f} us of kernel time in total (cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'kernel':40s} {'launches':>8s} {'mean_us':>10s} {'share_%':>8s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:40]:40s} {len(v):8d} {sum(v) / len(v) / 1e3:10.1f} {100 * sum(v) / tot:8.1f}\n")
    first = [i for i, x in enumerate(rows) if "k_classify_events" in x["Kernel Name"]]
    if len(first) >= 6:
        s, e = first[4], first[5]
        with open(os.path.join(DST, f"{tag}_step_sequence.txt"), "w") as f:
            f.write("# kernels of one step of the resident pipeline, launch order (ncu durations)\n")
            tot = 0.0
            for x in rows[s:e]:
                v = float(x["Metric Value"]) / 1e3
                tot += v
                f.write(f"{x['Kernel Name'].split('(')[0][:40]:40s} grid={x['Grid Size']:>14s} block={x['Block Size']:>12s} {v:9.1f} us\n")
            f.write(f"{'sum':40s} {tot:50.1f} us\n")


def full(tag: str):
    rep = os.path.join(OUT, f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(DST, f"{tag}_kernels.csv"), "w", newline="") as f:
        w = csv.writer(f)
        cols = [m for m in METRICS if m in idx]
        w.writerow(["kernel"] + [f"{m} [{units[idx[m]]}]" for m in cols])
        for d in data:
            w.writerow([d[idx["Kernel Name"]].split("(")[0]] + [d[idx[m]] for m in cols])
    # per-launch DRAM traffic of every captured kernel, for bench.py's roofline.traffic
    import json
    tpath = os.path.join(DST, "traffic.json")
    traffic = {"tag": tag, "kernels": {}}
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
        traffic["tag"] = tag
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        rd = float(d[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
        wr = float(d[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
        prev = traffic["kernels"].get(name)
        if prev is None or prev.get("tag") != tag or rd + wr > prev["dram_bytes"]:   # keep the largest launch of a capture
            traffic["kernels"][name] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "tag": tag}
    with open(tpath, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    pipe_keys = [h for h in hdr if "pipe" in h and h.endswith(".avg.pct_of_peak_sustained_active")] + \
                [h for h in hdr if "pipe_xu" in h or h.startswith("l1tex__data_pipe_lsu_wavefronts.avg")]
    stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    with open(os.path.join(DST, f"{tag}_pipes_stalls.txt"), "w") as f:
        for d in data:
            f.write(f"== {d[idx['Kernel Name']].split('(')[0]}  {d[idx['gpu__time_duration.sum']]} {units[idx['gpu__time_duration.sum']]}\n")
            for title, keys in (("pipes (% of peak)", pipe_keys), ("stalls (warps per issue-active cycle)", stall_keys)):
                vals = []
                for k in set(keys):
                    try:
                        vals.append((float(d[idx[k]]), k))
                    except ValueError:
                        pass
                f.write(f"  {title}\n")
                for v, k in sorted(vals, reverse=True)[:6]:
                    f.write(f"    {v:9.2f}  {k}\n")


def main():
    tag = sys.argv[1]
    os.makedirs(DST, exist_ok=True)
    b = os.path.join(OUT, f"bench_{tag}.json")
    if os.path.exists(b):
        shutil.copy(b, os.path.join(DST, f"{tag}_bench.json"))
    c = os.path.join(OUT, f"clocks_{tag}.csv")
    if os.path.exists(c):
        lines = open(c).read().splitlines()
        with open(os.path.join(DST, f"{tag}_clocks.txt"), "w") as f:
            f.write(lines[0] + "\n")
            sm = sorted(float(l.split(",")[1].split()[0]) for l in lines[1:] if l.strip())
            act = sorted(set(l.split(",")[4].strip() for l in lines[1:] if l.strip()))
            f.write(f"# {len(sm)} samples during pytest + bench: clocks.sm min/median/max = {sm[0]:.0f}/{sm[len(sm) // 2]:.0f}/{sm[-1]:.0f} MHz; "
                    f"clocks_event_reasons.active values seen: {act}\n")
    launches(tag)
    full(tag)


if __name__ == "__main__":
    main()
