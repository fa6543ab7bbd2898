#!/usr/bin/env python
"""profiles/fabric_timeline.py — where a multi-GPU step spends its time, measured INSIDE graph-replayed steps from the
device-side barrier log (rala_b200_multi_barrier_log): for every barrier of the last step, the time the rank computed
before it (previous barrier left -> this barrier entered) and the time it waited in it for its peers.
Run under torchrun like bench.py (`--workload`, default c3 per GPU); every rank prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import bench_multi  # noqa: E402
from rala_b200 import multi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3", choices=list(bench.WORKLOADS))
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
torch.cuda.set_device(local_rank)
device = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=device)
records, piles, t0, n_total = bench_multi._global_dataset(args, rank, world, device)
fg = multi.FabricGraph(local_rank, rank, world)
fg.set_inputs(records, piles, None, t0)
fg.plan()
for _ in range(5):
    fg.run()
fg.M.synchronize()
dist.barrier()
fg.M.event_record(0)
for _ in range(args.steps):
    fg.run()
fg.M.event_record(1)
ms = fg.M.event_elapsed_ms() / args.steps
log = fg.M.barrier_log(0).astype(np.int64)
B = fg.M.BARRIERS_PER_STEP
names = ["first pass over the records + events routed", "containment resolved (one persistent kernel) + survivors pass",
         "final pass classified + events routed", "final containment resolved + edges emitted + routed",
         "rows built + slice pushed", "transitive pass + marks routed"]
last = log[-B:]
prev_exit = log[-B - 1, 1]
rows, compute_us, wait_us = [], 0.0, 0.0
for i in range(B):
    comp, wait = (last[i, 0] - prev_exit) / 1e3, (last[i, 1] - last[i, 0]) / 1e3
    rows.append([names[i], round(float(comp), 1), round(float(wait), 1)])
    compute_us += comp
    wait_us += wait
    prev_exit = last[i, 1]
print(json.dumps({"rank": rank, "world": world, "ms_per_step": ms, "barriers_per_step": B, "compute_us": round(float(compute_us), 1),
                  "wait_us": round(float(wait_us), 1),
                  "resolution sweeps [open victims at start, us since kernel start at end]": [[int(a), round(float(b) / 1e3, 1)] for a, b in fg.M.sweep_log(0, 0)],
                  "phases [name, compute us before the barrier, wait us in it]": rows}), flush=True)
dist.barrier()
fg.close()
dist.destroy_process_group()
