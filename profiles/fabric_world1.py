#!/usr/bin/env python
"""profiles/fabric_world1.py — the multi-GPU session's kernels with ONE rank (every exchange goes to the rank itself), on
configs[2]: what the routing / gather / resolution / row-build / push kernels cost by themselves.  Meant for
`ncu --metrics gpu__time_duration.sum` (a launch list); several ranks cannot be profiled, their barrier kernels wait for
kernels that ncu has not let run yet."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rala_b200 import api  # noqa: E402

ds = bench.make_dataset("c3", 1)
M = api.Multi([0])
M.set_piles(ds.flat_piles()).set_shards(ds.records).plan()
M.use_cuda_graph(False)
for _ in range(2):
    M.run()
M.synchronize()
M.event_record(0)
for _ in range(5):
    M.run()
M.event_record(1)
print(json.dumps({"world": 1, "ms_per_step_eager": M.event_elapsed_ms() / 5, "counts": M.counts(), "stage_ms": M.stage_ms(0)}))
M.close()
