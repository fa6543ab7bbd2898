#!/usr/bin/env python
"""profiles/skew_k3.py [n_hubs] [spokes] — the transitive pass (K3) on a SKEWED degree distribution at the K3 boundary
(BASELINE configs[4]: node degrees > 2 k), for the ncu evidence the north star asks for (warp divergence and occupancy
of the light / heavy paths).  Run under ncu with profiles/capture_skew.sh; prints one JSON line with the degree
histogram and the wall time per call (host copies included: this is the stateless C-ABI entry point)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from rala_b200 import api, synth  # noqa: E402

n_hubs = int(sys.argv[1]) if len(sys.argv) > 1 else 48
spokes = int(sys.argv[2]) if len(sys.argv) > 2 else 2600
n_nodes, e = synth.hub_graph(n_hubs=n_hubs, spokes=spokes, links_per_spoke=6, seed=8)
# a power-law background on top of the hubs, ids shifted behind the hub graph's nodes
rng = np.random.Generator(np.random.PCG64(77))
n_bg = 200_000
w = 1.0 / np.arange(1, 2 * n_bg + 1) ** 0.9
w /= w.sum()
a = rng.choice(2 * n_bg, 1_500_000, p=w)
b = rng.integers(0, 2 * n_bg, a.shape[0])
keep = (a >> 1) != (b >> 1)
a, b = a[keep] + n_nodes, b[keep] + n_nodes
bg = np.empty((2 * a.shape[0], 3), np.uint32)
bg[0::2] = np.stack([a, b, rng.integers(10, 4000, a.shape[0])], 1)
bg[1::2] = np.stack([b ^ 1, a ^ 1, rng.integers(10, 4000, a.shape[0])], 1)
edges = np.ascontiguousarray(np.concatenate([e, bg]))
n_nodes += 2 * n_bg
deg = np.bincount(edges[:, 0], minlength=n_nodes)
ctx = api.Context(0)
times, pairs = [], 0
for _ in range(4):
    t0 = time.perf_counter()
    marked, pairs = ctx.transitive_reduce(n_nodes, edges)
    times.append(time.perf_counter() - t0)
hist = {f"<= {k}": int((deg <= k).sum()) for k in (1, 16, 64, 256, 1024)}
hist["> 1024"] = int((deg > 1024).sum())
print(json.dumps({"n_nodes": int(n_nodes), "n_edges": int(edges.shape[0]), "max_degree": int(deg.max()), "degree_histogram": hist,
                  "transitive_pairs": int(pairs), "marked_edges": int(marked.sum()), "s_per_call_incl_copies": min(times)}))
ctx.close()
