#!/usr/bin/env python
"""tests/fabric_check.py [world ...] — quick GPU check of the multi-GPU session with several ranks on device 0:
parity against the oracle, then timing of graph-replayed steps on a 20 Mbp batch.  Prints one JSON line."""
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # several ranks on one device: one hardware queue per stream
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import oracle as O  # noqa: E402
from rala_b200 import api, synth  # noqa: E402

worlds = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]
if len(worlds) > 1:   # one process per world size: a fabric that died in one must not take the others with it
    import subprocess
    merged = {}
    for wd in worlds:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(wd)], capture_output=True, text=True, timeout=240)
            merged.update(json.loads(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else {f"world{wd}": {"error": r.stderr[-800:]}})
        except subprocess.TimeoutExpired:
            merged[f"world{wd}"] = {"error": "timeout after 240 s"}
        print(json.dumps({k: v for k, v in merged.items() if k == f"world{wd}"})[:1500], file=sys.stderr, flush=True)
    print(json.dumps(merged))
    sys.exit(0)
out = {}
ds = synth.generate(genome_len=3_000_000, coverage=30, read_len=9000, len_sd=2500, seed=71, noise=50, dual=True, min_ovl=900)
flags = (np.random.Generator(np.random.PCG64(1)).random(ds.n_reads) < 0.05).astype(np.uint8) * 2
w = O.Pipeline(ds.records, ds.flat_piles(), flags).run()
big = synth.generate(20_000_000, 40, 10000, seed=3)
G = api.Graph(api.Context(0))
G.set_piles(big.flat_piles()).set_hills(None).set_overlaps(big.records)
for _ in range(4):
    G.run()
G.ctx.synchronize()
G.ctx.event_record(0)
for _ in range(50):
    G.run()
G.ctx.event_record(1)
out["single_ms_per_step_20Mbp"] = G.ctx.event_elapsed_ms() / 50
big_edges, big_marked = G.edges(), G.marked()
for world in worlds:
    r = {}
    try:
        M = api.Multi([0] * world)
        M.set_barrier_timeout_ms(1500)
        M.set_piles(ds.flat_piles(), flags).set_shards(ds.records).plan()
        for i in range(4):
            M.run()
        c = M.counts()
        e, mk = M.all_edges()
        r["parity"] = {"edges": bool(np.array_equal(e, w.edges)), "marked": bool(np.array_equal(mk, w.marked)),
                       "piles": bool(np.array_equal(M.piles(), w.piles)), "nodes": bool(np.array_equal(M.seq_to_node(), w.seq_to_node)),
                       "pairs": c["n_transitive_pairs"] == w.n_pairs}
        r["counts"] = c
        if not r["parity"]["edges"]:
            r["n_edges"] = [int(e.shape[0]), int(w.edges.shape[0])]
            k = min(e.shape[0], w.edges.shape[0])
            bad = np.nonzero((e[:k] != w.edges[:k]).any(1))[0]
            r["first_bad_edges"] = bad[:5].tolist()
        elif not r["parity"]["marked"]:
            bad = np.nonzero(mk != w.marked)[0]
            r["bad_marks"] = [int(bad.shape[0]), bad[:8].tolist(), mk[bad[:8]].tolist()]
        M.close()
        M = api.Multi([0] * world)
        M.set_barrier_timeout_ms(1500)
        M.set_piles(big.flat_piles()).set_shards(big.records).plan()
        r["caps"] = M.caps.tolist() if hasattr(M, "caps") else None
        for _ in range(4):
            M.run()
        M.synchronize()
        M.event_record(0)
        for _ in range(50):
            M.run()
        M.event_record(1)
        r["ms_per_step_20Mbp"] = M.event_elapsed_ms() / 50
        e, mk = M.all_edges()
        r["parity_20Mbp"] = bool(np.array_equal(e, big_edges) and np.array_equal(mk, big_marked))
        r["rounds"] = [M.counts()["n_rounds"], M.counts()["n_final_rounds"]]
        M.use_cuda_graph(False)
        M.run()
        r["stage_ms_rank0"] = M.stage_ms(0)
        M.close()
    except Exception as exc:  # noqa: BLE001
        r["error"] = f"{type(exc).__name__}: {exc}"
    out[f"world{world}"] = r
print(json.dumps(out))
