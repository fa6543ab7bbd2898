"""Regenerates tests/golden/*.npz from the compiled, unmodified reference (oracle/_ref/rala_ref).

Run in the build container (where /root/reference exists):
    make -C oracle && python tests/golden/make_golden.py

The fixtures hold, for every named dataset in tests/datasets.py:GOLDEN, the hot path's inputs as
the reference's own front end (Graph::initialize) produced them, the state at every stage
boundary of the staged driver (which `rala_ref dump` proves equal to Graph::construct), the
reference's edge list, removed-edge set, adjacency and transitive_edges_.  Plus known-answer
vectors for the pure functions (trim/type, comparable) and injected graphs for the transitive pass.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402
from rala_b200 import synth  # noqa: E402
from tests import datasets  # noqa: E402

KEEP = ("in.", "stage.s", "ref.edges", "ref.removed", "ref.node_seq", "ref.transitive_pairs", "ref.piles",
        "ref.suffix", "ref.after.prefix")


def trimtype_vectors(seed: int = 3, n: int = 4000) -> np.ndarray:
    """Rows: ab ae bb be ori pa0 pa1 pb0 pb1 | ok ab' ae' bb' be' type (type -1 -> 255)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = []
    for _ in range(n):
        la, lb = rng.integers(1500, 20000, 2)
        pa0 = int(rng.integers(0, 600)); pa1 = int(la - rng.integers(0, 600))
        pb0 = int(rng.integers(0, 600)); pb1 = int(lb - rng.integers(0, 600))
        kind = rng.integers(0, 4)
        if kind == 0:      # arbitrary
            ab, ae = sorted(rng.integers(0, la + 1, 2).tolist()); bb, be = sorted(rng.integers(0, lb + 1, 2).tolist())
        elif kind == 1:    # suffix of a / prefix of b, near-equal spans
            span = int(rng.integers(84, min(la, lb)))
            ab, ae = int(la - span), int(la); bb, be = 0, span + int(rng.integers(-20, 21))
        elif kind == 2:    # containment-like
            span = int(rng.integers(84, min(la, lb)))
            ab = int(rng.integers(0, la - span + 1)); ae = ab + span; bb = int(rng.integers(0, 40)); be = int(lb - rng.integers(0, 40))
        else:              # tiny overhang boundary cases
            span = int(rng.integers(700, min(la, lb) - 100))
            ab = int(rng.integers(0, 100)); ae = ab + span; bb = int(rng.integers(0, 100)); be = bb + span
        be = max(0, min(int(be), int(lb)))
        rows.append([ab, ae, bb, be, int(rng.integers(0, 2)), pa0, pa1, pb0, pb1])
    # hand-made boundaries: span*8 == 7*(span+overhang); |delta| == min_extension; a_begin == b_begin
    rows += [
        [700, 10000, 0, 9300, 0, 0, 10000, 0, 10000],
        [1000, 8000, 0, 7000, 0, 0, 8000, 0, 8000],          # 7000 vs 8000*0.875
        [1001, 8000, 0, 6999, 0, 0, 8000, 0, 8000],
        [500, 10000, 0, 9500, 0, 0, 10000, 0, 10000],        # offset == 0.05*len
        [499, 10000, 0, 9501, 0, 0, 10000, 0, 10000],
        [501, 10000, 0, 9499, 0, 0, 10000, 0, 10000],
        [300, 9000, 300, 9000, 0, 0, 10000, 0, 10000],       # a_begin == b_begin
        [300, 9000, 1000, 9700, 1, 0, 10000, 0, 10000],
        [15, 9985, 15, 9985, 1, 15, 9985, 15, 9985],
        [0, 5000, 5000, 10000, 0, 15, 9985, 15, 9985],
        [0, 5000, 5000, 10000, 1, 15, 9985, 15, 9985],
        [0, 98, 0, 98, 0, 15, 9985, 15, 9985],               # < 84 after clipping
        [0, 99, 0, 99, 0, 15, 9985, 15, 9985],
        [100, 9000, 2000, 10900, 0, 3000, 9000, 0, 11000],   # heavy clipping of a shifts b
        [100, 9000, 2000, 10900, 1, 3000, 9000, 0, 11000],
    ]
    q = np.asarray(rows, dtype=np.uint32)
    text = "\n".join(" ".join(str(v) for v in r) for r in q.tolist()) + "\n"
    out = np.asarray([[int(x) for x in ln.split()] for ln in O.ref_run(["trimtype"], stdin=text).splitlines()],
                     dtype=np.int64)
    out[:, 5] = np.where(out[:, 5] < 0, 255, out[:, 5])
    return np.concatenate([q, out.astype(np.uint32)], axis=1)


def comparable_vectors(seed: int = 4) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = [[88, 100], [87, 100], [89, 100], [112, 100], [113, 100], [100, 88], [100, 112], [100, 114], [0, 0], [0, 1],
            [1, 0], [4294967295, 4294967295], [4294967295, 3779571220], [22, 25], [25, 22], [25, 28], [28, 25]]
    b = rng.integers(1, 60000, 3000)
    for bb in b.tolist():
        for f in (0.88, 1.12, 1 / 0.88, 1 / 1.12):
            c = int(bb * f)
            rows += [[c - 1, bb], [c, bb], [c + 1, bb]]
    q = np.asarray(rows, dtype=np.uint32)
    text = "\n".join(f"{a} {b}" for a, b in q.tolist()) + "\n"
    res = np.asarray([int(x) for x in O.ref_run(["comparable"], stdin=text).split()], dtype=np.uint32)
    return np.concatenate([q, res[:, None]], axis=1)


def injected_graphs(tmp: str) -> dict:
    """Graphs fed straight to the reference's remove_transitive_edges."""
    out = {}
    rng = np.random.Generator(np.random.PCG64(9))

    def run(tag, n_nodes, edges):
        edges = np.ascontiguousarray(edges, dtype=np.uint32)
        path = os.path.join(tmp, tag + ".edges.u32")
        edges.tofile(path)
        summ = json.loads(O.ref_run(["transitive", path, n_nodes, os.path.join(tmp, tag)]).strip())
        out[tag + ".edges"] = edges
        out[tag + ".n_nodes"] = np.asarray([n_nodes], dtype=np.uint32)
        out[tag + ".removed"] = np.fromfile(os.path.join(tmp, tag + ".removed.u32"), dtype=np.uint32).astype(np.uint8)
        out[tag + ".n_pairs"] = np.asarray([summ["transitive_pairs"]], dtype=np.uint32)

    # random bidirected graph with parallel edges, asymmetric lengths, self-pair edges (a -> a^1)
    n_reads = 300
    rows = []
    for _ in range(2500):
        a, b = rng.integers(0, 2 * n_reads, 2).tolist()
        if a >> 1 == b >> 1 and rng.random() < 0.9:
            continue
        l1, l2 = rng.integers(50, 3000, 2).tolist()
        rows.append((a, b, l1)); rows.append((b ^ 1, a ^ 1, l2))
        if rng.random() < 0.1:   # parallel duplicate with another length
            rows.append((a, b, l1 + int(rng.integers(-30, 400)) % 5000)); rows.append((b ^ 1, a ^ 1, l2))
    run("rand", 2 * n_reads, np.asarray(rows))
    # local "overlap-like" graph: node i -> i+1..i+k with consistent lengths (+ jitter so T(e) != T(e^1) sometimes)
    rows = []
    pos = np.cumsum(rng.integers(200, 1500, 400))
    for i in range(400):
        for j in range(i + 1, min(400, i + 9)):
            d = int(pos[j] - pos[i])
            if d > 8000:
                break
            rows.append((2 * i, 2 * j, d + int(rng.integers(-120, 121)))); rows.append((2 * j + 1, 2 * i + 1, d + int(rng.integers(-120, 121))))
    run("chain", 800, np.asarray(rows))
    n_nodes, e = synth.hub_graph(n_hubs=2, spokes=700, links_per_spoke=5, seed=5)
    run("hub", n_nodes, e)
    return out


def duplicate_filter_vectors(tmp: str) -> dict:
    """is_valid_overlap_ of the reference's own Graph::initialize (graph.cpp:273-303, :340-361) on query groups full of
    repeated targets, length ties, self overlaps, unresolved names and one 1 500-record group (rala_ref dupfilter)."""
    d = synth.generate_duplicate_groups(big_groups=1, big_size=1500)
    d.write_fasta(os.path.join(tmp, "dups.fasta"))
    d.write_paf(os.path.join(tmp, "dups.paf"))
    out = os.path.join(tmp, "dups.u32")
    summ = json.loads(O.ref_run(["dupfilter", os.path.join(tmp, "dups.fasta"), os.path.join(tmp, "dups.paf"), out, 4]).strip().splitlines()[-1])
    rows = np.fromfile(out, dtype=np.uint32).reshape(-1, 5)
    assert rows.shape[0] == d.n == summ["records"]
    a, b, ln = d.columns()
    known = rows[:, 3] == 1
    assert np.array_equal(known, (a & 0x80000000) == 0) and np.array_equal(rows[known, 0], a[known])
    assert np.array_equal(rows[known, 1], b[known]) and np.array_equal(rows[:, 2], ln)
    return {"a": a, "b": b, "length": ln, "valid": rows[:, 4].astype(np.uint8)}


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle"
    with tempfile.TemporaryDirectory() as tmp:
        if "--dups-only" in sys.argv:
            np.savez_compressed(os.path.join(datasets.GOLDEN_DIR, "dups.npz"), **duplicate_filter_vectors(tmp))
            print("dups", "%.2f MB" % (os.path.getsize(os.path.join(datasets.GOLDEN_DIR, "dups.npz")) / 1e6))
            return
        np.savez_compressed(os.path.join(datasets.GOLDEN_DIR, "dups.npz"), **duplicate_filter_vectors(tmp))
        for name in datasets.GOLDEN:
            d = datasets.run_reference(name, tmp)
            summ = d.pop("summary")
            keep = {k: v for k, v in d.items() if k.startswith(KEEP)}
            np.savez_compressed(os.path.join(datasets.GOLDEN_DIR, name + ".npz"), summary_json=json.dumps(summ), **keep)
            print(name, summ, "%.2f MB" % (os.path.getsize(os.path.join(datasets.GOLDEN_DIR, name + ".npz")) / 1e6))
        np.savez_compressed(os.path.join(datasets.GOLDEN_DIR, "kat.npz"), trimtype=trimtype_vectors(),
                            comparable=comparable_vectors(), **injected_graphs(tmp))
        print("kat", "%.2f MB" % (os.path.getsize(os.path.join(datasets.GOLDEN_DIR, "kat.npz")) / 1e6))


if __name__ == "__main__":
    main()
