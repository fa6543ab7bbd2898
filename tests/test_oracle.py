"""CPU suite: the plain-C oracle against the golden vectors produced by the compiled reference
(tests/golden/*.npz) and, when oracle/_ref is built, against the reference itself."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import datasets
from tests.replay import assert_same

GOLDEN = list(datasets.GOLDEN)


def replay_oracle(st):
    P = O.Pipeline(st.records, st.piles, st.pflags, st.hills)
    P.classify()
    assert_same(P.ovl, st.lst("s1", "ovl"), "s1 overlaps")
    assert_same(P.int, st.lst("s1", "int"), "s1 internals")
    assert_same(P.piles, st.stage_piles("s1"), "s1 piles")
    assert_same(P.hill_cov, st.hill_cov, "hill coverage")
    P.set_piles(st.stage_piles("s2")).retrim()
    assert_same(P.ovl, st.lst("s2", "ovl"), "s2 overlaps")
    assert_same(P.int, st.lst("s2", "int"), "s2 internals")
    changed = True
    for r in range(st.pit_rounds):
        P.set_piles(st.stage_piles(f"s3r{r}"))
        changed = P.retrim_promote()
        assert_same(P.ovl, st.lst(f"s3r{r}", "ovl"), f"s3r{r} overlaps")
        assert_same(P.int, st.lst(f"s3r{r}", "int"), f"s3r{r} internals")
    assert not changed
    P.final_containment()
    assert_same(P.ovl, st.lst("s4", "ovl"), "s4 overlaps")
    assert_same(P.int, st.lst("s4", "int"), "s4 internals")
    assert_same(P.piles, st.stage_piles("s4"), "s4 piles")
    P.build_edges()
    assert P.n_nodes == st.n_nodes
    assert_same(P.edges, st.edges, "edges")
    P.transitive()
    assert P.n_pairs == st.n_pairs
    assert_same(P.marked, st.removed, "removed")
    assert_same(O.transitive_pairs(P.edges, P.marked), st.transitive_pairs, "transitive_edges_")
    off, ids = O.adjacency(P.n_nodes, P.edges, None, 0)
    assert_same(off, st.d["ref.suffix_off"], "suffix offsets")
    assert_same(ids, st.d["ref.suffix_ids"], "suffix ids (ascending edge id)")
    off, ids = O.adjacency(P.n_nodes, P.edges, P.marked, 1)
    assert_same(off, st.d["ref.after.prefix_off"], "prefix offsets after removal")
    assert_same(ids, st.d["ref.after.prefix_ids"], "prefix ids after removal (stable)")
    return P


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_golden_stage_by_stage(name):
    replay_oracle(datasets.Stages(datasets.load_golden(name)))


def test_golden_sets_cover_the_hard_cases():
    noisy = datasets.Stages(datasets.load_golden("g_noisy"))
    assert noisy.hills.shape[0] > 0 and (noisy.pflags & 2).sum() > 0        # hills + chimeric regions
    assert noisy.lst("s1", "int").shape[0] > 0                               # internals
    assert noisy.pit_rounds >= 2                                             # the pit loop iterated
    assert noisy.lst("s3r0", "ovl").shape[0] > noisy.lst("s2", "ovl").shape[0] - 50   # promotions happened
    dual = datasets.Stages(datasets.load_golden("g_dual"))
    e = dual.edges
    key = e[:, 0].astype(np.uint64) << np.uint64(32) | e[:, 1].astype(np.uint64)
    assert np.unique(key).shape[0] < key.shape[0]                            # parallel edges
    jit = datasets.Stages(datasets.load_golden("g_jitter"))
    m, ed = jit.removed, jit.edges
    # strand-asymmetric pairs exist: T(e) != T(e^1) for some marked pair (checked with the oracle's T)
    assert m[0::2].sum() == m[1::2].sum() == jit.n_pairs


def test_trim_type_known_answers():
    kat = np.load(os.path.join(datasets.GOLDEN_DIR, "kat.npz"))["trimtype"]
    for row in kat:
        rec = [0, 1, row[0], row[1], row[2], row[3], row[4]]
        ok, r, t = O.trim_type(rec, (row[5], row[6]), (row[7], row[8]))
        assert int(ok) == row[9], row
        if ok:
            assert r[2:6].tolist() == row[10:14].tolist(), row
            assert t == row[14], row


def test_comparable_known_answers():
    kat = np.load(os.path.join(datasets.GOLDEN_DIR, "kat.npz"))["comparable"]
    for a, b, want in kat.tolist():
        assert int(O.comparable(a, b)) == want, (a, b)
    assert O.comparable(88, 100) and not O.comparable(87, 100)


@pytest.mark.parametrize("tag", ["rand", "chain", "hub"])
def test_transitive_injected_graphs(tag):
    kat = np.load(os.path.join(datasets.GOLDEN_DIR, "kat.npz"))
    marked, n_pairs = O.transitive(int(kat[tag + ".n_nodes"][0]), kat[tag + ".edges"])
    assert n_pairs == int(kat[tag + ".n_pairs"][0])
    assert_same(marked, kat[tag + ".removed"], tag)


def test_empty_inputs():
    P = O.Pipeline(np.zeros((0, 7), np.uint32), np.array([[15, 9985]], np.uint32))
    P.run()
    assert P.n_nodes == 2 and P.edges.shape[0] == 0 and P.n_pairs == 0
    marked, n = O.transitive(0, np.zeros((0, 3), np.uint32))
    assert n == 0 and marked.shape[0] == 0


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_hotpath_frozen_piles(tmp_path):
    """Binary-input path of the harness (what bench.py --impl reference times) == oracle.Pipeline.run()."""
    from rala_b200 import synth
    ds = synth.generate(600_000, 30, 10000, len_sd=2500, seed=21, noise=60, dual=True)
    piles = ds.flat_piles()
    rng = np.random.Generator(np.random.PCG64(5))
    flags = (rng.random(ds.n_reads) < 0.05).astype(np.uint8) * 2     # some containers "have a chimeric region"
    prefix = str(tmp_path / "hp")
    O.write_hotpath_inputs(prefix, ds.records, piles, flags, None, ds.read_len)
    O.ref_run(["hotpath", prefix, prefix])
    P = O.Pipeline(ds.records, piles, flags).run()
    assert_same(P.edges, O.load_u32(prefix + ".stage.edges.u32", 3), "edges")
    assert_same(P.marked, O.load_u32(prefix + ".stage.removed.u32").astype(np.uint8), "removed")
    assert_same(P.ovl, O.load_u32(prefix + ".stage.s4.ovl.u32", 7), "final overlaps")


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_config5_repeat_hubs_reference_builds_degrees_above_2000(tmp_path):
    """BASELINE.json configs[4]: the high-repeat dataset pushed through the REFERENCE's own hot path (trim, type, ordered
    containment, Node / Edge creation, remove_transitive_edges) keeps node degrees > 2 000, and the oracle agrees with
    the reference on it bit for bit."""
    from rala_b200 import synth
    ds = synth.generate_repeat_hubs(genome_len=3_000_000, n_hubs=3)
    piles = ds.flat_piles()
    prefix = str(tmp_path / "c5")
    O.write_hotpath_inputs(prefix, ds.records, piles, None, None, ds.read_len)
    O.ref_run(["hotpath", prefix, prefix])
    edges = O.load_u32(prefix + ".stage.edges.u32", 3)
    removed = O.load_u32(prefix + ".stage.removed.u32").astype(np.uint8)
    deg = np.bincount(edges[:, 0])
    assert deg.max() > 2000 and int((deg > 2000).sum()) == 3, f"max out-degree {deg.max()}"
    hub = int(np.argmax(deg))
    out = np.nonzero(edges[:, 0] == hub)[0]
    assert 1000 < int(removed[out].sum()) < out.shape[0], "some, not all, of a hub's edges are transitive"
    P = O.Pipeline(ds.records, piles).run()
    assert_same(P.edges, edges, "edges")
    assert_same(P.marked, removed, "removed")
    assert P.int.shape[0] > 1000, "repeat-induced overlaps between different copies are internal"
