"""Stage-by-stage replay of a reference dump (tests/datasets.py:Stages) through any implementation
of the session interface (the CUDA session `rala_b200.api.Graph`, or the oracle pipeline).  The host
Pile operations the reference performs between the passes are replayed from the dump's pile tables."""
from __future__ import annotations

import numpy as np


def assert_same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    if not np.array_equal(a, b):
        bad = np.nonzero(a.reshape(a.shape[0], -1) != b.reshape(b.shape[0], -1))[0]
        raise AssertionError(f"{what}: {len(set(bad.tolist()))} differing rows, first at {bad[0]}: {a[bad[0]]} != {b[bad[0]]}")


def replay_cuda(G, st):
    """G: rala_b200.api.Graph; st: Stages.  Asserts bit-exact equality at every stage boundary."""
    G.set_piles(st.piles, st.pflags).set_hills(st.hills).set_overlaps(st.records)
    G.classify()
    ovl, inl = G.lists()
    assert_same(ovl, st.lst("s1", "ovl"), "s1 overlaps")
    assert_same(inl, st.lst("s1", "int"), "s1 internals")
    assert_same(G.hill_coverage(), st.hill_cov, "chimeric hill coverage")
    assert_same(G.piles(), st.stage_piles("s1"), "s1 piles")
    assert_same(G.connections(), st.lst("s1", "ovl")[:, :2], "connections")

    G.set_piles(st.stage_piles("s2"), st.stage_pflags("s2")).retrim()
    ovl, inl = G.lists()
    assert_same(ovl, st.lst("s2", "ovl"), "s2 overlaps")
    assert_same(inl, st.lst("s2", "int"), "s2 internals")

    changed = True
    for r in range(st.pit_rounds):
        tag = f"s3r{r}"
        G.set_piles(st.stage_piles(tag), st.stage_pflags(tag))
        changed = G.retrim_promote()
        ovl, inl = G.lists()
        assert_same(ovl, st.lst(tag, "ovl"), f"{tag} overlaps")
        assert_same(inl, st.lst(tag, "int"), f"{tag} internals")
    assert not changed, "the reference left the pit loop here"

    G.finalize()
    ovl, inl = G.lists()
    assert_same(ovl, st.lst("s4", "ovl"), "s4 overlaps")
    assert_same(inl, st.lst("s4", "int"), "s4 internals")
    assert_same(G.piles(), st.stage_piles("s4"), "s4 piles")

    G.build()
    c = G.counts()
    assert c["n_nodes"] == st.n_nodes
    assert_same(G.edges(), st.edges, "edge list")
    s2n = G.seq_to_node()
    alive = s2n != 0xFFFFFFFF
    assert_same(np.nonzero(alive)[0].astype(np.uint32), st.node_seq[0::2], "node -> sequence id")
    assert_same(s2n[alive], np.arange(0, st.n_nodes, 2, dtype=np.uint32), "sequence id -> node")

    n_pairs = G.remove_transitive_edges()
    assert n_pairs == st.n_pairs
    assert_same(G.removed, st.removed, "removed edge set")
    assert_same(G.transitive_edges, st.transitive_pairs, "transitive_edges_")
    return G.counts()
