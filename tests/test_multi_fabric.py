"""GPU suite: the multi-GPU session of the C ABI (rala_b200_multi: orchestration inside the library, exchanges as
kernels over peer memory) against the single-process oracle.  Several ranks share device 0 here, so world = 2 .. 8
runs on a one-GPU box: the same kernels, barriers and exchange buffers as one rank per GPU, only the peer pointers
are local.  tests/test_multi_gpu.py covers one process per GPU (CUDA IPC) on boxes that have them."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from rala_b200 import api, synth  # noqa: E402

pytestmark = pytest.mark.gpu

KW = dict(genome_len=3_000_000, coverage=30, read_len=9000, len_sd=2500, seed=71, noise=50, dual=True, min_ovl=900)


def _dataset():
    ds = synth.generate(**KW)
    # 5 % of the piles "have a chimeric region": their containments are not applied in the first pass (graph.cpp:470,476),
    # so the final pass (graph.cpp:831-866) has real work
    flags = (np.random.Generator(np.random.PCG64(1)).random(ds.n_reads) < 0.05).astype(np.uint8) * 2
    return ds, flags


@pytest.fixture(scope="module")
def want():
    ds, flags = _dataset()
    w = O.Pipeline(ds.records, ds.flat_piles(), flags).run()
    assert w.edges.shape[0] > 5000
    return ds, flags, w


def _check(M, w, label):
    c = M.counts()
    edges, marked = M.all_edges()
    assert np.array_equal(edges, w.edges), f"{label}: edge list"
    assert np.array_equal(marked, w.marked), f"{label}: removed-edge set"
    assert np.array_equal(M.piles(), w.piles), f"{label}: pile liveness"
    assert np.array_equal(M.seq_to_node(), w.seq_to_node), f"{label}: node ids"
    assert c["n_transitive_pairs"] == w.n_pairs and c["n_nodes"] == w.n_nodes and c["n_edges"] == w.edges.shape[0]
    return c


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_fabric_matches_oracle(world, want):
    ds, flags, w = want
    M = api.Multi([0] * world)
    M.set_piles(ds.flat_piles(), flags).set_shards(ds.records).plan()
    M.use_cuda_graph(False)
    M.run()
    c = _check(M, w, f"world {world} eager")
    assert c["n_final_candidates"] > 0, "the final containment pass should have events on this dataset"
    M.use_cuda_graph(True)
    for i in range(4):   # eager, captured + replayed, replayed, replayed
        M.run()
        _check(M, w, f"world {world} graph run {i}")
    # new inputs of the same shape through the same session: pile table re-uploaded (the other graph variant)
    for i in range(3):
        M.set_piles(ds.flat_piles(), flags)
        M.run()
        _check(M, w, f"world {world} after set_piles {i}")
    M.close()


def test_fabric_uneven_and_empty_shards(want):
    """Shards of very different sizes, one of them empty."""
    ds, flags, w = want
    n = ds.n_overlaps
    cuts = [0, 1000, 1000, n // 3 // 4 * 4, n]
    M = api.Multi([0] * 4)
    M.set_piles(ds.flat_piles(), flags).set_shards(ds.records, bounds=list(zip(cuts[:-1], cuts[1:]))).plan()
    M.run()
    _check(M, w, "uneven shards")
    M.close()


def test_fabric_small_buffers_are_reported_and_regrown(want):
    """Capacities far too small: the step must say so (never silently truncate), and a larger reservation must work."""
    ds, flags, w = want
    M = api.Multi([0] * 2)
    M.set_piles(ds.flat_piles(), flags).set_shards(ds.records)
    caps = M.default_caps()
    small = caps.copy()
    small[0], small[1], small[2], small[3] = 256, 256, 512, 1
    M.reserve(small)
    M.run().synchronize()
    need, fits = M.demand()
    assert not fits and need[0] > 256 and need[1] > 256
    with pytest.raises(api.RalaB200Error):
        M.counts()
    M.reserve(caps)
    M.run()
    _check(M, w, "after regrowing")
    M.close()
