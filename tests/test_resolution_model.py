"""CPU model of the DISTRIBUTED ordered-containment resolution (rala_b200/csrc/fabric.cu: k_fabric_resolve), in numpy.

The reference kills piles while it streams the overlap file (graph.cpp:469-480): a containment event (victim v, container c,
time t) fires iff both piles are still alive at t.  The CUDA path computes the same death times as a fixed point over
MONOTONE per-pile states (open with a lower bound / settled with a death time); on several GPUs every rank resolves the
piles it owns, pushes their state changes to all replicas, and reads foreign states as they were at the last exchange.
This model runs that scheme in lock-step rounds (an upper bound on what the asynchronous kernels need) and checks, for
1 .. 8 ranks and both ownership schemes that were tried, that it ends in the sequential answer — the property the kernels'
parity tests rely on — and that the block-cyclic ownership the library uses is balanced where the id ranges are not."""
import numpy as np
import pytest

from rala_b200 import synth

INF = np.iinfo(np.int64).max


def synthetic_events(n_piles, n_events, seed):
    """Random containment events with the structure that matters: containers are reads that lie NEAR the victim on the
    genome, read ids are shuffled against genome position, and times are file positions of a file that lists every pair under
    its lower id (so an event's time grows with min(victim id, container id))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    read_at = rng.permutation(n_piles)                               # genome rank -> read id
    vp = rng.integers(0, n_piles, n_events)
    cp = (vp + rng.integers(1, 40, n_events) * rng.choice([-1, 1], n_events)) % n_piles
    keep = vp != cp
    v, c = read_at[vp[keep]], read_at[cp[keep]]
    q = np.minimum(v, c)
    order = np.lexsort((rng.random(v.shape[0]), q))                  # grouped by query id, arbitrary order inside a group
    v, c = v[order], c[order]
    return v.astype(np.int64), c.astype(np.int64), np.arange(v.shape[0], dtype=np.int64)


def sequential(v, c, t, n):
    alive = np.ones(n, bool)
    D = np.full(n, INF)
    for i in range(v.shape[0]):
        if alive[v[i]] and alive[c[i]]:
            alive[v[i]] = False
            D[v[i]] = t[i]
    return D


def distributed(v, c, t, n, world, owner):
    """Rounds of: every rank sweeps its open victims to a local fixed point (own states live, foreign states as of the last
    exchange), then all states are exchanged.  Returns death times, rounds, open victims per rank at the start of each round."""
    o = np.lexsort((t, v))
    sv, sc, st = v[o], c[o], t[o]
    start, end = np.searchsorted(sv, np.arange(n)), np.searchsorted(sv, np.arange(n), side="right")
    settled = start == end
    D = np.full(n, INF)
    L = np.zeros(n, np.int64)
    L[~settled] = st[start[~settled]]           # initial lower bound: the earliest event (k_fabric_prepare)
    ptr = start.copy()
    snap = (settled.copy(), D.copy(), L.copy())
    history = []
    while not settled.all():
        history.append(np.bincount(owner[~settled], minlength=world).tolist())
        assert len(history) < 200, "the resolution does not converge"
        while True:
            u = np.nonzero(~settled)[0]
            done = ptr[u] == end[u]
            changed = bool(done.any())
            settled[u[done]] = True             # every event found its container dead: never killed
            u = u[~done]
            if u.size == 0:
                break
            cc, tt = sc[ptr[u]], st[ptr[u]]
            same = owner[cc] == owner[u]
            c_set = np.where(same, settled[cc], snap[0][cc])
            c_D = np.where(same, D[cc], snap[1][cc])
            c_L = np.where(same, L[cc], snap[2][cc])
            fire = (c_set & (c_D > tt)) | (~c_set & (c_L > tt))      # container alive at tt for sure
            dead = c_set & (c_D <= tt)                                 # container died first: the event never fires
            settled[u[fire]] = True
            D[u[fire]] = tt[fire]
            ptr[u[dead]] += 1
            blocked = ~fire & ~dead
            raised = np.maximum(L[u[blocked]], tt[blocked])
            changed |= bool(fire.any() or dead.any() or (raised != L[u[blocked]]).any())
            L[u[blocked]] = raised
            if not changed:
                break
        snap = (settled.copy(), D.copy(), L.copy())
    return D, len(history), history


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("scheme", ["cyclic32", "range"])
def test_distributed_resolution_reaches_the_sequential_death_times(world, scheme):
    n = 20_000
    v, c, t = synthetic_events(n, 60_000, seed=7)
    want = sequential(v, c, t, n)
    assert (want != INF).sum() > n // 4
    per = (n + world - 1) // world
    owner = (np.arange(n) // 32) % world if scheme == "cyclic32" else np.arange(n) // per
    got, rounds, history = distributed(v, c, t, n, world, owner)
    assert np.array_equal(got, want)
    assert rounds <= (1 if world == 1 else 40)
    if world == 8 and len(history) > 2:
        spread = max(history[1]) / max(1, min(history[1]))
        if scheme == "cyclic32":
            assert spread < 1.5, f"block-cyclic ownership should leave every rank about the same work: {history[1]}"


def test_block_cyclic_ownership_balances_what_id_ranges_do_not():
    """Why piles are owned in cyclic blocks of 32: a pair is listed under its lower id, so low ids settle in the first sweep
    while high ids wait for them.  With id ranges the last rank starts round 2 with several times the open victims of the
    first; with cyclic blocks every rank has the same (profiles/r02g vs r02h: 1310 -> 1083 us of compute on the slow rank)."""
    n, world = 40_000, 8
    v, c, t = synthetic_events(n, 150_000, seed=9)
    per = (n + world - 1) // world
    _, _, h_range = distributed(v, c, t, n, world, np.arange(n) // per)
    _, _, h_cyc = distributed(v, c, t, n, world, (np.arange(n) // 32) % world)
    assert len(h_range) > 1 and len(h_cyc) > 1
    assert max(h_range[1]) > 2 * min(h_range[1])
    assert max(h_cyc[1]) < 1.3 * min(h_cyc[1])


def test_events_of_a_real_batch_resolve_the_same_way():
    """The same check on the containment structure of a synthetic read set: death times = which piles the oracle kills."""
    from oracle import oracle as O
    ds = synth.generate(600_000, 30, 10000, seed=33)
    piles = ds.flat_piles()
    P = O.Pipeline(ds.records, piles).classify()
    # events from the oracle's own types, in file order
    _, types = O.trim_type_batch(ds.records[:40_000], piles)
    rec = ds.records[:40_000]
    isb, isa = types == O.KB, types == O.KA
    v = np.concatenate([rec[isb, 0], rec[isa, 1]]).astype(np.int64)
    c = np.concatenate([rec[isb, 1], rec[isa, 0]]).astype(np.int64)
    t = np.concatenate([np.nonzero(isb)[0], np.nonzero(isa)[0]]).astype(np.int64)
    o = np.argsort(t, kind="stable")
    v, c, t = v[o], c[o], t[o]
    n = ds.n_reads
    want = sequential(v, c, t, n)
    for world in (2, 8):
        got, _, _ = distributed(v, c, t, n, world, (np.arange(n) // 32) % world)
        assert np.array_equal(got, want)
    # and the sequential model itself is the reference's loop: on the whole file it kills exactly the piles the oracle kills
    _, types_all = O.trim_type_batch(ds.records, piles) if ds.n_overlaps <= 200_000 else (None, None)
    if types_all is not None:
        isb, isa = types_all == O.KB, types_all == O.KA
        v = np.concatenate([ds.records[isb, 0], ds.records[isa, 1]]).astype(np.int64)
        c = np.concatenate([ds.records[isb, 1], ds.records[isa, 0]]).astype(np.int64)
        t = np.concatenate([np.nonzero(isb)[0], np.nonzero(isa)[0]]).astype(np.int64)
        o = np.argsort(t, kind="stable")
        D = sequential(v[o], c[o], t[o], n)
        assert np.array_equal(D != INF, P.piles[:, 1] == 0)
