"""The CUDA kernels evaluate the reference's three fp64 products (overlap.cpp:221-222, 236-237 and
comparable(), graph.cpp:26-29) in EXACT integer arithmetic (rala_b200/csrc/common.cuh).  These tests
pin the equivalences against IEEE doubles (numpy float64 = the reference's `double`):

    (double)s <  (double)t * 0.875              <=>  8 s < 7 t
    (double)d <  (double)L * 0.01               <=>  100 d < L
    (u32)(0.05 * (double)M)                      ==  M // 20
    a >= b * (1 - 0.12)  and  a <= b * (1 + 0.12) <=>  25 a >= 22 b  and  25 a <= 28 b
    comparable(a, b)  (both clauses)             <=>  25 a >= 22 b  and  22 a <= 25 b   (union of overlapping intervals)

For an integer left-hand side the two forms can only differ where the exact rational product is an
integer (b multiple of 25, L of 100, M of 20): away from those the product is >= 0.01 from any integer
while the rounding error is < 1e-6.  So the boundary multiples are checked EXHAUSTIVELY over the whole
u32 range, everything else by dense random sampling around the thresholds.
"""
import numpy as np

N = 1 << 32
CHUNK = 1 << 24
LO, HI = 1 - 0.12, 1 + 0.12     # exactly how graph.cpp:27-28 forms them


def _chunks(limit):
    for s in range(0, limit, CHUNK):
        yield np.arange(s, min(s + CHUNK, limit), dtype=np.uint64)


def comparable_int(a, b):
    """common.cuh comparable(): the two clauses are overlapping intervals of a, their union is the hull"""
    return (25 * a >= 22 * b) & (22 * a <= 25 * b)


def comparable_interval(a, b):
    """common.cuh comparable_interval(): a - lo <= range in u32 arithmetic"""
    lo = (22 * b + 24) // 25
    hi = np.minimum((25 * b) // 22, np.uint64(N - 1))
    return ((a - lo) & np.uint64(N - 1)) <= (hi - lo)


def comparable_f64(a, b):
    af, bf = a.astype(np.float64), b.astype(np.float64)
    return ((af >= bf * LO) & (af <= bf * HI)) | ((bf >= af * LO) & (bf <= af * HI))


def test_comparable_every_boundary_multiple_of_25():
    bad = 0
    for m in _chunks(N // 25 + 1):
        b = 25 * m
        b = b[b < N]
        m = m[: b.shape[0]]
        bf = b.astype(np.float64)
        for k, prod, ge in ((22 * m, bf * LO, True), (28 * m, bf * HI, False)):
            kf = k.astype(np.float64)
            if ge:    # a >= b * lo holds exactly from a = 22 m upwards
                bad += int((~(kf >= prod)).sum()) + int(((kf - 1) >= prod)[kf >= 1].sum())
            else:     # a <= b * hi holds exactly up to a = 28 m
                bad += int((~(kf <= prod)).sum()) + int(((kf + 1) <= prod).sum())
    assert bad == 0


def test_min_extension_every_multiple_of_20():
    for m in _chunks(N // 20 + 1):
        M = 20 * m
        M = M[M < N]
        got = np.trunc(0.05 * M.astype(np.float64)).astype(np.uint64)     # (uint32_t)(0.05 * max(...)), overlap.cpp:237
        assert np.array_equal(got, M // 20)


def test_length_tolerance_every_multiple_of_100():
    for m in _chunks(N // 100 + 1):
        L = 100 * m
        L = L[L < N]
        k = (L // 100).astype(np.float64)
        y = L.astype(np.float64) * 0.01                                    # overlap.cpp:236
        assert not (k < y).any()            # d = L/100 is NOT below the threshold (100 d < L is false)
        assert ((k - 1) < y)[k >= 1].all()  # d = L/100 - 1 is


def test_random_values_around_every_threshold():
    rng = np.random.default_rng(2026)
    for _ in range(8):
        b = rng.integers(0, N, 1 << 21, dtype=np.uint64)
        s = rng.integers(0, N, 1 << 21, dtype=np.uint64)
        # comparable: a within +-2 of both thresholds, and arbitrary pairs
        for base in ((22 * b) // 25, (28 * b) // 25, (25 * b) // 22 % N, (25 * b) // 28):
            for d in (-2, -1, 0, 1, 2):
                a = np.clip(base.astype(np.int64) + d, 0, N - 1).astype(np.uint64)
                assert np.array_equal(comparable_int(a, b), comparable_f64(a, b))
                assert np.array_equal(comparable_interval(a, b), comparable_f64(a, b))
        assert np.array_equal(comparable_int(s, b), comparable_f64(s, b))
        assert np.array_equal(comparable_interval(s, b), comparable_f64(s, b))
        # 0.875: s < t * 0.875 around t = 8 s / 7, and arbitrary pairs
        for d in (-1, 0, 1):
            t = np.clip(((8 * s) // 7).astype(np.int64) + d, 0, N - 1).astype(np.uint64)
            assert np.array_equal(s.astype(np.float64) < t.astype(np.float64) * 0.875, 8 * s < 7 * t)
        assert np.array_equal(s.astype(np.float64) < b.astype(np.float64) * 0.875, 8 * s < 7 * b)
        # 0.01 and 0.05 on arbitrary lengths
        for d in (-1, 0, 1):
            dd = np.clip((b // 100).astype(np.int64) + d, 0, None).astype(np.uint64)
            assert np.array_equal(dd.astype(np.float64) < b.astype(np.float64) * 0.01, 100 * dd < b)
        assert np.array_equal(np.trunc(0.05 * b.astype(np.float64)).astype(np.uint64), b // 20)


def test_small_lengths_exhaustive():
    """every (a, b) with both below 1200: the scale of real edge lengths' low end"""
    v = np.arange(1200, dtype=np.uint64)
    a, b = np.meshgrid(v, v, indexing="ij")
    assert np.array_equal(comparable_int(a, b), comparable_f64(a, b))
    assert np.array_equal(comparable_interval(a, b), comparable_f64(a, b))
    assert np.array_equal(a.astype(np.float64) < b.astype(np.float64) * 0.875, 8 * a < 7 * b)
    assert np.array_equal(a.astype(np.float64) < b.astype(np.float64) * 0.01, 100 * a < b)


def test_event_code_straight_line_form_vs_oracle():
    """common.cuh event_code() — trim + type of the first pass over the records as straight-line predicate
    arithmetic — and the kernels' trim() / classify(), compiled for the CPU (tests/host_event_code.cu, host-only
    nvcc build), against the oracle's fp64 trim + type on boundary-heavy random geometries: every type class,
    dead piles, garbage coordinates (u32 wrap-around), pile ends up to the 2^30 limit."""
    import json
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else shutil.which("nvcc")
    assert nvcc, "nvcc is needed to compile the host harness"
    out_dir = os.path.join(root, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "host_event_code")
    subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(root, "tests", "host_event_code.cu"), os.path.join(root, "oracle", "rala_oracle.c")],
                   check=True, capture_output=True)
    for seed in (1, 2, 3):
        r = subprocess.run([exe, "4000000", str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        res = json.loads(r.stdout.strip().splitlines()[-1])
        assert res["mismatches"] == 0
        assert all(n > 100000 for n in res["by_type"]), res   # kX, kA, kB, kAB, kBA and rejections all occur
