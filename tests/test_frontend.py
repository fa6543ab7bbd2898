"""SURVEY.md 8(f) row 1, first half: the duplicate filter of the reference's front end (Graph::initialize,
graph.cpp:273-303 driven by :340-361).

CPU: the oracle's literal restatement of the loops against is_valid_overlap_ of the compiled reference
(tests/golden/dups.npz, made by `rala_ref dupfilter`; re-made on the spot when oracle/_ref is here), and the
kernel's per-record decision function, compiled for the CPU, against that oracle.
GPU (-m gpu): rala_b200_filter_duplicates through the C ABI against both."""
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle as O
from rala_b200 import synth
from tests import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(datasets.GOLDEN_DIR, "dups.npz")


def _golden():
    z = np.load(GOLDEN)
    return z["a"], z["b"], z["length"], z["valid"]


def test_oracle_duplicate_filter_matches_the_reference_golden():
    a, b, ln, valid = _golden()
    assert 0 < int(valid.sum()) < a.shape[0] // 2          # most records of these groups are duplicates
    assert int(((a & 0x80000000) != 0).sum()) > 100        # unresolved names occur, also inside groups
    assert int((a == (b & 0x7FFFFFFF)).sum()) > 100        # and self overlaps
    assert np.array_equal(O.filter_duplicates(a, b, ln), valid)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (compiled reference) not present")
def test_oracle_duplicate_filter_matches_the_reference_itself():
    """Other seeds than the committed fixture, run through the reference's own initialize() right here."""
    for seed, big in ((5, 0), (6, 2)):
        d = synth.generate_duplicate_groups(n_reads=200, n_groups=300, big_groups=big, big_size=800, seed=seed)
        with tempfile.TemporaryDirectory() as tmp:
            d.write_fasta(os.path.join(tmp, "r.fasta"))
            d.write_paf(os.path.join(tmp, "o.paf"))
            O.ref_run(["dupfilter", os.path.join(tmp, "r.fasta"), os.path.join(tmp, "o.paf"), os.path.join(tmp, "o.u32"), 2])
            rows = np.fromfile(os.path.join(tmp, "o.u32"), dtype=np.uint32).reshape(-1, 5)
        a, b, ln = d.columns()
        assert np.array_equal(O.filter_duplicates(a, b, ln), rows[:, 4].astype(np.uint8))


def test_filter_decision_function_vs_oracle_on_the_cpu():
    """common.cuh duplicate_filter_keeps() (one thread's work in k_filter_duplicates: 'the last longest record per
    query group and target survives'), compiled for the CPU, against the oracle's nested loops on adversarial short
    files: tiny id and length pools, unresolved records anywhere, queries that come back."""
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else shutil.which("nvcc")
    assert nvcc, "nvcc is needed to compile the host harness"
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "host_dupfilter")
    subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_dupfilter.cu"), os.path.join(ROOT, "oracle", "rala_oracle.c")],
                   check=True, capture_output=True)
    for seed in (1, 2):
        r = subprocess.run([exe, "100000", str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        res = json.loads(r.stdout.strip().splitlines()[-1])
        assert res["mismatches"] == 0 and res["kept"] > 100000 and res["records"] > 2 * res["kept"]


@pytest.mark.gpu
def test_filter_duplicates_on_the_device(ctx):
    a, b, ln, valid = _golden()
    before = ctx.launch_count
    got = ctx.filter_duplicates(a, b, ln)
    assert ctx.launch_count == before + 1
    assert np.array_equal(got, valid), "device filter differs from the reference's is_valid_overlap_"
    # orientation bits in b must not matter; a record must not see across an unresolved record's id
    assert np.array_equal(ctx.filter_duplicates(a, b | np.uint32(0x80000000), ln), valid)
    # bigger, against the oracle: ordinary groups, and groups of thousands of records (repeats)
    for kw in (dict(n_reads=5000, n_groups=20000, seed=21), dict(n_reads=400, n_groups=300, big_groups=40, big_size=4000, seed=22)):
        d = synth.generate_duplicate_groups(**kw)
        a2, b2, l2 = d.columns()
        got2, ms = ctx.filter_duplicates(a2, b2, l2, with_time=True)
        assert np.array_equal(got2, O.filter_duplicates(a2, b2, l2))
        assert ms > 0.0
    # edges of the input space
    assert ctx.filter_duplicates(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32)).shape == (0,)
    one = ctx.filter_duplicates(np.array([3], np.uint32), np.array([4], np.uint32), np.array([10], np.uint32))
    assert one.tolist() == [1]
    ghosts = ctx.filter_duplicates(np.full(100, 0x80000001, np.uint32), np.zeros(100, np.uint32), np.ones(100, np.uint32))
    assert not ghosts.any()
    same = ctx.filter_duplicates(np.full(1000, 7, np.uint32), np.full(1000, 9, np.uint32), np.full(1000, 500, np.uint32))
    assert same.sum() == 1 and same[-1] == 1              # all tied: the LAST one stays (graph.cpp:299-303)
