"""CPU suite: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/rala_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from rala_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rala_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rala_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(api.EXPORTS) == set(names), set(api.EXPORTS) ^ set(names)
    assert lib.rala_b200_abi_version() == 2


def test_no_device_means_failure_not_fallback():
    """Without a CUDA device the context constructor fails loudly; nothing routes to a CPU path."""
    import torch
    if torch.cuda.is_available():
        return
    lib = ctypes.CDLL(build.build())
    handle = ctypes.c_void_p()
    assert lib.rala_b200_create(ctypes.byref(handle), 0) != 0 and not handle.value
    try:
        api.Context(0)
    except api.RalaB200Error:
        pass
    else:
        raise AssertionError("Context() must raise without a GPU")


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under rala_b200/ may import, include or load it."""
    pkg = os.path.join(ROOT, "rala_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include\s*[\"<].*oracle|liboracle|oracle/_ref", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_records_to_columns_layout():
    """api.records_to_columns: the host-side marshalling of rala_b200_graph_set_overlaps_columns (bit 31 of a_id =
    invalid record or an id the 32-bit device ids cannot name, bit 31 of b_id = orientation; coordinates verbatim)."""
    import numpy as np
    from rala_b200 import api
    rec = np.array([[5, 9, 10, 900, 20, 910, 0],
                    [6, 7, 0, 5000, 100, 5100, 1],           # reverse complement
                    [8, 3, 1, 2, 3, 4, 2],                   # invalid
                    [2, 4, 11, 12, 13, 14, 3],               # invalid + reverse complement
                    [0x80000001, 4, 1, 2, 3, 4, 0],          # id beyond 2^31: cannot name a pile
                    [1, 0x80000002, 1, 2, 3, 4, 1]], dtype=np.uint32)
    c = api.records_to_columns(rec)
    assert c.shape == (6, 6) and c.dtype == np.uint32 and c.flags["C_CONTIGUOUS"]
    top = 0x80000000
    assert [int(x) & top != 0 for x in c[0]] == [False, False, True, True, True, True]
    assert [int(x) & ~top & 0xFFFFFFFF for x in c[0]] == [5, 6, 8, 2, 1, 1]
    assert [int(x) >> 31 for x in c[1]] == [0, 1, 0, 1, 0, 1]
    assert [int(x) & 0x7FFFFFFF for x in c[1]] == [9, 7, 3, 4, 4, 2]
    assert np.array_equal(c[2:6].T, rec[:, 2:6])
    assert api.records_to_columns(np.zeros((0, 7), np.uint32)).shape == (6, 0)


def test_records_to_packed_round_trips_through_the_column_layout():
    """api.records_to_packed: host side of rala_b200_graph_set_overlaps_packed (12 B per record).  Expanding it on the host
    the way k_unpack_records does must give the column layout of records_to_columns for every valid record, and both spans
    0 (which Overlap::trim rejects) for every invalid one."""
    import numpy as np
    from rala_b200 import api, synth
    ds = synth.generate(400_000, 30, 9000, len_sd=2500, seed=21, noise=60, dual=True)
    rec = ds.records.copy()
    rec[0, 6] |= 2                      # invalid first record
    rec[50:53, 6] |= 2                  # a run of invalid records inside a group
    starts = np.nonzero(rec[1:, 0] != rec[:-1, 0])[0] + 1
    rec[starts[3], 6] |= 2              # invalid first record of a group
    p = api.records_to_packed(rec)
    assert p is not None and p.n == rec.shape[0] and p.nbytes < 12.5 * rec.shape[0]
    assert np.all(np.diff(p.group_end.astype(np.int64)) > 0) and int(p.group_end[-1]) == rec.shape[0]
    a = np.repeat(p.query_id, np.diff(np.concatenate([[0], p.group_end]).astype(np.int64)))
    cols = api.records_to_columns(rec)
    bad = (cols[0] >> 31) == 1
    assert bad.sum() == 5
    assert np.array_equal(a[~bad], cols[0][~bad]) and np.array_equal(p.b_id[~bad], cols[1][~bad])
    assert np.array_equal((p.a_span & 0xFFFF)[~bad], cols[2][~bad]) and np.array_equal((p.a_span >> 16)[~bad], cols[3][~bad])
    assert np.array_equal((p.b_span & 0xFFFF)[~bad], cols[4][~bad]) and np.array_equal((p.b_span >> 16)[~bad], cols[5][~bad])
    assert not p.a_span[bad].any() and not p.b_span[bad].any() and (a[bad] < ds.n_reads).all()
    # coordinates that need more than 16 bits, or ids that need bit 31: the packer declines
    big = rec.copy()
    big[10, 5] = 65536
    assert api.records_to_packed(big) is None
    assert api.records_to_packed(np.zeros((0, 7), np.uint32)) is None
