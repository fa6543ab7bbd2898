"""The drop-in itself: rala's own CLI and front end (unmodified reference sources, compiled where they
lie) with Graph::construct / Graph::remove_transitive_edges routed through host/graph_b200.cpp and the
C ABI into the CUDA kernels (host/_build/rala_b200), against the unmodified reference CLI
(oracle/_ref/rala) on the same FASTA + PAF.

Both binaries are built in the container that has /root/reference (host/Makefile, oracle/Makefile) and
travel to the GPU box with the snapshot; nothing here reads /root/reference at run time.
"""
import os
import re
import subprocess

import pytest

from rala_b200 import synth
from tests import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "host", "_build", "rala_b200")
REFCLI = os.path.join(ROOT, "oracle", "_ref", "rala")

needs_binaries = pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(REFCLI)),
                                    reason="host/_build/rala_b200 or oracle/_ref/rala not built (needs /root/reference at build time)")

COUNTERS = ("number of nodes", "number of edges", "number of transitive edges", "number of tips", "number of bubbles")


def _write(tmp, name, **spec):
    ds = synth.generate(**spec)
    fa, paf = os.path.join(tmp, name + ".fasta"), os.path.join(tmp, name + ".paf")
    ds.write_fasta(fa)
    ds.write_paf(paf)
    return fa, paf


def _run(binary, args, env=None):
    return subprocess.run([binary] + args, capture_output=True, text=True, env=env, timeout=900)


def _counters(stderr):
    out = {}
    for key in COUNTERS:
        m = re.search(re.escape(key) + r" = (\d+)", stderr)
        if m:
            out[key] = int(m.group(1))
    return out


@needs_binaries
def test_cli_is_the_reference_cli():
    assert _run(DROPIN, ["--version"]).stdout == _run(REFCLI, ["--version"]).stdout == "v1.0.0\n"
    assert _run(DROPIN, ["-h"]).stdout == _run(REFCLI, ["-h"]).stdout
    r = _run(DROPIN, [])
    assert r.returncode == 1 and "[rala::] error: missing input file(s)!" in r.stderr


@needs_binaries
def test_fails_loudly_without_a_device(tmp_path):
    """no CPU fallback: with no usable device the reference's print-and-exit(1) convention applies"""
    fa, paf = _write(str(tmp_path), "tiny", genome_len=60_000, coverage=20, read_len=5000, seed=5)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = _run(DROPIN, ["-t", "2", fa, paf], env=env)
    assert r.returncode == 1
    assert "[rala::Graph::construct] error: no usable B200" in r.stderr
    assert r.stdout == ""


@pytest.mark.gpu
@needs_binaries
@pytest.mark.parametrize("name", ["g_clean", "g_noisy", "g_dual"])
def test_dropin_matches_reference_cli(tmp_path, name):
    fa, paf = _write(str(tmp_path), name, **datasets.GOLDEN[name])
    threads = str(min(os.cpu_count() or 1, 16))
    # -p: everything Graph::construct leaves behind that the CLI can print (deterministic in the reference)
    ref_p, got_p = _run(REFCLI, ["-p", "-t", threads, fa, paf]), _run(DROPIN, ["-p", "-t", threads, fa, paf])
    assert ref_p.returncode == 0 and got_p.returncode == 0, got_p.stderr[-2000:]
    assert _counters(got_p.stderr) == _counters(ref_p.stderr) and "number of edges" in _counters(got_p.stderr)
    assert got_p.stdout == ref_p.stdout
    # full run: construct + simplify + contigs
    ref, ref2, got = (_run(b, ["-t", threads, fa, paf]) for b in (REFCLI, REFCLI, DROPIN))
    assert ref.returncode == 0 and got.returncode == 0, got.stderr[-2000:]
    want, have = _counters(ref.stderr), _counters(got.stderr)
    for key in ("number of nodes", "number of edges", "number of transitive edges"):
        assert have[key] == want[key], (key, have, want)
    # the reference's layout is seeded from std::random_device (graph.cpp:1113): contigs are only comparable
    # where the reference agrees with itself
    if ref.stdout == ref2.stdout and _counters(ref.stderr) == _counters(ref2.stderr):
        assert have == want
        assert got.stdout == ref.stdout


@pytest.mark.gpu
@needs_binaries
def test_dropin_on_two_ranks_matches_reference_cli(tmp_path):
    """RALA_B200_DEVICES=0,0: the multi-GPU session (rala_b200_multi_*) behind the same CLI, two ranks sharing device 0 here
    (two GPUs in production).  Clean data: the pile table is frozen, which is what the multi-GPU session covers."""
    fa, paf = _write(str(tmp_path), "g_clean", **datasets.GOLDEN["g_clean"])
    threads = str(min(os.cpu_count() or 1, 16))
    env = dict(os.environ, RALA_B200_DEVICES="0,0", RALA_B200_REPORT="1")
    ref, got = _run(REFCLI, ["-p", "-t", threads, fa, paf]), _run(DROPIN, ["-p", "-t", threads, fa, paf], env=env)
    assert ref.returncode == 0 and got.returncode == 0, got.stderr[-2000:]
    assert '"ranks": 2' in got.stderr, "the multi-GPU session did not run"
    assert _counters(got.stderr) == _counters(ref.stderr) and "number of edges" in _counters(got.stderr)
    assert got.stdout == ref.stdout
    ref, got = _run(REFCLI, ["-t", threads, fa, paf]), _run(DROPIN, ["-t", threads, fa, paf], env=env)
    assert ref.returncode == 0 and got.returncode == 0, got.stderr[-2000:]
    for key in ("number of nodes", "number of edges", "number of transitive edges"):
        assert _counters(got.stderr)[key] == _counters(ref.stderr)[key], key


@pytest.mark.gpu
@needs_binaries
def test_dropin_falls_back_to_one_device_when_piles_change(tmp_path):
    """Chimeric pits / hills need the host's pile breaking between the passes: the CLI says so and uses one device."""
    fa, paf = _write(str(tmp_path), "g_noisy", **datasets.GOLDEN["g_noisy"])
    threads = str(min(os.cpu_count() or 1, 16))
    env = dict(os.environ, RALA_B200_DEVICES="0,0", RALA_B200_REPORT="1")
    ref, got = _run(REFCLI, ["-p", "-t", threads, fa, paf]), _run(DROPIN, ["-p", "-t", threads, fa, paf], env=env)
    assert ref.returncode == 0 and got.returncode == 0, got.stderr[-2000:]
    assert "using one device instead of 2" in got.stderr and '"ranks": 1' in got.stderr
    assert _counters(got.stderr) == _counters(ref.stderr)
    assert got.stdout == ref.stdout
