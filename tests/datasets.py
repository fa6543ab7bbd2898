"""Named synthetic datasets shared by the golden-vector generator and the parity tests.

`run_reference(name, workdir)` pushes a dataset through the UNMODIFIED reference front end and
hot path (oracle/_ref/rala_ref dump) and returns every stage boundary as numpy arrays;
`load_golden(name)` returns the same dictionary from the committed fixture.
"""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402  (tests are allowed to use the oracle)
from rala_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# small sets: committed as golden fixtures
GOLDEN = {
    "g_clean": dict(genome_len=1_000_000, coverage=30, read_len=10000, seed=11, min_ovl=1000),
    "g_noisy": dict(genome_len=1_000_000, coverage=45, read_len=10000, len_sd=3000, seed=12, min_ovl=1000, noise=30,
                    chimera_frac=0.03, adapter_frac=0.05, repeats=(1, 8, 4000)),
    "g_dual": dict(genome_len=400_000, coverage=30, read_len=10000, seed=13, min_ovl=1000, dual=True),
    "g_jitter": dict(genome_len=800_000, coverage=30, read_len=10000, len_sd=3000, seed=14, min_ovl=1000, noise=400),
}

# BASELINE.json configs that the reference CLI itself can run (need oracle/_ref at test time)
CONFIGS = {
    # configs[0]: 5 Mbp, 30x, 10 kbp reads
    "c1": dict(genome_len=5_000_000, coverage=30, read_len=10000, seed=1, min_ovl=1000),
    # configs[1]: 5 Mbp, 60x, repeats + chimeras + adapters
    "c2": dict(genome_len=5_000_000, coverage=60, read_len=10000, len_sd=3000, seed=2, min_ovl=1000, noise=30,
               chimera_frac=0.03, adapter_frac=0.05, repeats=(1, 12, 4000)),
}


def make(name: str) -> synth.Dataset:
    spec = GOLDEN.get(name) or CONFIGS[name]
    return synth.generate(**spec)


def _collect(prefix: str) -> dict:
    out = {}
    for path in glob.glob(prefix + ".*.u32"):
        key = os.path.basename(path)[len(os.path.basename(prefix)) + 1:-4]
        out[key] = np.fromfile(path, dtype=np.uint32)
    return out


def run_reference(name: str, workdir: str, threads: int = 0) -> dict:
    """FASTA + PAF -> rala_ref dump -> dict of arrays (+ 'summary')."""
    if not O.have_ref():
        raise RuntimeError("oracle/_ref/rala_ref is not built")
    ds = make(name)
    prefix = os.path.join(workdir, name)
    ds.write_fasta(prefix + ".fasta")
    ds.write_paf(prefix + ".paf")
    threads = threads or min(os.cpu_count() or 1, 16)
    summary = json.loads(O.ref_run(["dump", prefix + ".fasta", prefix + ".paf", prefix, threads]).strip().splitlines()[-1])
    assert summary["staged_driver_matches_reference"] is True
    d = _collect(prefix)
    d["summary"] = summary
    for ext in (".fasta", ".paf"):
        os.remove(prefix + ext)
    return d


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "summary_json"}
    d["summary"] = json.loads(str(z["summary_json"]))
    return d


class Stages:
    """Typed view over a dump dictionary."""

    def __init__(self, d: dict):
        self.d = d
        self.summary = d["summary"]
        self.records = d["in.records"].reshape(-1, 7)
        p = d["in.piles"].reshape(-1, 4)
        self.piles = np.ascontiguousarray(p[:, :2])
        self.pflags = p[:, 2].astype(np.uint8)
        self.medians = p[:, 3]
        self.hills = d["in.hills"].reshape(-1, 4)[:, :3]
        self.read_len = d["in.read_len"]
        self.hill_cov = d["stage.s1.hills"].reshape(-1, 4)[:, 3]
        self.pit_rounds = self.summary["pit_rounds"]
        self.edges = d["ref.edges"].reshape(-1, 3)
        self.removed = d["ref.removed"].astype(np.uint8)
        self.n_pairs = self.summary["transitive_pairs"]
        self.n_nodes = self.summary["nodes"]
        self.node_seq = d["ref.node_seq"]
        self.transitive_pairs = d["ref.transitive_pairs"].reshape(-1, 2)

    def lst(self, tag: str, which: str):
        return self.d[f"stage.{tag}.{which}"].reshape(-1, 7)

    def stage_piles(self, tag: str):
        return np.ascontiguousarray(self.d[f"stage.{tag}.piles"].reshape(-1, 4)[:, :2])

    def stage_pflags(self, tag: str):
        return self.d[f"stage.{tag}.piles"].reshape(-1, 4)[:, 2].astype(np.uint8)
