#!/usr/bin/env python
"""bench.py — graph edges/sec (construct + transitive reduction), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3]

A step is one pass of the hot path (classify -> ordered containment -> re-trim -> final containment ->
edge list + CSR -> transitive reduction) over one batch of synthetic overlaps:
  N = 1 : BASELINE.json configs[2] — 100 Mbp genome, 40x, 10 kbp reads (~400k reads, ~14M overlap
          records listed once per pair), read ids shuffled, flat pile table [15, len-15).
  N > 1 : the same shape per GPU (weak scaling): genome = N x 100 Mbp, records sharded by contiguous
          range, CSR replicated by all-gather, marks merged by all-reduce (rala_b200/multi.py).
`value` = E / t with the inputs resident in HBM; `e2e` = the same through the C ABI with HOST (pinned)
buffers: H2D of records + piles and D2H of edges + marks inside the timed region.
`--impl reference` times the reference's own CPU implementation of the path (oracle/_ref/rala_ref
hotpath: the unmodified reference's Overlap::trim/type, Graph::Node/Edge and
remove_transitive_edges driven in memory) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (genome bp per GPU, coverage, read length, description)
    "c3": (100_000_000, 40, 10000, "configs[2]: synthetic 100 Mbp genome, 40x 10 kbp reads (~400k reads), shuffled read ids"),
    "c1": (5_000_000, 30, 10000, "configs[0]: synthetic 5 Mbp genome, 30x 10 kbp reads (~15k reads)"),
    "c2": (5_000_000, 60, 10000, "configs[1]: synthetic 5 Mbp genome, 60x, injected repeats, chimeras, adapters, +-30 bp noise (pile trimming, promotion)"),
    "c5": (20_000_000, 30, 10000, "configs[4]: 20 Mbp genome, 30x, segmental repeats: 8 hub reads with ~2 600 spokes each (node degrees > 2 000)"),
    "c4s": (387_500_000, 30, 10000, "configs[3] shard: 3.1 Gbp / 8 per GPU, 30x 10 kbp reads (31.4 M overlap records per GPU, pairs listed once)"),
}
K1_BYTES_PER_OVERLAP = 41      # 24 read + 16 trimmed coords + 1 type (+ pile table amortised)
K3_BYTES_PER_VISIT = 8         # each two-hop visit streams one (dst, len)


def workload_config(workload: str, n_gpus: int) -> dict:
    """The `config` object of a bench line: what the workload IS (identical in this arm and in --impl reference, for any N);
    what a run counted on it goes into `counts`."""
    genome, cov, rl, desc = WORKLOADS[workload]
    return {"workload": desc if n_gpus == 1 else f"{n_gpus} x ({desc}), read ids shuffled globally",
            "genome_bp": genome * n_gpus, "coverage": cov, "read_len": rl, "n_gpus": n_gpus,
            "l2": "inputs larger than L2: every GPU streams its whole record shard (hundreds of MB) in every step",
            "scaling": "weak: the genome grows with the number of GPUs"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clocks + throttle reasons sampled DURING the timed region: NVML every 5 ms (nvidia_ml_py), else one
    nvidia-smi query every 100 ms."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.th = index, [], threading.Event(), None
        self.nvml, self.handle, self.max_mhz = None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _run_nvml(self):
        n = self.nvml
        bits = [getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4))]
        while not self.stop_flag.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                self.samples.append([mhz, self.max_mhz, watts] + [bool(r & b) for b in bits])
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def _run_smi(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                if len(f) >= 7:
                    self.samples.append([float(f[0]), float(f[1]), float(f[2])] + [x.lower().startswith("active") for x in f[3:7]])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        self.th = threading.Thread(target=self._run_nvml if self.nvml else self._run_smi, daemon=True)
        self.th.start()

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[3 + i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "power_w_max": max(s[2] for s in self.samples),
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml else "nvidia-smi"}


def ncu_traffic(kernels):
    """dram__bytes_read + write per launch of the named kernels, from the committed `ncu --set full` summary
    (profiles/traffic.json, written by profiles/summarize.py); None when no capture has them."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    if not all(k in t["kernels"] for k in kernels):
        return None, t.get("tag")
    tags = sorted({t["kernels"][k].get("tag", t.get("tag")) for k in kernels})   # the capture(s) the numbers come from
    return sum(t["kernels"][k]["dram_bytes"] for k in kernels), "+".join(tags)


def make_dataset(workload: str, n_gpus: int, seed: int = 3):
    from rala_b200 import synth
    genome, cov, rl, _ = WORKLOADS[workload]
    if workload == "c2":   # BASELINE.json configs[1] (tests/datasets.py CONFIGS["c2"]): only meaningful through the CLI (--cli), the
        #                    pile table changes between the passes
        return synth.generate(genome * n_gpus, cov, rl, len_sd=3000, seed=2, min_ovl=1000, noise=30, chimera_frac=0.03,
                              adapter_frac=0.05, repeats=(1, 12, 4000))
    if workload == "c5":   # BASELINE.json configs[4]: 8 hub reads with ~2 600 spokes each on the uniform background
        return synth.generate_repeat_hubs(genome * n_gpus, cov, rl, seed=5)
    return synth.generate(genome * n_gpus, cov, rl, seed=seed)


def cpu_reference_prepare(ds, piles, tmpdir: str):
    """Write the batch once in the form oracle/_ref/rala_ref reads (binary records + pile table)."""
    from oracle import oracle as O
    if not O.have_ref():
        return None
    prefix = os.path.join(tmpdir, "w")
    O.write_hotpath_inputs(prefix, ds.records, piles, None, None, ds.read_len)
    return prefix


def cpu_reference_run(ds, piles, prefix=None) -> dict:
    """One pass of the reference's own CPU path on in-memory inputs (oracle/_ref), else the plain-C oracle port."""
    from oracle import oracle as O
    cores_used = 1   # the reference runs this path on its main thread whatever -t is (SURVEY.md finding 1)
    if O.have_ref():
        with tempfile.TemporaryDirectory() as tmp:
            if prefix is None:
                prefix = cpu_reference_prepare(ds, piles, tmp)
            r = json.loads(O.ref_run(["hotpath", prefix, "-", 1]).strip().splitlines()[-1])
        t = r["t_classify"] + r["t_preprocess"] + r["t_nodes"] + r["t_edges"] + r["t_transitive"]
        return {"kind": "reference", "edges": r["edges"], "seconds": t, "cores": cores_used, "phases": r}
    t0 = time.perf_counter()
    P = O.Pipeline(ds.records, piles).run()
    t = time.perf_counter() - t0
    return {"kind": "port", "edges": int(P.edges.shape[0]), "seconds": t, "cores": cores_used, "phases": {}}


REFERENCE_ARM_BUDGET_S = 150.0   # wall-clock budget of one `--impl reference` run


def run_reference_arm(args):
    """The reference's own CPU implementation of the path, timed on this box's host cores.  Each step is one pass
    over a BOUNDED sample of the workload: a genome of the same coverage and read length, sized from a probe run
    so that warmup + steps passes fit REFERENCE_ARM_BUDGET_S (the whole batch when that fits)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rala_b200 import synth
    genome, cov, rl, desc = WORKLOADS[args.workload]
    genome *= max(args.gpus, 1)            # weak scaling: the workload of the N-GPU arm is N x the genome
    passes = max(args.steps + args.warmup, 1)
    with tempfile.TemporaryDirectory() as tmp:
        probe_g = min(genome, 5_000_000)
        probe = synth.generate(probe_g, cov, rl, seed=3)
        ppre = cpu_reference_prepare(probe, probe.flat_piles(), tmp)
        t0 = time.perf_counter()
        cpu_reference_run(probe, probe.flat_piles(), ppre)
        wall_per_bp = (time.perf_counter() - t0) / probe_g
        sample_g = int(min(genome, max(probe_g, REFERENCE_ARM_BUDGET_S / passes / wall_per_bp)))
        ds = probe if sample_g == probe_g else synth.generate(sample_g, cov, rl, seed=3)
        piles = ds.flat_piles()
        prefix = ppre if ds is probe else cpu_reference_prepare(ds, piles, tmp)
        times, edges, kind = [], 0, "port"
        for i in range(args.warmup + args.steps):
            r = cpu_reference_run(ds, piles, prefix)
            edges, kind = r["edges"], r["kind"]
            if i >= args.warmup:
                times.append(r["seconds"])
    t = sum(times) / len(times)
    value = edges / t
    sample = (f"{'full batch' if sample_g == genome else 'bounded sample'}: {sample_g / 1e6:.1f} Mbp genome of the same coverage / read length "
              f"({ds.n_overlaps} overlap records, {ds.n_reads} reads, {edges} edges) per step; workload genome {genome / 1e6:.0f} Mbp")
    print(json.dumps({
        "impl": "reference", "metric": "graph_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
        "config": workload_config(args.workload, max(args.gpus, 1)),
        "counts": {"n_overlaps": ds.n_overlaps, "n_reads": ds.n_reads, "edges": edges,
                   "note": "counts of the bounded sample one step runs on (cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "note": "the reference runs this path single-threaded regardless of -t (graph.cpp:443-518, 576-632, 1281-1335); "
                                 "its time includes Node / Edge allocation, the transitive_edges_ sort and remove_marked_objects, which "
                                 "the GPU arm's e2e (flat edge rows + marks in host memory) does not"},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_single(args):
    import torch
    from rala_b200 import api

    ds = make_dataset(args.workload, 1)
    piles = ds.flat_piles()
    n_ovl, n_reads = ds.n_overlaps, ds.n_reads
    ctx = api.Context(0)
    G = api.Graph(ctx)

    # pinned host buffers for the end-to-end path
    rec_pin = torch.from_numpy(ds.records).pin_memory()
    piles_pin = torch.from_numpy(piles).pin_memory()

    # ---- resident: inputs uploaded once, K steps of the whole device pipeline -------------------
    G.set_piles(piles_pin).set_hills(None).set_overlaps(rec_pin)
    for _ in range(args.warmup):
        G.run()
    ctx.synchronize()
    counts = G.counts()
    E = counts["n_edges"]
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = ctx.launch_count
    stage_acc = {}
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.event_record(0)
    for _ in range(args.steps):
        G.run()     # classify() restores the pile table it started from (device copy, inside the timed region)
    ctx.event_record(1)
    dev_ms = ctx.event_elapsed_ms()
    ctx.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = ctx.launch_count - launches0
    ms_per_step = dev_ms / args.steps
    value = E / (ms_per_step * 1e-3)

    # per-kernel durations averaged over a few extra (untimed) steps for the roofline object; the stage timers are
    # CUDA events between the kernels, which only exist in the eager chain (the timed steps replay a CUDA graph)
    G.use_cuda_graph(False)
    k1, k3 = [], []
    for _ in range(max(3, min(args.steps, 10))):
        G.run()
        s = G.stage_ms()
        k1.append(s["k1_classify_kernel"] + s["k1_survivors_kernel"])
        k3.append(s["k3_transitive_kernels"])
        for k, v in s.items():
            stage_acc.setdefault(k, []).append(v)
    G.use_cuda_graph(True)
    clocks = sampler.stop()
    k1_ms, k3_ms = float(np.mean(k1)), float(np.mean(k3))
    peak, peak_src = measured_peaks()
    k1_bytes = K1_BYTES_PER_OVERLAP * n_ovl
    k3_bytes = K3_BYTES_PER_VISIT * counts["n_two_hop"] + 9 * E + 8 * counts["n_nodes"]
    k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
    k3_gbs = k3_bytes / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else 0.0
    dominant = "k_classify_first" if k1_ms >= k3_ms else "k_transitive"
    traffic, traffic_tag = ncu_traffic(["k_classify_events", "k_classify_survivors"] if k1_ms >= k3_ms
                                       else ["k_transitive_group", "k_transitive_light", "k_transitive_heavy"])
    step_bytes = k1_bytes + k3_bytes + 20 * counts["n_candidates"] + 4 * n_reads + 25 * counts["n_overlaps"] + 48 * E
    roof = {"bound": "hbm", "kernel": dominant, "achieved": k1_gbs if k1_ms >= k3_ms else k3_gbs, "peak": peak,
            "unit": "GB/s", "frac": (k1_gbs if k1_ms >= k3_ms else k3_gbs) / peak, "traffic": traffic,
            "traffic_source": f"profiles/{traffic_tag}_kernels.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" if traffic else None,
            "peak_source": peak_src,
            "kernel_members": ["k_classify_events", "k_classify_survivors (+ k_scan_runs, k_relocate_runs)"] if k1_ms >= k3_ms
                              else ["k_transitive_group", "k_transitive_light", "k_transitive_heavy"],
            "whole_step": {"algorithmic_bytes": int(step_bytes), "gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                           "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
            "kernels": {"k_classify_first": {"ms": k1_ms, "algorithmic_bytes": k1_bytes, "gbs": k1_gbs, "frac": k1_gbs / peak,
                                             "note": "41 B / record is SURVEY.md 8(d)'s model (24 read + 16 trimmed coordinates + 1 type written); "
                                                     "this design writes coordinates only for the ~7 % survivors and reads the two id columns twice: "
                                                     "its own minimum is ~32 B / record, i.e. the same time is " +
                                                     f"{32 * n_ovl / (k1_ms * 1e-3) / 1e9 / peak:.3f} of the peak on that count"},
                        "k_transitive": {"ms": k3_ms, "algorithmic_bytes": k3_bytes, "gbs": k3_gbs, "frac": k3_gbs / peak}},
            # the per-stage rates reported next to the metric (eager-chain stage timers)
            "rates": {"k1_overlaps_per_s": n_ovl / (k1_ms * 1e-3) if k1_ms > 0 else None,
                      "k3_edges_per_s": E / (k3_ms * 1e-3) if k3_ms > 0 else None,
                      "k3_two_hop_visits_per_s": counts["n_two_hop"] / (k3_ms * 1e-3) if k3_ms > 0 else None},
            "stage_ms": {k: float(np.mean(v)) for k, v in stage_acc.items()}}

    # ---- end to end through the C ABI with HOST buffers ----------------------------------------
    # What the host shim does per batch (host/graph_b200.cpp): records marshalled column-wise in pinned memory
    # (24 B / record, rala_b200_graph_set_overlaps_columns), pile table, one run, results written by the GPU
    # straight into pinned host buffers (rala_b200_graph_set_outputs: edge rows leave while the transitive pass
    # runs), counts read back.  Every copy is inside the timed region.
    e2e = None
    if not args.skip_e2e:
        cols_pin = torch.from_numpy(api.records_to_columns(ds.records)).pin_memory()
        packed = api.records_to_packed(ds.records)     # 12 B / record when the coordinates fit 16 bits (they do for 10 kbp reads)
        packed_pin = packed.pin() if packed is not None else None
        edges_pin = torch.empty((max(E, 1), 3), dtype=torch.int32).pin_memory()
        marked_pin = torch.empty(max(E, 1), dtype=torch.uint8).pin_memory()
        e2e_steps = max(3, min(args.steps, 10))

        def e2e_step_columns():        # the column form: 24 B / record
            G.set_piles(piles_pin).set_overlaps_columns(cols_pin)
            G.run()
            return G.counts()          # synchronises: edges_pin / marked_pin are complete

        def e2e_step():
            if packed_pin is None:
                return e2e_step_columns()
            G.set_piles(piles_pin).set_overlaps_packed(packed_pin)
            G.run()
            return G.counts()

        def e2e_step_rows():           # the row form of the same call sequence (28 B / record + explicit downloads)
            G.set_piles(piles_pin).set_overlaps(rec_pin)
            G.run()
            G.edges(out=edges_pin)
            G.marked(out=marked_pin)

        def timed(fn, n):
            for _ in range(3):         # shape change -> eager run -> graph capture -> replay
                fn()
            ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            ctx.synchronize()
            return (time.perf_counter() - t0) / n

        rows_s = timed(e2e_step_rows, 3)
        edges_rows = edges_pin.numpy().copy()
        marked_rows = marked_pin.numpy().copy()
        edges_pin.zero_()
        marked_pin.zero_()
        G.set_outputs(edges_pin, marked_pin)
        cols_s = timed(e2e_step_columns, 3)
        e2e_s = timed(e2e_step, e2e_steps)
        c2 = e2e_step()
        G.set_outputs(None, None)
        assert c2["n_edges"] == E and c2["n_transitive_pairs"] == counts["n_transitive_pairs"], (c2, counts)
        assert np.array_equal(edges_pin.numpy(), edges_rows) and np.array_equal(marked_pin.numpy(), marked_rows), \
            "direct-to-host outputs differ from get_edges / get_marked"
        h2d = (packed_pin.nbytes if packed_pin is not None else cols_pin.numel() * 4) + piles.nbytes
        e2e = {"value": E / e2e_s, "unit": "edges/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(12 * E + E + 4 * 34), "ms_per_step": 1e3 * e2e_s,
               "path": ("set_piles + set_overlaps_packed (pinned, 12 B/record: b_id + two 16|16-bit spans + one (query, end) pair per query group, "
                        "expanded on the device)" if packed_pin is not None else "set_piles + set_overlaps_columns (pinned, 24 B/record)") +
                       " + run + outputs written to pinned host memory by the GPU + counts",
               "column_form_ms_per_step": 1e3 * cols_s, "column_form_h2d_bytes": int(cols_pin.numel() * 4 + piles.nbytes),
               "row_form_ms_per_step": 1e3 * rows_s}

    # ---- CPU baseline on the same batch (rank 0, N = 1) -----------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_run(ds, piles)
        cpu = {"value": r["edges"] / r["seconds"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"full batch ({n_ovl} overlap records, {r['edges']} edges) in {r['seconds']:.2f} s",
               "host_cores_available": os.cpu_count(), "phases_s": {k: v for k, v in r["phases"].items() if k.startswith("t_")},
               "scope_note": "the CPU arm's time includes Node / Edge object allocation, the sort of transitive_edges_ and "
                             "remove_marked_objects (oracle/ref_harness.cpp); the GPU arm's e2e ends with flat edge rows and marks "
                             "in pinned host memory: not like for like at the object level (whole programs: cli_baseline)"}
        assert r["edges"] == E, f"CPU baseline built {r['edges']} edges, GPU {E}"

    line = {
        "metric": "graph_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args.workload, 1),
        "counts": {"n_overlaps": n_ovl, "n_reads": n_reads, "edges": E,
                   "nodes": counts["n_nodes"], "two_hop_visits": counts["n_two_hop"], "transitive_pairs": counts["n_transitive_pairs"],
                   "containment_events": counts["n_candidates"], "fixpoint_rounds": counts["n_rounds"],
                   "heavy_items": counts["n_heavy_items"], "retrim_passes_executed": 0,
                   "record_bytes_streamed_per_step": int(ds.records.nbytes * 6 // 7), "parallelism": "1 GPU"},
        "wall_ms_per_step": wall_ms / args.steps,
        "e2e": e2e,
        "gpu_launches": int(launches), "lib": os.path.relpath(api.LIB_PATH, ROOT),
        "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        # the honest drop-in number: whole-program wall time of both CLIs on configs[0] (configs[1] / [2]: bench.py --cli)
        line["cli_baseline"] = cli_compare("c1")
    print(json.dumps(line))
    G.close()
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------
# The drop-in CLI (host/_build/rala_b200: the reference's own front end and CLI with the hot path on the GPU) next to
# the unmodified reference CLI (oracle/_ref/rala), on FASTA + PAF files of a workload.  This is the only way to time
# the NOISY path (configs[1]): hill breaking, pit rounds and promotion need the reference's Pile code between the passes.
# ---------------------------------------------------------------------------------------------------------------
DROPIN = os.path.join(ROOT, "host", "_build", "rala_b200")
REFCLI = os.path.join(ROOT, "oracle", "_ref", "rala")


def _logger_lines(stderr: str) -> dict:
    import re
    out = {}
    for line in stderr.splitlines():
        m = re.match(r"\[(rala::[\w:]+)\] ?(.*?) (\d+\.\d+) s$", line.strip())
        if m:
            out[(m.group(1) + " " + m.group(2)).strip()] = float(m.group(3))
        m = re.search(r"(number of [\w ]+) = (\d+)", line)
        if m:
            out[m.group(1)] = int(m.group(2))
    return out


def cli_compare(workload: str, devices: str | None = None, keep_dir: str | None = None) -> dict:
    """Wall time and logger phase lines of both CLIs on the same files; the drop-in also reports its device stages."""
    if not (os.path.exists(DROPIN) and os.path.exists(REFCLI)):
        return {"unavailable": "host/_build/rala_b200 or oracle/_ref/rala not built (needs /root/reference at build time)"}
    ds = make_dataset(workload, 1)
    threads = str(os.cpu_count() or 1)
    with tempfile.TemporaryDirectory(dir=keep_dir) as tmp:
        fa, paf = os.path.join(tmp, "reads.fasta"), os.path.join(tmp, "overlaps.paf")
        t0 = time.perf_counter()
        ds.write_fasta(fa)
        ds.write_paf(paf)
        t_write = time.perf_counter() - t0
        env = dict(os.environ, RALA_B200_REPORT="1")
        if devices:
            env["RALA_B200_DEVICES"] = devices
        t0 = time.perf_counter()
        got = subprocess.run([DROPIN, "-t", threads, fa, paf], capture_output=True, text=True, env=env)
        t_dropin = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = subprocess.run([REFCLI, "-t", threads, fa, paf], capture_output=True, text=True)   # CPU baseline leg
        t_ref = time.perf_counter() - t0
    if got.returncode != 0 or ref.returncode != 0:
        return {"error": (got.stderr[-600:] if got.returncode else ref.stderr[-600:])}
    reports = [json.loads(l.split("] ", 1)[1]) for l in got.stderr.splitlines() if l.startswith("[rala_b200::report]")]
    lg, lr = _logger_lines(got.stderr), _logger_lines(ref.stderr)
    same = all(lg.get(k) == lr.get(k) for k in ("number of nodes", "number of edges", "number of transitive edges"))
    return {"workload": WORKLOADS[workload][3], "threads": int(threads), "reads": ds.n_reads, "overlaps": ds.n_overlaps,
            "fasta_paf_written_s": round(t_write, 2), "dropin_wall_s": round(t_dropin, 2), "reference_wall_s": round(t_ref, 2),
            "same_node_edge_transitive_counts": same, "dropin_devices": devices or "one",
            "dropin_report": reports, "dropin_logger": lg, "reference_logger": lr,
            "note": "whole-program wall times: FASTA / PAF parsing, pile analysis and sequence handling are the reference's own "
                    "host code in both; the logger lines show where the hot path sits inside them"}


def duplicate_columns(n_groups: int, mean_group: int, seed: int = 31, repeat_groups: int = 0, repeat_size: int = 5000):
    """Columns (a_id, b_id, length) of a PAF in file order for the front end's duplicate filter: query groups of
    ~mean_group records, a fifth of them hitting a target the group already has (repeats), a few length ties;
    optionally some groups of thousands of records over few targets (repeat-heavy reads)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sizes = rng.integers(1, 2 * mean_group, n_groups)
    if repeat_groups:
        sizes[rng.choice(n_groups, size=repeat_groups, replace=False)] = repeat_size
    gid = np.repeat(np.arange(n_groups, dtype=np.int64), sizes)
    n = gid.shape[0]
    n_reads = max(1000, n // mean_group)
    a = (rng.permutation(n_groups).astype(np.int64) % n_reads)[gid]            # consecutive groups have different queries
    fresh = rng.integers(0, n_reads, n)
    pool = (a * 7919 + rng.integers(0, 6, n)) % n_reads                         # a handful of targets per query
    b = np.where(rng.random(n) < 0.2, pool, fresh)
    b = np.where(b == a, (b + 1) % n_reads, b)
    ln = 1000 + (gid % 4000) + rng.integers(0, 3, n)
    return a.astype(np.uint32), b.astype(np.uint32), ln.astype(np.uint32)


def run_frontend(args):
    """`bench.py --frontend`: SURVEY.md 8(f) row 1, first half — the duplicate filter of Graph::initialize
    (graph.cpp:273-303) through rala_b200_filter_duplicates, host columns in, validity bytes out, next to the oracle's
    literal restatement of the reference loops on one core (the reference runs them on its main thread)."""
    from oracle import oracle as O
    from rala_b200 import api
    out = {"metric": "overlaps_filtered_per_sec", "unit": "records/s", "n_gpus": 1, "higher_is_better": True, "dtype": "u32",
           "data": "synthetic", "cases": []}
    peak, peak_src = measured_peaks()
    ctx = api.Context(0)
    for name, kw in (("configs[2]-like: 400 k query groups of ~36 records", dict(n_groups=400000, mean_group=36)),
                     ("repeat-heavy: the same plus 400 groups of 5 000 records over a handful of targets",
                      dict(n_groups=100000, mean_group=36, repeat_groups=400, repeat_size=5000))):
        a, b, ln = duplicate_columns(**kw)
        n = a.shape[0]
        ctx.filter_duplicates(a[:1000], b[:1000], ln[:1000])
        best = None
        for _ in range(max(3, args.warmup)):
            t0 = time.perf_counter()
            valid, ms = ctx.filter_duplicates(a, b, ln, with_time=True)
            wall = 1e3 * (time.perf_counter() - t0)
            best = (ms, wall) if best is None or ms < best[0] else best
        sample = min(n, 4_000_000)
        t0 = time.perf_counter()
        want = O.filter_duplicates(a[:sample], b[:sample], ln[:sample])
        cpu_s = time.perf_counter() - t0
        # the sample's last group may be cut: compare up to the last group boundary inside it
        cut = sample if sample == n else int(np.flatnonzero(a[:sample] != a[sample - 1])[-1]) + 1
        same = bool(np.array_equal(valid[:cut], want[:cut]))
        out["cases"].append({"workload": name, "records": int(n), "kept": int(valid.sum()), "kernel_ms": best[0],
                             "call_ms_host_buffers": best[1], "records_per_s_kernel": n / (best[0] * 1e-3),
                             "records_per_s_call": n / (best[1] * 1e-3), "algorithmic_gbs": 13.0 * n / (best[0] * 1e-3) / 1e9,
                             "frac_of_hbm_peak": 13.0 * n / (best[0] * 1e-3) / 1e9 / peak,
                             "cpu_oracle_records_per_s": cut and sample / cpu_s, "cpu_sample_records": int(sample),
                             "same_as_oracle_on_sample": same})
        if not same:
            raise SystemExit("duplicate filter differs from the oracle")
    c0 = out["cases"][0]
    out.update({"value": c0["records_per_s_kernel"], "config": {"workload": c0["workload"]},
                "e2e": {"value": c0["records_per_s_call"], "unit": "records/s", "h2d_bytes_per_step": 12 * c0["records"],
                        "d2h_bytes_per_step": c0["records"]},
                "roofline": {"bound": "hbm", "kernel": "k_filter_duplicates", "achieved": c0["algorithmic_gbs"], "peak": peak,
                             "unit": "GB/s", "frac": c0["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                             "note": "13 B / record (three 4-byte columns read, one byte written); every thread re-reads its "
                                     "group's columns from L1 / L2, so the kernel is bound by the load pipe, not by HBM"},
                "cpu_baseline": {"value": c0["cpu_oracle_records_per_s"], "unit": "records/s", "cores": 1, "kind": "port",
                                 "sample": f"the first {c0['cpu_sample_records']} records"},
                "gpu_launches": int(ctx.launch_count)})
    print(json.dumps(out))


def run_cli(args):
    r = cli_compare(args.workload, args.devices)
    line = {"metric": "drop-in CLI vs reference CLI", "cli_baseline": r}
    rep = {x.get("stage"): x for x in r.get("dropin_report", [])}
    c = rep.get("construct")
    if c and "device_ms" in c:
        ms = c["device_ms"]
        t = rep.get("remove_transitive_edges", {}).get("device_ms", {}).get("transitive", 0.0)
        total = sum(ms.values()) + t
        peak, peak_src = measured_peaks()
        retrim_gbs = K1_BYTES_PER_OVERLAP * c["list_entries_retrimmed"] / (ms["retrim"] * 1e-3) / 1e9 if ms["retrim"] > 0 else 0.0
        line.update({"value": c["edges"] / (total * 1e-3) if total > 0 else None, "unit": "edges/s (device time of the hot-path stages)",
                     "config": {"workload": WORKLOADS[args.workload][3], "retrim_passes_executed": c["retrim_passes_executed"],
                                "pit_rounds": c["pit_rounds"], "n_overlaps": c["records"], "edges": c["edges"]},
                     "device_ms": dict(ms, transitive=t, total=total),
                     "roofline": {"bound": "hbm", "kernel": "k_list_pass (re-trim / promote passes)", "achieved": retrim_gbs, "peak": peak,
                                  "unit": "GB/s", "frac": retrim_gbs / peak, "traffic": None, "peak_source": peak_src,
                                  "algorithmic_bytes": K1_BYTES_PER_OVERLAP * c["list_entries_retrimmed"],
                                  "note": "41 B per list entry and pass (SURVEY.md 8d); the lists hold ~10^5 entries on this workload, so the "
                                          "passes are launch-latency bound, far from the bandwidth roofline"}})
    print(json.dumps(line))


def run_ab(args):
    """A/B of single optimisations: the product library and the builds with ONE switch turned off each
    (rala_b200/build.py VARIANTS, common.cuh RB_OPT_*), same batch, same process, device-resident steps."""
    from rala_b200 import api, build as B
    ds = make_dataset(args.workload, 1)
    piles = ds.flat_piles()
    libs = [("product", api.LIB_PATH, None)] + [(k, B.variant_path(k), None) for k in list(B.VARIANTS) + list(B.TUNINGS)
                                                if os.path.exists(B.variant_path(k))]
    if args.variants:
        keep = set(args.variants.split(","))
        libs = [l for l in libs if l[0] == "product" or l[0] in keep]
    prev = B.variant_path("prev")   # the previous commit's library, when profiles/capture_ab.sh built it
    if os.path.exists(prev):
        libs.append(("prev", prev, None))
    out = {}
    for rnd in range(2):            # two rounds, interleaved: drift shows up as a difference between them
        for name, path, _env in libs:
            ctx = api.Context(0, lib=api.load_path(path))
            G = api.Graph(ctx)
            G.set_piles(piles).set_hills(None).set_overlaps(ds.records)
            for _ in range(max(args.warmup, 3)):
                G.run()
            ctx.synchronize()
            ctx.event_record(0)
            for _ in range(args.steps):
                G.run()
            ctx.event_record(1)
            ms = ctx.event_elapsed_ms() / args.steps
            c = G.counts()
            import zlib
            crc = zlib.crc32(G.marked().tobytes(), zlib.crc32(G.edges().tobytes()))   # an experiment must not change a single bit
            rec = out.setdefault(name, {"ms_per_step": [], "edges": c["n_edges"], "pairs": c["n_transitive_pairs"],
                                        "edges_marks_crc32": crc})
            rec["ms_per_step"].append(ms)
            G.use_cuda_graph(False)   # stage timers exist in the eager chain only: which stage an experiment moved
            acc = {}
            for _ in range(5):
                G.run()
                for k, v in G.stage_ms().items():
                    acc.setdefault(k, []).append(v)
            rec["stage_us_eager"] = {k: round(1e3 * float(np.median(v)), 1) for k, v in acc.items()}
            G.close()
            ctx.close()
    base = min(out["product"]["ms_per_step"])
    for name, v in out.items():
        v["delta_us_vs_product"] = 1e3 * (min(v["ms_per_step"]) - base)
        v["same_result_as_product"] = v["edges_marks_crc32"] == out["product"]["edges_marks_crc32"]
    print(json.dumps({"ab": out, "steps": args.steps, "workload": WORKLOADS[args.workload][3]}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="device-resident timing only")
    ap.add_argument("--skip-parity", action="store_true", help="N > 1: do not verify the assembled result on rank 0")
    ap.add_argument("--parity-oracle-max", type=int, default=60_000_000,
                    help="N > 1: run the plain-C oracle on the whole batch when it has at most this many records")
    ap.add_argument("--frontend", action="store_true", help="time the front end's duplicate filter (SURVEY 8(f) row 1) against the oracle")
    ap.add_argument("--cli", action="store_true", help="time the drop-in CLI against the reference CLI on FASTA + PAF files of the workload")
    ap.add_argument("--devices", default=None, help="--cli: RALA_B200_DEVICES for the drop-in (e.g. 0,1)")
    ap.add_argument("--force-multi", action="store_true", help="run the multi-GPU path even with one rank (under torchrun)")
    ap.add_argument("--variants", default="", help="with --ab: only these variants (comma separated)")
    ap.add_argument("--ab", action="store_true", help="time the product library against the one-switch-off builds (rala_b200/variants/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.ab:
        run_ab(args)
        return
    if args.frontend:
        run_frontend(args)
        return
    if args.cli:
        run_cli(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started as plain `python bench.py --gpus N`: one rank per GPU needs the launcher the driver uses
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                  "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
                                  os.path.abspath(__file__)] + sys.argv[1:])
    if args.gpus > 1 or world > 1 or args.force_multi:
        import bench_multi
        bench_multi.bench_main(args)
        return
    run_single(args)


if __name__ == "__main__":
    main()
