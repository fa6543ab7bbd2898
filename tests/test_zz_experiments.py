"""GPU suite, LAST file on purpose: the experimental kernels that were written at the end of round 1 without GPU time
(DESIGN.md "Prepared, not measured yet").  They are NOT the default path; each runs in its own process (a faulting
kernel must not take the suite's CUDA context with it) and is marked xfail(strict=False): the suite stays green either
way and the XPASS / XFAIL line records whether the experiment is bit-exact against the oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
REASON = "experimental kernel, off by default, never run on a GPU before: this records whether it is bit-exact"


@pytest.mark.xfail(strict=False, reason=REASON)
@pytest.mark.parametrize("env", [{"RALA_B200_EV_V2": "1"}, {"RALA_B200_EV_V2": "2"}, {"RALA_B200_SURV_V2": "1"},
                                 {"RALA_B200_AGG_ATOMICS": "1"},
                                 {"RALA_B200_EV_V2": "1", "RALA_B200_SURV_V2": "1", "RALA_B200_AGG_ATOMICS": "1"}],
                         ids=["events_v2", "events_v2_2blk", "survivors_v2", "agg_atomics", "all_three"])
def test_experimental_kernel_is_bit_exact(env):
    """tests/quick_check.py: whole pipeline on a noisy dual-record batch against the oracle (lists, piles, edges,
    marks), then 50 timed steps of a 20 Mbp batch; the timing is printed for the log."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "quick_check.py")], capture_output=True, text=True,
                       timeout=180, env={**os.environ, **env, "QUICK_ONLY_PRODUCT": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])["libs"]["product"]
    print(f"experiment {env}: {res}")
    assert "error" not in res, res
    assert res["parity"] is True


@pytest.mark.xfail(strict=False, reason=REASON)
def test_experimental_edge_pairs_exchange_world1():
    """RALA_B200_EDGE_PAIRS=1: the multi-GPU orchestration with edge blocks as reverse-complement pairs (world = 1
    exercises export + import of the pair layout), eager and as a captured step graph."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_multi_gpu.py"), "-m", "gpu", "-q", "-x",
                        "-k", "1-eager or 1-step_graph"], capture_output=True, text=True, timeout=420, cwd=ROOT,
                       env={**os.environ, "RALA_B200_EDGE_PAIRS": "1"})
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    assert "2 passed" in r.stdout
