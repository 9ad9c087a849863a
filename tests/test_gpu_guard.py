"""Checking mode of the library (-m gpu): GJKEPA_GUARD=1 brackets every device allocation the library makes with pattern
bands and hands the payload out filled with 0xFF bytes (include/gjkepa.h, csrc/gjkepa_capi.cu `namespace guard`).

compute-sanitizer cannot be run on the boxes this library is tested on, so this is the memory check it gets: the host
pipeline and parity suites are run once more in a child process in checking mode; tests/conftest.py then asserts after EVERY
test that no band changed (no kernel or copy wrote past either end of a workspace), and the suites' own bit-for-bit and
oracle comparisons show that nothing depended on the zeros a fresh cudaMalloc happens to hold.  The first test shows the
detector firing on a planted one-byte overrun."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

PLANT = r"""
import json, ctypes as C, numpy as np
from collision_detect_gjk_epa_b200 import Engine, guard_check, workloads as W, _lib
L = _lib.lib()
eng = Engine(0)
b = W.make_mixed(5000, 0.3, seed=5)
eng.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="full", want_contacts=True)
clean = guard_check()
assert L.gjkepa_guard_selftest(1) == 0
after_end = guard_check()
assert L.gjkepa_guard_selftest(-1) == 0
before_start = guard_check()
eng.close()
print(json.dumps({"clean": clean, "after_end": after_end, "before_start": before_start, "closed": guard_check()}))
"""


def child_env():
    env = dict(os.environ, GJKEPA_GUARD="1")
    env["PYTHONPATH"] = str(ROOT) + os.pathsep + env.get("PYTHONPATH", "")
    return env


def test_outside_checking_mode_there_is_nothing_to_report():
    from collision_detect_gjk_epa_b200 import guard_check
    if os.environ.get("GJKEPA_GUARD", "0") in ("", "0"):
        assert guard_check() is None


def test_the_detector_fires_on_a_planted_overrun_and_only_then():
    r = subprocess.run([sys.executable, "-c", PLANT], capture_output=True, text=True, env=child_env(), cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    n, bad = out["clean"]
    assert n >= 10 and bad == 0                      # a context's workspaces are all bracketed, and a real batch left them intact
    assert out["after_end"][1] == 1                  # one byte past the end of a 1000-byte allocation
    assert out["before_start"][1] == 2               # one byte before the start of another
    assert out["closed"][1] == 2                     # freeing the context's buffers found nothing more


def test_host_pipeline_and_parity_suites_in_checking_mode():
    if os.environ.get("GJKEPA_GUARD", "0") not in ("", "0"):
        pytest.skip("already the checking-mode child")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_host_pipeline.py", "tests/test_gpu_parity.py",
                        "tests/test_gpu_f32_mode.py", "tests/test_broadphase.py", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "not full_size"],
                       capture_output=True, text=True, env=child_env(), cwd=ROOT, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert "guard: checked" in r.stdout, tail        # conftest's per-test band check really ran
