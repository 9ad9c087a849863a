import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure, oracle/)."""
    from oracle import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def W():
    from collision_detect_gjk_epa_b200 import workloads
    return workloads


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "gjkepa_golden.npz")


@pytest.fixture(scope="session")
def host_sim():
    """Host (g++) build of the product's GJK core, L=1 instantiation."""
    import ctypes as C
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", str(ROOT / "tests" / "host_sim"), "-s"], check=True, env=env)
    return C.CDLL(str(ROOT / "tests" / "host_sim" / "libgjk_host_sim.so"))


@pytest.fixture(scope="session")
def engine():
    """CUDA engine through the C ABI; GPU tests only.  Fails loudly when the library is missing."""
    from collision_detect_gjk_epa_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()


_guard_checks = {"tests": 0, "allocations": 0}


@pytest.fixture(autouse=True)
def _guard_bands(request):
    """Checking mode (GJKEPA_GUARD=1, tests/test_gpu_guard.py): after every GPU test no guard band of any device
    allocation of the library may have changed."""
    yield
    if os.environ.get("GJKEPA_GUARD", "0") in ("", "0") or request.node.get_closest_marker("gpu") is None:
        return
    from collision_detect_gjk_epa_b200 import guard_check
    got = guard_check()
    assert got is not None, "GJKEPA_GUARD is set but the library is not in checking mode"
    _guard_checks["tests"] += 1
    _guard_checks["allocations"] = max(_guard_checks["allocations"], got[0])
    assert got[1] == 0, f"{got[1]} guard band(s) overwritten (of {got[0]} live allocations) by the end of {request.node.nodeid}"


def pytest_terminal_summary(terminalreporter):
    if _guard_checks["tests"]:
        terminalreporter.write_line(f"guard: checked the bands of up to {_guard_checks['allocations']} device allocations after "
                                    f"{_guard_checks['tests']} tests, none overwritten")
