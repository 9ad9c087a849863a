"""examples/gjkepa_example.c: the library called from plain C99 (what any FFI stub does) -- the reference's documented cube case
through the drop-in call, a batch with its contact list on a multi-device context, and a resident store with compact rows."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_c_example_builds_and_gives_the_documented_answers(tmp_path):
    lib = ROOT / "collision_detect_gjk_epa_b200"
    exe = tmp_path / "gjkepa_example"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I", str(ROOT / "include"), str(ROOT / "examples" / "gjkepa_example.c"),
                        "-L", str(lib), "-lgjkepa_b200", f"-Wl,-rpath,{lib}", "-lm", "-o", str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "ok" and lines[0].startswith("pair:     collision 1") and "depth 0.5 " in lines[0]
