"""Pins the oracle to the REFERENCE ITSELF wherever a Fortran compiler exists.

This image has no gfortran, so both tests skip here and on the GPU box; on any machine with `gfortran` on PATH and the
reference checkout mounted (REF=/path or /root/reference) they

  1. build the reference's own src/GCLIB_GJKEPA.f90 by the committed recipe oracle/ref_patch/build.sh (four source fixes
     for Intel-only syntax, stand-ins for the three un-vendored modules) into oracle/_ref/gjkepa_ref, run it over the
     golden fixture's pairs and require: collision flags BIT-EXACT with the oracle's, depth within 1e-9 * max(d, 1) and
     normal 1 - cos <= 1e-9 on the pairs both solved cleanly;
  2. compile the ISO_C_BINDING shim fortran/gclib_gjkepa.f90 (syntax and interface check; it needs no GPU to compile).

A pass on such a machine lifts the "parity unpinned" cap of DESIGN.md §5 for the GJK flags without further work.
"""
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("REF", "/root/reference"))
FC = shutil.which(os.environ.get("FC", "gfortran"))

needs_fortran = pytest.mark.skipif(FC is None, reason="no Fortran compiler on PATH (gfortran): the reference cannot be built here")


@needs_fortran
@pytest.mark.skipif(not (REF / "src" / "GCLIB_GJKEPA.f90").exists(), reason="reference checkout not mounted")
def test_reference_fortran_agrees_with_the_oracle_on_the_golden_pairs(orc, golden, tmp_path):
    r = subprocess.run(["bash", str(ROOT / "oracle" / "ref_patch" / "build.sh")], capture_output=True, text=True,
                       env=dict(os.environ, REF=str(REF)), timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    exe = ROOT / "oracle" / "_ref" / "gjkepa_ref"
    assert exe.exists()
    verts, hull_off = golden["verts"], golden["hull_off"].astype(np.int64)
    pa, pb = golden["pair_a"].astype(np.int32), golden["pair_b"].astype(np.int32)
    n = len(pa)
    fin, fout = tmp_path / "pairs.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        np.array([len(hull_off) - 1, hull_off[-1], n], dtype=np.int64).tofile(f)
        hull_off.tofile(f); verts.astype(np.float64).tofile(f); pa.tofile(f); pb.tofile(f)
    r = subprocess.run([str(exe), str(fin), str(fout)], capture_output=True, text=True, timeout=1200, stdin=subprocess.DEVNULL)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    raw = np.fromfile(fout, dtype=np.uint8)
    o = 0
    col = raw[o:o + n].view(np.int8); o += n
    ctype = raw[o:o + 4 * n].view(np.int32); o += 4 * n
    nearest = raw[o:o + 48 * n].view(np.float64).reshape(n, 6); o += 48 * n
    normal = raw[o:o + 24 * n].view(np.float64).reshape(n, 3); o += 24 * n
    point = raw[o:o + 24 * n].view(np.float64).reshape(n, 3); o += 24 * n
    depth = raw[o:o + 8 * n].view(np.float64)
    ref = orc.batch(verts, hull_off, pa, pb, mode=orc.MODE_FULL)
    assert np.array_equal(col != 0, ref["collision"] != 0), "collision flags: reference Fortran vs oracle"
    clean = (ref["collision"] > 0) & ((ref["status"] >> 24) == 0)
    d = ref["depth"][clean]
    assert np.all(np.abs(depth[clean] - d) <= 1e-9 * np.maximum(d, 1.0))
    # tied depths may report either face (the reference's QuickHull is not in its repository): compare |cos|
    cosd = np.abs(np.sum(normal[clean] * ref["normal"][clean], axis=1))
    assert np.mean(1.0 - cosd <= 1e-9) > 0.99
    assert np.array_equal(ctype[clean], ref["colli_type"][clean]) or np.mean(ctype[clean] == ref["colli_type"][clean]) > 0.99


@needs_fortran
def test_iso_c_binding_shim_compiles(tmp_path):
    # default (gnu) standard: the shim keeps the reference's own INTEGER*4 / REAL*8 / LOGICAL*1 dummy declarations (F:44-52)
    r = subprocess.run([FC, "-c", "-Wall", "-fopenmp", "-J", str(tmp_path), str(ROOT / "fortran" / "gclib_gjkepa.f90"),
                        "-o", str(tmp_path / "gclib_gjkepa.o")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-4000:]
