"""ctypes loader for the CPU oracle (oracle/gjkepa_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libgjkepa_oracle.so"

MODE_GJK, MODE_EPA, MODE_FULL = 0, 1, 2

EXIT_NAMES = {
    0: "none", 1: "sphere", 2: "init_cap", 3: "dup3", 4: "tri_origin", 5: "coplanar4",
    6: "init_contain", 7: "loop_cap", 8: "loop_collinear", 9: "loop_coplanar",
    10: "loop_contain", 11: "loop_stall", 12: "stop",
}
ST_EPA_CAP, ST_STOP, ST_HULL_FAIL, ST_BAD_VERSION, ST_H1 = (1 << 24, 1 << 25, 1 << 26, 1 << 27, 1 << 28)


class OrcResult(C.Structure):
    _fields_ = [("collision", C.c_int32), ("colli_type", C.c_int32), ("nearest", C.c_double * 6),
                ("normal", C.c_double * 3), ("point", C.c_double * 3), ("depth", C.c_double),
                ("status", C.c_uint32), ("n_support", C.c_int32), ("simplex", C.c_double * 12)]


def build(force: bool = False) -> Path:
    src = [_HERE / "gjkepa_oracle.c", _HERE / "gjkepa_oracle.h", _HERE / "dirtable.inc"]
    if force or not _SO.exists() or any(s.stat().st_mtime > _SO.stat().st_mtime for s in src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True, env=env)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        dp, ip, i64p, u8p, u32p = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_uint8), C.POINTER(C.c_uint32))
        L.orc_gjkepa_pair.argtypes = [C.c_int32, C.c_double, dp, C.c_int32, dp, C.c_int32, C.c_int32,
                                      C.POINTER(OrcResult)]
        L.orc_gjkepa_pair.restype = None
        L.orc_gjkepa_batch.argtypes = [C.c_int32, C.c_double, dp, i64p, C.c_int64, ip, ip, C.c_int32, C.c_int32,
                                       u8p, ip, dp, dp, dp, dp, u32p, ip]
        L.orc_gjkepa_batch.restype = C.c_int
        L.orc_support.argtypes = [dp, C.c_int32, dp, C.c_int32, dp, dp, ip, ip]
        L.orc_support.restype = None
        L.orc_is_inside_pf.argtypes = [dp, C.c_int32, dp]
        L.orc_is_inside_pf.restype = C.c_int
        L.orc_point_in_simplex.argtypes = [dp, dp]
        L.orc_point_in_simplex.restype = C.c_int
        L.orc_sphere_test.argtypes = [dp, C.c_int32, dp, C.c_int32]
        L.orc_sphere_test.restype = C.c_int
        L.orc_quickhull.argtypes = [dp, C.c_int32, dp, ip, C.c_int32]
        L.orc_quickhull.restype = C.c_int
        L.orc_dir_table.argtypes = [C.c_int32, dp]
        L.orc_dir_table.restype = None
        L.orc_epilogue.argtypes = [C.c_int32, C.c_double, dp, C.c_int32, dp, C.c_int32, dp, dp, dp, dp, ip]
        L.orc_epilogue.restype = None
        L.orc_max_threads.restype = C.c_int
        L.orc_sphere_pairs.argtypes = [dp, C.POINTER(C.c_int64), C.c_int64, C.c_int64, ip, ip]
        L.orc_sphere_pairs.restype = C.c_int64
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fblock(p):
    """(n,3) vertices -> Fortran column-major block [x|y|z]."""
    p = np.asarray(p, dtype=np.float64)
    return np.ascontiguousarray(p.T).reshape(-1), p.shape[0]


def pair(p1, p2, version=1, tol_ff=1.0, mode=MODE_FULL) -> dict:
    """Run one pair given (n,3) vertex arrays; returns a dict of the reference outputs."""
    b1, n1 = _fblock(p1)
    b2, n2 = _fblock(p2)
    r = OrcResult()
    lib().orc_gjkepa_pair(version, tol_ff, _dp(b1), n1, _dp(b2), n2, mode, C.byref(r))
    st = int(r.status)
    return dict(collision=bool(r.collision), colli_type=int(r.colli_type),
                nearest=np.array(r.nearest).reshape(3, 2).T.copy(),   # (2,3)
                normal=np.array(r.normal), point=np.array(r.point), depth=float(r.depth),
                status=st, exit=EXIT_NAMES[st & 0xff], gjk_iters=(st >> 8) & 0xff,
                epa_iters=(st >> 16) & 0xff, n_support=int(r.n_support),
                simplex=np.array(r.simplex).reshape(4, 3))


def batch(verts, hull_off, pair_a, pair_b, version=1, tol_ff=1.0, mode=MODE_FULL, n_threads=0) -> dict:
    verts = np.ascontiguousarray(verts, dtype=np.float64)
    hull_off = np.ascontiguousarray(hull_off, dtype=np.int64)
    pair_a = np.ascontiguousarray(pair_a, dtype=np.int32)
    pair_b = np.ascontiguousarray(pair_b, dtype=np.int32)
    n = pair_a.shape[0]
    out = dict(collision=np.zeros(n, np.uint8), colli_type=np.zeros(n, np.int32),
               nearest=np.zeros((n, 6)), normal=np.zeros((n, 3)), point=np.zeros((n, 3)),
               depth=np.zeros(n), status=np.zeros(n, np.uint32), n_support=np.zeros(n, np.int32))
    used = lib().orc_gjkepa_batch(
        version, tol_ff, _dp(verts), hull_off.ctypes.data_as(C.POINTER(C.c_int64)), n,
        pair_a.ctypes.data_as(C.POINTER(C.c_int32)), pair_b.ctypes.data_as(C.POINTER(C.c_int32)),
        mode, n_threads,
        out["collision"].ctypes.data_as(C.POINTER(C.c_uint8)),
        out["colli_type"].ctypes.data_as(C.POINTER(C.c_int32)),
        _dp(out["nearest"]), _dp(out["normal"]), _dp(out["point"]), _dp(out["depth"]),
        out["status"].ctypes.data_as(C.POINTER(C.c_uint32)),
        out["n_support"].ctypes.data_as(C.POINTER(C.c_int32)))
    out["threads"] = used
    return out


def epilogue(p1, p2, normal, version=1, tol_ff=1.0) -> dict:
    """The reference's contact epilogue (F:326-343: nearest points, collision point by `version`, collision type)
    evaluated on a GIVEN collision normal.  p1, p2: (n,3) vertex arrays."""
    (b1, n1), (b2, n2) = _fblock(p1), _fblock(p2)
    nrm = np.ascontiguousarray(normal, dtype=np.float64)
    near, pt, nout = np.zeros(6), np.zeros(3), np.zeros(3)
    ct = C.c_int32(0)
    lib().orc_epilogue(int(version), float(tol_ff), _dp(b1), n1, _dp(b2), n2, _dp(nrm), _dp(near), _dp(pt), _dp(nout), C.byref(ct))
    return {"nearest": near, "point": pt, "normal": nout, "colli_type": int(ct.value)}


def quickhull(pts) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    cap = 640
    tris = np.zeros((cap, 9))
    nt = C.c_int32(0)
    rc = lib().orc_quickhull(_dp(pts), pts.shape[0], _dp(tris), C.byref(nt), cap)
    if rc:
        raise RuntimeError(f"orc_quickhull rc={rc}")
    return tris[:nt.value].reshape(-1, 3, 3).copy()


def ground_truth(p1, p2):
    """Independent ground truth via scipy/Qhull on all pairwise differences:
    returns (inside, depth, normal, second_smallest_face_distance)."""
    from scipy.spatial import ConvexHull
    p1 = np.asarray(p1, dtype=np.float64)
    p2 = np.asarray(p2, dtype=np.float64)
    md = (p1[:, None, :] - p2[None, :, :]).reshape(-1, 3)
    h = ConvexHull(md)
    eq = h.equations                      # n.x + off <= 0 inside
    off = eq[:, 3]
    inside = bool(np.all(off < 0))
    k = int(np.argmax(off))               # closest facet to the origin when inside
    nrm = eq[k, :3]
    # distinct planes only (Qhull triangulates coplanar facets)
    d = np.sort(-off)
    second = float(d[np.argmax(d > d[0] + 1e-9)]) if np.any(d > d[0] + 1e-9) else float("inf")
    return inside, float(-off[k]), nrm.copy(), second


def apply_pose(shape_verts, shape_off, body_shape, pose):
    """World-space hull store from body-space shapes and poses, with the arithmetic the product's pose kernel is
    specified to use (include/gjkepa.h): x' = ((r00*x + r01*y) + r02*z) + tx, every operation rounded (numpy
    evaluates element-wise products and sums separately: no FMA).  Returns (verts, hull_off).  Test
    infrastructure, like everything under oracle/."""
    shape_off = np.asarray(shape_off, np.int64); body_shape = np.asarray(body_shape, np.int64)
    pose = np.asarray(pose, np.float64).reshape(-1, 12)
    n = (shape_off[1:] - shape_off[:-1])[body_shape]
    hull_off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    verts = np.empty(3 * int(hull_off[-1]))
    for size in np.unique(n):                                   # vectorised per shape size
        idx = np.nonzero(n == size)[0]
        so = shape_off[body_shape[idx]]
        g = (3 * so[:, None] + np.arange(3 * size)[None, :])
        blk = np.asarray(shape_verts)[g].reshape(len(idx), 3, size)
        x, y, z = blk[:, 0], blk[:, 1], blk[:, 2]
        p = pose[idx]
        wx = ((p[:, 0, None] * x + p[:, 1, None] * y) + p[:, 2, None] * z) + p[:, 9, None]
        wy = ((p[:, 3, None] * x + p[:, 4, None] * y) + p[:, 5, None] * z) + p[:, 10, None]
        wz = ((p[:, 6, None] * x + p[:, 7, None] * y) + p[:, 8, None] * z) + p[:, 11, None]
        w = np.stack([wx, wy, wz], axis=1).reshape(len(idx), 3 * size)
        verts[(3 * hull_off[idx][:, None] + np.arange(3 * size)[None, :])] = w
    return verts, hull_off


def sphere_pairs(verts, hull_off):
    """All pairs (a < b) passing the reference's envelope-sphere test (F:1185), by brute force over all pairs, in
    lexicographic order: the checker for the product's grid broad phase."""
    verts = np.ascontiguousarray(verts, dtype=np.float64)
    hull_off = np.ascontiguousarray(hull_off, dtype=np.int64)
    nh = hull_off.shape[0] - 1
    ip32, ip64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    z = np.empty(1, np.int32)
    total = lib().orc_sphere_pairs(_dp(verts), hull_off.ctypes.data_as(ip64), nh, 0, z.ctypes.data_as(ip32), z.ctypes.data_as(ip32))
    pa, pb = np.empty(max(total, 1), np.int32), np.empty(max(total, 1), np.int32)
    lib().orc_sphere_pairs(_dp(verts), hull_off.ctypes.data_as(ip64), nh, total, pa.ctypes.data_as(ip32), pb.ctypes.data_as(ip32))
    return pa[:total], pb[:total]
