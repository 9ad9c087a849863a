"""Host-side mirror of the reference interface for the GJK+EPA path.

``GJKEPA(version_, TOL_FF_, p1_, p2_)`` keeps the argument names and meaning of
``SUBROUTINE GJKEPA`` (reference src/GCLIB_GJKEPA.f90:39-52) and returns its six outputs;
``Engine`` is the batch form (one context per GPU).  Everything computes in the CUDA library
through the C ABI of include/gjkepa.h -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import MODE_EPA, MODE_FULL, MODE_GJK

EXIT_NAMES = {
    0: "none", 1: "sphere", 2: "init_cap", 3: "dup3", 4: "tri_origin", 5: "coplanar4",
    6: "init_contain", 7: "loop_cap", 8: "loop_collinear", 9: "loop_coplanar",
    10: "loop_contain", 11: "loop_stall", 12: "stop", 13: "early_out",
}
ST_EPA_CAP, ST_STOP, ST_HULL_FAIL, ST_BAD_VERSION, ST_H1, ST_TOPO_WARN = (
    1 << 24, 1 << 25, 1 << 26, 1 << 27, 1 << 28, 1 << 29)

_MODES = {"gjk": MODE_GJK, "epa": MODE_EPA, "full": MODE_FULL,
          MODE_GJK: MODE_GJK, MODE_EPA: MODE_EPA, MODE_FULL: MODE_FULL}


class GjkEpaError(RuntimeError):
    pass


def _np_ptr(a: Optional[np.ndarray]):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class Engine:
    """One narrow-phase context: bound to one GPU (gjkepa_create), or spread over several GPUs of the box
    (``devices=[...]`` or ``devices="all"`` -> gjkepa_create_multi: the host-buffer batch calls then split the pair
    list into one contiguous range per device and every device copies its results into the caller's arrays)."""

    def __init__(self, device: int = 0, lanes: int = 0, devices=None):
        self._L = _lib.lib()
        h = C.c_void_p()
        if devices is not None:
            devs = [] if devices == "all" else [int(d) for d in devices]
            arr = (C.c_int32 * max(len(devs), 1))(*devs)
            rc = self._L.gjkepa_create_multi(C.byref(h), arr if devs else None, len(devs))
            device = devs[0] if devs else 0
        else:
            rc = self._L.gjkepa_create(C.byref(h), int(device))
        if rc != _lib.OK or not h:
            raise GjkEpaError(
                f"gjkepa_create(device={device}) failed with code {rc}: no usable sm_100a CUDA device "
                "(this package has no CPU fallback)")
        self._h = h
        self.device = int(device)
        self.n_devices = int(self._L.gjkepa_device_count(h))
        if lanes:
            self.set_lanes(lanes)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gjkepa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != _lib.OK:
            msg = self._L.gjkepa_last_error(self._h)
            raise GjkEpaError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

    def set_lanes(self, lanes: int):
        self._check(self._L.gjkepa_set_lanes(self._h, int(lanes)), "gjkepa_set_lanes")

    def set_max_hull_vertices(self, n: int):
        self._check(self._L.gjkepa_set_max_hull_vertices(self._h, int(n)), "gjkepa_set_max_hull_vertices")

    TUNE_WALK_REFILL, TUNE_SINGLE_KERNEL, TUNE_EPA_GENERAL, TUNE_TILE_MAX_VERTICES, TUNE_EARLY_OUT, TUNE_HOST_CHUNKS = 1, 2, 3, 4, 5, 6
    TUNE_LOCALITY_SORT, TUNE_F32_FILTER, TUNE_F32_ARITH = 7, 8, 9
    ST_F32_DECIDED = 32 << 16

    def f32_filter_stats(self):
        """(pairs of the last fp32 batch, pairs its binary32 separation filter proved apart -- GJKEPA_TUNE_F32_FILTER -- or its
        binary32 walk called separated -- GJKEPA_TUNE_F32_ARITH)."""
        out = (C.c_int64 * 2)()
        self._check(self._L.gjkepa_f32_filter_stats(self._h, out), "gjkepa_f32_filter_stats")
        return int(out[0]), int(out[1])

    def set_tuning(self, key: int, value: int):
        """GJKEPA_TUNE_* knobs of include/gjkepa.h (walk re-pack threshold, single-kernel A/B switch, host chunks)."""
        self._check(self._L.gjkepa_set_tuning(self._h, int(key), int(value)), "gjkepa_set_tuning")

    def set_stream(self, cuda_stream: int):
        self._check(self._L.gjkepa_set_stream(self._h, C.c_void_p(int(cuda_stream))), "gjkepa_set_stream")

    def set_profiling(self, on: bool):
        self._check(self._L.gjkepa_set_profiling(self._h, 1 if on else 0), "gjkepa_set_profiling")

    def kernel_ms(self) -> dict:
        """Durations (ms) of K0 hull-prep, K1 GJK, K3 EPA of the last batch (needs set_profiling(True))."""
        ms = (C.c_double * 3)()
        self._check(self._L.gjkepa_kernel_ms(self._h, ms), "gjkepa_kernel_ms")
        return {"hull_prep": ms[0], "gjk": ms[1], "epa": ms[2]}

    def sync(self):
        self._check(self._L.gjkepa_sync(self._h), "gjkepa_sync")

    def launch_count(self, reset: bool = False) -> int:
        return int(self._L.gjkepa_launch_count(self._h, 1 if reset else 0))

    # ---- the output block of a host-buffer batch call and the call itself (shared by the four entry points) ----
    @staticmethod
    def _host_out(n: int, mode: int, real, want_status: bool, want_contacts: bool, compact: bool) -> dict:
        out = {"collision": np.empty(n, np.uint8)}
        if want_status:
            out["status"] = np.empty(n, np.uint32)
        if mode >= MODE_EPA:
            out["depth"] = np.empty(n, real)
            out["normal"] = np.empty((n, 3), real)
        if mode >= MODE_FULL:
            out["colli_type"] = np.empty(n, np.int32)
            out["nearest"] = np.empty((n, 6), real)
            out["point"] = np.empty((n, 3), real)
        if want_contacts or compact:
            out["contact_index"] = np.empty(n, np.int32)
        return out

    def _host_call(self, fn, what: str, head: tuple, n: int, mode: int, real, out, want_status, want_contacts, compact) -> dict:
        if out is None:
            out = self._host_out(n, mode, real, want_status, want_contacts, compact)
        for k in ("depth", "normal", "nearest", "point"):
            assert k not in out or out[k].dtype == real, f"{k} must be {np.dtype(real).name}"
        counted = want_contacts or compact or "contact_index" in out
        nc = C.c_int64(0)
        rc = fn(self._h, *head, mode | (_lib.MODE_COMPACT if compact else 0),
                _np_ptr(out.get("collision")), _np_ptr(out.get("colli_type")), _np_ptr(out.get("nearest")),
                _np_ptr(out.get("normal")), _np_ptr(out.get("point")), _np_ptr(out.get("depth")),
                _np_ptr(out.get("status")),
                C.cast(C.byref(nc), C.c_void_p) if counted else C.c_void_p(0),
                _np_ptr(out.get("contact_index")))
        self._check(rc, what)
        if counted:
            out["n_contacts"] = int(nc.value)
        return out

    # ---- HOST buffers (numpy): copies in, runs, copies out -- the reference-facing call ----
    def batch_host(self, verts, hull_off, pair_a, pair_b, mode="full", version: int = 1,
                   tol_ff: float = 1.0, out: Optional[dict] = None, want_status: bool = True,
                   want_contacts: bool = False, compact: bool = False) -> dict:
        """gjkepa_batch_f64.  compact=True (GJKEPA_MODE_COMPACT): depth/normal/point/nearest/colli_type/status hold one
        row per CONTACT (pair contact_index[k], ascending; n_contacts rows are meaningful), collision stays per pair."""
        mode = _MODES[mode]
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        hull_off = np.ascontiguousarray(hull_off, dtype=np.int64)
        pair_a = np.ascontiguousarray(pair_a, dtype=np.int32)
        pair_b = np.ascontiguousarray(pair_b, dtype=np.int32)
        n, nh = int(pair_a.shape[0]), int(hull_off.shape[0] - 1)
        head = (int(version), float(tol_ff), nh, _np_ptr(verts), _np_ptr(hull_off), n, _np_ptr(pair_a), _np_ptr(pair_b))
        return self._host_call(self._L.gjkepa_batch_f64, "gjkepa_batch_f64", head, n, mode, np.float64, out, want_status, want_contacts, compact)

    # ---- fp32 mode: binary32 at the boundary (gjkepa_batch_f32) ----
    def batch_host_f32(self, verts, hull_off, pair_a, pair_b, mode="full", version: int = 1, tol_ff: float = 1.0,
                       out: Optional[dict] = None, want_status: bool = True, want_contacts: bool = False) -> dict:
        """Same call as batch_host with a float32 hull store and float32 real-valued results (half the PCIe bytes).
        The device widens the store and runs the binary64 kernels: flags/status are those of batch_host on the
        float-rounded vertices, the reals are its doubles rounded once (include/gjkepa.h)."""
        mode = _MODES[mode]
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        hull_off = np.ascontiguousarray(hull_off, dtype=np.int64)
        pair_a = np.ascontiguousarray(pair_a, dtype=np.int32)
        pair_b = np.ascontiguousarray(pair_b, dtype=np.int32)
        n, nh = int(pair_a.shape[0]), int(hull_off.shape[0] - 1)
        head = (int(version), float(tol_ff), nh, _np_ptr(verts), _np_ptr(hull_off), n, _np_ptr(pair_a), _np_ptr(pair_b))
        return self._host_call(self._L.gjkepa_batch_f32, "gjkepa_batch_f32", head, n, mode, np.float32, out, want_status, want_contacts, False)

    # ---- hull store resident on the device(s): pair lists in, results out (gjkepa_store_upload / gjkepa_batch_resident_f64) ----
    def upload_store(self, verts, hull_off):
        """World-space hull store copied to every device of the context once (envelope spheres computed there)."""
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        hull_off = np.ascontiguousarray(hull_off, dtype=np.int64)
        self._check(self._L.gjkepa_store_upload(self._h, int(hull_off.shape[0] - 1), _np_ptr(verts), _np_ptr(hull_off)), "gjkepa_store_upload")

    def batch_resident(self, pair_a, pair_b, mode="full", version: int = 1, tol_ff: float = 1.0, out: Optional[dict] = None,
                       want_status: bool = True, want_contacts: bool = False, compact: bool = False) -> dict:
        """batch_host against the uploaded store: 8 B per pair go to the device, the results (or, with compact=True,
        only the contacts' rows) come back."""
        mode = _MODES[mode]
        pair_a = np.ascontiguousarray(pair_a, dtype=np.int32)
        pair_b = np.ascontiguousarray(pair_b, dtype=np.int32)
        n = int(pair_a.shape[0])
        head = (int(version), float(tol_ff), n, _np_ptr(pair_a), _np_ptr(pair_b))
        return self._host_call(self._L.gjkepa_batch_resident_f64, "gjkepa_batch_resident_f64", head, n, mode, np.float64, out,
                               want_status, want_contacts, compact)

    # ---- shapes resident on the device, poses per call (gjkepa_shapes_upload / gjkepa_batch_posed_f64) ----
    def upload_shapes(self, shape_verts, shape_off):
        """Body-space hull store (same layout as `verts`/`hull_off`), copied to the device once."""
        shape_verts = np.ascontiguousarray(shape_verts, dtype=np.float64)
        shape_off = np.ascontiguousarray(shape_off, dtype=np.int64)
        self._check(self._L.gjkepa_shapes_upload(self._h, int(shape_off.shape[0] - 1), _np_ptr(shape_verts), _np_ptr(shape_off)),
                    "gjkepa_shapes_upload")

    def batch_posed(self, body_shape, pose, pair_a, pair_b, mode="full", version: int = 1, tol_ff: float = 1.0,
                    out: Optional[dict] = None, want_status: bool = True, want_contacts: bool = False, compact: bool = False) -> dict:
        """pose: (n_bodies, 12) = R row-major (9) + t (3); world vertex = ((r00*x + r01*y) + r02*z) + tx, applied on
        the device.  pair_a / pair_b are body numbers.  HOST buffers in and out."""
        mode = _MODES[mode]
        body_shape = np.ascontiguousarray(body_shape, dtype=np.int32)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        pair_a = np.ascontiguousarray(pair_a, dtype=np.int32)
        pair_b = np.ascontiguousarray(pair_b, dtype=np.int32)
        nb, n = int(body_shape.shape[0]), int(pair_a.shape[0])
        assert pose.size == 12 * nb
        head = (int(version), float(tol_ff), nb, _np_ptr(body_shape), _np_ptr(pose), n, _np_ptr(pair_a), _np_ptr(pair_b))
        return self._host_call(self._L.gjkepa_batch_posed_f64, "gjkepa_batch_posed_f64", head, n, mode, np.float64, out,
                               want_status, want_contacts, compact)

    # ---- candidate pairs on the device (gjkepa_broadphase_f64 / gjkepa_broadphase_posed_f64) ----
    def _broadphase(self, call, what: str, cap: int):
        """Runs `call(cap, pair_a, pair_b, n_pairs)`; grows the buffers once when the first guess was too small."""
        cap = max(int(cap), 0)
        while True:
            pa, pb = np.empty(cap, np.int32), np.empty(cap, np.int32)
            total = C.c_int64(0)
            self._check(call(cap, _np_ptr(pa) if cap else C.c_void_p(0), _np_ptr(pb) if cap else C.c_void_p(0),
                             C.cast(C.byref(total), C.c_void_p)), what)
            if total.value <= cap:
                return pa[:total.value], pb[:total.value]
            cap = int(total.value)

    def broadphase(self, verts, hull_off, cap: int = 0):
        """All pairs of hulls (a < b) that pass the reference's spherical-envelope test (F:1165-1188), found through
        a uniform grid on the device; returns (pair_a, pair_b), int32, in a deterministic order.  HOST buffers."""
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        hull_off = np.ascontiguousarray(hull_off, dtype=np.int64)
        nh = int(hull_off.shape[0] - 1)
        return self._broadphase(
            lambda c, pa, pb, tot: self._L.gjkepa_broadphase_f64(self._h, nh, _np_ptr(verts), _np_ptr(hull_off), c, pa, pb, tot),
            "gjkepa_broadphase_f64", cap or 8 * nh)

    def broadphase_posed(self, body_shape, pose, cap: int = 0):
        """Same over posed bodies of the resident shape store (upload_shapes); pose as in batch_posed."""
        body_shape = np.ascontiguousarray(body_shape, dtype=np.int32)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        nb = int(body_shape.shape[0])
        assert pose.size == 12 * nb
        return self._broadphase(
            lambda c, pa, pb, tot: self._L.gjkepa_broadphase_posed_f64(self._h, nb, _np_ptr(body_shape), _np_ptr(pose), c, pa, pb, tot),
            "gjkepa_broadphase_posed_f64", cap or 8 * nb)

    # ---- scene step: poses in, contacts out (gjkepa_scene_posed_f64) ----
    def scene_posed(self, body_shape, pose, mode="epa", version: int = 1, tol_ff: float = 1.0, cap: int = 0) -> dict:
        """Pose application + candidate pairs + GJK/EPA over them in one call; returns only the CONTACTS:
        body_a, body_b (a < b, rows sorted by (body_a, body_b)), depth, normal (mode >= "epa"), point, nearest, colli_type
        (mode "full"), n_candidates, n_contacts.  Same results as broadphase_posed -> batch_posed -> keep collisions."""
        mode = _MODES[mode]
        body_shape = np.ascontiguousarray(body_shape, dtype=np.int32)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        nb = int(body_shape.shape[0])
        assert pose.size == 12 * nb
        cap = int(cap) or 4 * nb
        while True:
            out = {"body_a": np.empty(cap, np.int32), "body_b": np.empty(cap, np.int32)}
            if mode >= MODE_EPA:
                out["depth"] = np.empty(cap, np.float64)
                out["normal"] = np.empty((cap, 3), np.float64)
            if mode >= MODE_FULL:
                out["point"] = np.empty((cap, 3), np.float64)
                out["nearest"] = np.empty((cap, 6), np.float64)
                out["colli_type"] = np.empty(cap, np.int32)
            ncand, nc = C.c_int64(0), C.c_int64(0)
            rc = self._L.gjkepa_scene_posed_f64(
                self._h, int(version), float(tol_ff), nb, _np_ptr(body_shape), _np_ptr(pose), mode, cap,
                _np_ptr(out["body_a"]), _np_ptr(out["body_b"]), _np_ptr(out.get("depth")), _np_ptr(out.get("normal")),
                _np_ptr(out.get("point")), _np_ptr(out.get("nearest")), _np_ptr(out.get("colli_type")),
                C.cast(C.byref(ncand), C.c_void_p), C.cast(C.byref(nc), C.c_void_p))
            self._check(rc, "gjkepa_scene_posed_f64")
            if nc.value <= cap:
                break
            cap = int(nc.value)
        n = int(nc.value)
        out = {k: v[:n] for k, v in out.items()}
        out["n_candidates"], out["n_contacts"] = int(ncand.value), n
        return out

    # ---- DEVICE buffers (torch CUDA tensors already resident in HBM), asynchronous ----
    def batch_device(self, verts, hull_off, pair_a, pair_b, mode="full", version: int = 1,
                     tol_ff: float = 1.0, out: Optional[dict] = None, want_status: bool = True,
                     want_contacts: bool = False) -> dict:
        import torch
        mode = _MODES[mode]
        dev = verts.device
        assert dev.type == "cuda" and verts.dtype == torch.float64 and hull_off.dtype == torch.int64
        assert self.n_devices == 1 and dev.index == self.device, "tensors must live on the engine's own device"
        assert all(t.device == dev for t in (hull_off, pair_a, pair_b)), "all inputs on one device"
        assert pair_a.dtype == torch.int32 and pair_b.dtype == torch.int32
        n = int(pair_a.shape[0])
        nh = int(hull_off.shape[0] - 1)
        if out is None:
            out = {"collision": torch.empty(n, dtype=torch.uint8, device=dev)}
            if want_status:
                out["status"] = torch.empty(n, dtype=torch.int32, device=dev)
            if mode >= MODE_EPA:
                out["depth"] = torch.empty(n, dtype=torch.float64, device=dev)
                out["normal"] = torch.empty((n, 3), dtype=torch.float64, device=dev)
            if mode >= MODE_FULL:
                out["colli_type"] = torch.empty(n, dtype=torch.int32, device=dev)
                out["nearest"] = torch.empty((n, 6), dtype=torch.float64, device=dev)
                out["point"] = torch.empty((n, 3), dtype=torch.float64, device=dev)
            if want_contacts:
                out["n_contacts"] = torch.zeros(1, dtype=torch.int64, device=dev)
                out["contact_index"] = torch.empty(n, dtype=torch.int32, device=dev)

        def p(t):
            return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

        self.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        rc = self._L.gjkepa_batch_device_f64(
            self._h, int(version), float(tol_ff), nh, p(verts), p(hull_off), n, p(pair_a), p(pair_b), mode,
            p(out.get("collision")), p(out.get("colli_type")), p(out.get("nearest")), p(out.get("normal")),
            p(out.get("point")), p(out.get("depth")), p(out.get("status")),
            p(out.get("n_contacts")), p(out.get("contact_index")))
        self._check(rc, "gjkepa_batch_device_f64")
        return out


def guard_check():
    """Checking mode (``GJKEPA_GUARD=1`` in the environment before the first call, include/gjkepa.h): returns
    ``(live device allocations, guard bands that no longer hold their pattern)``; ``None`` outside checking mode."""
    L = _lib.lib()
    n, bad = C.c_int64(0), C.c_int64(0)
    if L.gjkepa_guard_check(C.byref(n), C.byref(bad)) != _lib.OK:
        return None
    return int(n.value), int(bad.value)


def GJKEPA(version_: int, TOL_FF_: float, p1_, p2_):
    """Drop-in for ``CALL GJKEPA(version_, TOL_FF_, p1_, p2_, collision_, colliType_,
    nearest_points_, collision_normal_, collision_point_, penetration_depth_)`` (F:39-52).

    p1_, p2_: (n,3) vertex arrays.  Returns ``(collision_, colliType_, nearest_points_(2,3),
    collision_normal_(3), collision_point_(3), penetration_depth_)``.  Goes through the C ABI's
    ``gjkepa_pair_f64`` with every scalar passed by reference, as a Fortran caller would.
    """
    L = _lib.lib()
    p1 = np.asfortranarray(np.asarray(p1_, dtype=np.float64))      # column-major (n,3) = [x|y|z]
    p2 = np.asfortranarray(np.asarray(p2_, dtype=np.float64))
    if p1.ndim != 2 or p1.shape[1] != 3 or p2.ndim != 2 or p2.shape[1] != 3:
        raise ValueError("p1_, p2_ must be (n,3)")
    ver, tol = C.c_int32(int(version_)), C.c_double(float(TOL_FF_))
    n1, n2 = C.c_int32(p1.shape[0]), C.c_int32(p2.shape[0])
    col, typ, dep = C.c_uint8(0), C.c_int32(0), C.c_double(0.0)
    near = np.zeros((2, 3), order="F")
    nrm, pt = np.zeros(3), np.zeros(3)
    dp = C.POINTER(C.c_double)
    rc = L.gjkepa_pair_f64(C.byref(ver), C.byref(tol), p1.ctypes.data_as(dp), C.byref(n1),
                           p2.ctypes.data_as(dp), C.byref(n2), C.byref(col), C.byref(typ),
                           near.ctypes.data_as(dp), nrm.ctypes.data_as(dp), pt.ctypes.data_as(dp), C.byref(dep))
    if rc != _lib.OK:
        raise GjkEpaError(f"gjkepa_pair_f64 failed with code {rc} (no usable sm_100a CUDA device?)")
    return bool(col.value), int(typ.value), np.ascontiguousarray(near), nrm, pt, float(dep.value)
