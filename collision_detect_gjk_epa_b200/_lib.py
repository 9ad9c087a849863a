"""ctypes binding of libgjkepa_b200.so (the C ABI of include/gjkepa.h).

The library is built IN-TREE (collision_detect_gjk_epa_b200/libgjkepa_b200.so) by
``__graft_entry__.build()`` / ``make -C collision_detect_gjk_epa_b200/csrc``.  There is no fallback:
if the shared object is missing, or no sm_100a device can run it, importing callers get an error.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
SO_PATH = Path(os.environ["GJKEPA_LIB"]) if os.environ.get("GJKEPA_LIB") else _PKG / "libgjkepa_b200.so"   # env override: developer A/B builds
CSRC = _PKG / "csrc"

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM = 0, 1, 2, 3
MODE_GJK, MODE_EPA, MODE_FULL = 0, 1, 2
MODE_COMPACT = 16

# every symbol include/gjkepa.h declares
SYMBOLS = [
    "gjkepa_create", "gjkepa_destroy", "gjkepa_set_stream", "gjkepa_use_own_stream", "gjkepa_last_error",
    "gjkepa_launch_count", "gjkepa_set_lanes", "gjkepa_set_max_hull_vertices", "gjkepa_set_tuning", "gjkepa_set_profiling", "gjkepa_kernel_ms", "gjkepa_pair_f64", "gjkepa_batch_f64",
    "gjkepa_batch_device_f64", "gjkepa_sync", "gjkepa_debug_fp64", "gjkepa_shapes_upload", "gjkepa_batch_posed_f64", "gjkepa_broadphase_f64", "gjkepa_broadphase_posed_f64", "gjkepa_batch_f32", "gjkepa_scene_posed_f64",
    "gjkepa_create_multi", "gjkepa_device_count", "gjkepa_host_register", "gjkepa_host_unregister", "gjkepa_store_upload", "gjkepa_batch_resident_f64", "gjkepa_host_phase_ms", "gjkepa_f32_filter_stats",
    "gjkepa_guard_check", "gjkepa_guard_selftest",
]


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc")) + list(CSRC.glob("*.inl")) + list(CSRC.glob("*.h")) + \
        [_PKG.parent / "include" / "gjkepa.h"]
    stale = (not SO_PATH.exists()) or any(s.stat().st_mtime > SO_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        cmd = ["make", "-C", str(CSRC)] + (["-B"] if force else [])
        r = subprocess.run(cmd, env=env, capture_output=not verbose, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libgjkepa_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not SO_PATH.exists():
        raise RuntimeError(
            f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the GJK/EPA path)")
    L = C.CDLL(str(SO_PATH))
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.gjkepa_create.argtypes = [C.POINTER(vp), C.c_int32]
    L.gjkepa_create.restype = C.c_int
    L.gjkepa_guard_check.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.gjkepa_guard_check.restype = C.c_int
    L.gjkepa_guard_selftest.argtypes = [C.c_int64]
    L.gjkepa_guard_selftest.restype = C.c_int
    L.gjkepa_create_multi.argtypes = [C.POINTER(vp), ip, C.c_int32]
    L.gjkepa_create_multi.restype = C.c_int
    L.gjkepa_device_count.argtypes = [vp]
    L.gjkepa_device_count.restype = C.c_int32
    L.gjkepa_f32_filter_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    L.gjkepa_f32_filter_stats.restype = C.c_int
    L.gjkepa_host_phase_ms.argtypes = [vp, dp]
    L.gjkepa_host_phase_ms.restype = C.c_int
    L.gjkepa_host_register.argtypes = [vp, C.c_int64]
    L.gjkepa_host_register.restype = C.c_int
    L.gjkepa_host_unregister.argtypes = [vp]
    L.gjkepa_host_unregister.restype = C.c_int
    L.gjkepa_store_upload.argtypes = [vp, C.c_int64, vp, vp]
    L.gjkepa_store_upload.restype = C.c_int
    L.gjkepa_batch_resident_f64.argtypes = [vp, C.c_int32, C.c_double, C.c_int64, vp, vp, C.c_int32,
                                            vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gjkepa_batch_resident_f64.restype = C.c_int
    L.gjkepa_destroy.argtypes = [vp]
    L.gjkepa_destroy.restype = C.c_int
    L.gjkepa_set_stream.argtypes = [vp, vp]
    L.gjkepa_set_stream.restype = C.c_int
    L.gjkepa_use_own_stream.argtypes = [vp]
    L.gjkepa_use_own_stream.restype = C.c_int
    L.gjkepa_last_error.argtypes = [vp]
    L.gjkepa_last_error.restype = C.c_char_p
    L.gjkepa_launch_count.argtypes = [vp, C.c_int32]
    L.gjkepa_launch_count.restype = C.c_int64
    L.gjkepa_set_lanes.argtypes = [vp, C.c_int32]
    L.gjkepa_set_lanes.restype = C.c_int
    L.gjkepa_set_max_hull_vertices.argtypes = [vp, C.c_int32]
    L.gjkepa_set_max_hull_vertices.restype = C.c_int
    L.gjkepa_set_tuning.argtypes = [vp, C.c_int32, C.c_int32]
    L.gjkepa_set_tuning.restype = C.c_int
    L.gjkepa_set_profiling.argtypes = [vp, C.c_int32]
    L.gjkepa_set_profiling.restype = C.c_int
    L.gjkepa_kernel_ms.argtypes = [vp, dp]
    L.gjkepa_kernel_ms.restype = C.c_int
    L.gjkepa_sync.argtypes = [vp]
    L.gjkepa_sync.restype = C.c_int
    L.gjkepa_pair_f64.argtypes = [ip, dp, dp, ip, dp, ip, C.POINTER(C.c_uint8), ip, dp, dp, dp, dp]
    L.gjkepa_pair_f64.restype = C.c_int
    batch_args = [vp, C.c_int32, C.c_double, C.c_int64, vp, vp, C.c_int64, vp, vp, C.c_int32,
                  vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gjkepa_batch_f64.argtypes = batch_args
    L.gjkepa_batch_f64.restype = C.c_int
    L.gjkepa_batch_device_f64.argtypes = batch_args
    L.gjkepa_batch_device_f64.restype = C.c_int
    L.gjkepa_shapes_upload.argtypes = [vp, C.c_int64, vp, vp]
    L.gjkepa_shapes_upload.restype = C.c_int
    L.gjkepa_batch_posed_f64.argtypes = [vp, C.c_int32, C.c_double, C.c_int64, vp, vp, C.c_int64, vp, vp, C.c_int32,
                                         vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gjkepa_batch_posed_f64.restype = C.c_int
    L.gjkepa_batch_f32.argtypes = [vp, C.c_int32, C.c_float, C.c_int64, vp, vp, C.c_int64, vp, vp, C.c_int32,
                                   vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gjkepa_batch_f32.restype = C.c_int
    L.gjkepa_scene_posed_f64.argtypes = [vp, C.c_int32, C.c_double, C.c_int64, vp, vp, C.c_int32, C.c_int64,
                                         vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gjkepa_scene_posed_f64.restype = C.c_int
    L.gjkepa_broadphase_f64.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, vp, vp, vp]
    L.gjkepa_broadphase_f64.restype = C.c_int
    L.gjkepa_broadphase_posed_f64.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, vp, vp, vp]
    L.gjkepa_broadphase_posed_f64.restype = C.c_int
    L.gjkepa_debug_fp64.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp]
    L.gjkepa_debug_fp64.restype = C.c_int
    _lib = L
    return L
