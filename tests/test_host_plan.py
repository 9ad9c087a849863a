"""Host logic of the multi-GPU / chunked host batch (no GPU needed): csrc/host_plan.h, compiled with g++.

Pairs are independent (reference src/GCLIB_GJKEPA.f90:55-60), so the only thing the planning must guarantee is that
every pair lands in exactly one shard and one chunk, and that a chunk which brings its own hull range brings every hull
its pairs name."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def plan():
    env = dict(os.environ); env.pop("CC", None)
    subprocess.run(["make", "-C", str(ROOT / "tests" / "host_sim"), "-s", "libhost_plan_sim.so"], check=True, env=env)
    L = C.CDLL(str(ROOT / "tests" / "host_sim" / "libhost_plan_sim.so"))
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    L.plan_shard_range.argtypes = [C.c_int64, C.c_int32, C.c_int32, i64p, i64p]
    L.plan_shards_for.argtypes = [C.c_int64, C.c_int32]
    L.plan_shards_for.restype = C.c_int32
    L.plan_shard_c.argtypes = [C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int64, i64p, i32p, i32p, i64p, i64p]
    return L


def _plan(L, np_, forced, bpp, scan, hull_off, pa, pb):
    out = np.zeros(6, np.int64); ch = np.zeros(4 * 64, np.int64)
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    L.plan_shard_c(np_, forced, bpp, int(scan), len(hull_off) - 1, hull_off.ctypes.data_as(i64p), pa.ctypes.data_as(i32p),
                   pb.ctypes.data_as(i32p), out.ctypes.data_as(i64p), ch.ctypes.data_as(i64p))
    return dict(nchunk=int(out[0]), cp=int(out[1]), ranged=bool(out[2]), Hlo=int(out[3]), Hhi=int(out[4]), bad=bool(out[5]),
                chunks=ch[:4 * int(out[0])].reshape(-1, 4))


def test_shards_tile_the_pair_list(plan):
    lo, hi = C.c_int64(), C.c_int64()
    for n_pairs in list(range(0, 40)) + [4095, 4096, 12345, 10 ** 8 + 7]:
        for n in range(1, 10):
            edge, sizes = 0, []
            for s in range(n):
                plan.plan_shard_range(n_pairs, s, n, C.byref(lo), C.byref(hi))
                assert lo.value == edge and hi.value >= lo.value
                edge = hi.value; sizes.append(hi.value - lo.value)
            assert edge == n_pairs and max(sizes) - min(sizes) <= 1
    assert plan.plan_shards_for(100, 8) == 1 and plan.plan_shards_for(3 * 4096, 8) == 3 and plan.plan_shards_for(10 ** 6, 8) == 8


@pytest.mark.parametrize("forced", [0, 1, 3, 8, 64, 200])
def test_chunks_tile_a_shard_and_bring_the_hulls_their_pairs_name(plan, forced):
    rng = np.random.default_rng(forced)
    n = 300_000
    nv = rng.integers(4, 20, size=2 * n)
    hull_off = np.concatenate([[0], np.cumsum(nv)]).astype(np.int64)
    pa = (2 * np.arange(n) + rng.integers(-50, 50, size=n)).clip(0, 2 * n - 1).astype(np.int32)      # pair-private, roughly ordered
    pb = (2 * np.arange(n) + 1).astype(np.int32)
    P = _plan(plan, n, forced, 400.0, True, hull_off, pa, pb)
    assert not P["bad"] and 1 <= P["nchunk"] <= 64 and P["ranged"]
    if forced in (1, 3, 8, 64):
        assert P["nchunk"] == forced or P["cp"] % 1024 == 0
    edge = 0
    for a, b, hlo, hhi in P["chunks"]:
        assert a == edge and b > a
        edge = b
        assert hlo == min(pa[a:b].min(), pb[a:b].min()) and hhi == max(pa[a:b].max(), pb[a:b].max())
    assert edge == n
    assert P["Hlo"] == min(pa.min(), pb.min()) and P["Hhi"] == max(pa.max(), pb.max())


def test_lazy_plan_decides_from_a_sample_and_scans_chunk_by_chunk(plan):
    rng = np.random.default_rng(7)
    n = 600_000
    hull_off = (np.arange(2 * n + 1) * 8).astype(np.int64)
    pa = (2 * np.arange(n)).astype(np.int32); pb = pa + 1
    full = _plan(plan, n, 6, 400.0, 1, hull_off, pa, pb)
    lazy = _plan(plan, n, 6, 400.0, 2, hull_off, pa, pb)
    assert lazy["ranged"] and not lazy["bad"] and np.array_equal(lazy["chunks"], full["chunks"])
    ra = rng.integers(0, 2 * n, size=n).astype(np.int32)              # random list over the same store: the sample sees it
    assert not _plan(plan, n, 6, 400.0, 2, hull_off, ra, pb)["ranged"]
    bad = pa.copy(); bad[123_457] = 2 * n                             # not on the sampling grid: found by the exact scan of its chunk
    assert _plan(plan, n, 6, 400.0, 2, hull_off, bad, pb)["bad"]


def test_shared_and_random_stores_go_up_whole(plan):
    rng = np.random.default_rng(1)
    n, nh = 400_000, 300_000
    hull_off = (np.arange(nh + 1) * 16).astype(np.int64)
    pa = rng.integers(0, nh, size=n).astype(np.int32); pb = rng.integers(0, nh, size=n).astype(np.int32)
    P = _plan(plan, n, 4, 100.0, True, hull_off, pa, pb)         # scanned, but every chunk names the whole store
    assert not P["ranged"] and P["nchunk"] == 4
    P = _plan(plan, n, 0, 9.0, False, hull_off, pa, pb)          # resident / shared: not scanned, one piece (3.6 MB of traffic)
    assert not P["ranged"] and P["nchunk"] == 1 and (P["Hlo"], P["Hhi"]) == (0, nh - 1)
    P = _plan(plan, n, 0, 400.0, False, hull_off, pa, pb)        # 160 MB = 4.8 x 32 MiB: 4 chunks of at least 64k pairs
    assert P["nchunk"] == 4 and all(b - a >= 65536 for a, b, _, _ in P["chunks"][:-1])
    pa[7] = nh                                                   # one past the last hull
    assert _plan(plan, n, 2, 100.0, True, hull_off, pa, pb)["bad"]
