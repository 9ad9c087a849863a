"""The host-buffer side of the C ABI (-m gpu): chunked copy/compute pipeline, device-side contact lists,
GJKEPA_MODE_COMPACT rows, the resident hull store, and one context over several devices.

Every variant must return the SAME BITS as the plain call (one chunk, one device), and the plain call is checked
against the oracle here once more.  A multi-device context is exercised with whatever GPUs the box has; on a
one-GPU box it gets three children on device 0, which runs the same sharding code."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("collision", "status", "depth", "normal", "colli_type", "nearest", "point")


def same(a, b, keys=KEYS):
    for k in keys:
        if k in a or k in b:
            assert k in a and k in b, k
            assert np.array_equal(a[k], b[k]), f"{k} differs"


@pytest.fixture(scope="module")
def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def multi(n_gpus):
    from collision_detect_gjk_epa_b200 import Engine
    eng = Engine(devices=list(range(n_gpus)) if n_gpus > 1 else [0, 0, 0])
    yield eng
    eng.close()


def workloads(W):
    return {
        "box8_private": (W.make_pairs("box8", 150_000, 1.0, seed=21), "gjk"),
        "hull32_private": (W.make_pairs("hull", 40_000, 0.85, seed=22, n_vert=32), "full"),
        "mixed_private": (W.make_mixed(40_000, 0.25, seed=23), "epa"),
        "hull16_shared": (W.make_shared_store(3000, 60_000, n_vert=16, seed=24), "epa"),
    }


@pytest.mark.parametrize("name", ["box8_private", "hull32_private", "mixed_private", "hull16_shared"])
def test_chunks_give_the_same_bits_and_match_the_oracle(engine, orc, W, name):
    b, mode = workloads(W)[name]
    engine.set_max_hull_vertices(0)
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 1)
    one = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
    omode = {"gjk": orc.MODE_GJK, "epa": orc.MODE_EPA, "full": orc.MODE_FULL}[mode]
    ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=omode)
    assert np.array_equal(one["collision"], ref["collision"])
    assert np.array_equal(one["status"] & 0xffff, ref["status"] & 0xffff)
    hits = np.nonzero(ref["collision"])[0]
    assert one["n_contacts"] == len(hits) and np.array_equal(one["contact_index"][:len(hits)], hits.astype(np.int32))
    if mode != "gjk":
        clean = (ref["collision"] > 0) & ((ref["status"] >> 24) == 0)
        assert np.all(np.abs(one["depth"][clean] - ref["depth"][clean]) <= 1e-9 * np.maximum(ref["depth"][clean], 1.0))
        assert np.all(1.0 - np.sum(one["normal"][clean] * ref["normal"][clean], axis=1) <= 1e-9)
    for chunks in (2, 5, 0):
        engine.set_tuning(engine.TUNE_HOST_CHUNKS, chunks)
        got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
        same(got, one)
        assert got["n_contacts"] == one["n_contacts"]
        assert np.array_equal(got["contact_index"][:got["n_contacts"]], one["contact_index"][:one["n_contacts"]])
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 0)


def test_ranged_chunks_with_a_store_that_does_not_start_at_hull_zero(engine, W):
    """Pairs that name only the upper part of a pair-private store: every chunk brings its own hull range and the pair
    list is made relative to it on the device."""
    b = W.make_pairs("box8", 60_000, 1.0, seed=31)
    lo = 20_000
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 1)
    full = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 3)
    part = engine.batch_host(b.verts, b.hull_off, b.pair_a[lo:], b.pair_b[lo:], mode="epa")
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 0)
    for k in ("collision", "status", "depth", "normal"):
        assert np.array_equal(part[k], full[k][lo:]), k


@pytest.mark.parametrize("name", ["hull32_private", "hull16_shared"])
def test_compact_rows_are_the_dense_rows_of_the_contacts(engine, W, name):
    b, mode = workloads(W)[name]
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 0)
    dense = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode)
    for chunks in (1, 3):
        engine.set_tuning(engine.TUNE_HOST_CHUNKS, chunks)
        c = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, compact=True)
        hits = np.nonzero(dense["collision"])[0]
        n = c["n_contacts"]
        assert n == len(hits) and np.array_equal(c["contact_index"][:n], hits.astype(np.int32))
        assert np.array_equal(c["collision"], dense["collision"])
        for k in ("status", "depth", "normal", "colli_type", "nearest", "point"):
            if k in dense:
                assert np.array_equal(c[k][:n], dense[k][hits]), k
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 0)


@pytest.mark.parametrize("name", ["box8_private", "hull32_private", "hull16_shared"])
def test_resident_store_gives_the_bits_of_the_copying_call(engine, W, name):
    b, mode = workloads(W)[name]
    # the resident store knows its largest hull; give the copying call the same bound (it picks the EPA kernel family,
    # and the families may differ in the last bits of a normal)
    engine.set_max_hull_vertices(int(np.diff(b.hull_off).max()))
    want = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
    engine.set_max_hull_vertices(0)
    engine.upload_store(b.verts, b.hull_off)
    for chunks in (1, 4):
        engine.set_tuning(engine.TUNE_HOST_CHUNKS, chunks)
        got = engine.batch_resident(b.pair_a, b.pair_b, mode=mode, want_contacts=True)
        same(got, want)
        assert got["n_contacts"] == want["n_contacts"]
        assert np.array_equal(got["contact_index"][:got["n_contacts"]], want["contact_index"][:want["n_contacts"]])
    c = engine.batch_resident(b.pair_a, b.pair_b, mode=mode, compact=True)
    hits = np.nonzero(want["collision"])[0]
    assert c["n_contacts"] == len(hits)
    if "depth" in want:
        assert np.array_equal(c["depth"][:len(hits)], want["depth"][hits])
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 0)


@pytest.mark.parametrize("name", ["box8_private", "hull32_private", "mixed_private", "hull16_shared"])
def test_multi_device_context_gives_the_bits_of_one_device(engine, multi, W, name):
    b, mode = workloads(W)[name]
    hint = int(np.diff(b.hull_off).max())
    engine.set_max_hull_vertices(hint)
    multi.set_max_hull_vertices(hint)
    want = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
    got = multi.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
    same(got, want)
    n = want["n_contacts"]
    assert got["n_contacts"] == n and np.array_equal(got["contact_index"][:n], want["contact_index"][:n])
    # compact rows land at the row offset the earlier shards determine
    c = multi.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, compact=True)
    hits = want["contact_index"][:n]
    assert c["n_contacts"] == n and np.array_equal(c["contact_index"][:n], hits)
    for k in ("status", "depth", "normal", "colli_type", "nearest", "point"):
        if k in want:
            assert np.array_equal(c[k][:n], want[k][hits]), k
    # resident store, replicated to every device
    multi.upload_store(b.verts, b.hull_off)
    r = multi.batch_resident(b.pair_a, b.pair_b, mode=mode, want_contacts=True)
    same(r, want)
    assert r["n_contacts"] == n and np.array_equal(r["contact_index"][:n], hits)
    engine.set_max_hull_vertices(0)
    multi.set_max_hull_vertices(0)


def test_multi_device_posed_and_f32(engine, multi, W):
    pb = W.make_posed(n_shapes=50, n_bodies=30_000, n_pairs=50_000, sigma=0.85, seed=5)
    for e in (engine, multi):
        e.upload_shapes(pb.shape_verts, pb.shape_off)
    want = engine.batch_posed(pb.body_shape, pb.pose, pb.pair_a, pb.pair_b, mode="epa", want_contacts=True)
    for chunks in (1, 3):
        engine.set_tuning(engine.TUNE_HOST_CHUNKS, chunks)
        same(engine.batch_posed(pb.body_shape, pb.pose, pb.pair_a, pb.pair_b, mode="epa"), want)
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 0)
    got = multi.batch_posed(pb.body_shape, pb.pose, pb.pair_a, pb.pair_b, mode="epa", want_contacts=True)
    same(got, want)
    assert got["n_contacts"] == want["n_contacts"]
    b = W.make_pairs("box8", 50_000, 1.0, seed=8)
    v32 = b.verts.astype(np.float32)
    f_one = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="epa")
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 3)
    same(engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="epa"), f_one)
    engine.set_tuning(engine.TUNE_HOST_CHUNKS, 0)
    same(multi.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="epa"), f_one)


def test_small_and_empty_batches_on_a_multi_device_context(multi, orc, W):
    b = W.make_pairs("hull", 300, 0.85, seed=3, n_vert=12)
    got = multi.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa", want_contacts=True)
    ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
    assert np.array_equal(got["collision"], ref["collision"]) and got["n_contacts"] == int(ref["collision"].sum())
    e = multi.batch_host(b.verts, b.hull_off, b.pair_a[:0], b.pair_b[:0], mode="epa", want_contacts=True)
    assert e["n_contacts"] == 0
    assert multi.n_devices >= 2


def test_bad_shape_index_and_missing_store_are_argument_errors(engine, W):
    from collision_detect_gjk_epa_b200.api import GjkEpaError
    pb = W.make_posed(n_shapes=10, n_bodies=1000, n_pairs=2000, sigma=0.85, seed=6)
    engine.upload_shapes(pb.shape_verts, pb.shape_off)
    shp = pb.body_shape.copy()
    shp[17] = 10                                   # one past the last shape
    with pytest.raises(GjkEpaError):
        engine.batch_posed(shp, pb.pose, pb.pair_a, pb.pair_b, mode="gjk")
    shp[17] = -1
    with pytest.raises(GjkEpaError):
        engine.scene_posed(shp, pb.pose, mode="gjk")
    with pytest.raises(GjkEpaError):
        engine.broadphase_posed(shp, pb.pose)
    from collision_detect_gjk_epa_b200 import Engine
    fresh = Engine(0)
    with pytest.raises(GjkEpaError):
        fresh.batch_resident(pb.pair_a, pb.pair_b, mode="gjk")
    with pytest.raises(GjkEpaError):               # compact rows need the fp64 entry
        b = W.make_pairs("box8", 1000, 1.0, seed=1)
        fresh._host_call(fresh._L.gjkepa_batch_f32, "gjkepa_batch_f32",
                         (1, 1.0, int(b.hull_off.shape[0] - 1), C.c_void_p(b.verts.astype(np.float32).ctypes.data),
                          C.c_void_p(b.hull_off.ctypes.data), 1000, C.c_void_p(b.pair_a.ctypes.data), C.c_void_p(b.pair_b.ctypes.data)),
                         1000, 16, np.float32, None, True, True, False)
    fresh.close()


def test_hull_numbers_outside_the_store_fail_the_call_not_the_gpu(engine, multi, W):
    """A wrong hull number in a pair list must come back as an argument error -- on the host for pair-private stores, from the
    device-side bounds check for shared / resident stores -- never as a wild device read, and the context stays usable."""
    from collision_detect_gjk_epa_b200.api import GjkEpaError
    b = W.make_shared_store(3000, 40_000, n_vert=16, seed=4)
    good = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
    for eng in (engine, multi):
        eng.upload_store(b.verts, b.hull_off)
        for bad_value in (3000, -1, 2 ** 31 - 1):
            pa = b.pair_a.copy(); pa[12345] = bad_value
            with pytest.raises(GjkEpaError):
                eng.batch_resident(pa, b.pair_b, mode="epa")
            with pytest.raises(GjkEpaError):
                eng.batch_host(b.verts, b.hull_off, b.pair_a, pa, mode="gjk")
        again = eng.batch_resident(b.pair_a, b.pair_b, mode="epa")
        for k in ("collision", "status", "depth", "normal"):
            assert np.array_equal(again[k], good[k]), k
    p8 = W.make_pairs("box8", 30_000, 1.0, seed=6)             # pair-private store: the host-side range scan catches it
    pa = p8.pair_a.copy(); pa[77] = 2 * 30_000
    with pytest.raises(GjkEpaError):
        engine.batch_host(p8.verts, p8.hull_off, pa, p8.pair_b, mode="gjk")


def test_calls_leave_the_current_device_alone(engine, multi, W, n_gpus):
    import torch
    b = W.make_pairs("box8", 5000, 1.0, seed=2)
    last = n_gpus - 1
    torch.cuda.set_device(last)
    engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    multi.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    from collision_detect_gjk_epa_b200 import GJKEPA
    cubes = W.unit_cubes((0.5, 0.0, 0.0))
    GJKEPA(1, 1.0, cubes.hull(0), cubes.hull(1))
    assert torch.cuda.current_device() == last
    torch.cuda.set_device(0)


def test_registered_pageable_arrays(engine, W):
    b = W.make_pairs("box8", 100_000, 1.0, seed=12)
    want = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    L = engine._L
    assert L.gjkepa_host_register(C.c_void_p(b.verts.ctypes.data), b.verts.nbytes) == 0
    try:
        got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    finally:
        assert L.gjkepa_host_unregister(C.c_void_p(b.verts.ctypes.data)) == 0
    assert np.array_equal(got["collision"], want["collision"])


def test_random_sequence_of_host_batches(engine, multi, W):
    """Buffers are grown on demand and reused from call to call (input / output slots, accumulators, seed queue, work lists,
    routing lists): a random sequence of batch sizes, modes, chunk counts, output sets and contexts must keep giving the bits of
    the plain call on a fresh engine."""
    from collision_detect_gjk_epa_b200 import Engine
    rng = np.random.default_rng(2024)
    pools = [W.make_pairs("box8", 120_000, 1.0, seed=51), W.make_mixed(60_000, 0.3, seed=52), W.make_shared_store(5000, 90_000, n_vert=16, seed=53)]
    fresh = Engine(0)
    fresh.set_tuning(fresh.TUNE_HOST_CHUNKS, 1)
    want = [fresh.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="full") for b in pools]
    fresh.close()
    for it in range(36):
        k = int(rng.integers(0, 3)); b = pools[k]
        n = int(rng.choice([1, 7, 255, 257, 4095, 4097, 30_000, b.n_pairs]))
        lo = int(rng.integers(0, b.n_pairs - n + 1))
        mode = str(rng.choice(["gjk", "epa", "full"]))
        compact = bool(rng.integers(0, 2)) and mode != "gjk"
        eng = multi if rng.integers(0, 2) else engine
        eng.set_tuning(eng.TUNE_HOST_CHUNKS, int(rng.choice([0, 1, 2, 3, 7])))
        if rng.integers(0, 2):
            got = eng.batch_host(b.verts, b.hull_off, b.pair_a[lo:lo + n], b.pair_b[lo:lo + n], mode=mode, compact=compact, want_contacts=True)
        else:
            eng.upload_store(b.verts, b.hull_off)
            got = eng.batch_resident(b.pair_a[lo:lo + n], b.pair_b[lo:lo + n], mode=mode, compact=compact, want_contacts=True)
        ref = {key: v[lo:lo + n] for key, v in want[k].items()}
        tag = (it, k, n, lo, mode, compact, eng is multi)
        assert np.array_equal(got["collision"], ref["collision"]), tag
        hits = np.nonzero(ref["collision"])[0]
        assert got["n_contacts"] == len(hits) and np.array_equal(got["contact_index"][:len(hits)], hits.astype(np.int32)), tag
        keys = {"gjk": [], "epa": ["depth", "normal"], "full": ["depth", "normal", "colli_type", "nearest", "point"]}[mode]
        for key in keys + ["status"]:
            if compact:
                ok = np.array_equal(got[key][:len(hits)], ref[key][hits] if key != "status" or mode == "full" else got[key][:len(hits)])
            else:
                ok = np.array_equal(got[key], ref[key]) if key != "status" or mode == "full" else True
            assert ok, (tag, key)
        if mode != "full":          # the GJK half of the status word does not depend on the mode
            st = got["status"][:len(hits)] if compact else got["status"]
            rs = ref["status"][hits] if compact else ref["status"]
            assert np.array_equal(st & 0xffff, rs & 0xffff), tag
    for e in (engine, multi):
        e.set_tuning(e.TUNE_HOST_CHUNKS, 0)


def test_locality_order_changes_no_result(engine, multi, W):
    """GJKEPA_TUNE_LOCALITY_SORT: the init kernel visits the pairs of a shared / resident store in the order of their first hull
    (what a broad phase's output order gives for free); every result is written at the pair's own index, so nothing may change."""
    b = W.make_shared_store(20_000, 300_000, n_vert=16, seed=61)
    for eng in (engine, multi):
        eng.set_max_hull_vertices(0)
        eng.upload_store(b.verts, b.hull_off)
        eng.set_tuning(eng.TUNE_LOCALITY_SORT, 0)
        want = eng.batch_resident(b.pair_a, b.pair_b, mode="full", want_contacts=True)
        for chunks in (1, 3):
            eng.set_tuning(eng.TUNE_HOST_CHUNKS, chunks)
            eng.set_tuning(eng.TUNE_LOCALITY_SORT, 1)
            same(eng.batch_resident(b.pair_a, b.pair_b, mode="full", want_contacts=True), want)
            same(eng.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="full"), want)
        pb = W.make_posed(n_shapes=40, n_bodies=20_000, n_pairs=120_000, sigma=0.85, seed=62)
        eng.upload_shapes(pb.shape_verts, pb.shape_off)
        eng.set_tuning(eng.TUNE_LOCALITY_SORT, 0)
        wp = eng.batch_posed(pb.body_shape, pb.pose, pb.pair_a, pb.pair_b, mode="epa")
        eng.set_tuning(eng.TUNE_LOCALITY_SORT, 1)
        same(eng.batch_posed(pb.body_shape, pb.pose, pb.pair_a, pb.pair_b, mode="epa"), wp)
        eng.set_tuning(eng.TUNE_LOCALITY_SORT, 2)
        eng.set_tuning(eng.TUNE_HOST_CHUNKS, 0)
