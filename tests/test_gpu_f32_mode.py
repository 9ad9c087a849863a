"""fp32 mode (north_star: "1e-5 if an fp32 mode is offered"): gjkepa_batch_f32 takes a float32 hull
store and returns float32 reals; the device widens the store and runs the binary64 kernels (include/gjkepa.h says
why).  Two bars, both written here:
  1. against gjkepa_batch_f64 on the float-ROUNDED vertices: flags, status, colli_type, contact list IDENTICAL and
     every real equal to that call's double rounded once to float (bit-exact);
  2. against the oracle on the ORIGINAL binary64 hulls: depth |d32 - d64| <= 1e-5 * max(d64, 1) and normal
     1 - cos <= 1e-5 on every contact both agree on; flags may differ only for pairs whose gap/overlap is below
     the rounding of the vertices -- the mismatch rate is measured and bounded at 1e-4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5


@pytest.mark.parametrize("kind,n,sigma,mode", [("box8", 8, 1.0, "full"), ("hull", 32, 0.85, "epa"), ("hull", 15, 0.6, "full")])
def test_f32_boundary_equals_f64_on_rounded_hulls(W, orc, engine, kind, n, sigma, mode):
    b = W.make_pairs(kind, 30000, sigma, seed=100 + n, n_vert=n)
    v32 = b.verts.astype(np.float32)
    got = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
    same = engine.batch_host(v32.astype(np.float64), b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
    assert got["n_contacts"] == same["n_contacts"]
    nc = got["n_contacts"]
    assert np.array_equal(got["contact_index"][:nc], same["contact_index"][:nc])
    for k in ("collision", "status") + (("colli_type",) if mode == "full" else ()):
        assert np.array_equal(got[k], same[k]), k
    for k in ("depth", "normal") + (("nearest", "point") if mode == "full" else ()):
        assert got[k].dtype == np.float32 and np.array_equal(got[k], same[k].astype(np.float32)), k
    # bar 2: the oracle on the unrounded hulls
    ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
    mism = float(np.mean(got["collision"] != ref["collision"]))
    assert mism <= 1e-4, mism
    both = (got["collision"] > 0) & (ref["collision"] > 0) & ((ref["status"] >> 24) == 0) & ((got["status"] >> 24) == 0)
    assert both.mean() > 0.2
    d64, d32 = ref["depth"][both], got["depth"][both].astype(np.float64)
    assert np.all(np.abs(d32 - d64) <= F32_TOL * np.maximum(d64, 1.0))
    deep = both & (ref["depth"] > 1e-3)             # at depth ~ 0 neighbouring faces tie and the normal is not unique
    cosd = np.sum(got["normal"][deep].astype(np.float64) * ref["normal"][deep], axis=1)
    assert np.mean(1.0 - cosd <= F32_TOL) > 0.999   # tied depths between two faces may pick the other face


def test_f32_small_batch_empty_and_gjk_only(W, engine):
    c = W.unit_cubes((0.5, 0.25, 0.125))
    got = engine.batch_host_f32(c.verts, c.hull_off, c.pair_a, c.pair_b, mode="full")
    assert got["collision"][0] == 1 and got["depth"][0] == np.float32(0.5)
    assert np.array_equal(got["normal"][0], np.array([1, 0, 0], np.float32))
    z = np.zeros(0, np.int32)
    assert engine.batch_host_f32(c.verts, c.hull_off, z, z, mode="gjk", want_contacts=True)["n_contacts"] == 0
    b = W.make_pairs("box8", 5000, 1.0, seed=3)
    g = engine.batch_host_f32(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    assert set(g) == {"collision", "status"} and 0.3 < g["collision"].mean() < 0.8


@pytest.mark.parametrize("maker,kw,mode", [
    ("make_pairs", dict(kind="box8", n_pairs=200_000, sigma=1.0, seed=71), "gjk"),
    ("make_pairs", dict(kind="hull", n_pairs=60_000, sigma=0.85, seed=72, n_vert=32), "epa"),
    ("make_mixed", dict(n_pairs=60_000, sigma=0.6, seed=73), "full"),
    ("make_shared_store", dict(n_bodies=20_000, n_pairs=200_000, n_vert=16, seed=74), "epa"),
    ("make_touching_mix", dict(n_pairs=40_000, seed=75), "epa"),
])
def test_f32_arithmetic_filter_changes_no_flag(W, engine, multi_f32, maker, kw, mode):
    """GJKEPA_TUNE_F32_FILTER: binary32 arithmetic proves pairs apart before the binary64 kernels run.  It may only answer
    'separated': every flag, and every result of a pair it lets through, must be what the call without it returns (mismatch
    rate 0, not 'small'); the filtered pairs are exactly the ones whose exit path reads EARLY_OUT; touching and nearly touching
    pairs (gaps of 0 and 1e-7..1e-9, far below its margin) are never filtered."""
    b = getattr(W, maker)(**kw)
    v32 = b.verts.astype(np.float32)
    for eng in (engine, multi_f32):
        eng.set_max_hull_vertices(0)
        eng.set_tuning(eng.TUNE_F32_FILTER, 0)
        want = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
        for chunks in (1, 3):
            eng.set_tuning(eng.TUNE_HOST_CHUNKS, chunks)
            eng.set_tuning(eng.TUNE_F32_FILTER, 1)
            got = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode=mode, want_contacts=True)
            total, filtered = eng.f32_filter_stats()
            assert total == b.n_pairs
            assert np.array_equal(got["collision"], want["collision"]), "the filter changed a flag"
            assert got["n_contacts"] == want["n_contacts"]
            nc = got["n_contacts"]
            assert np.array_equal(got["contact_index"][:nc], want["contact_index"][:nc])
            out_path = (got["status"] & 0xff) == 13
            assert int(out_path.sum()) == filtered and not np.any(got["collision"][out_path])
            assert np.array_equal(got["status"][~out_path], want["status"][~out_path])
            for k in ("depth", "normal", "colli_type", "nearest", "point"):
                if k in want:
                    assert np.array_equal(got[k], want[k]), k
            if maker == "make_touching_mix":
                near = (want["collision"] == 0) & (((want["status"] & 0xff) == 9) | ((want["status"] & 0xff) == 8))
                assert near.any()
            else:
                assert filtered > 0.05 * b.n_pairs, (filtered, b.n_pairs)
        eng.set_tuning(eng.TUNE_F32_FILTER, 0)
        eng.set_tuning(eng.TUNE_HOST_CHUNKS, 0)


@pytest.fixture(scope="module")
def multi_f32():
    import torch
    from collision_detect_gjk_epa_b200 import Engine
    n = torch.cuda.device_count()
    eng = Engine(devices=list(range(n)) if n > 1 else [0, 0])
    yield eng
    eng.close()


def test_f32_filter_keeps_non_finite_pairs_for_the_binary64_kernels(W, engine):
    b = W.make_nonfinite("nan", 20000)
    v32 = b.verts.astype(np.float32)
    engine.set_tuning(engine.TUNE_F32_FILTER, 0)
    want = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    engine.set_tuning(engine.TUNE_F32_FILTER, 1)
    got = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    engine.set_tuning(engine.TUNE_F32_FILTER, 0)
    assert np.array_equal(got["collision"], want["collision"])


def _origin_depth(b, i):
    """Qhull on the Minkowski difference of pair i (binary64, float-rounded vertices): > 0 = overlap depth, < 0 = minus a lower
    bound of the gap."""
    from scipy.spatial import ConvexHull
    p1 = b.hull(int(b.pair_a[i])).astype(np.float32).astype(np.float64)
    p2 = b.hull(int(b.pair_b[i])).astype(np.float32).astype(np.float64)
    md = (p1[:, None, :] - p2[None, :, :]).reshape(-1, 3)
    return float((-ConvexHull(md).equations[:, 3]).min())


@pytest.mark.parametrize("maker,kw", [
    ("make_pairs", dict(kind="box8", n_pairs=300_000, sigma=1.0, seed=81)),
    ("make_pairs", dict(kind="hull", n_pairs=60_000, sigma=0.85, seed=82, n_vert=32)),
    ("make_mixed", dict(n_pairs=60_000, sigma=0.6, seed=83)),
    ("make_shared_store", dict(n_bodies=20_000, n_pairs=300_000, n_vert=16, seed=84)),
    ("make_touching_mix", dict(n_pairs=40_000, seed=85)),      # planted gaps of 0, 5e-9 and 1e-6: inside binary32's band by construction
])
def test_binary32_arithmetic_walk(W, engine, multi_f32, maker, kw):
    """GJKEPA_TUNE_F32_ARITH (SURVEY.md 8(f)2 as written): the GJK walk in binary32 arithmetic, same source text as the binary64
    kernels.  Flags are NOT guaranteed equal to the binary64 answer; what is asserted:
      * GJK mode: the mismatch rate against the binary64 kernels (same float-rounded vertices) is below 1e-4, and EVERY
        differing pair has a gap or overlap below 1e-4 (Qhull on its Minkowski difference) -- only nearly touching pairs flip;
      * EPA mode: the binary64 kernels confirm what the binary32 walk calls colliding, so no pair is flagged that the binary64
        call does not flag, and every contact both report has the binary64 call's depth and normal, bit for bit;
      * pairs decided by the binary32 walk say so in their status word; chunks and devices do not change a bit."""
    b = getattr(W, maker)(**kw)
    v32 = b.verts.astype(np.float32)
    engine.set_max_hull_vertices(0)
    engine.set_tuning(engine.TUNE_F32_ARITH, 0)
    want = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="epa", want_contacts=True)
    first = None
    for eng in (engine, multi_f32):
        for chunks in (1, 3):
            eng.set_max_hull_vertices(0)
            eng.set_tuning(eng.TUNE_HOST_CHUNKS, chunks)
            eng.set_tuning(eng.TUNE_F32_ARITH, 1)
            g = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk", want_contacts=True)
            total, decided = eng.f32_filter_stats()
            e = eng.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="epa", want_contacts=True)
            eng.set_tuning(eng.TUNE_F32_ARITH, 0)
            eng.set_tuning(eng.TUNE_HOST_CHUNKS, 0)
            assert total == b.n_pairs and decided == int((g["collision"] == 0).sum())
            assert np.all((g["status"] >> 16) & 0xff == 32)                       # every pair was decided in binary32
            assert g["n_contacts"] == int(g["collision"].sum())
            assert np.array_equal(g["contact_index"][:g["n_contacts"]], np.nonzero(g["collision"])[0].astype(np.int32))
            if first is None:
                first = (g, e)
                diff = np.nonzero(g["collision"] != want["collision"])[0]
                rate = len(diff) / b.n_pairs
                print(f"binary32 walk vs binary64 kernels on {maker}: {len(diff)} of {b.n_pairs} flags differ ({rate:.2e})")
                assert rate <= (1e-4 if maker != "make_touching_mix" else 1e-2), rate
                for i in diff:
                    d0 = _origin_depth(b, i)
                    assert abs(d0) < 1e-4, (int(i), d0, int(g["collision"][i]), int(want["collision"][i]))
                assert 0.05 < g["collision"].mean() < 0.98
            else:
                for k in ("collision", "status"):
                    assert np.array_equal(g[k], first[0][k]), k
                for k in ("collision", "status", "depth", "normal"):
                    assert np.array_equal(e[k], first[1][k]), k
            # EPA mode: binary64 has the last word on every contact
            assert not np.any((e["collision"] > 0) & (want["collision"] == 0))
            both = (e["collision"] > 0) & (want["collision"] > 0)
            assert both.sum() >= 0.999 * (want["collision"] > 0).sum()
            for k in ("depth", "normal", "status"):
                assert np.array_equal(e[k][both], want[k][both]), k
            sep32 = (e["collision"] == 0) & (((e["status"] >> 16) & 0xff) == 32)
            assert sep32.sum() >= 0.99 * (e["collision"] == 0).sum()              # the rest: survivors binary64 called separated
            assert np.all(e["depth"][e["collision"] == 0] == 0.0)


def test_binary32_walk_on_degenerate_and_non_finite_hulls(W, engine):
    """The binary32 walk on what a hull library would reject: 1-5 vertex 'hulls' (the initial-tetrahedron exits F:86, 123, 157),
    duplicated vertices, planar polygons, and hulls holding NaN / Inf.  It must terminate on one of the reference's exit paths
    for every pair, take the binary64 kernels' exits on non-finite hulls (NaN: always FALSE), and on the degenerate families disagree with
    them only marginally (their outcomes sit ON the thresholds, which the two precisions place differently)."""
    engine.set_max_hull_vertices(0)
    for b, bound in ((W.make_mixed(40000, 0.5, seed=3, n_lo=1, n_hi=5), 1e-3), (W.make_nonfinite("nan", 20000), 0.0), (W.make_nonfinite("inf", 20000), 0.0)):
        v32 = b.verts.astype(np.float32)
        engine.set_tuning(engine.TUNE_F32_ARITH, 0)
        want = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
        engine.set_tuning(engine.TUNE_F32_ARITH, 1)
        got = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
        engine.set_tuning(engine.TUNE_F32_ARITH, 0)
        path = got["status"] & 0xff
        assert path.min() >= 1 and path.max() <= 12
        assert np.array_equal(got["collision"] > 0, np.isin(path, (4, 6, 10)))
        bad = ~np.isfinite(v32)
        if bad.any():                                  # NaN: the sphere test fails (F:1185) in both; Inf: the same exits in both
            hull_bad = np.add.reduceat(bad.astype(np.int64), 3 * b.hull_off[:-1]) > 0
            pair_bad = hull_bad[b.pair_a] | hull_bad[b.pair_b]
            assert pair_bad.sum() > 1000
            assert np.array_equal(got["collision"], want["collision"])
            assert np.array_equal(path[pair_bad], (want["status"] & 0xff)[pair_bad])
            if np.isnan(v32).any():
                assert not got["collision"][pair_bad].any()
        else:
            rate = float(np.mean(got["collision"] != want["collision"]))
            print(f"binary32 walk on 1-5 vertex hulls: {rate:.2e} of the flags differ from the binary64 kernels'")
            assert rate <= bound, rate


def test_binary32_walk_ignores_a_bound_the_data_violates(W, engine):
    """gjkepa_set_max_hull_vertices only sizes the binary32 kernel's columns: hulls above it take the 4-lane kernel, and every
    pair gets the answer it gets without the bound."""
    b = W.make_mixed(30000, 0.5, seed=31, n_lo=4, n_hi=40)
    v32 = b.verts.astype(np.float32)
    engine.set_tuning(engine.TUNE_F32_ARITH, 1)
    engine.set_max_hull_vertices(0)
    want = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
    for hint in (8, 12, 16, 40):
        engine.set_max_hull_vertices(hint)
        got = engine.batch_host_f32(v32, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
        assert np.array_equal(got["collision"], want["collision"]) and np.array_equal(got["status"], want["status"]), hint
    engine.set_max_hull_vertices(0)
    engine.set_tuning(engine.TUNE_F32_ARITH, 0)
