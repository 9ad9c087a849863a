"""Candidate pairs on the device: gjkepa_broadphase_f64 / gjkepa_broadphase_posed_f64 must emit
exactly the pairs (a < b) that pass the reference's own rough test, RoughCollisionDetection_SphericalEnvelope
(reference src/GCLIB_GJKEPA.f90:1165-1188) -- the SET is bit-exact against the oracle's all-pairs scan
(oracle.orc.sphere_pairs), so a pair the grid leaves out is one GJKEPA would have returned FALSE for at F:77."""
import numpy as np
import pytest


def _scene(W, n_bodies, n_vert, spacing, seed):
    b = W.make_shared_store(n_bodies, 1, n_vert=n_vert, seed=seed, spacing=spacing)
    return b.verts, b.hull_off


def _as_set(pa, pb):
    return set(zip(pa.tolist(), pb.tolist()))


def test_oracle_all_pairs_scan_is_the_sphere_test(W, orc):
    """The checker itself: orc.sphere_pairs == orc_sphere_test (F:1185) applied to every pair, lexicographic."""
    verts, off = _scene(W, 400, 12, 2.5, 3)
    pa, pb = orc.sphere_pairs(verts, off)
    import ctypes as C
    dp = C.POINTER(C.c_double)
    want = []
    for a in range(400):
        for b in range(a + 1, 400):
            va, vb = verts[3 * off[a]:3 * off[a + 1]], verts[3 * off[b]:3 * off[b + 1]]
            if orc.lib().orc_sphere_test(va.ctypes.data_as(dp), 12, vb.ctypes.data_as(dp), 12):
                want.append((a, b))
    assert list(zip(pa.tolist(), pb.tolist())) == want
    assert 0 < len(want) < 400 * 399 // 2


@pytest.mark.gpu
@pytest.mark.parametrize("n_bodies,n_vert,spacing", [(5000, 16, 2.5), (20000, 8, 4.0), (3000, 32, 1.0), (64, 4, 100.0)])
def test_grid_broadphase_equals_all_pairs_scan(W, orc, engine, n_bodies, n_vert, spacing):
    verts, off = _scene(W, n_bodies, n_vert, spacing, seed=n_bodies)
    ra, rb = orc.sphere_pairs(verts, off)
    ga, gb = engine.broadphase(verts, off)
    assert np.all(ga < gb)
    assert len(ga) == len(ra) and _as_set(ga, gb) == _as_set(ra, rb)
    ga2, gb2 = engine.broadphase(verts, off, cap=max(len(ra) // 3, 1))     # too small a first guess: grows, same list
    assert np.array_equal(ga, ga2) and np.array_equal(gb, gb2)              # and the order is deterministic


@pytest.mark.gpu
def test_broadphase_scene_extremes(W, orc, engine):
    """Everything in one cell; two far-apart clusters (cell numbers in the hundreds of thousands); bodies of very
    different sizes (one large radius sets the cell size); negative coordinates."""
    rng = np.random.default_rng(5)
    verts, off = _scene(W, 600, 8, 0.05, 9)                                 # one clump: ~all pairs pass
    ra, rb = orc.sphere_pairs(verts, off)
    ga, gb = engine.broadphase(verts, off)
    assert len(ra) > 0.9 * 600 * 599 / 2 and _as_set(ga, gb) == _as_set(ra, rb)
    v = verts.reshape(600, 3, 8).copy()
    v[300:] += np.array([4.0e6, -3.0e6, 2.5e6])[None, :, None]             # second cluster far away
    v[::7] *= rng.uniform(2.0, 6.0, size=(len(v[::7]), 1, 1))              # some much larger bodies
    v -= 1.0e3
    verts2 = v.reshape(-1)
    ra, rb = orc.sphere_pairs(verts2, off)
    ga, gb = engine.broadphase(verts2, off)
    assert len(ra) > 0 and _as_set(ga, gb) == _as_set(ra, rb)


@pytest.mark.gpu
def test_broadphase_feeds_the_narrow_phase(W, orc, engine):
    """Candidate list -> gjkepa_batch_f64 gives the oracle's flags on those pairs; every pair the grid left out is
    a sphere reject in the oracle (exit path 1, FALSE)."""
    n = 4000
    verts, off = _scene(W, n, 16, 1.6, 21)
    ga, gb = engine.broadphase(verts, off)
    got = engine.batch_host(verts, off, ga, gb, mode="epa")
    ref = orc.batch(verts, off, ga, gb, mode=orc.MODE_EPA)
    assert np.array_equal(got["collision"], ref["collision"]) and np.all((ref["status"] & 0xff) != 1)
    assert 0.02 < got["collision"].mean() < 0.98
    rng = np.random.default_rng(1)
    a = rng.integers(0, n, 20000).astype(np.int32); b = rng.integers(0, n, 20000).astype(np.int32)
    keep = a < b
    a, b = a[keep], b[keep]
    cand = _as_set(ga, gb)
    out = np.array([(x, y) not in cand for x, y in zip(a.tolist(), b.tolist())])
    miss = orc.batch(verts, off, a[out], b[out], mode=orc.MODE_GJK)
    assert out.sum() > 1000 and np.all((miss["status"] & 0xff) == 1) and not miss["collision"].any()


@pytest.mark.gpu
def test_broadphase_over_posed_bodies(W, orc, engine):
    pb = W.make_posed(n_shapes=200, n_bodies=6000, n_pairs=10, kind="hull", n_lo=4, n_hi=24)
    pb.pose[:, 9:] *= 0.12                                                   # denser scene
    engine.upload_shapes(pb.shape_verts, pb.shape_off)
    ga, gb = engine.broadphase_posed(pb.body_shape, pb.pose)
    verts, off = orc.apply_pose(pb.shape_verts, pb.shape_off, pb.body_shape, pb.pose)
    ra, rb = orc.sphere_pairs(verts, off)
    assert len(ra) > 1000 and _as_set(ga, gb) == _as_set(ra, rb)
    ha, hb = engine.broadphase(verts, off)                                   # same list as the world-space entry
    assert np.array_equal(ga, ha) and np.array_equal(gb, hb)
    with pytest.raises(Exception):
        engine.broadphase_posed(pb.body_shape, pb.pose[:-1])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["gjk", "epa", "full"])
def test_scene_step_equals_broadphase_then_narrow_phase(W, orc, engine, mode):
    """gjkepa_scene_posed_f64 (poses in, contacts out) must return, row for row and bit for bit, what the separate calls
    return: broadphase_posed -> batch_posed on its pairs -> the rows whose collision_ is TRUE; and those flags are the
    oracle's on the posed hulls."""
    pb = W.make_posed(n_shapes=150, n_bodies=5000, n_pairs=10, kind="hull", n_lo=4, n_hi=20, seed=11)
    pb.pose[:, 9:] *= 0.1
    engine.upload_shapes(pb.shape_verts, pb.shape_off)
    engine.set_max_hull_vertices(8)          # a stale bound for ANOTHER store must not apply: the shape store knows its sizes
    sc = engine.scene_posed(pb.body_shape, pb.pose, mode=mode)
    engine.set_max_hull_vertices(0)
    ga, gb = engine.broadphase_posed(pb.body_shape, pb.pose)
    nar = engine.batch_posed(pb.body_shape, pb.pose, ga, gb, mode=mode)
    hit = nar["collision"] > 0
    assert sc["n_candidates"] == len(ga) and sc["n_contacts"] == int(hit.sum()) and 0.02 < hit.mean() < 0.9
    order = np.lexsort((gb[hit], ga[hit]))                                   # the scene step sorts its rows by (body_a, body_b)
    assert np.array_equal(sc["body_a"], ga[hit][order]) and np.array_equal(sc["body_b"], gb[hit][order])
    for k in ("depth", "normal", "point", "nearest", "colli_type"):
        assert (k in sc) == (k in nar)
        if k in sc:
            assert np.array_equal(sc[k], nar[k][hit][order]), k
    verts, off = orc.apply_pose(pb.shape_verts, pb.shape_off, pb.body_shape, pb.pose)
    ref = orc.batch(verts, off, ga, gb, mode=orc.MODE_GJK)
    assert np.array_equal(ref["collision"] > 0, hit)
    # shortcut E1 (opt-in): candidates shrink to the pairs whose spheres really touch, separated pairs leave early -- the
    # contact list must not change by a bit
    engine.set_tuning(engine.TUNE_EARLY_OUT, 1)
    fast = engine.scene_posed(pb.body_shape, pb.pose, mode=mode)
    engine.set_tuning(engine.TUNE_EARLY_OUT, 0)
    assert fast["n_candidates"] < sc["n_candidates"] and fast["n_contacts"] == sc["n_contacts"]
    for k in ("body_a", "body_b", "depth", "normal", "point", "nearest", "colli_type"):
        if k in sc:
            assert np.array_equal(fast[k], sc[k]), k
    small = engine.scene_posed(pb.body_shape, pb.pose, mode=mode, cap=7)        # too small a first guess: grows
    assert small["n_contacts"] == sc["n_contacts"] and np.array_equal(small["body_a"], sc["body_a"])


@pytest.mark.gpu
def test_scene_step_without_contacts(W, engine):
    pb = W.make_posed(n_shapes=10, n_bodies=200, n_pairs=10, kind="box", seed=12)
    pb.pose[:, 9:] = np.arange(200)[:, None] * 100.0                          # everything far apart: no candidates
    engine.upload_shapes(pb.shape_verts, pb.shape_off)
    sc = engine.scene_posed(pb.body_shape, pb.pose, mode="epa")
    assert sc["n_candidates"] == 0 and sc["n_contacts"] == 0 and len(sc["body_a"]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("poison", [np.nan, np.inf, -np.inf])
def test_non_finite_bodies_do_not_poison_the_grid(W, orc, engine, poison):
    """One NaN / Inf vertex used to reach the grid's global bounds (every body in cell 0: an all-pairs scan, or whole rows
    of cells dropped).  Such a body now takes no part; every other pair is found as before, and the contacts of a scene
    step are those of the scene without it (the narrow phase answers FALSE for any pair with a non-finite hull)."""
    verts, off = _scene(W, 4000, 16, 1.6, seed=77)
    bad = (17, 2500)
    v = verts.copy()
    for h in bad:
        v[3 * off[h] + 5] = poison
    ra, rb = orc.sphere_pairs(verts, off)                       # the clean scene
    want = {(a, b) for a, b in zip(ra.tolist(), rb.tolist()) if a not in bad and b not in bad}
    ga, gb = engine.broadphase(v, off)
    assert _as_set(ga, gb) == want
    # scene step: shapes = the hulls themselves, identity poses
    n = len(off) - 1
    pose = np.zeros((n, 12)); pose[:, 0] = pose[:, 4] = pose[:, 8] = 1.0
    shp = np.arange(n, dtype=np.int32)
    engine.upload_shapes(v, off)
    got = engine.scene_posed(shp, pose, mode="epa")
    engine.upload_shapes(verts, off)
    clean = engine.scene_posed(shp, pose, mode="epa")
    keep = ~(np.isin(clean["body_a"], bad) | np.isin(clean["body_b"], bad))
    assert got["n_contacts"] == int(keep.sum())
    assert np.array_equal(got["body_a"], clean["body_a"][keep]) and np.array_equal(got["body_b"], clean["body_b"][keep])
    assert np.array_equal(got["depth"], clean["depth"][keep])
