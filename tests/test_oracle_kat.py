"""Known-answer tests that pin the oracle: the reference has no tests of its
own, so the pins are analytic cube cases and the statement-level quirks of the source."""
import numpy as np
import pytest


def cubes(W, off):
    b = W.unit_cubes(off)
    return b.hull(0), b.hull(1)


@pytest.mark.parametrize("off", [(0.5, 0.0, 0.0), (0.5, 0.25, 0.125)])
def test_overlapping_unit_cubes_depth_and_normal(orc, W, off):
    # BASELINE.json configs[0]: two overlapping unit cubes, depth and normal checked
    a, b = cubes(W, off)
    r = orc.pair(a, b)
    assert r["collision"] and r["exit"] == "loop_contain"
    assert abs(r["depth"] - 0.5) < 1e-12
    assert np.allclose(r["normal"], [1.0, 0.0, 0.0], atol=1e-12)     # points from p1 towards p2 (Q9)
    assert r["colli_type"] == 2                                      # face-face with TOL_FF = 1


def test_separated_cubes_leave_through_coplanar_exit(orc, W):
    a, b = cubes(W, (1.5, 0.0, 0.0))
    r = orc.pair(a, b)
    assert not r["collision"] and r["exit"] == "loop_coplanar"       # F:203
    assert r["depth"] == 0.0 and np.all(r["normal"] == 0.0) and np.all(r["nearest"] == 0.0)


def test_exactly_touching_cubes_collide_with_zero_depth(orc, W):
    a, b = cubes(W, (1.0, 0.0, 0.0))                                  # quirk Q6
    r = orc.pair(a, b)
    assert r["collision"] and abs(r["depth"]) < 1e-12


def test_far_cubes_rejected_by_sphere_with_unit_slack(orc, W):
    a, b = cubes(W, (2.0 * np.sqrt(0.75) + 1.0 + 1e-9, 0.0, 0.0))     # r1 + r2 + 1.0 (Q7), just outside
    assert orc.pair(a, b)["exit"] == "sphere"
    a, b = cubes(W, (2.0 * np.sqrt(0.75) + 1.0 - 1e-9, 0.0, 0.0))
    assert orc.pair(a, b)["exit"] != "sphere"


def test_coincident_cubes_take_h1_path_with_defined_status(orc, W):
    a, b = cubes(W, (0.0, 0.0, 0.0))                                  # hazard H1: F:140 with stale row 4
    r = orc.pair(a, b)
    assert r["collision"] and r["exit"] == "tri_origin"
    assert r["status"] & orc.ST_H1 and r["status"] & orc.ST_STOP      # reference would STOP; oracle flags it
    assert r["depth"] == 0.0


@pytest.mark.parametrize("off", [(0.5, 0.0, 0.0), (0.5, 0.25, 0.125)])
def test_cube_answer_invariant_under_vertex_permutation(orc, W, off):
    a, b = cubes(W, off)
    rng = np.random.default_rng(1)
    for _ in range(200):
        r = orc.pair(a[rng.permutation(8)], b[rng.permutation(8)])
        assert r["collision"] and abs(r["depth"] - 0.5) < 1e-12
        assert np.allclose(r["normal"], [1, 0, 0], atol=1e-12)


def test_support_first_index_wins_ties(orc):
    import ctypes as C
    p1 = np.array([[1.0, 0, 0], [1.0, 0, 0], [0.0, 1, 0]])           # vertices 0 and 1 tie along +x
    p2 = np.array([[-1.0, 0, 0], [0.0, 0, 0], [-1.0, 0, 0]])         # vertices 0 and 2 tie along -x
    b1 = np.ascontiguousarray(p1.T).ravel(); b2 = np.ascontiguousarray(p2.T).ravel()
    d = np.array([1.0, 0.0, 0.0]); res = np.zeros(3)
    i1, i2 = C.c_int32(-1), C.c_int32(-1)
    dp = C.POINTER(C.c_double)
    orc.lib().orc_support(b1.ctypes.data_as(dp), 3, b2.ctypes.data_as(dp), 3, d.ctypes.data_as(dp),
                          res.ctypes.data_as(dp), C.byref(i1), C.byref(i2))
    assert (i1.value, i2.value) == (0, 0) and np.allclose(res, [2.0, 0, 0])


def test_is_inside_pf_quirks(orc):
    import ctypes as C
    dp = C.POINTER(C.c_double)

    def inside(tri, p):
        t = np.ascontiguousarray(tri, dtype=np.float64); q = np.asarray(p, dtype=np.float64)
        return bool(orc.lib().orc_is_inside_pf(t.ctypes.data_as(dp), 3, q.ctypes.data_as(dp)))

    ccw = [[0, 0, 0], [1, 0, 0], [0, 1, 0]]
    assert inside(ccw, [0.25, 0.25, 0]) and not inside(ccw, [2, 2, 0])
    # Q2: a clockwise (all-negative) XY set counts as "all zero" (F:1305-1310) and is re-projected on
    # XZ (F:1312-1321): this tilted triangle holds the point in XY but not in XZ -> FALSE
    cw = [[0, 0, 0], [0, 1, 1], [1, 0, 0]]
    assert not inside(cw, [0.25, 0.25, 1.0]) and inside(cw[::-1], [0.25, 0.25, 1.0])
    # Q3: c1 == 0 makes every product zero -> TRUE
    assert inside(ccw, [2.0, 0.0, 0.0])


def test_is_inside_pf_matches_second_transliteration(orc):
    import ctypes as C, sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    import ref_translit as R
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(11)
    for _ in range(3000):
        tri = rng.standard_normal((3, 3))
        if rng.random() < 0.3:
            tri[:, 2] = 0.0
        if rng.random() < 0.3:
            tri[:, 1] = tri[:, 0] * rng.standard_normal()            # collapses in XY
        p = tri.mean(0) * rng.choice([1.0, 1.0, 3.0]) + (rng.random() < 0.5) * 1e-13 * rng.standard_normal(3)
        t = np.ascontiguousarray(tri)
        got = bool(orc.lib().orc_is_inside_pf(t.ctypes.data_as(dp), 3, p.ctypes.data_as(dp)))
        assert got == R.IS_INSIDE_PF(tri.tolist(), p.tolist())


def test_direction_table_row_one_and_norms(orc):
    import ctypes as C
    d = np.zeros(3)
    orc.lib().orc_dir_table(1, d.ctypes.data_as(C.POINTER(C.c_double)))
    assert d.tolist() == [0.000001109357820885, 0.072093544214837393, 0.997397874913172555]   # F:1584
    for r in range(1, 101):
        orc.lib().orc_dir_table(r, d.ctypes.data_as(C.POINTER(C.c_double)))
        assert abs(np.linalg.norm(d) - 1.0) < 1e-12


def test_direction_tables_identical_and_match_reference_when_present():
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    a = (root / "oracle" / "dirtable.inc").read_text()
    b = (root / "collision_detect_gjk_epa_b200" / "csrc" / "dirtable.inc").read_text()
    assert a == b
    if Path("/root/reference/src/GCLIB_GJKEPA.f90").exists():
        import subprocess, sys
        subprocess.run([sys.executable, str(root / "tools" / "gen_dirtable.py"), "--check"], check=True)


def test_version3_and_bad_version(orc, W):
    a, b = cubes(W, (0.5, 0.25, 0.125))
    r3 = orc.pair(a, b, version=3)
    assert np.allclose(r3["normal"], [1, 0, 0]) and abs(r3["point"][2] - a[:, 2].mean()) < 1e-15
    r2 = orc.pair(a, b, version=2)
    assert not (r2["status"] & orc.ST_BAD_VERSION) and np.allclose(r2["point"], [0.5, 0, 0], atol=1e-12)
    r7 = orc.pair(a, b, version=7)                                    # the reference PAUSEs and STOPs (F:336-339)
    assert r7["status"] & orc.ST_BAD_VERSION and r7["collision"] and abs(r7["depth"] - 0.5) < 1e-12


def test_quickhull_standin_properties(orc):
    rng = np.random.default_rng(5)
    for n in (4, 8, 20, 60, 100):
        p = rng.standard_normal((n, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
        c0 = p.mean(0)
        pts = np.concatenate([p, p[:3], c0 + 0.1 * (p[:4] - c0)])      # duplicates and interior points
        tris = orc.quickhull(pts)
        assert len(tris) == 2 * n - 4                                 # every sphere point is extreme
        c = p.mean(0)
        nrm = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 1])
        assert np.all(np.einsum("ij,ij->i", nrm, tris[:, 0] - c) > 0)  # outward winding
    cube = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], float)
    grid = np.array([[x, y, z] for x in (0, .5, 1) for y in (0, .5, 1) for z in (0, .5, 1)], float)
    assert len(orc.quickhull(cube)) == 12 and len(orc.quickhull(grid)) == 12   # coplanar points swallowed


def test_apply_pose_is_the_stated_arithmetic(orc, W):
    """oracle/orc.py apply_pose (what a host would do before CALL GJKEPA, and what the pose kernel is specified to
    do): x' = ((r00*x + r01*y) + r02*z) + tx, operation by operation, against a scalar evaluation; layout kept."""
    pb = W.make_posed(40, 300, 10, kind="hull", n_lo=4, n_hi=12, seed=3)
    verts, hull_off = orc.apply_pose(pb.shape_verts, pb.shape_off, pb.body_shape, pb.pose)
    assert hull_off[0] == 0 and hull_off[-1] * 3 == verts.shape[0]
    for b in (0, 7, 123, 299):
        s = pb.body_shape[b]
        n = int(pb.shape_off[s + 1] - pb.shape_off[s])
        assert hull_off[b + 1] - hull_off[b] == n
        blk = pb.shape_verts[3 * pb.shape_off[s]:3 * pb.shape_off[s + 1]].reshape(3, n)
        w = verts[3 * hull_off[b]:3 * hull_off[b + 1]].reshape(3, n)
        p = pb.pose[b]
        for i in range(n):
            x, y, z = (float(v) for v in blk[:, i])
            for ax in range(3):
                want = ((float(p[3 * ax]) * x + float(p[3 * ax + 1]) * y) + float(p[3 * ax + 2]) * z) + float(p[9 + ax])
                assert w[ax, i] == want
