"""The C oracle against the committed golden fixture (tests/golden/gjkepa_golden.npz), which was
produced by a second, independently written transliteration of the reference source
(tests/golden/ref_translit.py + scipy/Qhull for the un-vendored QuickHull)."""
import numpy as np


def test_oracle_flags_and_gjk_path_bit_exact(orc, golden):
    g = golden
    o = orc.batch(g["verts"], g["hull_off"], g["pair_a"], g["pair_b"], mode=orc.MODE_GJK)
    st = o["status"]
    assert np.array_equal(o["collision"], g["collision"])
    assert np.array_equal(st & 0xff, g["exit_path"])
    assert np.array_equal((st >> 8) & 0xff, g["gjk_iters"])
    # every GJK exit path the fixture walks (sphere, coplanar4, init_contain, loop_coplanar, loop_contain)
    assert set(np.unique(g["exit_path"]).tolist()) >= {1, 5, 6, 9, 10}


def test_oracle_support_call_counts_match_on_gjk_only_pairs(orc, golden):
    g = golden
    o = orc.batch(g["verts"], g["hull_off"], g["pair_a"], g["pair_b"], mode=orc.MODE_GJK)
    nc = g["collision"] == 0
    assert np.array_equal(o["n_support"][nc], g["n_support"][nc])


def test_oracle_depth_and_normal_within_tolerance(orc, golden):
    g = golden
    o = orc.batch(g["verts"], g["hull_off"], g["pair_a"], g["pair_b"], mode=orc.MODE_EPA)
    m = g["epa_ok"] > 0
    assert m.sum() > 1000 and np.all((o["status"][m] >> 24) == 0)
    # tolerance stated by north_star: 1e-9 relative in real(8); relative to max(depth, hull scale = 1)
    tol = 1e-9 * np.maximum(g["depth"][m], 1.0)
    assert np.all(np.abs(o["depth"][m] - g["depth"][m]) <= tol)
    cosd = np.abs(np.sum(o["normal"][m] * g["normal"][m], axis=1))
    assert np.all(1.0 - cosd <= 1e-9)
    # the hull stand-ins differ (own QuickHull vs Qhull) but walk the same number of EPA steps almost always
    same = ((o["status"][m] >> 16) & 0xff) == g["epa_iters"][m]
    assert same.mean() > 0.85


def test_cube_rows_of_the_fixture(orc, golden):
    g = golden
    cubes = g["tag"] < 8
    assert g["collision"][cubes].tolist() == [1, 1, 0, 1, 1, 1, 0, 1]
    assert abs(g["depth"][0] - 0.5) < 1e-12 and abs(g["depth"][1] - 0.5) < 1e-12
