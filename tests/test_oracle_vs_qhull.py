"""Independent ground truth: scipy/Qhull hull of all pairwise differences."""
import numpy as np
import pytest


@pytest.mark.parametrize("kind,kw,n,nchk", [
    ("box8", dict(kind="box8", sigma=1.0), 4000, 400),
    ("hull32", dict(kind="hull", n_vert=32, sigma=0.85), 3000, 300),
    ("hull5", dict(kind="hull", n_vert=5, sigma=0.15), 4000, 400),
    ("deep64", dict(kind="hull", n_vert=64, sigma=0.05), 300, 60),
])
def test_oracle_matches_minkowski_hull(orc, W, kind, kw, n, nchk):
    b = W.make_pairs(n_pairs=n, seed=2024, **kw)
    o = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
    assert np.all((o["status"] >> 24) == 0)
    for i in range(nchk):
        inside, depth, nrm, second = orc.ground_truth(b.hull(b.pair_a[i]), b.hull(b.pair_b[i]))
        assert inside == bool(o["collision"][i])
        if inside:
            assert abs(depth - o["depth"][i]) <= 1e-9 * max(depth, 1.0)
            if second - depth > 1e-6:                      # tied depths make the normal ambiguous
                assert 1.0 - abs(float(np.dot(nrm, o["normal"][i]))) <= 1e-9
                assert float(np.dot(nrm, o["normal"][i])) > 0        # outward: from p1 towards p2 (Q9)


def test_mixed_sizes(orc, W):
    b = W.make_mixed(2000, 0.25, seed=9)
    o = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
    for i in range(150):
        inside, depth, nrm, second = orc.ground_truth(b.hull(b.pair_a[i]), b.hull(b.pair_b[i]))
        assert inside == bool(o["collision"][i])
        if inside:
            assert abs(depth - o["depth"][i]) <= 1e-9 * max(depth, 1.0)


def test_translate_by_depth_times_normal_separates(orc, W):
    # property: moving p2 by (depth + margin) * normal leaves the pair apart
    b = W.make_pairs("hull", 300, 0.5, seed=77, n_vert=16)
    o = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
    for i in np.nonzero(o["collision"])[0][:100]:
        p1, p2 = b.hull(b.pair_a[i]), b.hull(b.pair_b[i])
        moved = p2 + (o["depth"][i] + 1e-6) * o["normal"][i]
        assert not orc.pair(p1, moved, mode=orc.MODE_GJK)["collision"]
        closer = p2 + (o["depth"][i] - 1e-6) * o["normal"][i]
        assert orc.pair(p1, closer, mode=orc.MODE_GJK)["collision"]
