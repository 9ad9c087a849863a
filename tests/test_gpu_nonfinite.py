"""Garbage in: NaN / Inf / overflowing / underflowing vertices through the CUDA path.  The reference's flag path is plain
IEEE arithmetic plus comparisons, so it is DEFINED on such input (where it would print and STOP, F:1369, the oracle returns
FALSE + a STOP status), and the kernels must give the oracle's flags and exit paths in every variant; GJK+EPA on the same
input must come back (iteration caps, no hang) with the same flags."""
import numpy as np
import pytest

from test_gpu_parity import configure

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["nan", "inf", "infall", "huge2", "huge", "tiny", "tiny2", "far", "zero"])
def test_flag_path_on_nonfinite_inputs(W, orc, engine, kind):
    for n_vert, hint in ((8, 8), (20, 0), (40, 0)):
        b = W.make_nonfinite(kind, 20000, n_vert, seed=n_vert)
        with np.errstate(all="ignore"):
            ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_GJK)
        for variant in (0, "mono", 1, 2, 4):
            configure(engine, variant)
            engine.set_max_hull_vertices(hint)
            got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
            bad = np.nonzero((got["status"] & 0xffff) != (ref["status"] & 0xffff))[0]
            what = (kind, n_vert, variant, len(bad), sorted(set(zip((got["status"][bad] & 0xff).tolist(), (ref["status"][bad] & 0xff).tolist(),
                                                                        ((got["status"][bad] >> 8) & 0xff).tolist(), ((ref["status"][bad] >> 8) & 0xff).tolist())))[:12])
            assert np.array_equal(got["collision"], ref["collision"]), what
            assert len(bad) == 0, what
    engine.set_max_hull_vertices(0)


@pytest.mark.parametrize("kind", ["nan", "inf", "huge2", "tiny"])
def test_epa_returns_on_nonfinite_inputs(W, orc, engine, kind):
    configure(engine, 0)
    for n_vert in (8, 20, 40):
        b = W.make_nonfinite(kind, 10000, n_vert, seed=100 + n_vert)
        with np.errstate(all="ignore"):
            ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_GJK)
        got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="full")
        assert np.array_equal(got["collision"], ref["collision"])
        clean = np.isfinite(b.verts.reshape(-1, 3 * n_vert)).all(axis=1) & (np.abs(b.verts.reshape(-1, 3 * n_vert)).max(axis=1) < 1e100)
        ok = clean[b.pair_a] & clean[b.pair_b] & (got["collision"] > 0) & ((got["status"] >> 24) == 0)
        assert np.isfinite(got["depth"][ok]).all() and np.all(got["depth"][ok] >= 0.0)      # untouched pairs keep real results
