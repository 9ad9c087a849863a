"""The product's GJK core (csrc/gjk_core.cuh, L = 1 lane instantiation compiled by g++) against the
oracle: the logic the CUDA kernels share -- fused face normals, one-pass +-d supports, reused cross
products -- must reproduce flags, exit paths, walk iterations and support counts bit for bit."""
import ctypes as C

import numpy as np
import pytest


def run_sim(lib, b, early=False):
    n = b.n_pairs
    col = np.zeros(n, np.uint8); st = np.zeros(n, np.uint32); ns = np.zeros(n, np.int32); sx = np.zeros((n, 12))
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    (lib.hostsim_gjk_batch_early if early else lib.hostsim_gjk_batch)(P(b.verts, C.c_double), P(b.hull_off, C.c_int64), C.c_int64(n),
                          P(b.pair_a, C.c_int32), P(b.pair_b, C.c_int32),
                          P(col, C.c_uint8), P(st, C.c_uint32), P(ns, C.c_int32), P(sx, C.c_double))
    return col, st, ns, sx


@pytest.mark.parametrize("maker,kw", [
    ("make_pairs", dict(kind="box8", n_pairs=60000, sigma=1.0)),
    ("make_pairs", dict(kind="hull", n_pairs=30000, sigma=0.85, n_vert=32)),
    ("make_mixed", dict(n_pairs=15000, sigma=0.25)),
    ("make_pairs", dict(kind="hull", n_pairs=40000, sigma=0.15, n_vert=5)),
    ("make_pairs", dict(kind="hull", n_pairs=40000, sigma=0.5, n_vert=4)),
    ("make_pairs", dict(kind="hull", n_pairs=5000, sigma=0.05, n_vert=64)),
])
def test_core_bit_exact_with_oracle(orc, W, host_sim, maker, kw):
    b = getattr(W, maker)(**kw)
    col, st, ns, sx = run_sim(host_sim, b)
    o = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_GJK)
    assert np.array_equal(col, o["collision"])
    assert np.array_equal(st & 0xffff, o["status"] & 0xffff)
    assert np.array_equal(ns, o["n_support"])


@pytest.mark.parametrize("off", [(0.5, 0, 0), (0.5, 0.25, 0.125), (1.5, 0, 0), (1.0, 0, 0), (0, 0, 0),
                                 (0.5, 0.5, 0), (1, 1, 1), (1, 1, 0), (0.999999995, 0, 0)])
def test_core_on_axis_aligned_cubes(orc, W, host_sim, off):
    b = W.unit_cubes(off)
    col, st, ns, sx = run_sim(host_sim, b)
    r = orc.pair(b.hull(0), b.hull(1), mode=orc.MODE_GJK)
    assert bool(col[0]) == r["collision"] and (st[0] & 0xffff) == (r["status"] & 0xffff)
    assert np.array_equal(sx[0].reshape(4, 3), r["simplex"])           # final simplex, bit for bit


def test_core_golden_fixture(golden, host_sim, W):
    g = golden
    b = W.PairBatch(g["verts"], g["hull_off"], g["pair_a"], g["pair_b"])
    col, st, ns, sx = run_sim(host_sim, b)
    assert np.array_equal(col, g["collision"]) and np.array_equal(st & 0xff, g["exit_path"])
    assert np.array_equal((st >> 8) & 0xff, g["gjk_iters"])


def test_stall_test_equals_the_reference_form(host_sim):
    """F:219-234 with explicit simplex_last_1/_2 arrays (as the reference keeps them) against the product's
    form, which rebuilds simplex_last_2 from the previously dropped vertex.  No synthetic workload reaches
    this exit, so the cases are constructed: consistent walk histories with rows forced to coincide."""
    rng = np.random.default_rng(5)
    rows = {0: (0, 2, 3), 1: (0, 1, 3), 2: (0, 1, 2), 3: (1, 2, 3)}
    host_sim.hostsim_stall.restype = C.c_int
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    n_true = 0
    for trial in range(20000):
        have2 = trial % 5 != 0
        pool = rng.normal(size=(3, 3))                                  # few distinct points -> many coincidences
        pick = lambda: pool[rng.integers(0, 3)] + (rng.random(3) < 0.3) * rng.choice([0.0, 5e-9, 2e-8], size=3)
        last2 = np.array([pick() for _ in range(4)]) if have2 else np.zeros((4, 3))
        k1 = int(rng.integers(0, 4))
        last1 = np.zeros((4, 3))
        if have2:
            a1, b1, c1 = rows[k1]
            last1[0], last1[1], last1[2] = last2[a1], last2[b1], last2[c1]
            last1[3] = pick()
        else:
            last1 = np.array([pick() for _ in range(4)])
            if trial % 3 == 0:
                last1[:] = rng.choice([0.0, 3e-9], size=(4, 3))          # rows close to the zero history
        k = int(rng.integers(0, 4))
        a, b, c = rows[k]
        sp = pick() if trial % 4 else last1[3] + rng.choice([0.0, 4e-9, 2e-8], size=3)
        new = np.array([last1[a], last1[b], last1[c], sp])
        close = lambda u, v: bool(np.all(np.abs(u - v) < 1e-8))
        want = all(close(new[i], last1[i]) or close(new[i], last2[i]) for i in range(4))
        got = host_sim.hostsim_stall(P(np.ascontiguousarray(last1)), P(np.ascontiguousarray(last2)), int(have2), k1, k,
                                     P(np.ascontiguousarray(sp)))
        assert bool(got) == want, (trial, k1, k, have2)
        n_true += want
    assert n_true > 200                                                  # the constructed cases do reach the exit


def test_walk_history_bookkeeping_matches_explicit_arrays(W, host_sim):
    """Every walk iteration of real workloads: simplex_last_2 rebuilt from (dropped vertex, kept face) equals the
    explicitly kept array bit for bit, and the two-level stall test equals the reference's four-row form."""
    for b in (W.make_pairs("box8", 50000, 1.0, seed=11), W.make_pairs("hull", 20000, 0.85, seed=12, n_vert=32),
              W.make_mixed(10000, 0.25, seed=13), W.make_pairs("hull", 20000, 0.5, seed=14, n_vert=4)):
        run_sim(host_sim, b)
    checks, mism = C.c_longlong(0), C.c_longlong(0)
    host_sim.hostsim_shadow_counts(C.byref(checks), C.byref(mism))
    assert checks.value > 200000 and mism.value == 0


@pytest.mark.parametrize("maker,kw", [
    ("make_pairs", dict(kind="box8", n_pairs=300000, sigma=1.0, seed=31)),
    ("make_pairs", dict(kind="box8", n_pairs=100000, sigma=2.0, seed=32)),
    ("make_pairs", dict(kind="hull", n_pairs=60000, sigma=0.85, n_vert=32, seed=33)),
    ("make_pairs", dict(kind="hull", n_pairs=60000, sigma=1.2, n_vert=16, seed=34)),
    ("make_mixed", dict(n_pairs=40000, sigma=0.8, seed=35)),
    ("make_mixed", dict(n_pairs=40000, sigma=0.5, seed=36, n_lo=1, n_hi=5)),
    ("make_shared_store", dict(n_bodies=20000, n_pairs=200000, n_vert=16, seed=37)),
    ("make_touching_mix", dict(n_pairs=40000, seed=38)),
])
def test_early_out_e1_never_changes_a_flag(orc, W, host_sim, maker, kw):
    """Shortcut E1 (GJKEPA_TUNE_EARLY_OUT, default off): a support value s.d < -1e-6 proves separation.
    Flags must equal the oracle's (which has no shortcut) on every pair; only pairs the oracle calls non-colliding may
    leave through it, and they leave EARLIER (fewer supports)."""
    b = getattr(W, maker)(**kw)
    col, st, ns, _ = run_sim(host_sim, b, early=True)
    o = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_GJK)
    assert np.array_equal(col, o["collision"])
    e1 = (st & 0xff) == 13
    sph = (o["status"] & 0xff) == 1                                          # E1 also drops the sphere test: these walk on
    assert e1.sum() > 0 and not o["collision"][e1].any() and not col[sph].any()
    keep = ~e1 & ~sph
    assert np.array_equal(st[keep] & 0xffff, o["status"][keep] & 0xffff)     # everything else untouched
    w = e1 & ~sph
    assert np.all(ns[w] <= o["n_support"][w]) and ns[w].sum() < o["n_support"][w].sum()


@pytest.mark.parametrize("kind", ["nan", "inf", "infall", "huge2", "huge", "tiny", "tiny2", "far", "zero"])
@pytest.mark.parametrize("n_vert", [8, 20])
def test_core_follows_the_oracle_on_nonfinite_and_extreme_inputs(orc, W, host_sim, kind, n_vert):
    """NaN / Inf / overflowing / underflowing vertices: every shortcut of the core (shared face normals, one-pass +-d
    supports, sqrt of the max instead of max of sqrt, ...) must still give the oracle's flags, exit paths and support
    counts.  Inf inputs reach the reference's READ/STOP inside the loop (F:203 -> F:1392 -> F:1369), defined as FALSE +
    STOP status by the oracle."""
    b = W.make_nonfinite(kind, 20000, n_vert, seed=n_vert)
    with np.errstate(all="ignore"):
        col, st, ns, _ = run_sim(host_sim, b)
        o = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_GJK)
    assert np.array_equal(col, o["collision"])
    assert np.array_equal(st & 0xffff, o["status"] & 0xffff)
    assert np.array_equal(ns, o["n_support"])



@pytest.mark.parametrize("maker,kw", [
    ("make_pairs", dict(kind="box8", n_pairs=200000, sigma=1.0, seed=91)),
    ("make_pairs", dict(kind="hull", n_pairs=60000, sigma=0.85, n_vert=32, seed=92)),
    ("make_mixed", dict(n_pairs=40000, sigma=0.6, seed=93)),
    ("make_shared_store", dict(n_bodies=20000, n_pairs=200000, n_vert=16, seed=94)),
])
def test_binary32_instantiation_of_the_core(orc, W, host_sim, maker, kw):
    """csrc/gjk_core_f32.cuh (GJKEPA_TUNE_F32_ARITH): the same source text compiled with float operands, against
    the oracle on the same float-rounded vertices.  Flags are not promised equal; the mismatch rate is measured and bounded at
    1e-4, every walk terminates on one of the reference's exit paths, and separated pairs stay within the iteration cap."""
    b = getattr(W, maker)(**kw)
    v32 = b.verts.astype(np.float32)
    n = b.n_pairs
    col = np.zeros(n, np.uint8); st = np.zeros(n, np.uint32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    host_sim.hostsim_gjk32_batch(P(v32, C.c_float), P(b.hull_off, C.c_int64), C.c_int64(n), P(b.pair_a, C.c_int32), P(b.pair_b, C.c_int32),
                                 P(col, C.c_uint8), P(st, C.c_uint32))
    o = orc.batch(v32.astype(np.float64), b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_GJK)
    rate = float(np.mean(col != o["collision"]))
    print(f"binary32 core vs oracle on {maker}: mismatch rate {rate:.2e}; "
          f"mean walk iterations {((st >> 8) & 0xff).mean():.2f} (binary64: {((o['status'] >> 8) & 0xff).mean():.2f})")
    assert rate <= 1e-4, rate
    path = st & 0xff
    assert path.min() >= 1 and path.max() <= 12
    assert np.array_equal(col > 0, np.isin(path, (4, 6, 10)))
    assert ((st >> 8) & 0xff).max() <= 50
