"""The straight-line sqrt / division sequences (csrc/fp64_exact.cuh) must be bit-identical to IEEE
round-to-nearest results -- on the GPU against the compiler's own operators and against numpy on
the host -- wherever their range flag is up, and must lower the flag for every special operand."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def probe(engine, a, b):
    n = a.shape[0]
    outs = [np.empty(n) for _ in range(4)]
    flags = np.empty(n, np.uint8)
    p = lambda x: C.c_void_p(x.ctypes.data)
    rc = engine._L.gjkepa_debug_fp64(engine._h, n, p(a), p(b), *[p(o) for o in outs], p(flags))
    assert rc == 0
    return outs, flags


def test_fast_paths_bit_exact_on_random_operands(engine):
    rng = np.random.default_rng(123)
    n = 4_000_000
    # magnitudes spanning the path's working range and far beyond it
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    b = np.abs(rng.standard_normal(n)) * 10.0 ** rng.uniform(-30, 30, n) + 1e-300
    (qf, qr, sf, sr), fl = probe(engine, a, b)
    okd, oks = (fl & 1) > 0, (fl & 2) > 0
    assert okd.mean() > 0.999 and oks.mean() > 0.999
    assert np.array_equal(qf[okd].view(np.uint64), qr[okd].view(np.uint64))
    assert np.array_equal(sf[oks].view(np.uint64), sr[oks].view(np.uint64))
    with np.errstate(all="ignore"):
        assert np.array_equal(qr.view(np.uint64), (a / b).view(np.uint64))       # GPU '/' == host IEEE
        assert np.array_equal(sr.view(np.uint64), np.sqrt(b).view(np.uint64))


def test_unit_scale_operands_like_the_gjk_walk(engine):
    rng = np.random.default_rng(7)
    n = 4_000_000
    v = rng.standard_normal((n, 3)) * rng.choice([1.0, 1e-3, 1e-6, 1e-9], size=(n, 1))
    md = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    (qf, qr, sf, sr), fl = probe(engine, v[:, 0].copy(), md)
    assert np.all(fl == 3)
    assert np.array_equal(qf.view(np.uint64), (v[:, 0] / md).view(np.uint64))
    assert np.array_equal(sf.view(np.uint64), np.sqrt(md).view(np.uint64))


def test_special_operands_take_the_fallback(engine):
    tiny = np.nextafter(0.0, 1.0)
    a = np.array([1.0, 1.0, np.inf, np.nan, 1.0, 1e-300, tiny, 1.0, 1e300, 0.0, 0.0, 0.0, 0.0])
    b = np.array([0.0, np.inf, 1.0, 1.0, np.nan, 1.0, 1.0, tiny, 1e-300, 0.0, np.inf, np.nan, tiny])
    (qf, qr, sf, sr), fl = probe(engine, a, b)
    assert np.all((fl & 1) == 0)                      # every one of these leaves the division fast path
    with np.errstate(all="ignore"):
        want = a / b
    assert np.array_equal(np.isnan(qr), np.isnan(want)) and np.array_equal(qr[~np.isnan(want)], want[~np.isnan(want)])
    (_, _, sf, sr), fl = probe(engine, np.ones(5), np.array([0.0, tiny, 1e-300, np.inf, np.nan]))
    assert np.all((fl & 2) == 0)


def test_zero_numerator_is_answered_in_line_with_the_ieee_sign(engine):
    # exactly zero cross-product components are routine (degenerate tetrahedron faces, axis-aligned boxes)
    a = np.array([0.0, -0.0, 0.0, -0.0])
    b = np.array([1.5, 2.5, -3.0, -0.25])
    (qf, qr, sf, sr), fl = probe(engine, a, b)
    assert np.all((fl & 1) == 1)
    assert np.array_equal(qf.view(np.uint64), (a / b).view(np.uint64))
    assert np.array_equal(qf.view(np.uint64), qr.view(np.uint64))
