"""get_collisionPoint_02 (reference src/GCLIB_GJKEPA.f90:457-696, version_ = 2): the oracle's C restatement against
a second, independently written transliteration of the same Fortran in Python (lists instead of GCLIB_List, math.atan2,
the single-precision 2*pi of F:1547), on the normals the oracle's EPA produces."""
import math

import numpy as np
import pytest


def utz(v):
    md = math.sqrt(float(v[0] * v[0] + v[1] * v[1]) + float(v[2] * v[2]))
    return np.zeros(3) if md < 1e-12 else v / md


def dot(a, b):
    return float((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])


def all_supports(p, n, tol):
    d = [dot(n, v) for v in p]
    mx = max(d)
    return [v for v, t in zip(p, d) if t > mx - tol]


def foot_pl(pt, v1, v2):
    u = utz(v2 - v1)
    return v1 + dot(pt - v1, u) * u


def foot_ll(p1, q1, p2, q2):
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, b, c, e, f = dot(d1, d1), dot(d1, d2), dot(d1, r), dot(d2, d2), dot(d2, r)
    d = a * e - b ** 2
    if abs(d) < 1e-12:
        f1 = (p1 + q1) / 2.0
        return f1, foot_pl(f1, p2, q2)
    s, t = (b * f - c * e) / d, (a * f - b * c) / d
    return p1 + s * (q1 - p1), p2 + t * (q2 - p2)


def inside_pf(v, p):
    """IS_INSIDE_PF, F:1271-1337 (projection on XY, then XZ when no cross product is positive)."""
    n = len(v)
    c = [(v[(i + 1) % n][0] - v[i][0]) * (p[1] - v[i][1]) - (v[(i + 1) % n][1] - v[i][1]) * (p[0] - v[i][0]) for i in range(n)]
    c = [0.0 if abs(x) < 1e-12 else x for x in c]
    if not any(x > 1e-15 for x in c):
        c = [(v[(i + 1) % n][0] - v[i][0]) * (p[2] - v[i][2]) - (v[(i + 1) % n][2] - v[i][2]) * (p[0] - v[i][0]) for i in range(n)]
    return not any(c[0] * x < 0.0 for x in c)


def sort_clock(pts):
    n = len(pts)
    if all(np.all(np.abs(pts[i] - pts[j]) <= 1e-12) for i in range(n - 1) for j in range(i + 1, n)):
        return list(pts)
    cen = np.array([sum(float(p[k]) for p in pts) / n for k in range(3)])
    nrm = np.cross(pts[1] - pts[0], pts[2] - pts[0])
    two_pi = float(np.float32(2.0) * np.arccos(np.float32(-1.0)))
    out = [pts[0]]
    for i in range(1, n):
        best, idx = None, 0
        for j in range(n):
            if any(np.all(pts[j] == o) for o in out):
                continue
            v1, v2 = pts[j] - cen, out[-1] - cen
            ang = math.atan2(dot(nrm, np.cross(v2, v1)), dot(v1, v2))
            ang = math.fmod(ang + two_pi, two_pi)
            if ang < 0:
                ang += two_pi
            if best is None or ang < best:
                best, idx = ang, j
        out.append(pts[idx])
    return out


def point_v2(orc, p1, p2, nrm):
    s1, s2 = all_supports(p1, nrm, 1e-1), all_supports(p2, -nrm, 1e-1)
    n1, n2 = len(s1), len(s2)
    if n1 == 1 and n2 == 1:
        return (s1[0] + s2[0]) / 2.0
    if n1 == 1:
        return s1[0]
    if n2 == 1:
        return s2[0]
    if n1 == 2 and n2 == 2:
        f1, f2 = foot_ll(s1[0], s1[1], s2[0], s2[1])
        return (f1 + f2) / 2.0
    if n1 == 2 or n2 == 2:
        poly, seg = (s1, s2) if n1 >= 3 else (s2, s1)
        srt = np.array(sort_clock(poly))
        c = sum(1 for q in seg if inside_pf(srt, q))
        if c == 0:
            cen = np.array([sum(float(p[k]) for p in poly) / len(poly) for k in range(3)])
            return foot_pl(cen, seg[0], seg[1])
        return (seg[0] + seg[1]) * 0.5
    return np.array([sum(float(p[k]) for p in s1) / n1 for k in range(3)])


@pytest.mark.parametrize("maker,kw", [
    ("make_pairs", dict(kind="box8", n_pairs=1500, sigma=0.8, seed=41)),
    ("make_pairs", dict(kind="hull", n_pairs=1200, sigma=0.4, seed=42, n_vert=10)),
    ("make_touching_mix", dict(n_pairs=600, seed=43)),
])
def test_point_v2_against_second_transliteration(orc, W, maker, kw):
    b = getattr(W, maker)(**kw)
    o = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_FULL, version=2)
    hit = np.nonzero((o["collision"] > 0) & ((o["status"] >> 24) == 0))[0]
    assert len(hit) > 100
    cases = set()
    for i in hit:
        p1, p2 = b.hull(b.pair_a[i]), b.hull(b.pair_b[i])
        n = o["normal"][i]
        want = point_v2(orc, p1, p2, n)
        assert np.allclose(o["point"][i], want, rtol=0, atol=1e-12), i
        cases.add((min(len(all_supports(p1, n, 1e-1)), 3), min(len(all_supports(p2, -n, 1e-1)), 3)))
    assert len(cases) >= 4                                            # several branches of F:475-495 are reached


def test_point_v2_cube_face_contact(orc, W):
    b = W.unit_cubes((0.5, 0.25, 0.125))
    r = orc.pair(b.hull(0), b.hull(1), version=2)
    assert r["collision"] and np.allclose(r["point"], [0.5, 0.0, 0.0], atol=1e-12)      # case_05: mean of p1's face
