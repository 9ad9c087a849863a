"""Line-by-line Python transliteration of the reference module, used ONLY to generate the golden
fixtures in this directory (tests/golden/make_golden.py).  It is a second, independently written
restatement of src/GCLIB_GJKEPA.f90 ("F:<line>"): same statements, same order, plain float64
arithmetic (Python floats: IEEE binary64, no FMA), pure-Python loops -> small cases only.

The reference cannot be executed here (Fortran, no compiler, three un-vendored modules), so these
fixtures pin the C oracle against a second reading of the source, not against reference output:
parity stays "unpinned" in the sense of DESIGN.md.  The missing GCLIB_QuickHull is stood in for by
scipy.spatial.ConvexHull (Qhull, triangulated 'Qt'), GCLIB_DeHull by a unique-vertex scan.
"""
import math

import numpy as np

HUGE = 1.7976931348623157e308


def _load_table():
    import re
    from pathlib import Path
    txt = (Path(__file__).resolve().parents[2] / "oracle" / "dirtable.inc").read_text()
    rows = re.findall(r"\{\s*(-?[\d.]+),\s*(-?[\d.]+),\s*(-?[\d.]+)\s*\}", txt)
    assert len(rows) == 100
    return [[float(a), float(b), float(c)] for a, b, c in rows]


DATALIB = _load_table()          # GET_RANDOM_UNIT_VECTOR, F:1578-1689 (data)


def DOT(a, b):
    s = 0.0
    for i in range(3):
        s = s + a[i] * b[i]
    return s


def NORM2(v):
    return math.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])


def SUB(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def NEG(a):
    return [-a[0], -a[1], -a[2]]


def CROSS_PRODUCT_3D(A, B):                                  # F:1201-1212
    return [A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]]


def UTZVEC(v):                                               # F:1343-1352
    MD = NORM2(v)
    if MD < 1.0e-12:
        return [0.0, 0.0, 0.0]
    return [v[0] / MD, v[1] / MD, v[2] / MD]


def UNINML(P):                                               # F:1382-1394
    N1 = SUB(P[1], P[0])
    N2 = SUB(P[2], P[1])
    cross = CROSS_PRODUCT_3D(N1, N2)
    if any(abs(c) > 1.0e-12 for c in cross):
        m = NORM2(cross)
        return [cross[0] / m, cross[1] / m, cross[2] / m]
    return [0.0, 0.0, 0.0]


class StopReached(Exception):
    """the reference would WRITE, READ(5,*) and STOP here"""


def DIST_PF_SIGN(x, P):                                      # F:1357-1377
    n = UNINML(P)
    if all(abs(c) < 1.0e-12 for c in n):
        raise StopReached("DIST_PF_SIGN")
    return DOT(SUB(x, P[0]), n)


def VEC_PL(C, line):                                         # F:1423-1440
    A, B = line
    AB = SUB(B, A)
    AC = SUB(C, A)
    t = DOT(AC, AB) / NORM2(AB)
    u = UTZVEC(AB)
    D = [A[0] + t * u[0], A[1] + t * u[1], A[2] + t * u[2]]
    return UTZVEC(SUB(D, C))


def IS_INSIDE_PF(V, p):                                      # F:1271-1337
    NNODE = len(V)
    c = [0.0] * NNODE
    for i in range(NNODE):
        j = 0 if i == NNODE - 1 else i + 1
        c[i] = (V[j][0] - V[i][0]) * (p[1] - V[i][1]) - (V[j][1] - V[i][1]) * (p[0] - V[i][0])
    for i in range(NNODE):
        if abs(c[i]) < 1.0e-12:
            c[i] = 0.0
    zeroMask = [ci > 1.0e-15 for ci in c]
    if not any(zeroMask):
        for i in range(NNODE):
            j = 0 if i == NNODE - 1 else i + 1
            c[i] = (V[j][0] - V[i][0]) * (p[2] - V[i][2]) - (V[j][2] - V[i][2]) * (p[0] - V[i][0])
    for i in range(NNODE):
        if c[0] * c[i] < 0.0:
            return False
    return True


def support_mapping(p1, p2, d, counter=None):                # F:1030-1062
    maxDot1 = -HUGE
    maxIndex1 = 0
    for i in range(len(p1)):
        t = DOT(d, p1[i])
        if t > maxDot1:
            maxDot1 = t
            maxIndex1 = i
    nd = NEG(d)
    maxDot2 = -HUGE
    maxIndex2 = 0
    for i in range(len(p2)):
        t = DOT(nd, p2[i])
        if t > maxDot2:
            maxDot2 = t
            maxIndex2 = i
    if counter is not None:
        counter[0] += 1
    return SUB(p1[maxIndex1], p2[maxIndex2])


def RoughCollisionDetection_SphericalEnvelope(p1, p2):       # F:1165-1188
    def sphere(p):
        n = len(p)
        mp = []
        for k in range(3):
            s = 0.0
            for i in range(n):
                s = s + p[i][k]
            mp.append(s / n)
        r = max(NORM2(SUB(p[i], mp)) for i in range(n))
        return mp, r
    mp1, r1 = sphere(p1)
    mp2, r2 = sphere(p2)
    return NORM2(SUB(mp1, mp2)) <= r1 + r2 + 1.0


IDFC = [[0, 2, 3], [0, 1, 3], [0, 1, 2], [1, 2, 3]]        # idFc F:1227-1229 (column-major fill)


def _mean4(s):
    M = []
    for k in range(3):
        t = 0.0
        for i in range(4):
            t = t + s[i][k]
        M.append(t / 4.0)
    return M


def isPointInSimplex(p, s):                                  # F:1217-1265
    M = _mean4(s)
    nml = []
    for i in range(4):
        AB = SUB(s[IDFC[i][0]], s[IDFC[i][1]])
        BC = SUB(s[IDFC[i][1]], s[IDFC[i][2]])
        n = UTZVEC(CROSS_PRODUCT_3D(AB, BC))
        if DOT(n, SUB(s[i], M)) < 0.0:
            n = NEG(n)
        nml.append(n)
    dist = [DOT(SUB(s[i], p), nml[i]) for i in range(4)]
    for i in range(4):
        if abs(dist[i]) < 1.0e-8:
            if IS_INSIDE_PF([s[IDFC[i][j]] for j in range(3)], p):
                return True
    return all(d > 0.0 for d in dist)


def update_simplex_GJK(p1, p2, s, counter):                  # F:1070-1157
    M = _mean4(s)
    O = [0.0, 0.0, 0.0]
    faces = [(0, 2, 3, 0), (0, 1, 3, 0), (0, 1, 2, 0), (1, 2, 3, 1)]     # (A,B,C, reference vertex)
    nml, w = [], []
    for a, b, c, ref in faces:
        AB = SUB(s[a], s[b])
        BC = SUB(s[b], s[c])
        n = UTZVEC(CROSS_PRODUCT_3D(AB, BC))
        if DOT(n, SUB(s[ref], M)) < 0.0:
            n = NEG(n)
        nml.append(n)
        w.append(DOT(NEG(n), SUB(s[ref], O)))
    loc = 0
    for i in range(1, 4):
        if w[i] > w[loc]:
            loc = i
    SM = support_mapping(p1, p2, nml[loc], counter)
    keep = [(0, 2, 3), (0, 1, 3), (0, 1, 2), (1, 2, 3)][loc]
    return [list(s[keep[0]]), list(s[keep[1]]), list(s[keep[2]]), SM]


def _qhull(points):
    """stand-in for GCLIB_QuickHull.QuickHull: triangulated hull, faces as 3x3 coordinates"""
    from scipy.spatial import ConvexHull
    pts = np.array(points)
    h = ConvexHull(pts, qhull_options="Qt")
    return [[list(pts[i]) for i in simp] for simp in h.simplices]


def update_expandingPolytope_EPA(p1, p2, poly1, counter):    # F:863-1022
    O = [0.0, 0.0, 0.0]
    dist1 = [abs(DIST_PF_SIGN(O, f)) for f in poly1]
    k1 = min(range(len(dist1)), key=lambda i: (dist1[i], i))
    min_val_1 = dist1[k1]
    d = UNINML(poly1[k1])
    dot = DOT(SUB(poly1[k1][0], O), d)
    if abs(dot) < 1.0e-12:
        M = []
        for c in range(3):
            s = 0.0
            for v in range(3):
                for f in poly1:
                    s = s + f[v][c]
            M.append(s / (len(poly1) * 3))
        dot = DOT(SUB(poly1[k1][0], M), d)
    if dot <= -1.0e-12:
        d = NEG(d)
    SPMP = support_mapping(p1, p2, d, counter)
    scat = []
    for f in poly1:
        for v in f:
            if not any(v[0] == q[0] and v[1] == q[1] and v[2] == q[2] for q in scat):
                scat.append(list(v))
    scat.append(SPMP)
    if abs(min_val_1) < 1.0e-12:
        scat.append(support_mapping(p1, p2, NEG(d), counter))
    poly2 = _qhull(scat)
    dist2 = [abs(DIST_PF_SIGN(O, f)) for f in poly2]
    k2 = min(range(len(dist2)), key=lambda i: (dist2[i], i))
    min_val_2 = dist2[k2]
    d = UNINML(poly2[k2])
    if DOT(SUB(poly2[k2][0], O), d) < 0.0:
        d = NEG(d)
    if len(dist1) == len(dist2):
        a, b = sorted(dist1), sorted(dist2)
        if all(abs(x - y) < 1.0e-8 for x, y in zip(a, b)):
            return False, poly2, min_val_2, d
        return True, poly2, 0.0, [0.0, 0.0, 0.0]
    if len(dist1) > len(dist2):
        return False, poly2, min_val_2, d
    return True, poly2, 0.0, [0.0, 0.0, 0.0]


def EPA_solu(p1, p2, s, counter):                            # F:242-323
    poly = [[s[0], s[1], s[2]], [s[0], s[2], s[3]], [s[0], s[1], s[3]], [s[1], s[2], s[3]]]
    it = 0
    while True:
        it += 1
        if it > 99:
            return 0.0, [0.0, 0.0, 0.0], 99, "cap"
        isExpa, poly_res, depth, nrm = update_expandingPolytope_EPA(p1, p2, poly, counter)
        if not isExpa:
            return depth, nrm, it, "ok"
        poly = poly_res


def _close(a, b):
    return all(abs(a[k] - b[k]) < 1.0e-8 for k in range(3))


def GJKEPA(p1, p2, run_epa=True):
    """F:39-239.  Returns dict(collision, exit, gjk_iters, n_support, depth, normal, epa_iters, note)."""
    p1 = [list(map(float, v)) for v in p1]
    p2 = [list(map(float, v)) for v in p2]
    O = [0.0, 0.0, 0.0]
    cnt = [0]
    res = dict(collision=False, exit="", gjk_iters=0, depth=0.0, normal=[0.0, 0.0, 0.0], epa_iters=0, note="")

    def finish(exit_name, coll, s=None):
        res["collision"] = coll
        res["exit"] = exit_name
        if coll and run_epa:
            try:
                depth, nrm, it, note = EPA_solu(p1, p2, s, cnt)
                res.update(depth=depth, normal=nrm, epa_iters=it, note=note)
            except StopReached:
                res["note"] = "stop"
            except Exception as e:                             # Qhull refuses flat input etc.
                res["note"] = "hull:" + type(e).__name__
        res["n_support"] = cnt[0]
        return res

    if not RoughCollisionDetection_SphericalEnvelope(p1, p2):
        return finish("sphere", False)
    s = [[0.0] * 3 for _ in range(4)]
    it = 0
    while True:
        it += 1
        if it > 99:
            return finish("init_cap", False)
        d = list(DATALIB[it - 1])
        s[0] = support_mapping(p1, p2, d, cnt)
        d = NEG(d)
        s[1] = support_mapping(p1, p2, d, cnt)
        if _close(s[0], s[1]):
            continue
        break
    d = VEC_PL(O, (s[0], s[1]))
    s[2] = support_mapping(p1, p2, d, cnt)
    if _close(s[2], s[0]) or _close(s[2], s[1]):
        return finish("dup3", False)
    V12 = SUB(s[1], s[0])
    V23 = SUB(s[2], s[1])
    d = UTZVEC(CROSS_PRODUCT_3D(V12, V23))
    VO = SUB(O, s[2])
    if abs(DOT(VO, d)) < 1.0e-8:
        if IS_INSIDE_PF(s[0:3], O):
            return finish("tri_origin", True, s)
    if DOT(VO, d) < 0.0:
        d = NEG(d)
    s[3] = support_mapping(p1, p2, d, cnt)
    try:
        if abs(DIST_PF_SIGN(s[3], s[0:3])) < 1.0e-8:
            return finish("coplanar4", False)
    except StopReached:
        res["note"] = "stop"
        return finish("stop", False)
    if isPointInSimplex(O, s):
        return finish("init_contain", True, s)
    last1 = [[0.0] * 3 for _ in range(4)]
    last2 = [[0.0] * 3 for _ in range(4)]
    it = 0
    while True:
        it += 1
        if it > 50:
            res["gjk_iters"] = 50
            return finish("loop_cap", False)
        res["gjk_iters"] = it
        last2 = last1
        last1 = [list(r) for r in s]
        s = update_simplex_GJK(p1, p2, s, cnt)
        if NORM2(CROSS_PRODUCT_3D(SUB(s[1], s[0]), SUB(s[2], s[1]))) < 1.0e-8:
            return finish("loop_collinear", False)
        if abs(DIST_PF_SIGN(s[3], s[0:3])) < 1.0e-8:
            return finish("loop_coplanar", False)
        if isPointInSimplex(O, s):
            return finish("loop_contain", True, s)
        isOver = False
        for i in range(4):
            if _close(s[i], last1[i]) or _close(s[i], last2[i]):
                isOver = True
            else:
                isOver = False
                break
        if isOver:
            return finish("loop_stall", False)
