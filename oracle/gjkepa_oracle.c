/* gjkepa_oracle.c -- CPU restatement of GCLIB_GJKEPA (reference src/GCLIB_GJKEPA.f90 = "F:<line>").
 *
 * TEST INFRASTRUCTURE ONLY -- see gjkepa_oracle.h.  PARITY UNPINNED for the EPA hull (the
 * reference's QuickHull is an un-vendored dependency); GJK flags follow F:39-239 statement by
 * statement.
 *
 * Arithmetic contract: IEEE binary64, round-to-nearest, NO fused
 * multiply-add (build with -ffp-contract=off, no -ffast-math), true division, left-to-right
 * DOT_PRODUCT/SUM starting from 0, strict '>' scans (first index wins), first-occurrence
 * MAXLOC/MINLOC.  NORM2 is sqrt((x*x+y*y)+z*z) (define ORC_NORM2_SCALED for the running-scale
 * variant gfortran is believed to inline).
 */
#include "gjkepa_oracle.h"
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ direction table, F:1578-1689 */
static const double DIR_TABLE[100][3] = {
#include "dirtable.inc"
};

void orc_dir_table(int32_t row1, double d[3])
{
    d[0] = DIR_TABLE[row1 - 1][0]; d[1] = DIR_TABLE[row1 - 1][1]; d[2] = DIR_TABLE[row1 - 1][2];
}

/* ------------------------------------------------------------------ L0 math tools */
static inline double dot3(const double a[3], const double b[3])
{   /* DOT_PRODUCT: ((0 + a1*b1) + a2*b2) + a3*b3 */
    double s = 0.0;
    s = s + a[0] * b[0];
    s = s + a[1] * b[1];
    s = s + a[2] * b[2];
    return s;
}

static inline void cross3(const double a[3], const double b[3], double r[3])
{   /* CROSS_PRODUCT_3D, F:1208-1210 */
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

static inline double norm2_3(const double v[3])
{
#ifdef ORC_NORM2_SCALED
    double scale = 1.0, r = 0.0;
    for (int k = 0; k < 3; ++k) {
        if (v[k] != 0.0) {
            double a = fabs(v[k]);
            if (a > scale) { double t = scale / a; r = 1.0 + r * t * t; scale = a; }
            else           { double t = a / scale; r += t * t; }
        }
    }
    return scale * sqrt(r);
#else
    return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
#endif
}

static inline void sub3(const double a[3], const double b[3], double r[3])
{ r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }

static inline void neg3(const double a[3], double r[3])
{ r[0] = -a[0]; r[1] = -a[1]; r[2] = -a[2]; }

/* UTZVEC, F:1343-1352: v/|v|, zero vector if |v| < 1e-12 */
static inline void utzvec(const double v[3], double r[3])
{
    double md = norm2_3(v);
    if (md < 1.0e-12) { r[0] = r[1] = r[2] = 0.0; }
    else { r[0] = v[0] / md; r[1] = v[1] / md; r[2] = v[2] / md; }
}

/* UNINML, F:1382-1394: unit normal of (P2-P1)x(P3-P2); zero unless some |c| > 1e-12.
 * tri is 3 points, row-major. */
static inline void uninml(const double tri[9], double r[3])
{
    double n1[3], n2[3], c[3];
    sub3(tri + 3, tri + 0, n1);
    sub3(tri + 6, tri + 3, n2);
    cross3(n1, n2, c);
    if (fabs(c[0]) > 1.0e-12 || fabs(c[1]) > 1.0e-12 || fabs(c[2]) > 1.0e-12) {
        double md = norm2_3(c);
        r[0] = c[0] / md; r[1] = c[1] / md; r[2] = c[2] / md;
    } else { r[0] = r[1] = r[2] = 0.0; }
}

/* DIST_PF_SIGN, F:1357-1377.  Returns 0 and *res on success; 1 on the reference's STOP path
 * (all |n| < 1e-12). */
static inline int dist_pf_sign(const double x[3], const double tri[9], double *res)
{
    double n[3], d[3];
    uninml(tri, n);
    if (fabs(n[0]) < 1.0e-12 && fabs(n[1]) < 1.0e-12 && fabs(n[2]) < 1.0e-12) { *res = 0.0; return 1; }
    sub3(x, tri, d);
    *res = dot3(d, n);
    return 0;
}

/* VEC_PL, F:1423-1440: unit vector from point C to its foot D on line (A,B) (quirk Q4) */
static void vec_pl(const double c[3], const double a[3], const double b[3], double r[3])
{
    double ab[3], ac[3], u[3], d[3], dc[3];
    sub3(b, a, ab);
    sub3(c, a, ac);
    double t = dot3(ac, ab) / norm2_3(ab);
    utzvec(ab, u);
    d[0] = a[0] + t * u[0]; d[1] = a[1] + t * u[1]; d[2] = a[2] + t * u[2];
    sub3(d, c, dc);
    utzvec(dc, r);
}

/* IS_INSIDE_PF, F:1271-1337 (quirks Q2, Q3).  v: nnode points row-major. */
int orc_is_inside_pf(const double *v, int32_t nnode, const double p[3])
{
    double c[64];
    int i;
    if (nnode > 64) nnode = 64;
    for (i = 0; i < nnode; ++i) {
        int j = (i == nnode - 1) ? 0 : i + 1;
        c[i] = (v[3 * j + 0] - v[3 * i + 0]) * (p[1] - v[3 * i + 1])
             - (v[3 * j + 1] - v[3 * i + 1]) * (p[0] - v[3 * i + 0]);
    }
    for (i = 0; i < nnode; ++i) if (fabs(c[i]) < 1.0e-12) c[i] = 0.0;
    int any_pos = 0;
    for (i = 0; i < nnode; ++i) if (c[i] > 1.0e-15) any_pos = 1;
    if (!any_pos) {                 /* F:1310-1322: re-project on XZ (no re-zeroing afterwards) */
        for (i = 0; i < nnode; ++i) {
            int j = (i == nnode - 1) ? 0 : i + 1;
            c[i] = (v[3 * j + 0] - v[3 * i + 0]) * (p[2] - v[3 * i + 2])
                 - (v[3 * j + 2] - v[3 * i + 2]) * (p[0] - v[3 * i + 0]);
        }
    }
    for (i = 0; i < nnode; ++i) {
        double t = c[0] * c[i];
        if (t < 0.0) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------ L1 kernels */
typedef struct { const double *x, *y, *z; int n; } hull_t;

static inline hull_t mk_hull(const double *p, int n)
{ hull_t h; h.x = p; h.y = p + n; h.z = p + 2 * (int64_t)n; h.n = n; return h; }

/* support_mapping, F:1030-1062.  Strict '>' from -HUGE: first index wins ties. */
static void support(const hull_t *a, const hull_t *b, const double dir[3], double res[3],
                    int *ia, int *ib, int32_t *counter)
{
    double best = -DBL_MAX; int i1 = 0, i2 = 0, i;
    double v[3], nd[3];
    for (i = 0; i < a->n; ++i) {
        v[0] = a->x[i]; v[1] = a->y[i]; v[2] = a->z[i];
        double t = dot3(dir, v);
        if (t > best) { best = t; i1 = i; }
    }
    neg3(dir, nd);
    best = -DBL_MAX;
    for (i = 0; i < b->n; ++i) {
        v[0] = b->x[i]; v[1] = b->y[i]; v[2] = b->z[i];
        double t = dot3(nd, v);
        if (t > best) { best = t; i2 = i; }
    }
    res[0] = a->x[i1] - b->x[i2];
    res[1] = a->y[i1] - b->y[i2];
    res[2] = a->z[i1] - b->z[i2];
    if (ia) *ia = i1;
    if (ib) *ib = i2;
    if (counter) ++*counter;
}

void orc_support(const double *p1, int32_t n1, const double *p2, int32_t n2,
                 const double dir[3], double res[3], int32_t *i1, int32_t *i2)
{
    hull_t a = mk_hull(p1, n1), b = mk_hull(p2, n2);
    int ia, ib;
    support(&a, &b, dir, res, &ia, &ib, 0);
    if (i1) *i1 = ia;
    if (i2) *i2 = ib;
}

/* RoughCollisionDetection_SphericalEnvelope, F:1165-1188 */
static void hull_sphere(const hull_t *h, double mp[3], double *r)
{
    double sx = 0.0, sy = 0.0, sz = 0.0; int i;
    for (i = 0; i < h->n; ++i) { sx = sx + h->x[i]; sy = sy + h->y[i]; sz = sz + h->z[i]; }
    mp[0] = sx / (double)h->n; mp[1] = sy / (double)h->n; mp[2] = sz / (double)h->n;
    double rr = -DBL_MAX;                      /* MAXVAL */
    for (i = 0; i < h->n; ++i) {
        double d[3] = { h->x[i] - mp[0], h->y[i] - mp[1], h->z[i] - mp[2] };
        double t = norm2_3(d);
        if (t > rr) rr = t;
    }
    *r = rr;
}

static int sphere_test(const hull_t *a, const hull_t *b)
{
    double m1[3], m2[3], r1, r2, d[3];
    hull_sphere(a, m1, &r1);
    hull_sphere(b, m2, &r2);
    sub3(m1, m2, d);
    return norm2_3(d) <= r1 + r2 + 1.0;       /* F:1185, TOL = 1.0 absolute (Q7) */
}

int orc_sphere_test(const double *p1, int32_t n1, const double *p2, int32_t n2)
{ hull_t a = mk_hull(p1, n1), b = mk_hull(p2, n2); return sphere_test(&a, &b); }

/* face vertex table shared by update_simplex_GJK (F:1093-1118) and isPointInSimplex
 * (idFc, F:1227-1229, column-major fill): faces (1,3,4) (1,2,4) (1,2,3) (2,3,4), 0-based here */
static const int FACE_V[4][3] = { {0, 2, 3}, {0, 1, 3}, {0, 1, 2}, {1, 2, 3} };

static void face_normal_raw(const double s[12], int f, double n[3])
{
    double ab[3], bc[3], c[3];
    sub3(s + 3 * FACE_V[f][0], s + 3 * FACE_V[f][1], ab);
    sub3(s + 3 * FACE_V[f][1], s + 3 * FACE_V[f][2], bc);
    cross3(ab, bc, c);
    utzvec(c, n);
}

static void simplex_mean(const double s[12], double m[3])
{   /* SUM(simplex_(:,i)) / 4, F:1086, F:1232 */
    for (int k = 0; k < 3; ++k) {
        double t = 0.0;
        t = t + s[0 + k]; t = t + s[3 + k]; t = t + s[6 + k]; t = t + s[9 + k];
        m[k] = t / 4.0;
    }
}

/* isPointInSimplex, F:1217-1265 */
int orc_point_in_simplex(const double p[3], const double s[12])
{
    double m[3], nml[4][3], dist[4], d[3];
    int i, j;
    simplex_mean(s, m);
    for (i = 0; i < 4; ++i) {
        face_normal_raw(s, i, nml[i]);
        sub3(s + 3 * i, m, d);                                   /* simplex_(i,:) - M, F:1240 */
        if (dot3(nml[i], d) < 0.0) neg3(nml[i], nml[i]);
    }
    for (i = 0; i < 4; ++i) { sub3(s + 3 * i, p, d); dist[i] = dot3(d, nml[i]); }
    for (i = 0; i < 4; ++i) {
        if (fabs(dist[i]) < 1.0e-8) {
            double tri[9];
            for (j = 0; j < 3; ++j) memcpy(tri + 3 * j, s + 3 * FACE_V[i][j], 3 * sizeof(double));
            if (orc_is_inside_pf(tri, 3, p)) return 1;
        }
    }
    return dist[0] > 0.0 && dist[1] > 0.0 && dist[2] > 0.0 && dist[3] > 0.0;
}

/* update_simplex_GJK, F:1070-1157: tetrahedron walk (not Johnson sub-simplex) */
static void update_simplex(const hull_t *a, const hull_t *b, const double s[12], double out[12],
                           int32_t *counter)
{
    static const int REF[4] = { 0, 0, 0, 1 };   /* flip / distance reference vertex per face */
    double m[3], nml[4][3], w[4], d[3], nn[3], sm[3];
    int f, k;
    simplex_mean(s, m);
    for (f = 0; f < 4; ++f) {
        face_normal_raw(s, f, nml[f]);
        sub3(s + 3 * REF[f], m, d);
        if (dot3(nml[f], d) < 0.0) neg3(nml[f], nml[f]);
        neg3(nml[f], nn);
        w[f] = dot3(nn, s + 3 * REF[f]);          /* DOT(-nml, simplex_(ref,:) - O), O = 0: x-0 = x */
    }
    k = 0;                                       /* MAXLOC: first maximum */
    for (f = 1; f < 4; ++f) if (w[f] > w[k]) k = f;
    support(a, b, nml[k], sm, 0, 0, counter);
    for (f = 0; f < 3; ++f) memcpy(out + 3 * f, s + 3 * FACE_V[k][f], 3 * sizeof(double));
    memcpy(out + 9, sm, 3 * sizeof(double));
}

/* ------------------------------------------------------------------ hull stand-in (NOT in the reference)
 * Replaces GCLIB_QuickHull.QuickHull (call F:950) and GCLIB_DeHull.getHullMeshesVertex (F:920).
 * A plain QuickHull: exact-duplicate removal, extreme-point tetrahedron, then repeatedly take the
 * lowest-index face with a non-empty outside set, insert its farthest point, delete every face
 * that sees it by more than QH_EPS, and fan the horizon to the point.  Points never more than
 * QH_EPS outside any face are swallowed (what the reference's own comment at F:1005 describes).
 * Output triangles are outward-wound ((v2-v1)x(v3-v2) points away from the hull).            */
#define QH_EPS   1.0e-10
#define QH_MAXP  160
#define QH_MAXF  640

typedef struct { int v[3]; double n[3]; double d; } qh_face;

static int qh_plane(const double (*p)[3], qh_face *f)
{
    double e1[3], e2[3], c[3];
    sub3(p[f->v[1]], p[f->v[0]], e1);
    sub3(p[f->v[2]], p[f->v[1]], e2);
    cross3(e1, e2, c);
    double md = norm2_3(c);
    if (!(md > 0.0)) return 1;
    f->n[0] = c[0] / md; f->n[1] = c[1] / md; f->n[2] = c[2] / md;
    f->d = dot3(f->n, p[f->v[0]]);
    return 0;
}

static inline double qh_dist(const qh_face *f, const double q[3]) { return dot3(f->n, q) - f->d; }

int orc_quickhull(const double *pts_in, int32_t npts, double *tris, int32_t *ntri, int32_t cap)
{
    double p[QH_MAXP][3];
    int n = 0, i, j, k;
    *ntri = 0;
    if (npts > QH_MAXP) return 2;
    for (i = 0; i < npts; ++i) {                         /* exact-duplicate removal, order kept */
        int dup = 0;
        for (j = 0; j < n; ++j)
            if (p[j][0] == pts_in[3 * i] && p[j][1] == pts_in[3 * i + 1] && p[j][2] == pts_in[3 * i + 2]) { dup = 1; break; }
        if (!dup) { p[n][0] = pts_in[3 * i]; p[n][1] = pts_in[3 * i + 1]; p[n][2] = pts_in[3 * i + 2]; ++n; }
    }
    if (n < 4) return 1;

    /* initial tetrahedron from extreme points */
    int ext[6];
    for (k = 0; k < 3; ++k) {
        int lo = 0, hi = 0;
        for (i = 1; i < n; ++i) { if (p[i][k] < p[lo][k]) lo = i; if (p[i][k] > p[hi][k]) hi = i; }
        ext[2 * k] = lo; ext[2 * k + 1] = hi;
    }
    int i0 = 0, i1 = 0; double best = -1.0;
    for (i = 0; i < 6; ++i) for (j = i + 1; j < 6; ++j) {
        double d[3]; sub3(p[ext[i]], p[ext[j]], d);
        double t = dot3(d, d);
        if (t > best) { best = t; i0 = ext[i]; i1 = ext[j]; }
    }
    if (!(best > QH_EPS * QH_EPS)) return 1;
    int i2 = -1; best = -1.0;
    {
        double ab[3]; sub3(p[i1], p[i0], ab);
        double ab2 = dot3(ab, ab);
        for (i = 0; i < n; ++i) {
            double ap[3], c[3]; sub3(p[i], p[i0], ap); cross3(ap, ab, c);
            double t = dot3(c, c) / ab2;
            if (t > best) { best = t; i2 = i; }
        }
    }
    if (!(best > QH_EPS * QH_EPS)) return 1;
    qh_face base; base.v[0] = i0; base.v[1] = i1; base.v[2] = i2;
    if (qh_plane(p, &base)) return 1;
    int i3 = -1; best = -1.0; double sgn = 0.0;
    for (i = 0; i < n; ++i) {
        double t = qh_dist(&base, p[i]);
        if (fabs(t) > best) { best = fabs(t); i3 = i; sgn = t; }
    }
    if (!(best > QH_EPS)) return 1;
    if (sgn > 0.0) { int t = i1; i1 = i2; i2 = t; }      /* make the base face look away from i3 */

    static const int T4[4][3] = { {0, 1, 2}, {1, 0, 3}, {2, 1, 3}, {0, 2, 3} };
    int tv[4] = { i0, i1, i2, i3 };
    qh_face F[QH_MAXF]; int nf = 4;
    for (i = 0; i < 4; ++i) {
        F[i].v[0] = tv[T4[i][0]]; F[i].v[1] = tv[T4[i][1]]; F[i].v[2] = tv[T4[i][2]];
        if (qh_plane(p, &F[i])) return 1;
    }

    int owner[QH_MAXP]; double odist[QH_MAXP];
    for (i = 0; i < n; ++i) {
        owner[i] = -1; odist[i] = 0.0;
        if (i == i0 || i == i1 || i == i2 || i == i3) continue;
        for (j = 0; j < nf; ++j) { double t = qh_dist(&F[j], p[i]); if (t > QH_EPS) { owner[i] = j; odist[i] = t; break; } }
    }

    for (;;) {
        int fsel = -1;
        for (i = 0; i < n; ++i) if (owner[i] >= 0 && (fsel < 0 || owner[i] < fsel)) fsel = owner[i];
        if (fsel < 0) break;
        int P = -1; best = -1.0;
        for (i = 0; i < n; ++i) if (owner[i] == fsel && odist[i] > best) { best = odist[i]; P = i; }

        unsigned char vis[QH_MAXF];
        for (j = 0; j < nf; ++j) vis[j] = qh_dist(&F[j], p[P]) > QH_EPS;

        /* horizon: directed edges of visible faces whose reverse is not in a visible face */
        int he[QH_MAXF][2], nh = 0;
        for (j = 0; j < nf; ++j) if (vis[j]) {
            for (k = 0; k < 3; ++k) {
                int a = F[j].v[k], b = F[j].v[(k + 1) % 3], twin = 0, g, e;
                for (g = 0; g < nf && !twin; ++g) if (vis[g] && g != j)
                    for (e = 0; e < 3; ++e) if (F[g].v[e] == b && F[g].v[(e + 1) % 3] == a) { twin = 1; break; }
                if (!twin) { if (nh >= QH_MAXF) return 2; he[nh][0] = a; he[nh][1] = b; ++nh; }
            }
        }

        /* rebuild: surviving faces keep their order, new faces follow */
        int map[QH_MAXF], nn = 0;
        qh_face G[QH_MAXF];
        for (j = 0; j < nf; ++j) { if (!vis[j]) { map[j] = nn; G[nn++] = F[j]; } else map[j] = -1; }
        int first_new = nn;
        for (j = 0; j < nh; ++j) {
            if (nn >= QH_MAXF) return 2;
            G[nn].v[0] = he[j][0]; G[nn].v[1] = he[j][1]; G[nn].v[2] = P;
            if (qh_plane(p, &G[nn])) return 3;
            ++nn;
        }
        owner[P] = -1;
        for (i = 0; i < n; ++i) if (owner[i] >= 0) {
            if (map[owner[i]] >= 0) { owner[i] = map[owner[i]]; continue; }
            owner[i] = -1;
            for (j = first_new; j < nn; ++j) { double t = qh_dist(&G[j], p[i]); if (t > QH_EPS) { owner[i] = j; odist[i] = t; break; } }
        }
        memcpy(F, G, (size_t)nn * sizeof(qh_face));
        nf = nn;
    }

    if (nf > cap) return 2;
    for (j = 0; j < nf; ++j)
        for (k = 0; k < 3; ++k) memcpy(tris + 9 * j + 3 * k, p[F[j].v[k]], 3 * sizeof(double));
    *ntri = nf;
    return 0;
}

/* unique vertices of a triangle list, first-appearance order (stand-in for getHullMeshesVertex) */
static int mesh_vertices(const double *tris, int nt, double (*out)[3])
{
    int n = 0, i, j;
    for (i = 0; i < 3 * nt; ++i) {
        const double *q = tris + 3 * i; int dup = 0;
        for (j = 0; j < n; ++j) if (out[j][0] == q[0] && out[j][1] == q[1] && out[j][2] == q[2]) { dup = 1; break; }
        if (!dup) { out[n][0] = q[0]; out[n][1] = q[1]; out[n][2] = q[2]; ++n; }
    }
    return n;
}

/* ------------------------------------------------------------------ L2 EPA */
#define EPA_MAXF 512

/* update_expandingPolytope_EPA, F:863-1022.  poly1: nf1 triangles (9 doubles each).
 * returns 0 ok, 1 STOP path (degenerate face), 2 hull stand-in failed */
static int epa_expand(const hull_t *a, const hull_t *b, const double *poly1, int nf1,
                      int *is_expa, double *poly2, int *nf2_out, double *depth, double nrm[3],
                      int32_t *counter)
{
    static const double O[3] = { 0.0, 0.0, 0.0 };
    double dist1[EPA_MAXF], dist2[EPA_MAXF];
    int i, j, k1 = 0, k2 = 0;
    *is_expa = 0; *depth = 0.0; nrm[0] = nrm[1] = nrm[2] = 0.0;

    for (i = 0; i < nf1; ++i) {                                  /* F:893-895 */
        double d; if (dist_pf_sign(O, poly1 + 9 * i, &d)) return 1;
        dist1[i] = fabs(d);
    }
    for (i = 1; i < nf1; ++i) if (dist1[i] < dist1[k1]) k1 = i;   /* MINLOC, first minimum */
    double min1 = dist1[k1];

    double dir[3], v1[3], dt;
    uninml(poly1 + 9 * k1, dir);                                 /* F:902 */
    sub3(poly1 + 9 * k1, O, v1);
    dt = dot3(v1, dir);
    if (fabs(dt) < 1.0e-12) {                                    /* F:905-908: re-base on the slot centroid */
        double m[3];
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int v = 0; v < 3; ++v) for (i = 0; i < nf1; ++i) s = s + poly1[9 * i + 3 * v + c];
            m[c] = s / (double)(nf1 * 3);
        }
        sub3(poly1 + 9 * k1, m, v1);
        dt = dot3(v1, dir);
    }
    if (dt <= -1.0e-12) neg3(dir, dir);                          /* F:910 */

    double sp[3];
    support(a, b, dir, sp, 0, 0, counter);                       /* F:914 */

    double scat[QH_MAXP][3];
    int ns = mesh_vertices(poly1, nf1, scat);                    /* F:920 */
    if (ns + 2 > QH_MAXP) return 2;
    memcpy(scat[ns++], sp, sizeof sp);                           /* F:927-932 (appended unconditionally) */
    if (fabs(min1) < 1.0e-12) {                                  /* F:935-944 */
        double nd[3]; neg3(dir, nd);
        support(a, b, nd, sp, 0, 0, counter);
        memcpy(scat[ns++], sp, sizeof sp);
    }

    int nf2 = 0;
    if (orc_quickhull(&scat[0][0], ns, poly2, &nf2, EPA_MAXF)) return 2;   /* F:950 */
    *nf2_out = nf2;

    for (i = 0; i < nf2; ++i) {                                  /* F:958-960 */
        double d; if (dist_pf_sign(O, poly2 + 9 * i, &d)) return 1;
        dist2[i] = fabs(d);
    }
    for (i = 1; i < nf2; ++i) if (dist2[i] < dist2[k2]) k2 = i;
    double min2 = dist2[k2];
    uninml(poly2 + 9 * k2, dir);                                 /* F:967-969 */
    sub3(poly2 + 9 * k2, O, v1);
    if (dot3(v1, dir) < 0.0) neg3(dir, dir);

    if (nf1 == nf2) {                                            /* F:972-1003 */
        int n = nf1;
        for (i = 1; i <= n - 1; ++i) for (j = 0; j < n - i; ++j) {
            if (dist1[j] > dist1[j + 1]) { double t = dist1[j]; dist1[j] = dist1[j + 1]; dist1[j + 1] = t; }
            if (dist2[j] > dist2[j + 1]) { double t = dist2[j]; dist2[j] = dist2[j + 1]; dist2[j + 1] = t; }
        }
        int same = 1;
        for (i = 0; i < n; ++i) if (!(fabs(dist1[i] - dist2[i]) < 1.0e-8)) { same = 0; break; }
        if (same) { *is_expa = 0; *depth = min2; memcpy(nrm, dir, sizeof dir); }
        else      { *is_expa = 1; }
    } else if (nf1 > nf2) {                                      /* F:1005-1009 */
        *is_expa = 0; *depth = min2; memcpy(nrm, dir, sizeof dir);
    } else {
        *is_expa = 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ contact epilogue, F:326-343 */
static inline void hull_pt(const hull_t *h, int i, double v[3]) { v[0] = h->x[i]; v[1] = h->y[i]; v[2] = h->z[i]; }

/* get_nearest_points, F:813-855 */
static void nearest_points(const hull_t *a, const hull_t *b, const double nrm[3], double res[6])
{
    double best = -DBL_MAX, v[3], nd[3]; int i1 = 0, i2 = 0, i;
    for (i = 0; i < a->n; ++i) { hull_pt(a, i, v); double t = dot3(nrm, v); if (t > best) { best = t; i1 = i; } }
    neg3(nrm, nd); best = -DBL_MAX;
    for (i = 0; i < b->n; ++i) { hull_pt(b, i, v); double t = dot3(nd, v); if (t > best) { best = t; i2 = i; } }
    res[0] = a->x[i1]; res[2] = a->y[i1]; res[4] = a->z[i1];     /* (2,3) column-major */
    res[1] = b->x[i2]; res[3] = b->y[i2]; res[5] = b->z[i2];
}

/* the "> max - 1e-8, last wins, previous winner becomes second" tracker of F:722-747 (Q8) */
static void near_max_pair(const hull_t *h, const double d[3], int idx[2])
{
    double mx = -DBL_MAX, v[3]; int i;
    idx[0] = idx[1] = 0;                                         /* 1-based, 0 = unset */
    for (i = 0; i < h->n; ++i) {
        hull_pt(h, i, v);
        double t = dot3(d, v);
        if (t > mx - 1.0e-8) { mx = t; idx[1] = idx[0]; idx[0] = i + 1; }
    }
    if (idx[1] == 0) idx[1] = idx[0];
}

/* get_collisionPoint_01, F:700-806 */
static void collision_point_01(const hull_t *a, const hull_t *b, const double nrm[3], double res[3])
{
    int s1[2], s2[2], i; double nd[3], v[3], w[3];
    res[0] = res[1] = res[2] = 0.0;
    near_max_pair(a, nrm, s1);
    neg3(nrm, nd);
    near_max_pair(b, nd, s2);
    int m1 = s1[0] != s1[1], m2 = s2[0] != s2[1];
    if (!m1 && !m2) {
        hull_pt(a, s1[0] - 1, v); hull_pt(b, s2[0] - 1, w);
        for (i = 0; i < 3; ++i) res[i] = (v[i] + w[i]) / 2.0;
    }
    if (m1 && !m2)      hull_pt(b, s2[0] - 1, res);
    else if (!m1 && m2) hull_pt(a, s1[0] - 1, res);
    if (m1 && m2) {                                              /* F:766-804 */
        double mx = -DBL_MAX, sx = 0.0, sy = 0.0, sz = 0.0; int c = 0;
        for (i = 0; i < a->n; ++i) { hull_pt(a, i, v); double t = dot3(nrm, v); if (t > mx) mx = t; }
        for (i = 0; i < a->n; ++i) {
            hull_pt(a, i, v);
            double t = dot3(nrm, v);
            if (t > mx - 1.0e-1) { ++c; sx = sx + v[0]; sy = sy + v[1]; sz = sz + v[2]; }
        }
        res[0] = sx / (double)c; res[1] = sy / (double)c; res[2] = sz / (double)c;
    }
}

/* ---- get_collisionPoint_02, F:457-696.  The reference keeps the support vertices in GCLIB_List's List_Array3d
 * (un-vendored; used only as an append-only array: reset / append / getSize / atIndex, F:466-468, F:526, F:540),
 * restated here as plain arrays in append order. ---- */
#define CP2_MAX 256

/* AddAllSupports, F:510-530: every vertex within tol of the strict maximum along nml, in index order */
static int all_supports(const hull_t *h, const double nml[3], double tol, double (*out)[3])
{
    double mx = -DBL_MAX, v[3]; int i, n = 0;
    for (i = 0; i < h->n; ++i) { hull_pt(h, i, v); double t = dot3(nml, v); if (t > mx) mx = t; }
    for (i = 0; i < h->n; ++i) {
        hull_pt(h, i, v);
        double t = dot3(nml, v);
        if (t > mx - tol && n < CP2_MAX) { memcpy(out[n], v, 3 * sizeof(double)); ++n; }
    }
    return n;
}

/* FOOT_PL, F:1496-1510: foot of P on the line through V1, V2 */
static void foot_pl(const double p[3], const double v1[3], const double v2[3], double r[3])
{
    double d[3], u[3], pv[3];
    sub3(v2, v1, d); utzvec(d, u); sub3(p, v1, pv);
    double t = dot3(pv, u);
    r[0] = v1[0] + t * u[0]; r[1] = v1[1] + t * u[1]; r[2] = v1[2] + t * u[2];
}

/* FOOT_LL, F:1445-1491: feet of the common perpendicular of two lines (parallel: midpoint of the first) */
static void foot_ll(const double p1[3], const double q1[3], const double p2[3], const double q2[3], double f1[3], double f2[3])
{
    double d1[3], d2[3], r[3];
    sub3(q1, p1, d1); sub3(q2, p2, d2); sub3(p1, p2, r);
    double a = dot3(d1, d1), b = dot3(d1, d2), c = dot3(d1, r), e = dot3(d2, d2), f = dot3(d2, r);
    double d = a * e - b * b;
    int i;
    if (fabs(d) < 1.0e-12) {
        for (i = 0; i < 3; ++i) f1[i] = (p1[i] + q1[i]) / 2.0;
        foot_pl(f1, p2, q2, f2);
    } else {
        double s = (b * f - c * e) / d, t = (a * f - b * c) / d;
        for (i = 0; i < 3; ++i) { f1[i] = p1[i] + s * (q1[i] - p1[i]); f2[i] = p2[i] + t * (q2[i] - p2[i]); }
    }
}

/* SORT_CLOCK, F:1517-1575: polygon vertices in counter-clockwise order around their centroid, starting from
 * the first.  Quirks kept: the angle is reduced with a SINGLE-precision 2*pi (F:1547); OVERLAP (all points within
 * 1e-12, F:1399-1418) returns before anything is written -- the result is then undefined in the reference, here
 * the input order.  IS_POINT_IN_ORDERED_POINTS compares against ALL rows of the result, written or not (F:1560-
 * 1571, uninitialised memory in the reference); only the rows already written are compared here. */
static void sort_clock(double (*pts)[3], int n, double (*out)[3])
{
    int i, j, k, overlap = 1;
    for (i = 0; i < n - 1 && overlap; ++i)
        for (j = i + 1; j < n && overlap; ++j)
            for (k = 0; k < 3; ++k) if (fabs(pts[i][k] - pts[j][k]) > 1.0e-12) { overlap = 0; break; }
    memcpy(out, pts, (size_t)n * 3 * sizeof(double));
    if (overlap) return;
    double cen[3] = {0.0, 0.0, 0.0}, v1[3], v2[3], nrm[3], cr[3];
    for (k = 0; k < 3; ++k) { double sum = 0.0; for (i = 0; i < n; ++i) sum = sum + pts[i][k]; cen[k] = sum / (double)n; }
    sub3(pts[1], pts[0], v1); sub3(pts[2], pts[0], v2); cross3(v1, v2, nrm);
    const double two_pi = (double)(2.0f * acosf(-1.0f));
    for (i = 1; i < n; ++i) {
        double min_angle = DBL_MAX; int index = -1;
        for (j = 0; j < n; ++j) {
            int used = 0, m;
            for (m = 0; m < i; ++m) if (pts[j][0] == out[m][0] && pts[j][1] == out[m][1] && pts[j][2] == out[m][2]) { used = 1; break; }
            if (used) continue;
            sub3(pts[j], cen, v1); sub3(out[i - 1], cen, v2);
            cross3(v2, v1, cr);
            double ang = atan2(dot3(nrm, cr), dot3(v1, v2));
            ang = fmod(ang + two_pi, two_pi);
            if (ang < 0.0) ang += two_pi;                            /* MODULO takes the sign of its second argument */
            if (ang < min_angle) { min_angle = ang; index = j; }
        }
        if (index < 0) index = 0;                                     /* duplicates exhausted the candidates: reference reads row -1 */
        memcpy(out[i], pts[index], 3 * sizeof(double));
    }
}

static void collision_point_02(const hull_t *a, const hull_t *b, const double nrm[3], double res[3])
{
    static _Thread_local double s1[CP2_MAX][3], s2[CP2_MAX][3], sorted[CP2_MAX][3];
    double nd[3];
    int i, k;
    res[0] = res[1] = res[2] = 0.0;
    neg3(nrm, nd);
    const int n1 = all_supports(a, nrm, 1.0e-1, s1), n2 = all_supports(b, nd, 1.0e-1, s2);
    if (n1 == 1 && n2 == 1) { for (i = 0; i < 3; ++i) res[i] = (s1[0][i] + s2[0][i]) / 2.0; }          /* case_01 */
    else if (n1 == 1 && n2 >= 2) memcpy(res, s1[0], 3 * sizeof(double));                                 /* case_02 */
    else if (n1 >= 2 && n2 == 1) memcpy(res, s2[0], 3 * sizeof(double));
    else if (n1 == 2 && n2 == 2) {                                                                        /* case_03 */
        double f1[3], f2[3];
        foot_ll(s1[0], s1[1], s2[0], s2[1], f1, f2);
        for (i = 0; i < 3; ++i) res[i] = (f1[i] + f2[i]) / 2.0;
    } else if ((n1 == 2 && n2 >= 3) || (n1 >= 3 && n2 == 2)) {                                            /* case_04 */
        double (*poly)[3] = (n1 >= 3) ? s1 : s2; const int np = (n1 >= 3) ? n1 : n2;
        double (*seg)[3] = (n1 >= 3) ? s2 : s1;
        sort_clock(poly, np, sorted);
        int c = 0;
        for (i = 0; i < 2; ++i) if (orc_is_inside_pf(&sorted[0][0], np, seg[i])) ++c;
        if (c == 0) {                                                                                     /* case_04_1 */
            double cen[3];
            for (k = 0; k < 3; ++k) { double sum = 0.0; for (i = 0; i < np; ++i) sum = sum + poly[i][k]; cen[k] = sum / (double)np; }
            foot_pl(cen, seg[0], seg[1], res);
        } else for (i = 0; i < 3; ++i) res[i] = (seg[0][i] + seg[1][i]) * 0.5;                            /* case_04_2 / _3 */
    } else if (n1 >= 3 && n2 >= 3) {                                                                      /* case_05 */
        for (k = 0; k < 3; ++k) { double sum = 0.0; for (i = 0; i < n1; ++i) sum = sum + s1[i][k]; res[k] = sum / (double)n1; }
    }
}

/* get_collisionPoint_03, F:426-452 */
static void collision_point_03(const hull_t *a, const hull_t *b, const double nrm[3], double res[3], double nnew[3])
{
    double mx = -DBL_MAX, nd[3], v[3], s = 0.0; int idx = 0, i;
    neg3(nrm, nd);
    for (i = 0; i < b->n; ++i) { hull_pt(b, i, v); double t = dot3(nd, v); if (t > mx - 1.0e-8) { mx = t; idx = i; } }
    hull_pt(b, idx, res);
    for (i = 0; i < a->n; ++i) s = s + a->z[i];
    res[2] = s / (double)(float)a->n;                            /* REAL(SIZE()) is single precision, F:446 */
    nnew[0] = nrm[0]; nnew[1] = nrm[1]; nnew[2] = 0.0;
    double md = norm2_3(nnew);
    nnew[0] = nnew[0] / md; nnew[1] = nnew[1] / md; nnew[2] = nnew[2] / md;   /* NaN if normal || z, as the reference */
}

/* get_info_collisionType, F:353-413 */
static int collision_type(const hull_t *a, const hull_t *b, const double nrm[3], double tol)
{
    double mx = -DBL_MAX, v[3], nd[3]; int c = 0, d = 0, i;
    for (i = 0; i < a->n; ++i) { hull_pt(a, i, v); double t = dot3(nrm, v); if (t > mx) mx = t; }
    for (i = 0; i < a->n; ++i) { hull_pt(a, i, v); double t = dot3(nrm, v); if (t > mx - tol) ++c; }
    neg3(nrm, nd); mx = -DBL_MAX;
    for (i = 0; i < b->n; ++i) { hull_pt(b, i, v); double t = dot3(nd, v); if (t > mx) mx = t; }
    for (i = 0; i < b->n; ++i) { hull_pt(b, i, v); double t = dot3(nd, v); if (t > mx - tol) ++d; }
    return (c >= 3 && d >= 3) ? 2 : 1;
}

/* EPA_solu, F:242-346 */
static void epa_solve(int version, double tol_ff, const hull_t *a, const hull_t *b, const double s[12],
                      int mode, orc_result *out)
{
    static const int SEED[4][3] = { {0, 1, 2}, {0, 2, 3}, {0, 1, 3}, {1, 2, 3} };   /* F:279-293 */
    double polyA[EPA_MAXF * 9], polyB[EPA_MAXF * 9];
    double *poly = polyA, *poly_res = polyB;
    int nf = 4, i, j, iter = 0;
    for (i = 0; i < 4; ++i) for (j = 0; j < 3; ++j) memcpy(poly + 9 * i + 3 * j, s + 3 * SEED[i][j], 3 * sizeof(double));

    for (;;) {
        ++iter;
        if (iter > 99) { out->status |= ORC_ST_EPA_CAP | (99u << 16); return; }   /* F:299-303: outputs stay zero */
        int is_expa, nf2 = 0; double depth, nrm[3];
        int rc = epa_expand(a, b, poly, nf, &is_expa, poly_res, &nf2, &depth, nrm, &out->n_support);
        if (rc) {
            out->status |= (rc == 1 ? ORC_ST_STOP : ORC_ST_HULL_FAIL) | ((uint32_t)iter << 16);
            return;
        }
        if (!is_expa) { out->depth = depth; memcpy(out->normal, nrm, sizeof nrm); break; }
        double *t = poly; poly = poly_res; poly_res = t; nf = nf2;                /* F:316-321 */
    }
    out->status |= (uint32_t)iter << 16;
    if (mode < ORC_MODE_FULL) return;

    nearest_points(a, b, out->normal, out->nearest);                              /* F:326 */
    if (version == 1) collision_point_01(a, b, out->normal, out->point);
    else if (version == 2) collision_point_02(a, b, out->normal, out->point);
    else if (version == 3) {
        double nnew[3];
        collision_point_03(a, b, out->normal, out->point, nnew);
        memcpy(out->normal, nnew, sizeof nnew);
    } else { out->status |= ORC_ST_BAD_VERSION; }                                  /* the reference PAUSEs and STOPs, F:336-339 */
    out->colli_type = collision_type(a, b, out->normal, tol_ff);                  /* F:343 */
}

/* The contact epilogue alone (F:326-343) on a GIVEN normal: what the reference computes once EPA has produced
 * collision_normal_.  For tests that check a result's nearest points / collision point / type against the reference's
 * own rules evaluated on that result's normal (ties between faces of equal depth may legitimately pick another normal). */
void orc_epilogue(int32_t version, double tol_ff, const double *p1, int32_t n1, const double *p2, int32_t n2,
                  const double normal_in[3], double nearest[6], double point[3], double normal_out[3], int32_t *colli_type)
{
    hull_t a = mk_hull(p1, n1), b = mk_hull(p2, n2);
    double nrm[3] = { normal_in[0], normal_in[1], normal_in[2] };
    point[0] = point[1] = point[2] = 0.0;
    nearest_points(&a, &b, nrm, nearest);
    if (version == 1) collision_point_01(&a, &b, nrm, point);
    else if (version == 2) collision_point_02(&a, &b, nrm, point);
    else if (version == 3) { double nnew[3]; collision_point_03(&a, &b, nrm, point, nnew); memcpy(nrm, nnew, sizeof nnew); }
    memcpy(normal_out, nrm, sizeof nrm);
    *colli_type = collision_type(&a, &b, nrm, tol_ff);
}

/* ------------------------------------------------------------------ L3 GJKEPA, F:39-239 */
static inline int all_close3(const double a[3], const double b[3])
{ return fabs(a[0] - b[0]) < 1.0e-8 && fabs(a[1] - b[1]) < 1.0e-8 && fabs(a[2] - b[2]) < 1.0e-8; }

void orc_gjkepa_pair(int32_t version, double tol_ff,
                     const double *p1, int32_t n1, const double *p2, int32_t n2,
                     int32_t mode, orc_result *out)
{
    static const double O[3] = { 0.0, 0.0, 0.0 };
    hull_t a = mk_hull(p1, n1), b = mk_hull(p2, n2);
    double s[12], last1[12], last2[12], dir[3], nd[3];
    int iter;
    memset(out, 0, sizeof *out);                                  /* F:65-71 */
    memset(s, 0, sizeof s);                                       /* H1: SAVEd simplex defined as zeros */

#define FINISH(path, coll) do { out->collision = (coll); out->status |= (uint32_t)(path); \
        memcpy(out->simplex, s, sizeof s); goto done; } while (0)

    if (!sphere_test(&a, &b)) FINISH(ORC_EXIT_SPHERE, 0);         /* F:76-77 */

    iter = 0;
    for (;;) {                                                    /* F:83-112 */
        ++iter;
        if (iter > 99) FINISH(ORC_EXIT_INIT_CAP, 0);
        orc_dir_table(iter, dir);
        support(&a, &b, dir, s + 0, 0, 0, &out->n_support);
        neg3(dir, nd);
        support(&a, &b, nd, s + 3, 0, 0, &out->n_support);
        if (all_close3(s + 0, s + 3)) continue;
        break;
    }

    vec_pl(O, s + 0, s + 3, dir);                                 /* F:116 */
    support(&a, &b, dir, s + 6, 0, 0, &out->n_support);
    if (all_close3(s + 6, s + 0) || all_close3(s + 6, s + 3)) FINISH(ORC_EXIT_DUP3, 0);   /* F:123 */

    {
        double v12[3], v23[3], c[3], vo[3];
        sub3(s + 3, s + 0, v12);
        sub3(s + 6, s + 3, v23);
        cross3(v12, v23, c);
        utzvec(c, dir);                                           /* F:135 */
        sub3(O, s + 6, vo);
        if (fabs(dot3(vo, dir)) < 1.0e-8) {                       /* F:140-148 */
            if (orc_is_inside_pf(s, 3, O)) {
                out->status |= ORC_ST_H1;
                if (mode >= ORC_MODE_EPA) epa_solve(version, tol_ff, &a, &b, s, mode, out);
                FINISH(ORC_EXIT_TRI_ORIGIN, 1);
            }
        }
        if (dot3(vo, dir) < 0.0) neg3(dir, dir);                  /* F:151 */
    }
    support(&a, &b, dir, s + 9, 0, 0, &out->n_support);           /* F:154 */
    {
        double d;
        if (dist_pf_sign(s + 9, s, &d)) { out->status |= ORC_ST_STOP; FINISH(ORC_EXIT_STOP, 0); }
        if (fabs(d) < 1.0e-8) FINISH(ORC_EXIT_COPLANAR4, 0);      /* F:157 */
    }
    if (orc_point_in_simplex(O, s)) {                             /* F:164 */
        if (mode >= ORC_MODE_EPA) epa_solve(version, tol_ff, &a, &b, s, mode, out);
        FINISH(ORC_EXIT_INIT_CONTAIN, 1);
    }

    memset(last1, 0, sizeof last1);                               /* F:178-179 */
    memset(last2, 0, sizeof last2);
    iter = 0;
    for (;;) {                                                    /* F:182-236 */
        ++iter;
        if (iter > 50) { out->status |= 50u << 8; FINISH(ORC_EXIT_LOOP_CAP, 0); }
        memcpy(last2, last1, sizeof last1);
        memcpy(last1, s, sizeof s);
        {
            double ns[12];
            update_simplex(&a, &b, s, ns, &out->n_support);
            memcpy(s, ns, sizeof s);
        }
        out->status = (out->status & ~0xff00u) | ((uint32_t)iter << 8);
        {
            double e1[3], e2[3], c[3], d;
            sub3(s + 3, s + 0, e1);
            sub3(s + 6, s + 3, e2);
            cross3(e1, e2, c);
            if (norm2_3(c) < 1.0e-8) FINISH(ORC_EXIT_LOOP_COLLINEAR, 0);            /* F:199 */
            if (dist_pf_sign(s + 9, s, &d)) { out->status |= ORC_ST_STOP; FINISH(ORC_EXIT_STOP, 0); }
            if (fabs(d) < 1.0e-8) FINISH(ORC_EXIT_LOOP_COPLANAR, 0);               /* F:203 */
        }
        if (orc_point_in_simplex(O, s)) {                         /* F:210 */
            if (mode >= ORC_MODE_EPA) epa_solve(version, tol_ff, &a, &b, s, mode, out);
            FINISH(ORC_EXIT_LOOP_CONTAIN, 1);
        }
        {
            int over = 0, i;                                      /* F:219-229 */
            for (i = 0; i < 4; ++i) {
                if (all_close3(s + 3 * i, last1 + 3 * i) || all_close3(s + 3 * i, last2 + 3 * i)) over = 1;
                else { over = 0; break; }
            }
            if (over) FINISH(ORC_EXIT_LOOP_STALL, 0);
        }
    }
#undef FINISH
done:
    return;
}

/* Every pair (a < b) of a hull store that passes F:1185, by testing ALL pairs (the checker for the product's
 * grid broad phase).  Spheres per hull first (F:1175-1182), then the F:1185 comparison per pair; pairs come out in
 * lexicographic order.  Returns the number of passing pairs; writes at most cap of them. */
int64_t orc_sphere_pairs(const double *verts, const int64_t *hull_off, int64_t n_hulls,
                         int64_t cap, int32_t *pair_a, int32_t *pair_b)
{
    double *sp = (double *)malloc((size_t)n_hulls * 4 * sizeof(double));
    int64_t *row = (int64_t *)calloc((size_t)n_hulls + 1, sizeof(int64_t));
    int64_t a, total = 0;
    if (!sp || !row) { free(sp); free(row); return -1; }
#pragma omp parallel for schedule(static)
    for (a = 0; a < n_hulls; ++a) {
        hull_t h = mk_hull(verts + 3 * hull_off[a], (int32_t)(hull_off[a + 1] - hull_off[a]));
        hull_sphere(&h, sp + 4 * a, sp + 4 * a + 3);
    }
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel for schedule(dynamic, 64)
        for (a = 0; a < n_hulls; ++a) {
            int64_t b, k = 0;
            for (b = a + 1; b < n_hulls; ++b) {
                double d[3];
                sub3(sp + 4 * a, sp + 4 * b, d);
                if (norm2_3(d) <= sp[4 * a + 3] + sp[4 * b + 3] + 1.0) {
                    if (pass == 1 && row[a] + k < cap) { pair_a[row[a] + k] = (int32_t)a; pair_b[row[a] + k] = (int32_t)b; }
                    ++k;
                }
            }
            if (pass == 0) row[a + 1] = k;
        }
        if (pass == 0) { for (a = 0; a < n_hulls; ++a) row[a + 1] += row[a]; total = row[n_hulls]; }
    }
    free(sp); free(row);
    return total;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_gjkepa_batch(int32_t version, double tol_ff,
                     const double *verts, const int64_t *hull_off,
                     int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b,
                     int32_t mode, int32_t n_threads,
                     uint8_t *collision, int32_t *colli_type, double *nearest,
                     double *normal, double *point, double *depth,
                     uint32_t *status, int32_t *n_support)
{
    int used = 1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    used = n_threads;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(n_threads)
#endif
    for (int64_t i = 0; i < n_pairs; ++i) {
        int32_t ha = pair_a[i], hb = pair_b[i];
        orc_result r;
        orc_gjkepa_pair(version, tol_ff,
                        verts + 3 * hull_off[ha], (int32_t)(hull_off[ha + 1] - hull_off[ha]),
                        verts + 3 * hull_off[hb], (int32_t)(hull_off[hb + 1] - hull_off[hb]),
                        mode, &r);
        if (collision)  collision[i] = (uint8_t)r.collision;
        if (colli_type) colli_type[i] = r.colli_type;
        if (nearest)    memcpy(nearest + 6 * i, r.nearest, sizeof r.nearest);
        if (normal)     memcpy(normal + 3 * i, r.normal, sizeof r.normal);
        if (point)      memcpy(point + 3 * i, r.point, sizeof r.point);
        if (depth)      depth[i] = r.depth;
        if (status)     status[i] = r.status;
        if (n_support)  n_support[i] = r.n_support;
    }
    return used;
}
