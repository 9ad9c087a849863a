// gjk_core.cuh -- GJK intersection test for one convex-hull pair, run by a sub-warp of L lanes.
//
// Path: GJKEPA driver (reference src/GCLIB_GJKEPA.f90:76-236, "F:<line>") with support_mapping
// (F:1030-1062), update_simplex_GJK (F:1070-1157), isPointInSimplex (F:1217-1265), IS_INSIDE_PF
// (F:1271-1337), UTZVEC/UNINML/DIST_PF_SIGN/VEC_PL (F:1343-1440).  This is not a translation:
//  * one pair is worked by L in {1,2,4} lanes; lane l owns tetrahedron faces l, l+L, ... and hull
//    vertices l, l+L, ...; the support function is a sub-warp (value,index) max-dot reduction;
//  * the four face normals are computed ONCE per simplex and shared by the containment test of
//    this walk step and the face pick of the next (the reference computes them twice), and the
//    degeneracy tests reuse face (1,2,3)'s cross product;
//  * the two opposite initial supports come from one pass (first arg-max and first arg-min).
// Flags must be bit-exact with the reference arithmetic, so every product/sum that feeds a
// comparison goes through mul()/add() (never contracted to FMA), divisions and square roots are
// IEEE, scans keep the first index on ties and reductions break ties towards the lower index.
//
// The L=1 instantiation is plain C++ and is also compiled by g++ (tests/host_sim) to check the
// shared logic against the oracle on a machine without a GPU.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define GJK_HD __host__ __device__ __forceinline__
#define GJK_HD_NOINLINE __host__ __device__ __noinline__
#else
#define GJK_HD inline
#define GJK_HD_NOINLINE
#endif

#include "fp64_exact.cuh"

namespace gjk {

// ---- exit paths (values = include/gjkepa.h GJKEPA_EXIT_*) ----
enum : int {
    EXIT_SPHERE = 1, EXIT_INIT_CAP = 2, EXIT_DUP3 = 3, EXIT_TRI_ORIGIN = 4, EXIT_COPLANAR4 = 5,
    EXIT_INIT_CONTAIN = 6, EXIT_LOOP_CAP = 7, EXIT_LOOP_COLLINEAR = 8, EXIT_LOOP_COPLANAR = 9,
    EXIT_LOOP_CONTAIN = 10, EXIT_LOOP_STALL = 11, EXIT_STOP = 12,
    EXIT_EARLY_OUT = 13          // not a reference exit: opt-in shortcut E1 below (a support value proves separation)
};
GJK_HD bool exit_is_collision(int e) { return e == EXIT_TRI_ORIGIN || e == EXIT_INIT_CONTAIN || e == EXIT_LOOP_CONTAIN; }

// ---- non-contracting arithmetic ----
GJK_HD double mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;            // host builds use -ffp-contract=off
#endif
}
GJK_HD double add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
GJK_HD double sub(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}

struct Vec3 { double x, y, z; };

GJK_HD Vec3 v3(double x, double y, double z) { Vec3 r; r.x = x; r.y = y; r.z = z; return r; }
GJK_HD Vec3 operator-(Vec3 a, Vec3 b) { return v3(sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)); }
// sign flip on the integer pipe: the FP64 pipe is the scarce unit (a DADD with -0 would do the same bits)
GJK_HD double negd(double x)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000u, __double2loint(x));
#else
    return -x;
#endif
}
GJK_HD Vec3 neg(Vec3 a) { return v3(negd(a.x), negd(a.y), negd(a.z)); }
GJK_HD double dot3(Vec3 a, Vec3 b)
{   // DOT_PRODUCT: ((a1*b1) + a2*b2) + a3*b3   (the leading 0+ only affects the sign of zero)
    return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z));
}
GJK_HD Vec3 cross3(Vec3 a, Vec3 b)
{   // F:1208-1210
    return v3(sub(mul(a.y, b.z), mul(a.z, b.y)),
              sub(mul(a.z, b.x), mul(a.x, b.z)),
              sub(mul(a.x, b.y), mul(a.y, b.x)));
}
GJK_HD double norm3(Vec3 v) { return sqrt(add(add(mul(v.x, v.x), mul(v.y, v.y)), mul(v.z, v.z))); }
GJK_HD Vec3 divs(Vec3 v, double s) { return v3(v.x / s, v.y / s, v.z / s); }
GJK_HD Vec3 utzvec_md(Vec3 v, double md) { return md < 1.0e-12 ? v3(0.0, 0.0, 0.0) : divs(v, md); }   // F:1349-1350
GJK_HD Vec3 utzvec(Vec3 v) { return utzvec_md(v, norm3(v)); }

// Two interchangeable arithmetic policies with bit-identical results: Exact uses the compiler's
// sqrt and '/', Fast the straight-line sequences of fp64_exact.cuh (shared reciprocal per vector,
// no control flow) and clears `ok` when an operand leaves their range -- the caller then redoes the
// block with Exact.
struct MathExact {
    static GJK_HD double nrm(Vec3 v, bool &) { return norm3(v); }
    static GJK_HD Vec3 dv(Vec3 v, double s, bool &) { return divs(v, s); }
    static GJK_HD void dv4(Vec3 v, double w, double s, Vec3 &qv, double &qw, bool &) { qv = divs(v, s); qw = w / s; }
};
// Same results again, for the rare redo after MathFast lowered `ok` (an exactly zero cross-product
// component of axis-aligned boxes, say): scalar out-of-line calls keep the call sites tiny so the
// cold code does not bloat the walk's instruction footprint.
GJK_HD_NOINLINE double cold_sqrt(double x) { return sqrt(x); }
GJK_HD_NOINLINE double cold_div(double a, double b) { return a / b; }
struct MathCold {
    static GJK_HD double nrm(Vec3 v, bool &) { return cold_sqrt(add(add(mul(v.x, v.x), mul(v.y, v.y)), mul(v.z, v.z))); }
    static GJK_HD Vec3 dv(Vec3 v, double s, bool &) { return v3(cold_div(v.x, s), cold_div(v.y, s), cold_div(v.z, s)); }
    static GJK_HD void dv4(Vec3 v, double w, double s, Vec3 &qv, double &qw, bool &)
    { qv = v3(cold_div(v.x, s), cold_div(v.y, s), cold_div(v.z, s)); qw = cold_div(w, s); }
};
struct MathFast {
    static GJK_HD double nrm(Vec3 v, bool &ok)
    { return sqrt_fast(add(add(mul(v.x, v.x), mul(v.y, v.y)), mul(v.z, v.z)), ok); }
    static GJK_HD Vec3 dv(Vec3 v, double s, bool &ok)
    {
        double r = rcp_div_seed(s);
        return v3(div_with_seed(v.x, s, r, ok), div_with_seed(v.y, s, r, ok), div_with_seed(v.z, s, r, ok));
    }
    static GJK_HD void dv4(Vec3 v, double w, double s, Vec3 &qv, double &qw, bool &ok)
    {
        double r = rcp_div_seed(s);
        qv = v3(div_with_seed(v.x, s, r, ok), div_with_seed(v.y, s, r, ok), div_with_seed(v.z, s, r, ok));
        qw = div_with_seed(w, s, r, ok);
    }
};
GJK_HD bool close3(Vec3 a, Vec3 b)
{ return fabs(sub(a.x, b.x)) < 1.0e-8 && fabs(sub(a.y, b.y)) < 1.0e-8 && fabs(sub(a.z, b.z)) < 1.0e-8; }
GJK_HD double nan_marker()
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}
GJK_HD Vec3 sel(bool c, Vec3 a, Vec3 b) { return v3(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z); }
GJK_HD Vec3 pick4(const Vec3 *s, int i)
{   // register-friendly s[i] for runtime i
    Vec3 r = s[0];
    r = sel(i == 1, s[1], r); r = sel(i == 2, s[2], r); r = sel(i == 3, s[3], r);
    return r;
}

// IS_INSIDE_PF for a triangle, F:1271-1337 (quirks Q2/Q3 kept).  Cold: only inside the 1e-8 bands.
GJK_HD_NOINLINE bool inside_pf3(Vec3 v0, Vec3 v1, Vec3 v2, Vec3 p)
{
    double c0 = sub(mul(sub(v1.x, v0.x), sub(p.y, v0.y)), mul(sub(v1.y, v0.y), sub(p.x, v0.x)));
    double c1 = sub(mul(sub(v2.x, v1.x), sub(p.y, v1.y)), mul(sub(v2.y, v1.y), sub(p.x, v1.x)));
    double c2 = sub(mul(sub(v0.x, v2.x), sub(p.y, v2.y)), mul(sub(v0.y, v2.y), sub(p.x, v2.x)));
    if (fabs(c0) < 1.0e-12) c0 = 0.0;
    if (fabs(c1) < 1.0e-12) c1 = 0.0;
    if (fabs(c2) < 1.0e-12) c2 = 0.0;
    if (!(c0 > 1.0e-15 || c1 > 1.0e-15 || c2 > 1.0e-15)) {        // F:1310: re-project on XZ
        c0 = sub(mul(sub(v1.x, v0.x), sub(p.z, v0.z)), mul(sub(v1.z, v0.z), sub(p.x, v0.x)));
        c1 = sub(mul(sub(v2.x, v1.x), sub(p.z, v1.z)), mul(sub(v2.z, v1.z), sub(p.x, v1.x)));
        c2 = sub(mul(sub(v0.x, v2.x), sub(p.z, v2.z)), mul(sub(v0.z, v2.z), sub(p.x, v2.x)));
    }
    if (mul(c0, c0) < 0.0) return false;
    if (mul(c0, c1) < 0.0) return false;
    if (mul(c0, c2) < 0.0) return false;
    return true;
}

// ---- hull access ----
// Global-memory view of a Fortran p(n,3) block [x(n) | y(n) | z(n)].
struct HullView {
    const double *x, *y, *z;     // y = x+n, z = x+2n
    int n;
    GJK_HD Vec3 get(int i) const
    {
#if defined(__CUDA_ARCH__)
        return v3(__ldg(x + i), __ldg(y + i), __ldg(z + i));
#else
        return v3(x[i], y[i], z[i]);
#endif
    }
    GJK_HD double comp(int ax, int i) const
    {
        const double *p = (ax == 0) ? x : (ax == 1 ? y : z);
#if defined(__CUDA_ARCH__)
        return __ldg(p + i);
#else
        return p[i];
#endif
    }
};

// The same block staged in shared memory, transposed across the CTA's pairs: element e of the block
// lives at base[e * stride] (stride = pairs per tile + 1, odd, so that consecutive threads hit
// consecutive 8-byte banks when they read the same element of their own pair).
template <int STRIDE>
struct HullViewStrided {
    const double *base;
    int n;
    GJK_HD Vec3 get(int i) const
    { return v3(base[i * STRIDE], base[(n + i) * STRIDE], base[(2 * n + i) * STRIDE]); }
};

// per-hull envelope sphere, F:1175-1182 (pair independent: computed once per hull by K0)
struct Sphere { double cx, cy, cz, r; };

// centroid = sums / REAL(n), F:1175.  For n a power of two the quotient is the product with 2^-k, bit for bit
// (both are the correctly rounded value of the same real number), which spares three IEEE divisions per hull.
GJK_HD void centroid_div(double sx, double sy, double sz, int n, Sphere &s)
{
    const double dn = (double)n;
    if ((n & (n - 1)) == 0 && n > 0) {
        const double inv = 1.0 / dn;                 // exact
        s.cx = mul(sx, inv); s.cy = mul(sy, inv); s.cz = mul(sz, inv);
    } else {
        s.cx = sx / dn; s.cy = sy / dn; s.cz = sz / dn;
    }
}

template <class HV>
GJK_HD Sphere hull_sphere(const HV &h)
{
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = 0; i < h.n; ++i) { Vec3 v = h.get(i); sx = add(sx, v.x); sy = add(sy, v.y); sz = add(sz, v.z); }
    Sphere s;
    centroid_div(sx, sy, sz, h.n, s);
    // MAXVAL(NORM2(v_i - c)): a correctly rounded sqrt is monotone, so the maximum of the norms is the
    // sqrt of the maximum squared norm, bit for bit -- one sqrt per hull instead of n
    double r2 = -1.0;
    Vec3 c = v3(s.cx, s.cy, s.cz);
    for (int i = 0; i < h.n; ++i) {
        Vec3 d = h.get(i) - c;
        double t = add(add(mul(d.x, d.x), mul(d.y, d.y)), mul(d.z, d.z));
        if (t > r2) r2 = t;
    }
    s.r = (h.n > 0) ? sqrt(r2) : -DBL_MAX;
    return s;
}
GJK_HD bool sphere_hit(const Sphere &a, const Sphere &b)
{   // F:1185: NORM2(mp1-mp2) <= r1 + r2 + 1.0
    return norm3(v3(a.cx, a.cy, a.cz) - v3(b.cx, b.cy, b.cz)) <= add(add(a.r, b.r), 1.0);
}

// ---- sub-warp of L lanes ----
template <int L> struct Lanes;

template <> struct Lanes<1> {
    static constexpr int LANES = 1;
    int lane;
    unsigned mask;
    GJK_HD Lanes() : lane(0), mask(0u) {}
    GJK_HD double bcast(double v, int) const { return v; }
    GJK_HD int bcast_i(int v, int) const { return v; }
    GJK_HD bool all(bool p) const { return p; }
    GJK_HD bool any(bool p) const { return p; }
    GJK_HD void argmax_first(double &, int &) const {}
    GJK_HD void argmin_first(double &, int &) const {}
};

#if defined(__CUDACC__)
template <int L> struct Lanes {
    static constexpr int LANES = L;
    int lane;            // lane within the group
    unsigned mask;       // warp mask of the group's lanes
    __device__ __forceinline__ Lanes()
    {
        int wl = threadIdx.x & 31;
        lane = wl & (L - 1);
        mask = (L == 32) ? 0xffffffffu : (((1u << (L & 31)) - 1u) << (wl & ~(L - 1)));
    }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(mask, v, src, L); }
    __device__ __forceinline__ int bcast_i(int v, int src) const { return __shfl_sync(mask, v, src, L); }
    __device__ __forceinline__ bool all(bool p) const { return (__ballot_sync(mask, p) & mask) == mask; }
    __device__ __forceinline__ bool any(bool p) const { return (__ballot_sync(mask, p) & mask) != 0u; }
    // larger value wins; equal values -> lower index (== first occurrence of the maximum)
    __device__ __forceinline__ void argmax_first(double &v, int &i) const
    {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(mask, v, o, L);
            int oi = __shfl_xor_sync(mask, i, o, L);
            bool take = (ov > v) || (ov == v && oi < i);
            v = take ? ov : v; i = take ? oi : i;
        }
    }
    __device__ __forceinline__ void argmin_first(double &v, int &i) const
    {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(mask, v, o, L);
            int oi = __shfl_xor_sync(mask, i, o, L);
            bool take = (ov < v) || (ov == v && oi < i);
            v = take ? ov : v; i = take ? oi : i;
        }
    }
};

// The same interface for a warp whose 32/L groups execute the SAME instructions (each on its own pair, bodies
// predicated): every collective takes the constant full mask, which spares the member-mask check nvcc wraps
// around each shuffle / vote with a run-time mask, and must be reached by the whole warp.
template <int L> struct LanesW {
    static constexpr int LANES = L;
    int lane, shift;
    __device__ __forceinline__ LanesW() { const int wl = threadIdx.x & 31; lane = wl & (L - 1); shift = wl & ~(L - 1); }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(0xffffffffu, v, src, L); }
    __device__ __forceinline__ int bcast_i(int v, int src) const { return __shfl_sync(0xffffffffu, v, src, L); }
    __device__ __forceinline__ unsigned bits(bool p) const { return (__ballot_sync(0xffffffffu, p) >> shift) & ((1u << L) - 1u); }
    __device__ __forceinline__ bool all(bool p) const { return bits(p) == ((1u << L) - 1u); }
    __device__ __forceinline__ bool any(bool p) const { return bits(p) != 0u; }
    __device__ __forceinline__ void argmax_first(double &v, int &i) const
    {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, v, o, L);
            int oi = __shfl_xor_sync(0xffffffffu, i, o, L);
            bool take = (ov > v) || (ov == v && oi < i);
            v = take ? ov : v; i = take ? oi : i;
        }
    }
    __device__ __forceinline__ void argmin_first(double &v, int &i) const
    {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, v, o, L);
            int oi = __shfl_xor_sync(0xffffffffu, i, o, L);
            bool take = (ov < v) || (ov == v && oi < i);
            v = take ? ov : v; i = take ? oi : i;
        }
    }
};
#endif

// Scans of one hull by lanes lane0, lane0+step, ...: first arg-max (and first arg-min) of d.v.  Views with
// a vectorised layout overload these (found by argument-dependent lookup at instantiation).
template <class HV>
GJK_HD void scan_argmax(const HV &h, Vec3 d, int lane0, int step, double &best, int &idx)
{
    best = -DBL_MAX; idx = 0;
    for (int v = lane0; v < h.n; v += step) { double t = dot3(d, h.get(v)); if (t > best) { best = t; idx = v; } }
}
template <class HV>
GJK_HD void scan_minmax(const HV &h, Vec3 d, int lane0, int step, double &hi, int &ihi, double &lo, int &ilo)
{
    hi = -DBL_MAX; lo = DBL_MAX; ihi = 0; ilo = 0;
    for (int v = lane0; v < h.n; v += step) {
        double t = dot3(d, h.get(v));
        if (t > hi) { hi = t; ihi = v; }
        if (t < lo) { lo = t; ilo = v; }
    }
}

// support_mapping, F:1030-1062: p1(first argmax d.p1) - p2(first argmax (-d).p2)
template <class G, class HV>
GJK_HD Vec3 support_i(const G &g, const HV &A, const HV &B, Vec3 d, int &iA, int &iB)
{
    constexpr int L = G::LANES;
    double bA, bB;
    scan_argmax(A, d, g.lane, L, bA, iA);
    scan_argmax(B, neg(d), g.lane, L, bB, iB);
    g.argmax_first(bA, iA);
    g.argmax_first(bB, iB);
    return A.get(iA) - B.get(iB);
}
template <class G, class HV>
GJK_HD Vec3 support(const G &g, const HV &A, const HV &B, Vec3 d)
{
    int iA, iB;
    return support_i(g, A, B, d, iA, iB);
}

// supports along +d and -d from one pass (F:97 and F:103): dot(-d,v) == -dot(d,v) bit for bit, so
// the first arg-max of dot(-d,.) is the first arg-min of dot(d,.).
// ia / ib: vertex numbers of the two rows in their hulls (row 0 in bits 0-7, row 1 in bits 8-15)
template <class G, class HV>
GJK_HD void support_both(const G &g, const HV &A, const HV &B, Vec3 d, Vec3 &sp, Vec3 &sm, unsigned &ia, unsigned &ib)
{
    constexpr int L = G::LANES;
    double hiA, loA, hiB, loB; int ihA, ilA, ihB, ilB;
    scan_minmax(A, d, g.lane, L, hiA, ihA, loA, ilA);
    scan_minmax(B, d, g.lane, L, hiB, ihB, loB, ilB);
    g.argmax_first(hiA, ihA); g.argmin_first(loA, ilA);
    g.argmax_first(hiB, ihB); g.argmin_first(loB, ilB);
    sp = A.get(ihA) - B.get(ilB);      // direction +d
    sm = A.get(ilA) - B.get(ihB);      // direction -d
    ia = (unsigned)ihA | ((unsigned)ilA << 8);
    ib = (unsigned)ilB | ((unsigned)ihB << 8);
}

#if defined(__CUDACC__)
__constant__ double c_dir_table[100][3] = {
#include "dirtable.inc"
};
#endif
static const double h_dir_table[100][3] = {
#include "dirtable.inc"
};
GJK_HD Vec3 dir_table(int row0)
{
#if defined(__CUDA_ARCH__)
    return v3(c_dir_table[row0][0], c_dir_table[row0][1], c_dir_table[row0][2]);
#else
    return v3(h_dir_table[row0][0], h_dir_table[row0][1], h_dir_table[row0][2]);
#endif
}

// faces (1,3,4) (1,2,4) (1,2,3) (2,3,4) of F:1093-1118 / idFc F:1227-1229, 0-based
GJK_HD void face_rows(int f, int &a, int &b, int &c)
{
    a = (f == 3) ? 1 : 0;
    b = (f == 0 || f == 3) ? 2 : 1;
    c = (f == 2) ? 2 : 3;
}

template <int L> struct Faces {
    static constexpr int FPL = 4 / L;
    Vec3 q[FPL];        // UTZVEC((Sa-Sb) x (Sb-Sc)), unflipped
    double md[FPL];     // |cross|
};

template <int L, class M, int SKIP>
GJK_HD bool faces_from_cross_impl(const Vec3 *cr, Faces<L> &F)
{
    bool ok = true;
#pragma unroll
    for (int j = 0; j < Faces<L>::FPL; ++j) if (j != SKIP) F.md[j] = M::nrm(cr[j], ok);
#pragma unroll
    for (int j = 0; j < Faces<L>::FPL; ++j) if (j != SKIP) {
        Vec3 q = M::dv(cr[j], F.md[j], ok);
        F.q[j] = (F.md[j] < 1.0e-12) ? v3(0.0, 0.0, 0.0) : q;       // UTZVEC, F:1349-1350
    }
    return ok;
}

GJK_HD Vec3 simplex_mean(const Vec3 *S)
{   // SUM(simplex_(:,i)) / 4, F:1086, F:1232
    return v3(add(add(add(S[0].x, S[1].x), S[2].x), S[3].x) / 4.0,
              add(add(add(S[0].y, S[1].y), S[2].y), S[3].y) / 4.0,
              add(add(add(S[0].z, S[1].z), S[2].z), S[3].z) / 4.0);
}

// tail of VEC_PL(O, (A, B)), F:1436-1438: D = A + (AC.AB / |AB|) * UTZVEC(AB); returns UTZVEC(D - O).
// With MathFast a NaN x-component tells the caller to redo the block with MathExact.
template <class M>
GJK_HD Vec3 vec_pl_tail(Vec3 A, Vec3 ab, double dab)
{
    bool ok = true;
    double nab = M::nrm(ab, ok);
    Vec3 u; double t;
    M::dv4(ab, dab, nab, u, t, ok);
    if (nab < 1.0e-12) u = v3(0.0, 0.0, 0.0);
    Vec3 D = v3(add(A.x, mul(t, u.x)), add(A.y, mul(t, u.y)), add(A.z, mul(t, u.z)));   // D - O = D
    double nd = M::nrm(D, ok);
    Vec3 r = M::dv(D, nd, ok);
    if (nd < 1.0e-12) r = v3(0.0, 0.0, 0.0);
    if (!ok) r.x = nan_marker();
    return r;
}

struct GjkOut {
    int exit_path;
    int iters;          // walk steps taken, F:182-236
    int n_support;
    bool stop_flag;     // reference STOP path (F:1369)
    int cold;           // times the straight-line sqrt/div block was redone with the IEEE operators
    int band;           // times IS_INSIDE_PF was entered (1e-8 bands)
};

// =====================================================================================================
// The flag path in two stages.
//
//   gjk_init     sphere test and the whole initial tetrahedron: line, triangle, fourth vertex and its
//                degenerate cases (F:76-157).  Every pair does the same amount of work here, so a warp of
//                32 pairs runs it in lockstep.
//   walk_round   one turn of the tetrahedron loop, cut where the loop is cheapest to leave:
//                    face normals of the current tetrahedron -> containment test  (F:164 / F:210, exit)
//                    -> stall / iteration cap (F:219-234, F:186)                   (exit)
//                    -> next face and direction (F:1093-1124) -> support (F:1127) -> new tetrahedron
//                    -> collinear / coplanar tests against the kept face (F:199-206, exit)
//                The containment test of the INITIAL tetrahedron and of every later one are the same
//                computation, so the whole walk has one support site and one normal-computation site
//                (the instruction cache, not the FP64 pipe, bounded the monolithic form), and the exit
//                separated pairs take -- F:203, ~all of them -- falls at the END of a round.
//
// State carried between rounds (struct Walk): the simplex, the unit normal and |cross| of its face
// (1,2,3) -- the face the last update kept, whose cross product has the operands of the new simplex's
// face (1,2,3), so it is never recomputed -- and what the stall test F:219-234 needs: simplex_last_1 is
// the simplex before the update; simplex_last_2 is rebuilt, in the rare case row 4 already matches, from
// the vertex the previous update dropped (D1) and the face it kept (k1).
// =====================================================================================================
enum : unsigned { WF_FIRST = 2u, WF_HAVE2 = 8u, WF_COLD = 16u, WF_STOP = 32u, WF_STALL = 64u, WF_EARLY = 128u };

// Shortcut E1 (opt-in, default off).  A support point s = support(d) along a unit direction d with
// s.d < -m proves the Minkowski difference lies in the half-space x.d < -m: the origin is outside, the hulls are at least
// m apart.  The reference can only answer TRUE with the origin inside -- or within its 1e-8 on-face band of -- a
// tetrahedron of Minkowski-difference points, and every other exit of it is FALSE, so it answers FALSE for such a pair
// whichever way its walk goes on; leaving here changes the exit path and the iteration count, not the flag.  The same
// argument makes the envelope-sphere test (F:76-77) redundant for the flag, so E1 drops it and its two passes over the hulls.  (Caveat kept
// from the survey: IS_INSIDE_PF tests the band on an XY projection, so a face within ~1e-8/m rad of vertical with the
// origin within 1e-8 of its plane could give a spurious TRUE the shortcut would not reproduce; m is 100x the band.)
constexpr double EARLY_OUT_MARGIN = 1.0e-6;
GJK_HD bool early_separated(Vec3 s, Vec3 d) { return dot3(s, d) < -EARLY_OUT_MARGIN; }

struct Walk {
    Vec3 S[4];          // simplex rows (same in every lane of the group)
    Vec3 qk;            // UTZVEC of face (1,2,3)'s cross product, unflipped
    double mk;          // |cross| of face (1,2,3)
    Vec3 D1;            // vertex dropped by the previous update
    int k1;             // face kept by the previous update
    int iters;          // reference `iter`
    int n_support;
    unsigned fl;
    unsigned ia, ib;    // after gjk_init: the vertex numbers behind the four rows, S[r] = A[(ia >> 8r) & 255] - B[(ib >> 8r) & 255]
                        // (hulls of at most 256 vertices; what the seed queue carries instead of the rows)
#if defined(GJK_SHADOW_HISTORY) && !defined(__CUDA_ARCH__)
    Vec3 sh1[4], sh2[4];    // test builds: simplex_last_1 / _2 kept explicitly, as the reference does (F:193-194)
#endif
};

// F:219-234 for rows 1..3 once row 4 has matched: p = simplex_last_1 (the simplex before the update that
// keeps face k and appends the support point), simplex_last_2 rebuilt from D1/k1.  Cold.
GJK_HD_NOINLINE bool stall_rows(Vec3 p0, Vec3 p1, Vec3 p2, Vec3 p3, int k, Vec3 D1, int k1, bool have2)
{
    const Vec3 O = v3(0.0, 0.0, 0.0);
    const Vec3 n0 = sel(k == 3, p1, p0), n1 = sel(k == 0 || k == 3, p2, p1), n2 = sel(k == 2, p2, p3);
    Vec3 l0 = O, l1 = O, l2 = O;                 // rows 0..2 of simplex_last_2 (zeros before the second iteration, F:178)
    if (have2) {                                 // previous update: last_2(a,b,c) -> last_1(0,1,2), last_2(d) dropped
        l0 = (k1 == 3) ? D1 : p0;                // row 0 is a of faces 0,1,2
        l1 = (k1 == 0) ? D1 : ((k1 == 3) ? p0 : p1);     // row 1: dropped by face 0; a of face 3; b of faces 1,2
        l2 = (k1 == 1) ? D1 : ((k1 == 2) ? p2 : p1);     // row 2: dropped by face 1; c of face 2; b of faces 0,3
    }
    return (close3(n0, p0) || close3(n0, l0)) && (close3(n1, p1) || close3(n1, l1)) && (close3(n2, p2) || close3(n2, l2));
}

// F:219-234 on the rows as they are before the update.  Row 4 is tested first: it is the new support
// point, which differs from the two before it unless the walk really cycles, so the other rows (and the
// rebuilt simplex_last_2) are almost never looked at.
GJK_HD bool stall_test(Vec3 p0, Vec3 p1, Vec3 p2, Vec3 p3, int k, Vec3 sp, Vec3 D1, int k1, bool have2)
{
    const Vec3 l23 = have2 ? sel(k1 == 2, D1, p2) : v3(0.0, 0.0, 0.0);
    if (!(close3(sp, p3) || close3(sp, l23))) return false;
    return stall_rows(p0, p1, p2, p3, k, D1, k1, have2);
}

#if defined(GJK_SHADOW_HISTORY) && !defined(__CUDA_ARCH__)
// Test builds: keep simplex_last_1/_2 as explicit arrays and compare, at every walk iteration, (a) the rows
// rebuilt from D1/k1 with simplex_last_2 and (b) the two-level stall test with the reference's four-row form.
static long long g_shadow_checks = 0, g_shadow_mismatch = 0;
inline bool same3(Vec3 a, Vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline void shadow_history_check(Walk &w, int k, Vec3 sp, bool have2)
{
    const Vec3 O = v3(0.0, 0.0, 0.0);
    if (w.fl & WF_FIRST) { for (int i = 0; i < 4; ++i) w.sh1[i] = w.sh2[i] = O; }     // F:178-179
    for (int i = 0; i < 4; ++i) { w.sh2[i] = w.sh1[i]; w.sh1[i] = w.S[i]; }           // F:193-194
    ++g_shadow_checks;
    const Vec3 p0 = w.S[0], p1 = w.S[1], p2 = w.S[2];
    Vec3 l0 = O, l1 = O, l2 = O, l3 = O;
    if (have2) {
        const int k1 = w.k1; const Vec3 D1 = w.D1;
        l0 = (k1 == 3) ? D1 : p0;
        l1 = (k1 == 0) ? D1 : ((k1 == 3) ? p0 : p1);
        l2 = (k1 == 1) ? D1 : ((k1 == 2) ? p2 : p1);
        l3 = (k1 == 2) ? D1 : p2;
    }
    if (!(same3(l0, w.sh2[0]) && same3(l1, w.sh2[1]) && same3(l2, w.sh2[2]) && same3(l3, w.sh2[3]))) ++g_shadow_mismatch;
    int a, b, c; face_rows(k, a, b, c);
    const Vec3 nw[4] = { w.S[a], w.S[b], w.S[c], sp };
    bool ref = true;
    for (int i = 0; i < 4; ++i) ref = ref && (close3(nw[i], w.sh1[i]) || close3(nw[i], w.sh2[i]));
    if (ref != stall_test(w.S[0], w.S[1], w.S[2], w.S[3], k, sp, w.D1, w.k1, have2)) ++g_shadow_mismatch;
}
#endif

// Initial tetrahedron.  Returns the exit path, or 0 with `w` ready for the first walk_round.
template <class G, class HV>
GJK_HD int gjk_init(const G &g, const HV &A, const HV &B, const Sphere &sa, const Sphere &sb, Walk &w, bool early = false)
{
    const Vec3 O = v3(0.0, 0.0, 0.0);
    w.S[0] = w.S[1] = w.S[2] = w.S[3] = O;
    w.qk = O; w.mk = 0.0; w.D1 = O; w.k1 = 2; w.iters = 0; w.n_support = 0; w.fl = WF_FIRST | (early ? WF_EARLY : 0u);
    w.ia = w.ib = 0u;

    // F:76-77.  With shortcut E1 the test is skipped (and the spheres need not be computed): the spheres enclose the hulls, so a
    // miss means the hulls are more than 1.0 apart, which the walk finds out on its own -- same flag, different exit path
    if (!early && !sphere_hit(sa, sb)) return EXIT_SPHERE;
    int it = 0;
    for (;;) {
        ++it;
        if (it > 99) return EXIT_INIT_CAP;                                    // F:86
        support_both(g, A, B, dir_table(it - 1), w.S[0], w.S[1], w.ia, w.ib);  // F:94-103
        w.n_support += 2;
        if (early && (early_separated(w.S[0], dir_table(it - 1)) || early_separated(w.S[1], neg(dir_table(it - 1))))) return EXIT_EARLY_OUT;
        if (!close3(w.S[0], w.S[1])) break;                                  // F:106
    }
    Vec3 dir;
    {   // VEC_PL(O, S0, S1), F:1431-1438 (quirk Q4: points from O to its foot on the line)
        Vec3 ab = w.S[1] - w.S[0], ac = O - w.S[0];
        double dab = dot3(ac, ab);
        dir = vec_pl_tail<MathFast>(w.S[0], ab, dab);
        if (dir.x != dir.x) { dir = vec_pl_tail<MathCold>(w.S[0], ab, dab); w.fl |= WF_COLD; }   // NaN marks "redo exactly"
    }
    { int iA, iB; w.S[2] = support_i(g, A, B, dir, iA, iB); w.ia |= (unsigned)iA << 16; w.ib |= (unsigned)iB << 16; }   // F:119
    ++w.n_support;
    if (early && early_separated(w.S[2], dir)) return EXIT_EARLY_OUT;
    if (close3(w.S[2], w.S[0]) || close3(w.S[2], w.S[1])) return EXIT_DUP3;  // F:123

    Vec3 c3 = cross3(w.S[1] - w.S[0], w.S[2] - w.S[1]);                      // F:132-135
    double md3; Vec3 q3;
    {
        bool ok = true;
        md3 = MathFast::nrm(c3, ok);
        q3 = MathFast::dv(c3, md3, ok);
        // cold redo with the IEEE operators; a NaN norm divides as IEEE does (c3 / NaN), only an exactly zero one is skipped
        if (!ok) { w.fl |= WF_COLD; bool dummy = true; md3 = MathCold::nrm(c3, dummy); q3 = (md3 == 0.0) ? O : MathCold::dv(c3, md3, dummy); }
    }
    dir = (md3 < 1.0e-12) ? O : q3;                                          // UTZVEC
    // the triangle is face (1,2,3) of the tetrahedron to come: (S1-S2)x(S2-S3) has the operands of c3 with
    // both factors negated, i.e. the same products -- its normal is not computed again
    w.qk = dir; w.mk = md3;
    {
        Vec3 vo = O - w.S[2];
        double dt = dot3(vo, dir);
        if (fabs(dt) < 1.0e-8 && inside_pf3(w.S[0], w.S[1], w.S[2], O)) return EXIT_TRI_ORIGIN;   // F:140
        if (dt < 0.0) dir = neg(dir);                                        // F:151
    }
    { int iA, iB; w.S[3] = support_i(g, A, B, dir, iA, iB); w.ia |= (unsigned)iA << 24; w.ib |= (unsigned)iB << 24; }   // F:154
    ++w.n_support;
    if (early && early_separated(w.S[3], dir)) return EXIT_EARLY_OUT;
    {   // DIST_PF_SIGN(S3, (S0,S1,S2)), F:157, with UNINML's own zero rule (F:1392)
        bool nz = fabs(c3.x) > 1.0e-12 || fabs(c3.y) > 1.0e-12 || fabs(c3.z) > 1.0e-12;
        Vec3 n = nz ? q3 : O;
        if (fabs(n.x) < 1.0e-12 && fabs(n.y) < 1.0e-12 && fabs(n.z) < 1.0e-12) { w.fl |= WF_STOP; return EXIT_STOP; }
        if (fabs(dot3(w.S[3] - w.S[0], n)) < 1.0e-8) return EXIT_COPLANAR4;
    }
    return 0;
}

// One turn of the tetrahedron loop in two halves (see the header of this section).
//
// Both halves are straight-line for the caller: neither skips work on an exit, they only report it.  A
// kernel that keeps several pairs per warp runs them on every lane, finished or not (`live` = false just
// gates the cold out-of-line paths and means "results ignored"), so no branch has to merge two versions
// of the walk state -- with early returns ptxas materialised the whole simplex in a second register set at
// every exit (30 % of the walk kernel's instructions were register moves, ncu).
struct Pick { int k; Vec3 qk; double mk; bool flip; };

// first half: face normals, containment, stall flag, iteration cap; chooses the next face.  Returns the exit
// path or 0; the simplex is not modified.
template <class G>
GJK_HD int walk_test(const G &g, Walk &w, Pick &pk, bool live)
{
    constexpr int L = G::LANES;
    constexpr int FPL = Faces<L>::FPL;
    const Vec3 O = v3(0.0, 0.0, 0.0);
    const bool first = (w.fl & WF_FIRST) != 0u;

    Faces<L> F;
    {
        Vec3 cr[FPL];
#pragma unroll
        for (int j = 0; j < FPL; ++j) if (!(L == 1 && j == 2)) {
            int f = g.lane + j * L, a, b, c;
            face_rows(f, a, b, c);
            Vec3 sa = pick4(w.S, a), sb = pick4(w.S, b), sc = pick4(w.S, c);
            cr[j] = cross3(sa - sb, sb - sc);
        }
        if (L == 1) { F.q[2 % FPL] = w.qk; F.md[2 % FPL] = w.mk; }     // face (1,2,3): kept by the last update, same operands
        const bool ok = faces_from_cross_impl<L, MathFast, (L == 1 ? 2 : -1)>(cr, F);
        if (!ok && live) { faces_from_cross_impl<L, MathCold, (L == 1 ? 2 : -1)>(cr, F); w.fl |= WF_COLD; }
    }
    const Vec3 M = simplex_mean(w.S);

    // containment (isPointInSimplex F:1236-1262) and the next face (update_simplex_GJK F:1093-1124) from the
    // same normals: a flipped normal only changes the sign of its dot products, bit for bit
    bool band_hit = false, all_pos = true;
    double wbest = 0.0; int fbest = 0; Vec3 qbest = O; double mbest = 0.0; bool flbest = false;
#pragma unroll
    for (int j = 0; j < FPL; ++j) {
        const int f = g.lane + j * L;
        const Vec3 q = F.q[j];
        const Vec3 refc = pick4(w.S, f);                                    // simplex_(i,:), F:1240
        const double sc = dot3(q, refc - M);
        const double uc = dot3(refc, q);
        const double dist = (sc < 0.0) ? negd(uc) : uc;                     // (simplex_(i,:) - O) . nml(i,:)
        all_pos = all_pos && (dist > 0.0);
        if (live && fabs(uc) < 1.0e-8) {                                    // F:1248
            int a, b, c;
            face_rows(f, a, b, c);
            if (inside_pf3(pick4(w.S, a), pick4(w.S, b), pick4(w.S, c), O)) band_hit = true;
        }
        double sk, uk;
        if (L == 1 && j == 0) { sk = sc; uk = uc; }                         // face (1,3,4): both tests refer to row 1
        else {
            const Vec3 refp = (f == 3) ? w.S[1] : w.S[0];
            sk = dot3(q, refp - M);
            uk = dot3(q, refp);
        }
        const bool flp = sk < 0.0;
        double wv = flp ? uk : negd(uk);                                    // DOT(-nml, simplex_(ref,:) - O)
        // MAXLOC split over lanes must stay the sequential rule (face 0 seeds, later faces need a strict >) when a value is
        // NaN, or the lanes would disagree on the winner: a NaN on face 0 wins outright, a NaN elsewhere never does
        if (L > 1 && wv != wv) wv = (f == 0) ? HUGE_VAL : -HUGE_VAL;
        if (j == 0 || wv > wbest) { wbest = wv; fbest = f; qbest = q; mbest = F.md[j]; flbest = flp; }   // MAXLOC: first maximum
    }
    g.argmax_first(wbest, fbest);
    const int owner = fbest % L;
    pk.k = fbest;
    pk.qk = v3(g.bcast(qbest.x, owner), g.bcast(qbest.y, owner), g.bcast(qbest.z, owner));
    pk.mk = g.bcast(mbest, owner);
    pk.flip = g.bcast_i(flbest ? 1 : 0, owner) != 0;

    // both votes are taken unconditionally: with LanesW every lane of the warp must reach every collective
    const bool on_band = g.any(band_hit), inside = g.all(all_pos);
    int e = 0;
    if (on_band || inside) e = first ? EXIT_INIT_CONTAIN : EXIT_LOOP_CONTAIN;                   // F:164 / F:210
    else if (w.fl & WF_STALL) e = EXIT_LOOP_STALL;                          // F:219-234, decided when the simplex was updated
    else if (w.iters + 1 > 50) { w.iters = 50; e = EXIT_LOOP_CAP; }         // F:186
    else ++w.iters;
    return e;
}

// F:203 calls DIST_PF_SIGN, whose UNINML (F:1392) returns the zero vector unless some |cross component| > 1e-12, and a
// zero normal is the reference's READ/STOP path (F:1369).  With finite vertices that cannot follow a norm >= 1e-8; with
// Inf/NaN vertices it can (norm NaN passes F:199, every component NaN fails F:1392).  Cold: only a NaN norm gets here.
GJK_HD_NOINLINE bool walk_stop_nan(Vec3 s0, Vec3 s1, Vec3 s2)
{
    const Vec3 c = cross3(s1 - s0, s2 - s1);
    return !(fabs(c.x) > 1.0e-12 || fabs(c.y) > 1.0e-12 || fabs(c.z) > 1.0e-12);
}

// second half: support along the chosen face's normal, new tetrahedron (in place), degenerate cases.
template <class G, class HV>
GJK_HD int walk_step(const G &g, const HV &A, const HV &B, Walk &w, const Pick &pk, bool live)
{
    const int k = pk.k;
    const Vec3 dirk = pk.flip ? neg(pk.qk) : pk.qk;
    const Vec3 sp = support(g, A, B, dirk);                                 // F:1127
    const bool sep = (w.fl & WF_EARLY) != 0u && early_separated(sp, dirk);
    ++w.n_support;
    const bool have2 = (w.fl & WF_HAVE2) != 0u;
#if defined(GJK_SHADOW_HISTORY) && !defined(__CUDA_ARCH__)
    shadow_history_check(w, k, sp, have2);
#endif
    const bool stall = live && stall_test(w.S[0], w.S[1], w.S[2], w.S[3], k, sp, w.D1, w.k1, have2);
    w.D1 = sel(k == 0, w.S[1], sel(k == 1, w.S[2], sel(k == 2, w.S[3], w.S[0])));
    w.k1 = k;
    // F:1130-1150 in place: rows (a,b,c) of the kept face move up, the support point becomes row 4
    w.S[0] = sel(k == 3, w.S[1], w.S[0]);
    w.S[1] = sel(k == 0 || k == 3, w.S[2], w.S[1]);
    w.S[2] = sel(k != 2, w.S[3], w.S[2]);
    w.S[3] = sp;
    w.qk = pk.qk; w.mk = pk.mk;
    w.fl = (w.fl & (WF_COLD | WF_STOP | WF_EARLY)) | WF_HAVE2 | (stall ? WF_STALL : 0u);   // from now on simplex_last_2 exists
    // degenerate cases against the kept face, F:199-206
    int e = 0;
    if (pk.mk < 1.0e-8) e = EXIT_LOOP_COLLINEAR;
    else if (pk.mk != pk.mk && live && walk_stop_nan(w.S[0], w.S[1], w.S[2])) { w.fl |= WF_STOP; e = EXIT_STOP; }
    else if (fabs(dot3(sp - w.S[0], pk.qk)) < 1.0e-8) e = EXIT_LOOP_COPLANAR;
    else if (sep) e = EXIT_EARLY_OUT;                                       // reference exits first: same status when both apply
    return e;
}

template <class G, class HV>
GJK_HD int walk_round(const G &g, const HV &A, const HV &B, Walk &w)
{
    Pick pk;
    const int e = walk_test(g, w, pk, true);
    if (e) return e;
    return walk_step(g, A, B, w, pk, true);
}

// The whole GJKEPA flag path for one pair.  S: 4 rows, final simplex on return (row 3 zero on the
// F:140 path, hazard H1).  All lanes of the group return the same values.
template <class G, class HV>
GJK_HD GjkOut gjk_pair(const G &g, const HV &A, const HV &B,
                       const Sphere &sa, const Sphere &sb, Vec3 *S, bool early = false)
{
    Walk w;
    int e = gjk_init(g, A, B, sa, sb, w, early);
    while (e == 0) e = walk_round(g, A, B, w);
    GjkOut o;
    o.exit_path = e; o.iters = w.iters; o.n_support = w.n_support;
    o.stop_flag = (w.fl & WF_STOP) != 0u; o.cold = (w.fl & WF_COLD) ? 1 : 0; o.band = 0;
    S[0] = w.S[0]; S[1] = w.S[1]; S[2] = w.S[2]; S[3] = w.S[3];
    return o;
}

}  // namespace gjk
