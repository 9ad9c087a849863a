// epa_kernel.cuh -- K3: EPA penetration depth + contact normal (and the contact epilogue K4) for the
// pairs GJK reported as colliding.  One warp per pair; polytope AND both hulls resident in shared
// memory.
//
// Path: EPA_solu (reference src/GCLIB_GJKEPA.f90:242-346) and update_expandingPolytope_EPA
// (F:863-1022).  The reference rebuilds the whole polytope from scratch with an external QuickHull
// on every iteration (F:920-950) and decides convergence by comparing bubble-sorted face-distance
// lists (F:972-1015).  Here the polytope is kept incrementally:
//   per iteration  closest face (warp arg-min over cached plane distances, F:893-898)
//                  -> outward direction with the reference's orientation rule (F:902-910)
//                  -> support point (32-lane max-dot reduction over the staged hulls, F:914)
//                  -> faces that see the point by more than HULL_EPS are marked in parallel,
//                     the horizon is the set of their edges whose twin is not marked, marked faces
//                     are compacted away and one new face per horizon edge is fanned to the point;
//   converged      when the support point(s) do not change the hull -- the geometric meaning of
//                  the reference test -- reporting the closest face's distance and outward normal
//                  (F:963-969, F:997-998).
// Plane normals/distances use the reference forms (UNINML F:1389-1392, DIST_PF_SIGN F:1375) so that
// depth and normal agree with the CPU path to rounding.
//
// Warp arg-max / arg-min: the (value, index) butterfly (5 steps x 3 shuffles + compares) was 23 % of
// this kernel's instructions (ncu source view, profiles/r1).  Doubles are mapped to order-preserving
// 64-bit keys and reduced with three REDUX instructions (max high word, max low word among the
// leaders, min index among the ties) -- first-index-wins semantics kept, -0 canonicalised.
//
// Capacity: the kernel is instantiated for (V_CAP, F_CAP) = (64, 128), which covers every pair of the
// benchmark configurations with 20 resident warps per SM, and for (128, 256) = the reference's hard
// bound (4 + 99 iterations); a pair that outgrows the small polytope is re-queued to the large one.
#pragma once
#include "gjk_kernels.cuh"
#include <cuda_runtime.h>

namespace epa {

using gjk::Vec3; using gjk::v3; using gjk::dot3; using gjk::cross3; using gjk::norm3; using gjk::neg;
using gjk::HullView; using gjk::Lanes;

constexpr int    WARPS = 4;
constexpr int    HULL_CAP = 128;       // staged hull vertices per pair (nA + nB); larger pairs read global memory
constexpr double HULL_EPS = 1.0e-10;   // same "swallow" tolerance as the oracle's hull stand-in
constexpr int    MAX_ITERS = 99;       // F:299

constexpr uint32_t ST_EPA_CAP = 1u << 24, ST_STOP = 1u << 25, ST_HULL_FAIL = 1u << 26,
                   ST_BAD_VERSION = 1u << 27, ST_TOPO_WARN = 1u << 29;

struct EpaArgs {
    const double *verts; const int64_t *hull_off;
    const int32_t *pair_a, *pair_b;
    const int32_t *wl_pair; const double *wl_simplex;
    const unsigned long long *wl_count; unsigned long long *queue_head;
    // optional indirection over work-list entries (the large-capacity pass runs the overflow list)
    const int32_t *item_list; const unsigned long long *item_count;
    int32_t *ovf_items; unsigned long long *ovf_count;
    int32_t version; double tol_ff; bool full;
    int32_t *colli_type; double *nearest, *normal, *point, *depth; uint32_t *status;
    double *result;                 // optional: normal xyz, depth (-1: no result) in the first 4 doubles of the entry's 12-double
                                    // work-list slot (dead once the pair is retired); input of the contact kernel
    // un-routed passes of the linked-mesh kernel over the whole work list (small batches, where two routing launches would cost
    // more than they save): a pass works the items of ITS size class only -- take_gt < nA + nB, and with skip_large nothing
    // beyond its staging buffer -- and leaves the others to the pass that owns them, so a pair meets the same kernel as in a
    // routed batch
    int32_t take_gt = -1; bool skip_large = false;
};

template <int VC, int FC>
struct alignas(16) Polytope {
    double vx[VC], vy[VC], vz[VC];
    double nx[FC], ny[FC], nz[FC], d[FC];     // outward unit normal, n.x = d on the face
    double hx[HULL_CAP], hy[HULL_CAP], hz[HULL_CAP];   // staged hulls: A at [0,nA), B at [nA,nA+nB)
    uint32_t fw[FC];                          // face vertices a | b<<8 | c<<16
    uint16_t hedge[FC];                       // horizon edges a | b<<8 (directed)
    uint8_t vis[FC];
    uint8_t vlist[FC];                        // indices of visible faces
    int nv, nf;
};

// generic-pointer hull view (shared or global memory)
struct HullPlain {
    const double *x, *y, *z; int n;
    __device__ __forceinline__ Vec3 get(int i) const { return v3(x[i], y[i], z[i]); }
};

// ---- warp arg-reductions with REDUX; first index wins ties ----
__device__ __forceinline__ int warp_argmax_first(double v, int i)
{
    const unsigned FULL = 0xffffffffu;
    long long b = __double_as_longlong(__dadd_rn(v, 0.0));               // -0 -> +0
    unsigned long long k = (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
    unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    unsigned mh = __reduce_max_sync(FULL, hi);
    unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    return __reduce_min_sync(FULL, (hi == mh && lo == ml) ? i : 0x7fffffff);
}
__device__ __forceinline__ int warp_argmin_first_nonneg(double v, int i)   // v >= 0 (or +inf)
{
    const unsigned FULL = 0xffffffffu;
    unsigned long long k = (unsigned long long)__double_as_longlong(v);
    unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    unsigned mh = __reduce_min_sync(FULL, hi);
    unsigned ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
    return __reduce_min_sync(FULL, (hi == mh && lo == ml) ? i : 0x7fffffff);
}

// support_mapping (F:1030-1062) by the whole warp
__device__ __forceinline__ Vec3 warp_support(const HullPlain &A, const HullPlain &B, Vec3 d, int lane)
{
    double bA = -DBL_MAX, bB = -DBL_MAX; int iA = 0, iB = 0;
    for (int v = lane; v < A.n; v += 32) { double t = dot3(d, A.get(v)); if (t > bA) { bA = t; iA = v; } }
    Vec3 nd = neg(d);
    for (int v = lane; v < B.n; v += 32) { double t = dot3(nd, B.get(v)); if (t > bB) { bB = t; iB = v; } }
    iA = warp_argmax_first(bA, iA);
    iB = warp_argmax_first(bB, iB);
    return A.get(iA) - B.get(iB);
}

template <int VC, int FC>
__device__ __forceinline__ Vec3 pv(const Polytope<VC, FC> &P, int i) { return v3(P.vx[i], P.vy[i], P.vz[i]); }

// plane of face (a,b,c): UNINML (F:1389-1392) and its signed offset; false if degenerate
template <int VC, int FC>
__device__ __forceinline__ bool face_plane(const Polytope<VC, FC> &P, int a, int b, int c, Vec3 &n, double &d)
{
    Vec3 va = pv(P, a), vb = pv(P, b), vc = pv(P, c);
    Vec3 cr = cross3(vb - va, vc - vb);
    if (!(fabs(cr.x) > 1.0e-12 || fabs(cr.y) > 1.0e-12 || fabs(cr.z) > 1.0e-12)) { n = v3(0, 0, 0); d = 0; return false; }
    bool ok = true;
    double md = gjk::MathFast::nrm(cr, ok);
    n = gjk::MathFast::dv(cr, md, ok);
    if (!ok) { bool dummy = true; md = gjk::MathCold::nrm(cr, dummy); n = gjk::MathCold::dv(cr, md, dummy); }
    d = dot3(va, n);
    return true;
}

// Insert point q.  Returns 1 if the hull changed, 0 if not, -1 on failure, -2 when out of capacity.
template <int VC, int FC>
__device__ int insert_point(Polytope<VC, FC> &P, Vec3 q, int lane, uint32_t &st)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const int nf0 = P.nf, nv0 = P.nv;
    // (a support point that repeats an existing vertex needs no special case: it lies on the hull, so no
    //  face sees it by more than HULL_EPS and the visibility pass below reports "unchanged")
    // visible faces
    int nvis = 0;
    for (int c = 0; c < nf0; c += 32) {
        int f = c + lane;
        bool v = false;
        if (f < nf0) {
            double dist = gjk::sub(dot3(v3(P.nx[f], P.ny[f], P.nz[f]), q), P.d[f]);
            v = dist > HULL_EPS;
            P.vis[f] = v ? 1 : 0;
        }
        unsigned b = __ballot_sync(FULL, v);
        if (v) P.vlist[nvis + __popc(b & lt)] = (uint8_t)f;
        nvis += __popc(b);
    }
    if (nvis == 0) return 0;
    if (nv0 >= VC) return -2;
    __syncwarp();

    // horizon: directed edges of visible faces whose reverse is not an edge of a visible face
    int nh = 0;
    for (int c = 0; c < 3 * nvis; c += 32) {
        int it = c + lane;
        bool hz = false; unsigned ea = 0, eb = 0;
        if (it < 3 * nvis) {
            unsigned w = P.fw[P.vlist[it / 3]]; int e = it % 3;
            unsigned a0 = w & 0xffu, a1 = (w >> 8) & 0xffu, a2 = (w >> 16) & 0xffu;
            ea = e == 0 ? a0 : (e == 1 ? a1 : a2);
            eb = e == 0 ? a1 : (e == 1 ? a2 : a0);
            const unsigned rev01 = eb | (ea << 8);                       // (g0,g1) == (eb,ea)
            hz = true;
            for (int j = 0; j < nvis; ++j) {
                unsigned gw = P.fw[P.vlist[j]];
                unsigned e01 = gw & 0xffffu, e12 = (gw >> 8) & 0xffffu, e20 = ((gw >> 16) & 0xffu) | ((gw & 0xffu) << 8);
                if (e01 == rev01 || e12 == rev01 || e20 == rev01) { hz = false; break; }
            }
        }
        unsigned b = __ballot_sync(FULL, hz);
        if (hz) {
            int pos = nh + __popc(b & lt);
            if (pos < FC) P.hedge[pos] = (uint16_t)(ea | (eb << 8));
        }
        nh += __popc(b);
    }
    if (nh != nvis + 2) st |= ST_TOPO_WARN;         // visible region is not a disk
    if (nf0 - nvis + nh > FC || nh > FC) return -2;
    __syncwarp();

    // drop visible faces (stable, in place)
    int base = 0;
    for (int c = 0; c < nf0; c += 32) {
        int f = c + lane;
        bool keep = f < nf0 && !P.vis[f];
        double nx = 0, ny = 0, nz = 0, dd = 0; unsigned w = 0;
        if (keep) { nx = P.nx[f]; ny = P.ny[f]; nz = P.nz[f]; dd = P.d[f]; w = P.fw[f]; }
        unsigned b = __ballot_sync(FULL, keep);
        __syncwarp();
        if (keep) {
            int pos = base + __popc(b & lt);
            P.nx[pos] = nx; P.ny[pos] = ny; P.nz[pos] = nz; P.d[pos] = dd; P.fw[pos] = w;
        }
        base += __popc(b);
        __syncwarp();
    }

    // new vertex and the fan of new faces
    const int vid = nv0;
    if (lane == 0) { P.vx[vid] = q.x; P.vy[vid] = q.y; P.vz[vid] = q.z; }
    __syncwarp();
    bool bad = false;
    for (int e = lane; e < nh; e += 32) {
        unsigned h = P.hedge[e];
        int a = h & 0xffu, b = (h >> 8) & 0xffu, f = base + e;
        Vec3 n; double d;
        if (!face_plane(P, a, b, vid, n, d)) bad = true;
        P.nx[f] = n.x; P.ny[f] = n.y; P.nz[f] = n.z; P.d[f] = d;
        P.fw[f] = (unsigned)a | ((unsigned)b << 8) | ((unsigned)vid << 16);
    }
    if (lane == 0) { P.nv = vid + 1; P.nf = base + nh; }
    __syncwarp();
    if (__any_sync(FULL, bad)) { st |= ST_STOP; return -1; }      // DIST_PF_SIGN's STOP path (F:1369)
    return 1;
}

// ---- contact epilogue (K4), F:326-343 ----
// The order-dependent scans of the reference (quirk Q8) stay sequential, but over dot products that the
// warp computed in parallel into shared scratch (tA[i] = n.p1(i), tB[i] = (-n).p2(i)); lane 0 walks hull
// A while lane 1 walks hull B.
__device__ __forceinline__ void near_max_pair(const double *t, int n, int idx[2])
{   // F:722-747 tracker: "> max - 1e-8", last wins, previous winner becomes second
    double mx = -DBL_MAX;
    idx[0] = idx[1] = 0;
    for (int i = 0; i < n; ++i) {
        double v = t[i];
        if (v > mx - 1.0e-8) { mx = v; idx[1] = idx[0]; idx[0] = i + 1; }
    }
    if (idx[1] == 0) idx[1] = idx[0];
}

__device__ void contact_epilogue(const EpaArgs &a, int p, const HullPlain &A, const HullPlain &B,
                                 Vec3 nrm, int lane, uint32_t &st, double *tA, double *tB)
{
    const unsigned FULL = 0xffffffffu;
    const Vec3 nd = neg(nrm);
    // dots along +-normal, once
    double bA = -DBL_MAX, bB = -DBL_MAX; int iA = 0, iB = 0;
    for (int v = lane; v < A.n; v += 32) { double t = dot3(nrm, A.get(v)); tA[v] = t; if (t > bA) { bA = t; iA = v; } }
    for (int v = lane; v < B.n; v += 32) { double t = dot3(nd, B.get(v)); tB[v] = t; if (t > bB) { bB = t; iB = v; } }
    iA = warp_argmax_first(bA, iA); iB = warp_argmax_first(bB, iB);
    __syncwarp();
    // get_nearest_points, F:813-855: the two support vertices along +-normal
    if (lane == 0 && a.nearest) {
        Vec3 pa = A.get(iA), pb = B.get(iB);
        double *o = a.nearest + 6 * (int64_t)p;
        o[0] = pa.x; o[1] = pb.x; o[2] = pa.y; o[3] = pb.y; o[4] = pa.z; o[5] = pb.z;
    }
    Vec3 nrm_out = nrm;
    if (a.version == 1) {                       // get_collisionPoint_01, F:700-806
        int s[2] = {0, 0};
        if (lane == 0) near_max_pair(tA, A.n, s);
        if (lane == 1) near_max_pair(tB, B.n, s);
        int a0 = __shfl_sync(FULL, s[0], 0), a1 = __shfl_sync(FULL, s[1], 0);
        int b0 = __shfl_sync(FULL, s[0], 1), b1 = __shfl_sync(FULL, s[1], 1);
        bool m1 = a0 != a1, m2 = b0 != b1;
        if (lane == 0 && a.point) {
            Vec3 r = v3(0, 0, 0);
            if (!m1 && !m2) {
                Vec3 u = A.get(a0 - 1), w = B.get(b0 - 1);
                r = v3(gjk::add(u.x, w.x) / 2.0, gjk::add(u.y, w.y) / 2.0, gjk::add(u.z, w.z) / 2.0);
            } else if (m1 && !m2) r = B.get(b0 - 1);
            else if (!m1 && m2) r = A.get(a0 - 1);
            else {                              // F:766-804: mean of p1 supports within 0.1 of the maximum
                const double mx = tA[iA];       // strict maximum of the same dot products (F:779-787)
                double sx = 0, sy = 0, sz = 0; int c = 0;
                for (int i = 0; i < A.n; ++i) {
                    if (tA[i] > mx - 1.0e-1) { Vec3 v = A.get(i); ++c; sx = gjk::add(sx, v.x); sy = gjk::add(sy, v.y); sz = gjk::add(sz, v.z); }
                }
                r = v3(sx / (double)c, sy / (double)c, sz / (double)c);
            }
            double *o = a.point + 3 * (int64_t)p;
            o[0] = r.x; o[1] = r.y; o[2] = r.z;
        }
    } else if (a.version == 3) {                // get_collisionPoint_03, F:426-452
        if (lane == 0) {
            double mx = -DBL_MAX; int idx = 0;
            for (int i = 0; i < B.n; ++i) { double t = tB[i]; if (t > mx - 1.0e-8) { mx = t; idx = i; } }
            Vec3 r = B.get(idx);
            double s = 0.0;
            for (int i = 0; i < A.n; ++i) s = gjk::add(s, A.z[i]);
            r.z = s / (double)(float)A.n;
            if (a.point) { double *o = a.point + 3 * (int64_t)p; o[0] = r.x; o[1] = r.y; o[2] = r.z; }
        }
        Vec3 nn = v3(nrm.x, nrm.y, 0.0);
        double md = norm3(nn);
        nrm_out = v3(nn.x / md, nn.y / md, nn.z / md);           // NaN when the normal is parallel to z, as the reference
        if (lane == 0 && a.normal) { double *o = a.normal + 3 * (int64_t)p; o[0] = nrm_out.x; o[1] = nrm_out.y; o[2] = nrm_out.z; }
    } else {
        st |= ST_BAD_VERSION;                   // version 2 needs the un-vendored GCLIB_List; others STOP (F:336)
    }
    // get_info_collisionType, F:353-413 (on the possibly re-projected normal, F:335-343)
    {
        const bool same = (a.version != 3);     // version 1: the dots above are the ones F:372/391 compute
        double mA = -DBL_MAX, mB = -DBL_MAX;
        const Vec3 nd2 = neg(nrm_out);
        __syncwarp();
        if (!same) {
            for (int v = lane; v < A.n; v += 32) tA[v] = dot3(nrm_out, A.get(v));
            for (int v = lane; v < B.n; v += 32) tB[v] = dot3(nd2, B.get(v));
            __syncwarp();
        }
        for (int v = lane; v < A.n; v += 32) { double t = tA[v]; if (t > mA) mA = t; }
        for (int v = lane; v < B.n; v += 32) { double t = tB[v]; if (t > mB) mB = t; }
        for (int o = 16; o > 0; o >>= 1) {
            double t = __shfl_xor_sync(FULL, mA, o); if (t > mA) mA = t;
            t = __shfl_xor_sync(FULL, mB, o); if (t > mB) mB = t;
        }
        int cA = 0, cB = 0;
        for (int v = lane; v < A.n; v += 32) cA += tA[v] > gjk::sub(mA, a.tol_ff);
        for (int v = lane; v < B.n; v += 32) cB += tB[v] > gjk::sub(mB, a.tol_ff);
        cA = __reduce_add_sync(FULL, cA); cB = __reduce_add_sync(FULL, cB);
        if (lane == 0 && a.colli_type) a.colli_type[p] = (cA >= 3 && cB >= 3) ? 2 : 1;
    }
}

#ifndef EPA_MIN_CTAS
#define EPA_MIN_CTAS 5
#endif
template <int VC, int FC>
__global__ void __launch_bounds__(WARPS * 32, EPA_MIN_CTAS)
epa_kernel(EpaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Polytope<VC, FC> &P = reinterpret_cast<Polytope<VC, FC> *>(smem_raw)[warp];
    const unsigned FULL = 0xffffffffu;
    const unsigned long long n_items = a.item_list ? *a.item_count : *a.wl_count;

    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.queue_head, 1ull);
        q = __shfl_sync(FULL, q, 0);
        if (q >= n_items) break;
        const long long item = a.item_list ? (long long)a.item_list[q] : (long long)q;

        const int p = a.wl_pair[item];
        const HullView GA = gjk::hull_view(a.verts, a.hull_off, a.pair_a[p]);
        const HullView GB = gjk::hull_view(a.verts, a.hull_off, a.pair_b[p]);
        uint32_t st = 0;
        int iters = 0;
        double depth = 0.0; Vec3 nrm = v3(0, 0, 0);
        bool ok = true, overflow = false;
        if (a.full && a.ovf_items && (GA.n > FC || GB.n > FC)) {       // epilogue scratch too small: large pass
            if (lane == 0) a.ovf_items[atomicAdd(a.ovf_count, 1ull)] = (int32_t)item;
            continue;
        }

        __syncwarp();                                   // previous item's shared state is dead
        // stage both hulls (coalesced: the block [x|y|z] is contiguous) unless they exceed the buffer
        HullPlain A, B;
        const bool staged = GA.n + GB.n <= HULL_CAP;
        if (staged) {
            for (int v = lane; v < GA.n; v += 32) { P.hx[v] = __ldg(GA.x + v); P.hy[v] = __ldg(GA.y + v); P.hz[v] = __ldg(GA.z + v); }
            for (int v = lane; v < GB.n; v += 32) {
                P.hx[GA.n + v] = __ldg(GB.x + v); P.hy[GA.n + v] = __ldg(GB.y + v); P.hz[GA.n + v] = __ldg(GB.z + v);
            }
            A.x = P.hx; A.y = P.hy; A.z = P.hz; A.n = GA.n;
            B.x = P.hx + GA.n; B.y = P.hy + GA.n; B.z = P.hz + GA.n; B.n = GB.n;
        } else {
            A.x = GA.x; A.y = GA.y; A.z = GA.z; A.n = GA.n;
            B.x = GB.x; B.y = GB.y; B.z = GB.z; B.n = GB.n;
        }
        // seed polytope from the GJK tetrahedron, faces as F:279-293, wound outwards
        if (lane < 4) {
            const double *s = a.wl_simplex + 12 * item + 3 * lane;
            P.vx[lane] = s[0]; P.vy[lane] = s[1]; P.vz[lane] = s[2];
        }
        if (lane == 0) { P.nv = 4; P.nf = 4; }
        __syncwarp();
        {
            bool bad = false;
            if (lane < 4) {
                int fa = (lane == 3) ? 1 : 0, fb = (lane == 0 || lane == 2) ? 1 : 2, fc = (lane == 0) ? 2 : 3;
                Vec3 M = v3(gjk::add(gjk::add(gjk::add(P.vx[0], P.vx[1]), P.vx[2]), P.vx[3]) / 4.0,
                            gjk::add(gjk::add(gjk::add(P.vy[0], P.vy[1]), P.vy[2]), P.vy[3]) / 4.0,
                            gjk::add(gjk::add(gjk::add(P.vz[0], P.vz[1]), P.vz[2]), P.vz[3]) / 4.0);
                Vec3 n; double d;
                if (!face_plane(P, fa, fb, fc, n, d)) bad = true;
                else if (dot3(n, pv(P, fa) - M) < 0.0) {           // inward: swap two vertices
                    int t = fb; fb = fc; fc = t;
                    if (!face_plane(P, fa, fb, fc, n, d)) bad = true;
                }
                P.nx[lane] = n.x; P.ny[lane] = n.y; P.nz[lane] = n.z; P.d[lane] = d;
                P.fw[lane] = (unsigned)fa | ((unsigned)fb << 8) | ((unsigned)fc << 16);
            }
            if (__any_sync(FULL, bad)) { st |= ST_STOP; ok = false; }   // degenerate face: reference STOPs (F:1369)
        }
        __syncwarp();

        while (ok) {
            ++iters;
            if (iters > MAX_ITERS) { st |= ST_EPA_CAP; iters = MAX_ITERS; ok = false; break; }   // F:299-303

            // closest face, F:893-898 (first minimum)
            const int nf = P.nf;
            double best = DBL_MAX; int k = 0;
            for (int f = lane; f < nf; f += 32) { double ad = fabs(P.d[f]); if (ad < best) { best = ad; k = f; } }
            k = warp_argmin_first_nonneg(best, k);
            const double dk = P.d[k];
            best = fabs(dk);
            Vec3 n = v3(P.nx[k], P.ny[k], P.nz[k]);
            double dt = dk;                                         // (v1 - O) . dir, F:903
            if (fabs(dt) < 1.0e-12) {                               // F:905-908: re-base on the slot centroid
                double sx = 0, sy = 0, sz = 0;
                for (int f = lane; f < nf; f += 32) {
                    unsigned w = P.fw[f];
                    Vec3 u = pv(P, w & 0xff), v = pv(P, (w >> 8) & 0xff), x = pv(P, (w >> 16) & 0xff);
                    sx += u.x + v.x + x.x; sy += u.y + v.y + x.y; sz += u.z + v.z + x.z;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    sx += __shfl_xor_sync(FULL, sx, o); sy += __shfl_xor_sync(FULL, sy, o); sz += __shfl_xor_sync(FULL, sz, o);
                }
                double cnt = (double)(nf * 3);
                dt = dot3(pv(P, P.fw[k] & 0xff) - v3(sx / cnt, sy / cnt, sz / cnt), n);
            }
            Vec3 dir = (dt <= -1.0e-12) ? neg(n) : n;               // F:910

            Vec3 sp = warp_support(A, B, dir, lane);                // F:914
            int ch = insert_point(P, sp, lane, st);
            if (ch >= 0 && best < 1.0e-12) {                        // F:935-944: origin on the face
                Vec3 sp2 = warp_support(A, B, neg(dir), lane);
                int ch2 = insert_point(P, sp2, lane, st);
                ch = (ch2 < 0) ? ch2 : (ch | ch2);
            }
            if (ch == -2) { overflow = true; ok = false; break; }
            if (ch < 0) { ok = false; break; }
            if (ch == 0) {                                          // hull unchanged: converged (F:994-998)
                depth = best;
                nrm = (dk < 0.0) ? neg(n) : n;                      // F:967-969
                break;
            }
        }

        if (overflow) {
            if (a.ovf_items) {                                      // retry with the large polytope
                if (lane == 0) a.ovf_items[atomicAdd(a.ovf_count, 1ull)] = (int32_t)item;
                continue;
            }
            st |= ST_HULL_FAIL;
        }
        st |= (uint32_t)iters << 16;
        if (ok) {
            if (lane == 0) {
                if (a.depth) a.depth[p] = depth;
                if (a.normal) { double *o = a.normal + 3 * (int64_t)p; o[0] = nrm.x; o[1] = nrm.y; o[2] = nrm.z; }
            }
            if (a.full) {
                __syncwarp();
                if (A.n <= FC && B.n <= FC) contact_epilogue(a, p, A, B, nrm, lane, st, P.nx, P.ny);     // planes are dead now
                else st |= ST_HULL_FAIL;                                                            // hulls beyond the scratch: epilogue skipped
            }
        }
        if (lane == 0 && a.result) {
            double *r = a.result + 12 * item;
            r[0] = nrm.x; r[1] = nrm.y; r[2] = nrm.z; r[3] = ok ? depth : -1.0;
        }
        if (lane == 0 && a.status) atomicOr(a.status + p, st);
    }
}

template <int VC, int FC>
inline int launch_one(const EpaArgs &a, int sm_count, cudaStream_t st)
{
    const size_t smem = sizeof(Polytope<VC, FC>) * WARPS;
    cudaError_t e = cudaFuncSetAttribute(epa_kernel<VC, FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epa_kernel<VC, FC>, WARPS * 32, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    epa_kernel<VC, FC><<<sm_count * per_sm, WARPS * 32, smem, st>>>(a);
    return 0;
}

// two passes: small polytopes for everything, then the pairs that outgrew them with the full bound.
// queue heads: counters[1] (pass 1) and counters[3] (pass 2); overflow count: counters[4]
inline int launch(EpaArgs a, int sm_count, cudaStream_t st, int32_t *ovf_items, unsigned long long *counters, int *launches)
{
    a.item_list = nullptr; a.item_count = nullptr;
    a.queue_head = counters + 1;
    a.ovf_items = ovf_items; a.ovf_count = counters + 4;
    int rc = launch_one<64, 128>(a, sm_count, st);
    if (rc) return rc;
    EpaArgs b = a;
    b.item_list = ovf_items; b.item_count = counters + 4;
    b.queue_head = counters + 3;
    b.ovf_items = nullptr; b.ovf_count = nullptr;
    rc = launch_one<128, 256>(b, sm_count, st);
    if (launches) *launches = 2;
    return rc;
}

}  // namespace epa
