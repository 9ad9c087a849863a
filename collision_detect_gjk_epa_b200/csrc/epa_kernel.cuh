// epa_kernel.cuh -- K3: EPA penetration depth + contact normal (and the contact epilogue K4) for the
// pairs GJK reported as colliding.  One warp per pair, polytope resident in shared memory.
//
// Path: EPA_solu (reference src/GCLIB_GJKEPA.f90:242-346) and update_expandingPolytope_EPA
// (F:863-1022).  The reference rebuilds the whole polytope from scratch with an external QuickHull
// on every iteration (F:920-950) and decides convergence by comparing bubble-sorted face-distance
// lists (F:972-1015).  Here the polytope is kept incrementally:
//   per iteration  closest face (warp arg-min over cached plane distances, F:893-898)
//                  -> outward direction with the reference's orientation rule (F:902-910)
//                  -> support point (32-lane max-dot reduction, F:914)
//                  -> faces that see the point by more than HULL_EPS are marked in parallel,
//                     the horizon is the set of their edges whose twin is not marked, marked faces
//                     are compacted away and one new face per horizon edge is fanned to the point;
//   converged      when the support point(s) do not change the hull -- the geometric meaning of
//                  the reference test -- reporting the closest face's distance and outward normal
//                  (F:963-969, F:997-998).
// Plane normals/distances use the reference forms (UNINML F:1389-1392, DIST_PF_SIGN F:1375) so that
// depth and normal agree with the CPU path to rounding.
#pragma once
#include "gjk_kernels.cuh"
#include <cuda_runtime.h>

namespace epa {

using gjk::Vec3; using gjk::v3; using gjk::dot3; using gjk::cross3; using gjk::norm3; using gjk::neg;
using gjk::HullView; using gjk::Lanes;

constexpr int    V_CAP = 128;          // 4 + 99 iterations + slack
constexpr int    F_CAP = 256;          // 2V-4
constexpr int    WARPS = 4;
constexpr double HULL_EPS = 1.0e-10;   // same "swallow" tolerance as the oracle's hull stand-in
constexpr int    MAX_ITERS = 99;       // F:299

constexpr uint32_t ST_EPA_CAP = 1u << 24, ST_STOP = 1u << 25, ST_HULL_FAIL = 1u << 26,
                   ST_BAD_VERSION = 1u << 27, ST_TOPO_WARN = 1u << 29;

struct EpaArgs {
    const double *verts; const int64_t *hull_off;
    const int32_t *pair_a, *pair_b;
    const int32_t *wl_pair; const double *wl_simplex;
    const unsigned long long *wl_count; unsigned long long *queue_head;
    int32_t version; double tol_ff; bool full;
    int32_t *colli_type; double *nearest, *normal, *point, *depth; uint32_t *status;
};

struct alignas(16) Polytope {
    double vx[V_CAP], vy[V_CAP], vz[V_CAP];
    double nx[F_CAP], ny[F_CAP], nz[F_CAP], d[F_CAP];     // outward unit normal, n.x = d on the face
    uint8_t fa[F_CAP], fb[F_CAP], fc[F_CAP];
    uint8_t vis[F_CAP];
    uint8_t vlist[F_CAP];                                  // indices of visible faces
    uint8_t ha[F_CAP], hb[F_CAP];                          // horizon edges (directed)
    int nv, nf;
};

__device__ __forceinline__ Vec3 pv(const Polytope &P, int i) { return v3(P.vx[i], P.vy[i], P.vz[i]); }

// plane of face (a,b,c): UNINML (F:1389-1392) and its signed offset; false if degenerate
__device__ __forceinline__ bool face_plane(const Polytope &P, int a, int b, int c, Vec3 &n, double &d)
{
    Vec3 va = pv(P, a), vb = pv(P, b), vc = pv(P, c);
    Vec3 cr = cross3(vb - va, vc - vb);
    if (!(fabs(cr.x) > 1.0e-12 || fabs(cr.y) > 1.0e-12 || fabs(cr.z) > 1.0e-12)) { n = v3(0, 0, 0); d = 0; return false; }
    n = gjk::divs(cr, norm3(cr));
    d = dot3(va, n);
    return true;
}

// Insert point q.  Returns 1 if the hull changed, 0 if not, <0 on failure.  Warp-synchronous.
__device__ int insert_point(Polytope &P, Vec3 q, int lane, uint32_t &st)
{
    const unsigned FULL = 0xffffffffu;
    // exact duplicate of an existing vertex (the reference appends unconditionally, F:927; its hull then drops it)
    bool dup = false;
    for (int v = lane; v < P.nv; v += 32) dup = dup || (P.vx[v] == q.x && P.vy[v] == q.y && P.vz[v] == q.z);
    if (__any_sync(FULL, dup)) return 0;

    // visible faces
    int nvis = 0;
    for (int c = 0; c < P.nf; c += 32) {
        int f = c + lane;
        bool v = false;
        if (f < P.nf) {
            double dist = gjk::sub(dot3(v3(P.nx[f], P.ny[f], P.nz[f]), q), P.d[f]);
            v = dist > HULL_EPS;
            P.vis[f] = v ? 1 : 0;
        }
        unsigned b = __ballot_sync(FULL, v);
        if (v) P.vlist[nvis + __popc(b & ((1u << lane) - 1u))] = (uint8_t)f;
        nvis += __popc(b);
    }
    if (nvis == 0) return 0;
    if (P.nv >= V_CAP) { st |= ST_HULL_FAIL; return -1; }
    __syncwarp();

    // horizon: directed edges of visible faces whose reverse is not an edge of a visible face
    int nh = 0;
    for (int c = 0; c < 3 * nvis; c += 32) {
        int it = c + lane;
        bool hz = false; int ea = 0, eb = 0;
        if (it < 3 * nvis) {
            int f = P.vlist[it / 3], e = it % 3;
            int a0 = P.fa[f], a1 = P.fb[f], a2 = P.fc[f];
            ea = e == 0 ? a0 : (e == 1 ? a1 : a2);
            eb = e == 0 ? a1 : (e == 1 ? a2 : a0);
            hz = true;
            for (int j = 0; j < nvis; ++j) {
                int g = P.vlist[j];
                int g0 = P.fa[g], g1 = P.fb[g], g2 = P.fc[g];
                if ((g0 == eb && g1 == ea) || (g1 == eb && g2 == ea) || (g2 == eb && g0 == ea)) { hz = false; break; }
            }
        }
        unsigned b = __ballot_sync(FULL, hz);
        if (hz) {
            int pos = nh + __popc(b & ((1u << lane) - 1u));
            if (pos < F_CAP) { P.ha[pos] = (uint8_t)ea; P.hb[pos] = (uint8_t)eb; }
        }
        nh += __popc(b);
    }
    if (nh != nvis + 2) st |= ST_TOPO_WARN;         // visible region is not a disk
    if (P.nf - nvis + nh > F_CAP || nh > F_CAP) { st |= ST_HULL_FAIL; return -1; }
    __syncwarp();

    // drop visible faces (stable, in place)
    int base = 0;
    for (int c = 0; c < P.nf; c += 32) {
        int f = c + lane;
        bool keep = f < P.nf && !P.vis[f];
        double nx = 0, ny = 0, nz = 0, dd = 0; uint8_t a = 0, b_ = 0, c_ = 0;
        if (keep) { nx = P.nx[f]; ny = P.ny[f]; nz = P.nz[f]; dd = P.d[f]; a = P.fa[f]; b_ = P.fb[f]; c_ = P.fc[f]; }
        unsigned b = __ballot_sync(FULL, keep);
        __syncwarp();
        if (keep) {
            int pos = base + __popc(b & ((1u << lane) - 1u));
            P.nx[pos] = nx; P.ny[pos] = ny; P.nz[pos] = nz; P.d[pos] = dd; P.fa[pos] = a; P.fb[pos] = b_; P.fc[pos] = c_;
        }
        base += __popc(b);
        __syncwarp();
    }

    // new vertex and the fan of new faces
    const int vid = P.nv;
    if (lane == 0) { P.vx[vid] = q.x; P.vy[vid] = q.y; P.vz[vid] = q.z; }
    __syncwarp();
    bool bad = false;
    for (int e = lane; e < nh; e += 32) {
        int a = P.ha[e], b = P.hb[e], f = base + e;
        Vec3 n; double d;
        if (!face_plane(P, a, b, vid, n, d)) bad = true;
        P.nx[f] = n.x; P.ny[f] = n.y; P.nz[f] = n.z; P.d[f] = d;
        P.fa[f] = (uint8_t)a; P.fb[f] = (uint8_t)b; P.fc[f] = (uint8_t)vid;
    }
    if (lane == 0) { P.nv = vid + 1; P.nf = base + nh; }
    __syncwarp();
    if (__any_sync(FULL, bad)) { st |= ST_STOP; return -1; }      // DIST_PF_SIGN's STOP path (F:1369)
    return 1;
}

// ---- contact epilogue (K4), F:326-343 ----
__device__ void near_max_pair(const HullView &h, Vec3 d, int idx[2])
{   // F:722-747 tracker (quirk Q8): "> max - 1e-8", last wins, previous winner becomes second
    double mx = -DBL_MAX;
    idx[0] = idx[1] = 0;
    for (int i = 0; i < h.n; ++i) {
        double t = dot3(d, h.get(i));
        if (t > mx - 1.0e-8) { mx = t; idx[1] = idx[0]; idx[0] = i + 1; }
    }
    if (idx[1] == 0) idx[1] = idx[0];
}

__device__ void contact_epilogue(const EpaArgs &a, int p, const HullView &A, const HullView &B,
                                 Vec3 nrm, int lane, uint32_t &st)
{
    const Lanes<32> g;
    // get_nearest_points, F:813-855: the two support vertices along +-normal
    {
        double bA = -DBL_MAX, bB = -DBL_MAX; int iA = 0, iB = 0;
        for (int v = lane; v < A.n; v += 32) { double t = dot3(nrm, A.get(v)); if (t > bA) { bA = t; iA = v; } }
        Vec3 nd = neg(nrm);
        for (int v = lane; v < B.n; v += 32) { double t = dot3(nd, B.get(v)); if (t > bB) { bB = t; iB = v; } }
        g.argmax_first(bA, iA); g.argmax_first(bB, iB);
        if (lane == 0 && a.nearest) {
            Vec3 pa = A.get(iA), pb = B.get(iB);
            double *o = a.nearest + 6 * (int64_t)p;
            o[0] = pa.x; o[1] = pb.x; o[2] = pa.y; o[3] = pb.y; o[4] = pa.z; o[5] = pb.z;
        }
    }
    Vec3 nrm_out = nrm;
    if (a.version == 1) {                       // get_collisionPoint_01, F:700-806
        int s[2] = {0, 0};
        if (lane == 0) near_max_pair(A, nrm, s);
        if (lane == 1) near_max_pair(B, neg(nrm), s);
        int a0 = __shfl_sync(0xffffffffu, s[0], 0), a1 = __shfl_sync(0xffffffffu, s[1], 0);
        int b0 = __shfl_sync(0xffffffffu, s[0], 1), b1 = __shfl_sync(0xffffffffu, s[1], 1);
        bool m1 = a0 != a1, m2 = b0 != b1;
        if (lane == 0 && a.point) {
            Vec3 r = v3(0, 0, 0);
            if (!m1 && !m2) {
                Vec3 u = A.get(a0 - 1), w = B.get(b0 - 1);
                r = v3(gjk::add(u.x, w.x) / 2.0, gjk::add(u.y, w.y) / 2.0, gjk::add(u.z, w.z) / 2.0);
            } else if (m1 && !m2) r = B.get(b0 - 1);
            else if (!m1 && m2) r = A.get(a0 - 1);
            else {                              // F:766-804: mean of p1 supports within 0.1 of the maximum
                double mx = -DBL_MAX;
                for (int i = 0; i < A.n; ++i) { double t = dot3(nrm, A.get(i)); if (t > mx) mx = t; }
                double sx = 0, sy = 0, sz = 0; int c = 0;
                for (int i = 0; i < A.n; ++i) {
                    Vec3 v = A.get(i);
                    if (dot3(nrm, v) > mx - 1.0e-1) { ++c; sx = gjk::add(sx, v.x); sy = gjk::add(sy, v.y); sz = gjk::add(sz, v.z); }
                }
                r = v3(sx / (double)c, sy / (double)c, sz / (double)c);
            }
            double *o = a.point + 3 * (int64_t)p;
            o[0] = r.x; o[1] = r.y; o[2] = r.z;
        }
    } else if (a.version == 3) {                // get_collisionPoint_03, F:426-452
        if (lane == 0) {
            double mx = -DBL_MAX; int idx = 0; Vec3 nd = neg(nrm);
            for (int i = 0; i < B.n; ++i) { double t = dot3(nd, B.get(i)); if (t > mx - 1.0e-8) { mx = t; idx = i; } }
            Vec3 r = B.get(idx);
            double s = 0.0;
            for (int i = 0; i < A.n; ++i) s = gjk::add(s, A.get(i).z);
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
        double mA = -DBL_MAX, mB = -DBL_MAX;
        Vec3 nd = neg(nrm_out);
        for (int v = lane; v < A.n; v += 32) { double t = dot3(nrm_out, A.get(v)); if (t > mA) mA = t; }
        for (int v = lane; v < B.n; v += 32) { double t = dot3(nd, B.get(v)); if (t > mB) mB = t; }
        for (int o = 16; o > 0; o >>= 1) {
            double t = __shfl_xor_sync(0xffffffffu, mA, o); if (t > mA) mA = t;
            t = __shfl_xor_sync(0xffffffffu, mB, o); if (t > mB) mB = t;
        }
        int cA = 0, cB = 0;
        for (int v = lane; v < A.n; v += 32) cA += dot3(nrm_out, A.get(v)) > gjk::sub(mA, a.tol_ff);
        for (int v = lane; v < B.n; v += 32) cB += dot3(nd, B.get(v)) > gjk::sub(mB, a.tol_ff);
        for (int o = 16; o > 0; o >>= 1) { cA += __shfl_xor_sync(0xffffffffu, cA, o); cB += __shfl_xor_sync(0xffffffffu, cB, o); }
        if (lane == 0 && a.colli_type) a.colli_type[p] = (cA >= 3 && cB >= 3) ? 2 : 1;
    }
}

__global__ void __launch_bounds__(WARPS * 32)
epa_kernel(EpaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Polytope &P = reinterpret_cast<Polytope *>(smem_raw)[warp];
    const unsigned FULL = 0xffffffffu;
    const Lanes<32> g;
    const unsigned long long n_items = *a.wl_count;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.queue_head, 1ull);
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_items) break;

        const int p = a.wl_pair[item];
        const HullView A = gjk::hull_view(a.verts, a.hull_off, a.pair_a[p]);
        const HullView B = gjk::hull_view(a.verts, a.hull_off, a.pair_b[p]);
        uint32_t st = 0;
        int iters = 0;
        double depth = 0.0; Vec3 nrm = v3(0, 0, 0);
        bool ok = true;

        // seed polytope from the GJK tetrahedron, faces as F:279-293, wound outwards
        __syncwarp();
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
                P.fa[lane] = (uint8_t)fa; P.fb[lane] = (uint8_t)fb; P.fc[lane] = (uint8_t)fc;
            }
            if (__any_sync(FULL, bad)) { st |= ST_STOP; ok = false; }   // degenerate face: reference STOPs (F:1369)
        }
        __syncwarp();

        while (ok) {
            ++iters;
            if (iters > MAX_ITERS) { st |= ST_EPA_CAP; iters = MAX_ITERS; ok = false; break; }   // F:299-303

            // closest face, F:893-898 (first minimum)
            double best = DBL_MAX; int k = 0;
            for (int f = lane; f < P.nf; f += 32) { double ad = fabs(P.d[f]); if (ad < best) { best = ad; k = f; } }
            g.argmin_first(best, k);
            Vec3 n = v3(P.nx[k], P.ny[k], P.nz[k]);
            double dt = P.d[k];                                     // (v1 - O) . dir, F:903
            if (fabs(dt) < 1.0e-12) {                               // F:905-908: re-base on the slot centroid
                double sx = 0, sy = 0, sz = 0;
                for (int f = lane; f < P.nf; f += 32) {
                    Vec3 u = pv(P, P.fa[f]), v = pv(P, P.fb[f]), w = pv(P, P.fc[f]);
                    sx += u.x + v.x + w.x; sy += u.y + v.y + w.y; sz += u.z + v.z + w.z;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    sx += __shfl_xor_sync(FULL, sx, o); sy += __shfl_xor_sync(FULL, sy, o); sz += __shfl_xor_sync(FULL, sz, o);
                }
                double cnt = (double)(P.nf * 3);
                dt = dot3(pv(P, P.fa[k]) - v3(sx / cnt, sy / cnt, sz / cnt), n);
            }
            Vec3 dir = (dt <= -1.0e-12) ? neg(n) : n;               // F:910

            Vec3 sp = gjk::support<32>(g, A, B, dir);               // F:914
            int ch = insert_point(P, sp, lane, st);
            if (ch >= 0 && fabs(best) < 1.0e-12) {                  // F:935-944: origin on the face
                Vec3 sp2 = gjk::support<32>(g, A, B, neg(dir));
                int ch2 = insert_point(P, sp2, lane, st);
                ch = (ch2 < 0) ? ch2 : (ch | ch2);
            }
            if (ch < 0) { ok = false; break; }
            if (ch == 0) {                                          // hull unchanged: converged (F:994-998)
                depth = best;
                nrm = (P.d[k] < 0.0) ? neg(n) : n;                  // F:967-969
                break;
            }
        }

        st |= (uint32_t)iters << 16;
        if (ok) {
            if (lane == 0) {
                if (a.depth) a.depth[p] = depth;
                if (a.normal) { double *o = a.normal + 3 * (int64_t)p; o[0] = nrm.x; o[1] = nrm.y; o[2] = nrm.z; }
            }
            if (a.full) contact_epilogue(a, p, A, B, nrm, lane, st);
        }
        if (lane == 0 && a.status) a.status[p] |= st;
    }
}

inline int launch(const EpaArgs &a, int sm_count, cudaStream_t st, int *blocks_out)
{
    const size_t smem = sizeof(Polytope) * WARPS;
    cudaError_t e = cudaFuncSetAttribute(epa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epa_kernel, WARPS * 32, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    int blocks = sm_count * per_sm;
    epa_kernel<<<blocks, WARPS * 32, smem, st>>>(a);
    if (blocks_out) *blocks_out = blocks;
    return 0;
}

}  // namespace epa
