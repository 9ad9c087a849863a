// epa_v3.cuh -- K3 fast path: EPA penetration depth + contact normal (and the contact epilogue K4) with the
// polytope kept as a linked triangle mesh in shared memory, one pair per GROUP of LP lanes (LP = 16: two
// pairs per warp; LP = 32: one).
//
// Path: EPA_solu (reference src/GCLIB_GJKEPA.f90:242-346) and update_expandingPolytope_EPA (F:863-1022).
// The reference re-hulls all polytope vertices with an external QuickHull on every iteration; epa_kernel.cuh
// (the general kernel, kept as the last pass for whatever this one declines) updates the face list in place:
// visible faces, horizon by an all-pairs edge match, stable compaction, fan.  ncu on that kernel (cfg3, profiles/):
// 750 warp instructions per iteration of which 13 % FP64, the fan's face planes computed by 4 of 32 lanes, 20
// warps per SM because each pair holds 9.7 KB.  Here:
//   * every face stores its three neighbours, so a visible face's edge is on the horizon iff the neighbour
//     across it is not visible -- no edge matching;
//   * faces never move: the fan's faces take the slots of the faces they replace (the visible region is a
//     disk: nh = nvis + 2, so the mesh grows by two slots per vertex and F = 2V - 4 slots suffice);
//   * planes are one 32-byte record (unit normal, offset), hulls are read from shared memory by plain LDS;
//   * a pair needs 5.4 KB (V = 36) and 16 lanes, so an SM holds 40 pairs in 20 warps, and the two pairs of a
//     warp run in LOCKSTEP: one instruction stream (trip counts = the larger of the two, bodies predicated)
//     serves both, every warp collective keeps the constant full mask (a run-time member mask makes nvcc wrap
//     each shuffle/vote/redux in a MATCH + vote + divergence check), segmented reductions are one REDUX per
//     segment.  A group that finishes fetches its next pair while the other keeps iterating.
//   * the contact epilogue is a kernel of its own (contact_kernel.cuh).
// Results are the same quantities as epa_kernel.cuh computes (closest face of the converged polytope: F:963-969,
// F:997-998) with the reference's plane forms (UNINML F:1389-1392, DIST_PF_SIGN F:1375); the order of faces in
// memory differs, which only matters for which of several EXACTLY tied faces is reported.
// Anything unusual -- capacity, a visible region that is not a disk, hulls larger than the staging buffer --
// re-queues the pair for the next pass.
#pragma once
#include "epa_kernel.cuh"

namespace epa3 {

using gjk::Vec3; using gjk::v3; using gjk::dot3; using gjk::cross3; using gjk::neg;
using epa::EpaArgs; using epa::HULL_EPS; using epa::MAX_ITERS;
using epa::ST_EPA_CAP; using epa::ST_STOP; using epa::ST_HULL_FAIL; using epa::ST_BAD_VERSION;

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned FV_VIS = 1u << 30;

// ---- a group of LP consecutive lanes inside a warp whose 32/LP groups execute the same instructions ----
// Every collective takes the constant full mask and must be reached by the whole warp.
template <int LP> struct Grp {
    int lane, shift;
    __device__ __forceinline__ Grp() { const int wl = threadIdx.x & 31; lane = wl & (LP - 1); shift = wl & ~(LP - 1); }
    __device__ __forceinline__ unsigned ballot(bool p) const
    { return (LP == 32) ? __ballot_sync(FULL, p) : ((__ballot_sync(FULL, p) >> shift) & ((1u << (LP & 31)) - 1u)); }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(FULL, v, src, LP); }
    // segmented REDUX: one reduction per group, lanes of the other groups contribute the neutral element
    __device__ __forceinline__ unsigned rmax(unsigned v) const
    {
        if (LP == 32) return __reduce_max_sync(FULL, v);
        unsigned r = 0u;
#pragma unroll
        for (int s = 0; s < 32; s += LP) { const unsigned t = __reduce_max_sync(FULL, shift == s ? v : 0u); if (shift == s) r = t; }
        return r;
    }
    __device__ __forceinline__ unsigned rmin(unsigned v) const
    {
        if (LP == 32) return __reduce_min_sync(FULL, v);
        unsigned r = 0u;
#pragma unroll
        for (int s = 0; s < 32; s += LP) { const unsigned t = __reduce_min_sync(FULL, shift == s ? v : 0xffffffffu); if (shift == s) r = t; }
        return r;
    }
    // first index of the maximum / of the minimum of non-negative values (order-preserving 64-bit keys)
    __device__ __forceinline__ int argmax_first(double v, int i) const
    {
        long long b = __double_as_longlong(__dadd_rn(v, 0.0));               // -0 -> +0
        unsigned long long k = (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
        unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
        unsigned mh = rmax(hi);
        unsigned ml = rmax(hi == mh ? lo : 0u);
        return (int)rmin((hi == mh && lo == ml) ? (unsigned)i : 0x7fffffffu);
    }
    __device__ __forceinline__ int argmin_first_nonneg(double v, int i) const
    {
        unsigned long long k = (unsigned long long)__double_as_longlong(v);
        unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
        unsigned mh = rmin(hi);
        unsigned ml = rmin(hi == mh ? lo : 0xffffffffu);
        return (int)rmin((hi == mh && lo == ml) ? (unsigned)i : 0x7fffffffu);
    }
};
__device__ __forceinline__ int wmax(int v) { return __reduce_max_sync(FULL, v); }      // warp-uniform trip count

template <int VC, int FC, int HC>
struct alignas(32) Mesh {
    double4 pl[FC];                       // outward unit normal (x,y,z) and offset w: n.x = w on the face
    double vx[VC], vy[VC], vz[VC];        // polytope vertices (Minkowski-difference points)
    double hx[HC], hy[HC], hz[HC];        // staged hulls: A at [0,nA), B at [nA,nA+nB)
    uint32_t fv[FC];                      // a | b<<8 | c<<16 | FV_VIS
    uint32_t adj[FC];                     // neighbour across edge (a,b) | across (b,c) << 8 | across (c,a) << 16
    uint32_t hl[FC];                      // horizon edges of this insertion: a | b<<8 | outside face<<16 | its edge<<24
    uint8_t vlist[FC];                    // visible faces of this insertion
    uint8_t estart[VC], eend[VC];         // fan face whose horizon edge starts / ends at a vertex
};

struct HullS { const double *x, *y, *z; int n; };      // hull in shared memory

// support_mapping (F:1030-1062) for the groups with `on`; the others idle through the same instructions
template <int LP>
__device__ __forceinline__ Vec3 group_support(const Grp<LP> &g, const HullS &A, const HullS &B, Vec3 d, bool on)
{
    const int nA = on ? A.n : 0, nB = on ? B.n : 0;
    const int nAw = wmax(nA), nBw = wmax(nB);
    double bA = -DBL_MAX, bB = -DBL_MAX; int iA = 0, iB = 0;
    for (int v = g.lane; v < nAw; v += LP)
        if (v < nA) { double t = dot3(d, v3(A.x[v], A.y[v], A.z[v])); if (t > bA) { bA = t; iA = v; } }
    const Vec3 nd = neg(d);
    for (int v = g.lane; v < nBw; v += LP)
        if (v < nB) { double t = dot3(nd, v3(B.x[v], B.y[v], B.z[v])); if (t > bB) { bB = t; iB = v; } }
    iA = g.argmax_first(bA, iA);
    iB = g.argmax_first(bB, iB);
    return v3(A.x[iA], A.y[iA], A.z[iA]) - v3(B.x[iB], B.y[iB], B.z[iB]);
}

// v / |v| with the compiler's IEEE operators: the redo after an operand left the straight-line range.  Out of line.
__device__ __noinline__ Vec3 unit_cold(Vec3 v)
{
    const double md = sqrt(gjk::add(gjk::add(gjk::mul(v.x, v.x), gjk::mul(v.y, v.y)), gjk::mul(v.z, v.z)));
    return v3(v.x / md, v.y / md, v.z / md);
}

// plane of triangle (va,vb,vc): UNINML (F:1389-1392) and its offset; false if degenerate
__device__ __forceinline__ bool tri_plane(Vec3 va, Vec3 vb, Vec3 vc, double4 &pl)
{
    const Vec3 cr = cross3(vb - va, vc - vb);
    if (!(fabs(cr.x) > 1.0e-12 || fabs(cr.y) > 1.0e-12 || fabs(cr.z) > 1.0e-12)) { pl = make_double4(0, 0, 0, 0); return false; }
    bool ok = true;
    double md = gjk::MathFast::nrm(cr, ok);
    Vec3 n = gjk::MathFast::dv(cr, md, ok);
    if (!ok) n = unit_cold(cr);
    pl = make_double4(n.x, n.y, n.z, dot3(va, n));
    return true;
}

// Insert point q into the meshes of the groups with `on`.  Per group: 1 hull changed, 0 unchanged (also for
// groups without `on`), -1 degenerate face (reference STOP path), -2 declined (capacity, or a visible region
// that is not a disk; the mesh is left untouched in both cases).
template <int LP, int VC, int FC, int HC>
__device__ __forceinline__ int mesh_insert(const Grp<LP> &g, Mesh<VC, FC, HC> &P, int &nv, int &nf, Vec3 q, bool on)
{
    const unsigned lt = (1u << g.lane) - 1u;
    // faces that see q
    const int nf_on = on ? nf : 0;
    const int nfw = wmax(nf_on);
    int nvis = 0;
    for (int c = 0; c < nfw; c += LP) {
        const int f = c + g.lane;
        bool v = false;
        if (f < nf_on) {
            const double4 pl = P.pl[f];
            const unsigned w = P.fv[f];
            v = gjk::sub(dot3(v3(pl.x, pl.y, pl.z), q), pl.w) > HULL_EPS;
            P.fv[f] = v ? (w | FV_VIS) : (w & ~FV_VIS);
        }
        const unsigned b = g.ballot(v);
        if (v) P.vlist[nvis + __popc(b & lt)] = (uint8_t)f;
        nvis += __popc(b);
    }
    int code = 0;
    if (nvis > 0 && nv >= VC) code = -2;
    const bool go = nvis > 0 && code == 0;
    __syncwarp();
    // horizon: edges of visible faces whose neighbour is not visible
    const int ne = go ? 3 * nvis : 0;
    const int new_ = wmax(ne);
    int nh = 0;
    for (int c = 0; c < new_; c += LP) {
        const int it = c + g.lane;
        bool hz = false; unsigned rec = 0;
        if (it < ne) {
            const int f = P.vlist[it / 3], e = it - 3 * (it / 3);
            const unsigned w = P.fv[f];
            const unsigned nb = (P.adj[f] >> (8 * e)) & 0xffu;
            hz = !(P.fv[nb] & FV_VIS);
            if (hz) {
                const unsigned a0 = w & 0xffu, a1 = (w >> 8) & 0xffu, a2 = (w >> 16) & 0xffu;
                const unsigned ea = e == 0 ? a0 : (e == 1 ? a1 : a2), eb = e == 0 ? a1 : (e == 1 ? a2 : a0);
                const unsigned an = P.adj[nb];
                const unsigned eg = ((an & 0xffu) == (unsigned)f) ? 0u : ((((an >> 8) & 0xffu) == (unsigned)f) ? 1u : 2u);
                rec = ea | (eb << 8) | (nb << 16) | (eg << 24);
            }
        }
        const unsigned b = g.ballot(hz);
        if (hz) { const int pos = nh + __popc(b & lt); if (pos < FC) P.hl[pos] = rec; }
        nh += __popc(b);
    }
    if (go && (nh != nvis + 2 || nf + 2 > FC)) code = -2;     // not a disk, or out of slots
    const bool fan = go && code == 0;
    __syncwarp();
    // the fan: face i takes the slot of the i-th replaced face, the last two take new slots
    const int vid = nv;
    if (fan && g.lane == 0) { P.vx[vid] = q.x; P.vy[vid] = q.y; P.vz[vid] = q.z; }
    const int nfan = fan ? nh : 0;
    const int nfanw = wmax(nfan);
    bool bad = false;
    for (int c = 0; c < nfanw; c += LP) {
        const int i = c + g.lane;
        if (i < nfan) {
            const unsigned rec = P.hl[i];
            const int a = rec & 0xffu, b = (rec >> 8) & 0xffu, nb = (rec >> 16) & 0xffu, eg = rec >> 24;
            const int s = i < nvis ? (int)P.vlist[i] : nf + (i - nvis);
            double4 pl;
            if (!tri_plane(v3(P.vx[a], P.vy[a], P.vz[a]), v3(P.vx[b], P.vy[b], P.vz[b]), q, pl)) bad = true;
            P.pl[s] = pl;
            P.fv[s] = (unsigned)a | ((unsigned)b << 8) | ((unsigned)vid << 16);
            P.estart[a] = (uint8_t)s; P.eend[b] = (uint8_t)s;
            // the outside neighbour now borders this face (byte store: several fan faces may share a neighbour)
            reinterpret_cast<uint8_t *>(&P.adj[nb])[eg] = (uint8_t)s;
        }
    }
    __syncwarp();
    for (int c = 0; c < nfanw; c += LP) {
        const int i = c + g.lane;
        if (i < nfan) {
            const unsigned rec = P.hl[i];
            const int a = rec & 0xffu, b = (rec >> 8) & 0xffu, nb = (rec >> 16) & 0xffu;
            const int s = i < nvis ? (int)P.vlist[i] : nf + (i - nvis);
            // edges (a,b) | (b,vid) | (vid,a): outside face, the fan face starting at b, the fan face ending at a
            P.adj[s] = (unsigned)nb | ((unsigned)P.estart[b] << 8) | ((unsigned)P.eend[a] << 16);
        }
    }
    if (fan) { nv = vid + 1; nf += 2; code = 1; }
    __syncwarp();
    if (g.any(bad) && fan) code = -1;
    return code;
}

constexpr int WARPS = 4;

#ifndef EPA3_MIN_CTAS
#define EPA3_MIN_CTAS 5
#endif
template <int LP, int VC, int FC, int HC>
__global__ void __launch_bounds__(WARPS * 32, (HC > 64 ? 4 : EPA3_MIN_CTAS))      // shared memory holds 4 CTAs of the 128-vertex form
epa3_kernel(EpaArgs a)
{
    static_assert(FC <= 255 && VC <= 255, "ids are bytes");
    extern __shared__ __align__(32) unsigned char smem_raw3[];
    constexpr int GPW = 32 / LP;                               // groups per warp
    const Grp<LP> g;
    const int grp_id = (threadIdx.x >> 5) * GPW + ((threadIdx.x & 31) / LP);
    Mesh<VC, FC, HC> &P = reinterpret_cast<Mesh<VC, FC, HC> *>(smem_raw3)[grp_id];
    const unsigned long long n_items = a.item_list ? *a.item_count : *a.wl_count;

    // per-group state, identical in the group's lanes
    bool busy = false, drained = false;
    long long item = 0; int p = 0;
    int nv = 4, nf = 4, iters = 0; uint32_t st = 0;
    HullS A, B;
    A.x = P.hx; A.y = P.hy; A.z = P.hz; A.n = 0; B = A;

    // the pair leaves this kernel: results and status, or the next pass's list
    auto retire = [&](bool good, bool declined, double depth, Vec3 nrm) {
        if (g.lane == 0) {
            if (declined && a.ovf_items) a.ovf_items[atomicAdd(a.ovf_count, 1ull)] = (int32_t)item;
            else {
                uint32_t s2 = st | ((uint32_t)iters << 16) | (declined ? ST_HULL_FAIL : 0u);
                if (good) {
                    if (a.depth) a.depth[p] = depth;
                    if (a.normal) { double *o = a.normal + 3 * (int64_t)p; o[0] = nrm.x; o[1] = nrm.y; o[2] = nrm.z; }
                }
                if (a.result) { double *r = a.result + 12 * item; r[0] = nrm.x; r[1] = nrm.y; r[2] = nrm.z; r[3] = good ? depth : -1.0; }
                if (a.status) atomicOr(a.status + p, s2);          // fire and forget: no round trip for the GJK half of the word
            }
        }
        busy = false;
    };

    for (;;) {
        // ---- idle groups fetch their next pair ----
        const bool need = !busy && !drained;
        if (__any_sync(FULL, need)) {
            unsigned long long qi = 0;
            if (need && g.lane == 0) qi = atomicAdd(a.queue_head, 1ull);
            qi = g.shfl(qi, 0);
            bool start = false;
            if (need) {
                if (qi >= n_items) drained = true;
                else {
                    item = a.item_list ? (long long)a.item_list[qi] : (long long)qi;
                    p = a.wl_pair[item];
                    const gjk::HullView GA = gjk::hull_view(a.verts, a.hull_off, a.pair_a[p]);
                    const gjk::HullView GB = gjk::hull_view(a.verts, a.hull_off, a.pair_b[p]);
                    iters = 0; st = 0; nv = 4; nf = 4;
                    const int tot = GA.n + GB.n;
                    if (tot <= a.take_gt || (a.skip_large && tot > HC)) { /* another pass owns this pair */ }
                    else if (tot > HC) retire(false, true, 0.0, v3(0, 0, 0));             // hulls beyond the staging buffer
                    else {
                        start = true;
                        A.n = GA.n; B.x = P.hx + GA.n; B.y = P.hy + GA.n; B.z = P.hz + GA.n; B.n = GB.n;
                        for (int v = g.lane; v < GA.n; v += LP) { P.hx[v] = __ldg(GA.x + v); P.hy[v] = __ldg(GA.y + v); P.hz[v] = __ldg(GA.z + v); }
                        for (int v = g.lane; v < GB.n; v += LP) {
                            P.hx[GA.n + v] = __ldg(GB.x + v); P.hy[GA.n + v] = __ldg(GB.y + v); P.hz[GA.n + v] = __ldg(GB.z + v);
                        }
                        if (g.lane < 4) {
                            const double *s = a.wl_simplex + 12 * item + 3 * g.lane;
                            P.vx[g.lane] = s[0]; P.vy[g.lane] = s[1]; P.vz[g.lane] = s[2];
                        }
                    }
                }
            }
            __syncwarp();
            // seed from the GJK tetrahedron: faces as F:279-293, wound outwards, neighbours by the opposite vertex
            bool bad = false;
            if (start && g.lane < 4) {
                const int l = g.lane;
                int fa = (l == 3) ? 1 : 0, fb = (l == 0 || l == 2) ? 1 : 2, fc = (l == 0) ? 2 : 3;
                const Vec3 M = v3(gjk::add(gjk::add(gjk::add(P.vx[0], P.vx[1]), P.vx[2]), P.vx[3]) / 4.0,
                                  gjk::add(gjk::add(gjk::add(P.vy[0], P.vy[1]), P.vy[2]), P.vy[3]) / 4.0,
                                  gjk::add(gjk::add(gjk::add(P.vz[0], P.vz[1]), P.vz[2]), P.vz[3]) / 4.0);
                double4 pl;
                const Vec3 va = v3(P.vx[fa], P.vy[fa], P.vz[fa]);
                for (int turn = 0; turn < 2; ++turn) {                        // one plane site; second turn = swapped winding
                    if (!tri_plane(va, v3(P.vx[fb], P.vy[fb], P.vz[fb]), v3(P.vx[fc], P.vy[fc], P.vz[fc]), pl)) { bad = true; break; }
                    if (turn == 1 || !(dot3(v3(pl.x, pl.y, pl.z), va - M) < 0.0)) break;
                    const int t = fb; fb = fc; fc = t;                         // inward: swap two vertices
                }
                P.pl[l] = pl;
                P.fv[l] = (unsigned)fa | ((unsigned)fb << 8) | ((unsigned)fc << 16);
                // face l lacks vertex 3,1,2,0; the neighbour across an edge lacks the face's third vertex
                const unsigned face_without = 3u | (1u << 8) | (2u << 16) | (0u << 24);   // vertex -> face id
                P.adj[l] = ((face_without >> (8 * fc)) & 0xffu) | (((face_without >> (8 * fa)) & 0xffu) << 8)
                         | (((face_without >> (8 * fb)) & 0xffu) << 16);
            }
            const bool anybad = g.any(bad);
            if (start) {
                busy = true;
                if (anybad) { st |= ST_STOP; retire(false, false, 0.0, v3(0, 0, 0)); }     // degenerate face: reference STOPs (F:1369)
            }
            __syncwarp();
        }
        if (!__any_sync(FULL, busy)) {
            if (__all_sync(FULL, drained)) break;
            continue;
        }

        // ---- one EPA iteration for every busy group, in lockstep ----
        bool run = busy;
        if (run) {
            ++iters;
            if (iters > MAX_ITERS) { st |= ST_EPA_CAP; iters = MAX_ITERS; retire(false, false, 0.0, v3(0, 0, 0)); run = false; }   // F:299-303
        }
        // closest face, F:893-898
        const int nf_run = run ? nf : 0;
        const int nfw = wmax(nf_run);
        double best = DBL_MAX; int k = 0;
        for (int f = g.lane; f < nfw; f += LP)
            if (f < nf_run) { const double ad = fabs(P.pl[f].w); if (ad < best) { best = ad; k = f; } }
        k = g.argmin_first_nonneg(best, k);
        const double4 pk = P.pl[k];
        const double dk = pk.w;
        best = fabs(dk);
        const Vec3 n = v3(pk.x, pk.y, pk.z);
        double dt = dk;                                         // (v1 - O) . dir, F:903
        const bool rebase = run && fabs(dt) < 1.0e-12;          // F:905-908: re-base on the slot centroid
        if (__any_sync(FULL, rebase)) {
            double sx = 0, sy = 0, sz = 0;
            for (int f = g.lane; f < nfw; f += LP) {
                if (f < nf_run) {
                    const unsigned w = P.fv[f];
                    const int i0 = w & 0xff, i1 = (w >> 8) & 0xff, i2 = (w >> 16) & 0xff;
                    sx += P.vx[i0] + P.vx[i1] + P.vx[i2]; sy += P.vy[i0] + P.vy[i1] + P.vy[i2]; sz += P.vz[i0] + P.vz[i1] + P.vz[i2];
                }
            }
            for (int o = LP / 2; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(FULL, sx, o, LP); sy += __shfl_xor_sync(FULL, sy, o, LP); sz += __shfl_xor_sync(FULL, sz, o, LP);
            }
            if (rebase) {
                const double cnt = (double)(nf * 3);
                const int i0 = P.fv[k] & 0xff;
                dt = dot3(v3(P.vx[i0], P.vy[i0], P.vz[i0]) - v3(sx / cnt, sy / cnt, sz / cnt), n);
            }
        }
        const Vec3 dir = (dt <= -1.0e-12) ? neg(n) : n;         // F:910

        int ch = 0;
        bool ins = run;
        for (int rep = 0; rep < 2; ++rep) {                     // F:914; once more along -dir when the origin is on the face (F:935-944)
            if (rep == 1) {
                ins = run && ch >= 0 && best < 1.0e-12;
                if (!__any_sync(FULL, ins)) break;
            }
            const Vec3 sp = group_support<LP>(g, A, B, rep ? neg(dir) : dir, ins);
            const int c1 = mesh_insert<LP>(g, P, nv, nf, sp, ins);
            if (ins) ch = (rep == 0 || c1 < 0) ? c1 : (ch | c1);
        }
        if (run) {
            if (ch == -2) retire(false, true, 0.0, v3(0, 0, 0));                 // next pass (larger mesh / general kernel)
            else if (ch < 0) { st |= ST_STOP; retire(false, false, 0.0, v3(0, 0, 0)); }
            else if (ch == 0) retire(true, false, best, (dk < 0.0) ? neg(n) : n); // hull unchanged: converged (F:994-998, F:967-969)
        }
    }
}

template <int LP, int VC, int FC, int HC>
inline int launch_one(const EpaArgs &a, int sm_count, cudaStream_t st)
{
    constexpr int GPW = 32 / LP;
    const size_t smem = sizeof(Mesh<VC, FC, HC>) * WARPS * GPW;
    cudaError_t e = cudaFuncSetAttribute(epa3_kernel<LP, VC, FC, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epa3_kernel<LP, VC, FC, HC>, WARPS * 32, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    epa3_kernel<LP, VC, FC, HC><<<sm_count * per_sm, WARPS * 32, smem, st>>>(a);
    return 0;
}

// ---- per-pair routing by hull size (counters below are indices into the context's counter block) ----
// The colliding pairs reach the EPA stage in the order the GJK warps finished them.  route_count / route_fill bin the
// work list by nA + nB: "small" (<= 64 vertices: the linked-mesh kernel with a 64-vertex staging buffer), "big"
// (<= 128: the same kernel with a 128-vertex buffer, 4 CTAs per SM instead of 5) and "huge" (the general kernel, which
// reads such hulls from global memory); inside small and big the items are grouped in four sub-bins of 16 vertices, so
// that the two pairs sharing a warp in lockstep are of similar size (their loops run to the larger trip count).
// Which kernel family solves a pair therefore depends on the PAIR, not on the batch it came in or on a caller's hint.
constexpr int R_BINS = 8;                      // 0-3 small (<=16, <=32, <=48, <=64), 4-7 big (<=80, .., <=128)
constexpr int C_BIN = 16, C_CUR = 24, C_NSMALL = 32, C_NBIG = 33, C_HEAD_SMALL = 34, C_HEAD_BIG = 35;   // counter slots

__device__ __forceinline__ int route_bin(const EpaArgs &a, long long item)
{
    const int p = a.wl_pair[item];
    const int ha = a.pair_a[p], hb = a.pair_b[p];
    const int s = (int)(a.hull_off[ha + 1] - a.hull_off[ha]) + (int)(a.hull_off[hb + 1] - a.hull_off[hb]);
    return s > 128 ? R_BINS : (s <= 0 ? 0 : (s - 1) >> 4);
}

__global__ void __launch_bounds__(256)
route_count_kernel(EpaArgs a, unsigned long long *counters, uint8_t *bins)
{
    const unsigned long long n = *a.wl_count;
    const int lane = threadIdx.x & 31;
    for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; i0 < n; i0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = i0 + lane;
        const int bin = i < n ? route_bin(a, (long long)i) : -1;
        if (bin >= 0) bins[i] = (uint8_t)bin;               // the filling pass reads this instead of chasing the indices again
        const unsigned peers = __match_any_sync(FULL, bin);
        if (bin >= 0 && bin < R_BINS && lane == __ffs(peers) - 1) atomicAdd(counters + C_BIN + bin, (unsigned long long)__popc(peers));
    }
}

// item numbers into small_list / big_list (grouped by sub-bin) and, for the huge class, straight into the general kernel's list
__global__ void __launch_bounds__(256)
route_fill_kernel(EpaArgs a, unsigned long long *counters, const uint8_t *bins, int32_t *small_list, int32_t *big_list, int32_t *huge_list,
                  unsigned long long *huge_count)
{
    const unsigned long long n = *a.wl_count;
    const int lane = threadIdx.x & 31;
    unsigned long long base[R_BINS];
    {
        unsigned long long s = 0;
#pragma unroll
        for (int b = 0; b < R_BINS; ++b) { if (b == 4) s = 0; base[b] = s; s += counters[C_BIN + b]; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counters[C_NSMALL] = counters[C_BIN] + counters[C_BIN + 1] + counters[C_BIN + 2] + counters[C_BIN + 3];
        counters[C_NBIG] = counters[C_BIN + 4] + counters[C_BIN + 5] + counters[C_BIN + 6] + counters[C_BIN + 7];
    }
    for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; i0 < n; i0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = i0 + lane;
        const int bin = i < n ? (int)bins[i] : -1;
        const unsigned peers = __match_any_sync(FULL, bin);
        if (bin < 0) continue;
        const int leader = __ffs(peers) - 1;
        unsigned long long at = 0;
        if (lane == leader) at = atomicAdd(bin < R_BINS ? counters + C_CUR + bin : huge_count, (unsigned long long)__popc(peers));
        at = __shfl_sync(peers, at, leader) + __popc(peers & ((1u << lane) - 1u));
        if (bin >= R_BINS) huge_list[at] = (int32_t)i;
        else if (bin < 4) small_list[base[bin] + at] = (int32_t)i;
        else big_list[base[bin] + at] = (int32_t)i;
    }
}

// EPA stage.  route -> linked-mesh kernel over the small list -> linked-mesh kernel (128-vertex staging) over the big list
// -> general kernel (64,128) over what they declined (capacity, a visible region that is not a disk) plus the huge class
// -> general kernel (128,256) = the reference's hard bound (4 + 99 iterations) for what is left.
// lists: route (2 x route_stride ints: small | big, then route_stride bytes of bin numbers), ovf1 (declined + huge), ovf2
// (declined by the (64,128) pass).  all_small: the store's hulls are known to have at most 32 vertices, so every pair is
// "small" and of one sub-bin width -- the same kernel the routing would pick, without the two routing passes.
// counters: [4] ovf1 count, [3] (64,128) head, [8] ovf2 count, [9] (128,256) head, [16..35] routing (above)
inline int launch(EpaArgs a, int sm_count, cudaStream_t st, int32_t *ovf1, int32_t *ovf2, int32_t *route, int64_t route_stride,
                  unsigned long long *counters, bool all_small, bool unrouted, int classes, int *launches)
{
    int rc = 0, nl = 0;
    a.item_list = nullptr; a.item_count = nullptr;
    EpaArgs s = a;
    s.queue_head = counters + C_HEAD_SMALL;
    s.ovf_items = ovf1; s.ovf_count = counters + 4;
#ifndef EPA3_LP
#define EPA3_LP 16
#endif
    if (all_small) {
        rc = launch_one<EPA3_LP, 36, 68, 64>(s, sm_count, st);
        if (rc) return rc;
        nl += 1;
    } else if (unrouted) {
        // small batch: both size classes scan the whole work list and take their own items (EpaArgs::take_gt / skip_large);
        // `classes` (bit 0 small, bit 1 big or huge) lets a caller that knows its hull sizes skip a pass that would find nothing
        if (classes & 1) {
            s.skip_large = true;
            rc = launch_one<EPA3_LP, 36, 68, 64>(s, sm_count, st);
            if (rc) return rc;
            ++nl;
        }
        if (classes & 2) {
            EpaArgs g = s;
            g.skip_large = false; g.take_gt = 64; g.queue_head = counters + C_HEAD_BIG;
            rc = launch_one<EPA3_LP, 36, 68, 128>(g, sm_count, st);
            if (rc) return rc;
            ++nl;
        }
    } else {
        const int rb = sm_count * 8;
        uint8_t *bins = reinterpret_cast<uint8_t *>(route + 2 * route_stride);
        route_count_kernel<<<rb, 256, 0, st>>>(a, counters, bins);
        route_fill_kernel<<<rb, 256, 0, st>>>(a, counters, bins, route, route + route_stride, ovf1, counters + 4);
        s.item_list = route; s.item_count = counters + C_NSMALL;
        rc = launch_one<EPA3_LP, 36, 68, 64>(s, sm_count, st);
        if (rc) return rc;
        EpaArgs g = s;
        g.item_list = route + route_stride; g.item_count = counters + C_NBIG; g.queue_head = counters + C_HEAD_BIG;
        rc = launch_one<EPA3_LP, 36, 68, 128>(g, sm_count, st);
        if (rc) return rc;
        nl += 4;
    }
    if (unrouted) {
        // small batch: what the linked-mesh passes declined goes straight to the general kernel with the reference's hard
        // bound (same code as the (64,128) form, larger arrays: same arithmetic), one launch less on the single-pair path
        EpaArgs c = a;
        c.item_list = ovf1; c.item_count = counters + 4;
        c.queue_head = counters + 9;
        c.ovf_items = nullptr; c.ovf_count = nullptr;
        rc = epa::launch_one<128, 256>(c, sm_count, st);
        if (launches) *launches = nl + 1;
        return rc;
    }
    EpaArgs b = a;
    b.item_list = ovf1; b.item_count = counters + 4;
    b.queue_head = counters + 3;
    b.ovf_items = ovf2; b.ovf_count = counters + 8;
    rc = epa::launch_one<64, 128>(b, sm_count, st);
    if (rc) return rc;
    EpaArgs c = a;
    c.item_list = ovf2; c.item_count = counters + 8;
    c.queue_head = counters + 9;
    c.ovf_items = nullptr; c.ovf_count = nullptr;
    rc = epa::launch_one<128, 256>(c, sm_count, st);
    if (launches) *launches = nl + 2;
    return rc;
}

}  // namespace epa3
