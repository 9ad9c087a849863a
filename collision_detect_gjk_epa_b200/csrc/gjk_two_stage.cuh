// gjk_two_stage.cuh -- the GJK flag path (reference src/GCLIB_GJKEPA.f90:76-236) as two kernels with
// iteration compaction between them, one THREAD per pair (sm_100a).
//
//   K1i gjk_init_kernel   sphere test + initial tetrahedron (gjk_init, F:76-157).  Every pair costs the
//                         same here, so a warp runs 32 pairs in lockstep.  Pairs that are decided (sphere
//                         miss, F:123, F:140, F:157) write their flag; the others are appended, warp-
//                         aggregated, to the SEED QUEUE: pair number, hull locations, tetrahedron, the unit
//                         normal of its face (1,2,3).
//   K1w gjk_walk_kernel   persistent warps; every lane walks one tetrahedron (walk_round, F:164-236).  The
//                         walk takes 1..12 rounds depending on the pair, so lanes finish at different
//                         times: whenever `refill` lanes of a warp are idle they are re-packed with the
//                         next seeds of the queue, while the other lanes keep their state in registers.
//                         The walk loop therefore runs with nearly full warps (the monolithic kernel ran
//                         every tile until its slowest pair had finished: 57 % lane efficiency, ncu).
//
// Shared-memory layout ("T2"): each warp owns a region of col_units x 512 bytes; 16-byte unit u of lane
// j's column sits at region + u*512 + j*16.  A lane fills its own column with 16-byte cp.async copies
// straight from the hull block [x(n) | y(n) | z(n)] (each axis padded to an even count by repeating its
// last element, which can never win a strict-> scan), so staging costs ~1 instruction per pair and
// unit, not a cooperative loop over the 32 pairs; one LDS.128 then delivers the same coordinate of two
// consecutive vertices and all 32 lanes read conflict-free.
#pragma once
#include "gjk_kernels.cuh"

namespace gjk {

constexpr int T2_PITCH = 512;            // bytes between consecutive 16-byte units of one lane's column

// This lane's copy of one hull inside its column.
struct HullViewT2 {
    const char *col;                     // shared-memory address of the hull's first unit, lane offset included
    int n, h;                            // vertices; units per axis = (n + 1) / 2
    __device__ __forceinline__ Vec3 get(int i) const
    {
        const char *p = col + (i >> 1) * T2_PITCH + (i & 1) * 8;
        return v3(*reinterpret_cast<const double *>(p), *reinterpret_cast<const double *>(p + h * T2_PITCH),
                  *reinterpret_cast<const double *>(p + 2 * h * T2_PITCH));
    }
};

__device__ __forceinline__ void scan_argmax(const HullViewT2 &hv, Vec3 d, int, int, double &best, int &idx)
{
    best = -DBL_MAX; idx = 0;
    const char *p = hv.col;
    const int ay = hv.h * T2_PITCH, az = 2 * ay;
#pragma unroll 2
    for (int u = 0; u < hv.h; ++u, p += T2_PITCH) {
        const double2 X = *reinterpret_cast<const double2 *>(p);
        const double2 Y = *reinterpret_cast<const double2 *>(p + ay);
        const double2 Z = *reinterpret_cast<const double2 *>(p + az);
        const double t0 = add(add(mul(d.x, X.x), mul(d.y, Y.x)), mul(d.z, Z.x));
        const double t1 = add(add(mul(d.x, X.y), mul(d.y, Y.y)), mul(d.z, Z.y));
        if (t0 > best) { best = t0; idx = 2 * u; }
        if (t1 > best) { best = t1; idx = 2 * u + 1; }         // the padded repeat of the last vertex ties, never wins
    }
}
__device__ __forceinline__ void scan_minmax(const HullViewT2 &hv, Vec3 d, int, int, double &hi, int &ihi, double &lo, int &ilo)
{
    hi = -DBL_MAX; lo = DBL_MAX; ihi = 0; ilo = 0;
    const char *p = hv.col;
    const int ay = hv.h * T2_PITCH, az = 2 * ay;
    for (int u = 0; u < hv.h; ++u, p += T2_PITCH) {
        const double2 X = *reinterpret_cast<const double2 *>(p);
        const double2 Y = *reinterpret_cast<const double2 *>(p + ay);
        const double2 Z = *reinterpret_cast<const double2 *>(p + az);
        const double t0 = add(add(mul(d.x, X.x), mul(d.y, Y.x)), mul(d.z, Z.x));
        const double t1 = add(add(mul(d.x, X.y), mul(d.y, Y.y)), mul(d.z, Z.y));
        if (t0 > hi) { hi = t0; ihi = 2 * u; }
        if (t0 < lo) { lo = t0; ilo = 2 * u; }
        if (t1 > hi) { hi = t1; ihi = 2 * u + 1; }
        if (t1 < lo) { lo = t1; ilo = 2 * u + 1; }
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

// Every `on` lane copies its own hull block into its own column, starting at `dst` (lanes run in lockstep,
// each on different data; `tmax` is the warp-wide maximum of 3*h).
__device__ __forceinline__ void stage_own_hull(char *dst, const double *src, int n, int h, bool on, int tmax)
{
    const unsigned FULL = 0xffffffffu;
    const bool lin = ((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);   // one linear 16-byte stream
    const int t_end = on ? 3 * h : 0;
    if (__all_sync(FULL, lin || !on)) {
        for (int t = 0; t < tmax; ++t)
            if (t < t_end) cp_async16(dst + t * T2_PITCH, src + 2 * t);
    } else {
        const int odd = n & 1;
        for (int t = 0; t < tmax; ++t) {
            if (t < t_end) {
                const int ax = (t >= h) + (t >= 2 * h);
                const int u = t - ax * h;
                const int e0 = 2 * t - ax * odd;                    // = ax*n + 2u
                const int e1 = (2 * u + 1 < n) ? e0 + 1 : e0;       // odd n: repeat the axis' last element
                cp_async8(reinterpret_cast<double *>(dst + t * T2_PITCH), src + e0);
                cp_async8(reinterpret_cast<double *>(dst + t * T2_PITCH) + 1, src + e1);
            }
        }
    }
}

// envelope sphere F:1175-1182 from a T2 column (the padded repeats are skipped: the centroid sum is exact-order)
__device__ __forceinline__ Sphere hull_sphere_t2(const HullViewT2 &hv)
{
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = 0; i < hv.n; ++i) { Vec3 v = hv.get(i); sx = add(sx, v.x); sy = add(sy, v.y); sz = add(sz, v.z); }
    const double dn = (double)hv.n;
    Sphere s; s.cx = sx / dn; s.cy = sy / dn; s.cz = sz / dn;
    double r2 = -1.0;
    const Vec3 c = v3(s.cx, s.cy, s.cz);
    for (int i = 0; i < hv.n; ++i) {
        Vec3 dd = hv.get(i) - c;
        double t = add(add(mul(dd.x, dd.x), mul(dd.y, dd.y)), mul(dd.z, dd.z));
        if (t > r2) r2 = t;
    }
    s.r = (hv.n > 0) ? sqrt(r2) : -DBL_MAX;
    return s;
}

// ---- seed queue: what K1i hands to K1w (structure of arrays, `cap` entries each) ----
constexpr int SEED_F64 = 16;             // S0..S3 xyz, unit normal of face (1,2,3) xyz, its |cross|
constexpr unsigned SEED_COLD = 1u;       // fln = flags (4 bits) | nA << 4 | nB << 18
struct SeedQueue {
    double   *f64;                       // SEED_F64 arrays of cap doubles
    int32_t  *pair;
    uint32_t *fln;
    int64_t  *oa, *ob;                   // hull offsets (vertex units), so K1w needs no dependent index loads
    unsigned long long *count;           // entries appended by K1i
    unsigned long long *head;            // next entry K1w hands out
    int64_t cap;
};

struct TwoStageArgs {
    BatchArgs b;
    SeedQueue q;
    int32_t col_units;                   // capacity of one column, 16-byte units
    int32_t refill;                      // K1w re-packs a warp when this many lanes are idle
};

__device__ __forceinline__ uint32_t pack_status_walk(int exit_path, const Walk &w)
{
    return (uint32_t)exit_path | ((uint32_t)w.iters << 8) | ((w.fl & WF_STOP) ? (1u << 25) : 0u)
         | (exit_path == EXIT_TRI_ORIGIN ? (1u << 28) : 0u) | ((w.fl & WF_COLD) ? (1u << 30) : 0u);
}

// append the colliding lanes' tetrahedra to the EPA work list: one atomic per warp
__device__ __forceinline__ void push_contacts(const BatchArgs &a, bool coll, int64_t p, const Vec3 *S, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned m = __ballot_sync(FULL, coll);
    if (!m || !a.wl_pair) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.wl_count, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL, base, 0);
    if (coll) {
        const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
        a.wl_pair[slot] = (int32_t)p;
        double *dst = a.wl_simplex + 12 * slot;
#pragma unroll
        for (int r = 0; r < 4; ++r) { dst[3 * r] = S[r].x; dst[3 * r + 1] = S[r].y; dst[3 * r + 2] = S[r].z; }
    }
}

#ifndef GJK_INIT_MIN_CTAS
#define GJK_INIT_MIN_CTAS 4
#endif
#ifndef GJK_WALK_MIN_CTAS
#define GJK_WALK_MIN_CTAS 4
#endif

// ------------------------------------------------------------------ K1i
struct TileMeta { int ha, hb, nA, nB; long long oa, ob; bool valid; };
__device__ __forceinline__ TileMeta load_tile_meta(const BatchArgs &a, int64_t tile, int64_t n_tiles, int lane)
{
    TileMeta m; m.ha = m.hb = m.nA = m.nB = 0; m.oa = m.ob = 0; m.valid = false;
    const int64_t p = tile * 32 + lane;
    if (tile < n_tiles && p < a.n_pairs) {
        m.valid = true;
        m.ha = __ldg(a.pair_a + p); m.hb = __ldg(a.pair_b + p);
        m.oa = __ldg(a.hull_off + m.ha); m.nA = (int)(__ldg(a.hull_off + m.ha + 1) - m.oa);
        m.ob = __ldg(a.hull_off + m.hb); m.nB = (int)(__ldg(a.hull_off + m.hb + 1) - m.ob);
    }
    return m;
}

template <bool FUSE>
__global__ void __launch_bounds__(TILE, GJK_INIT_MIN_CTAS)
gjk_init_kernel(TwoStageArgs ta)
{
    extern __shared__ __align__(16) char sm2[];
    const BatchArgs &a = ta.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    char *col = sm2 + (size_t)warp * ta.col_units * T2_PITCH + lane * 16;
    const Lanes<1> g;
    const int64_t n_tiles = (a.n_pairs + 31) / 32;
    const int64_t n_warps = (int64_t)gridDim.x * (TILE / 32);

    int64_t tile = (int64_t)blockIdx.x * (TILE / 32) + warp;
    TileMeta nx = load_tile_meta(a, tile, n_tiles, lane);
    for (; tile < n_tiles; tile += n_warps) {
        const TileMeta m = nx;
        const int64_t p = tile * 32 + lane;
        const int hA = (m.nA + 1) >> 1, hB = (m.nB + 1) >> 1;
        const bool fits = m.valid && 3 * (hA + hB) <= ta.col_units;
        const int tmaxA = __reduce_max_sync(FULL, fits ? 3 * hA : 0), tmaxB = __reduce_max_sync(FULL, fits ? 3 * hB : 0);
        stage_own_hull(col, a.verts + 3 * m.oa, m.nA, hA, fits, tmaxA);
        stage_own_hull(col + 3 * hA * T2_PITCH, a.verts + 3 * m.ob, m.nB, hB, fits, tmaxB);
        nx = load_tile_meta(a, tile + n_warps, n_tiles, lane);       // the next tile's index chain overlaps this tile's work
        cp_async_wait_all();
        {   // hulls too large for a column: leave to the sub-warp kernels (one atomic per warp)
            const unsigned big = __ballot_sync(FULL, m.valid && !fits);
            if (big) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.ovf_count, (unsigned long long)__popc(big));
                base = __shfl_sync(FULL, base, 0);
                if (m.valid && !fits) a.ovf_list[base + __popc(big & ((1u << lane) - 1u))] = (int32_t)p;
            }
        }
        Walk w;
        int e = -1;                                         // -1: nothing to report for this lane
        if (fits) {
            HullViewT2 A, B;
            A.col = col; A.n = m.nA; A.h = hA;
            B.col = col + 3 * hA * T2_PITCH; B.n = m.nB; B.h = hB;
            Sphere sa, sb;
            if (FUSE) { sa = hull_sphere_t2(A); sb = hull_sphere_t2(B); }
            else      { sa = a.sphere[m.ha];    sb = a.sphere[m.hb]; }
            e = gjk_init<1>(g, A, B, sa, sb, w);
            if (e > 0) {
                if (a.collision) a.collision[p] = exit_is_collision(e) ? 1 : 0;
                if (a.status) a.status[p] = pack_status_walk(e, w);
            }
        }
        __syncwarp();
        push_contacts(a, e > 0 && exit_is_collision(e), p, w.S, lane);      // F:140: the origin on the first triangle
        const unsigned go = __ballot_sync(FULL, e == 0);
        if (go) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(ta.q.count, (unsigned long long)__popc(go));
            base = __shfl_sync(FULL, base, 0);
            if (e == 0) {
                const unsigned long long s = base + __popc(go & ((1u << lane) - 1u));
                double *f = ta.q.f64 + s;
                const int64_t cap = ta.q.cap;
#pragma unroll
                for (int r = 0; r < 4; ++r) { f[(3 * r) * cap] = w.S[r].x; f[(3 * r + 1) * cap] = w.S[r].y; f[(3 * r + 2) * cap] = w.S[r].z; }
                f[12 * cap] = w.qk.x; f[13 * cap] = w.qk.y; f[14 * cap] = w.qk.z;
                f[15 * cap] = w.mk;
                ta.q.pair[s] = (int32_t)p;
                ta.q.fln[s] = ((w.fl & WF_COLD) ? SEED_COLD : 0u) | ((unsigned)m.nA << 4) | ((unsigned)m.nB << 18);
                ta.q.oa[s] = m.oa; ta.q.ob[s] = m.ob;
            }
        }
    }
}

// ------------------------------------------------------------------ K1w
__global__ void __launch_bounds__(TILE, GJK_WALK_MIN_CTAS)
gjk_walk_kernel(TwoStageArgs ta)
{
    extern __shared__ __align__(16) char sm2[];
    const BatchArgs &a = ta.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    char *col = sm2 + (size_t)warp * ta.col_units * T2_PITCH + lane * 16;
    const Lanes<1> g;
    const unsigned long long count = *ta.q.count;
    const int64_t cap = ta.q.cap;

    Walk w;
    HullViewT2 A, B;
    A.col = col; A.n = 0; A.h = 0; B = A;
    int64_t p = 0;
    bool active = false, exhausted = false;
    w.S[0] = w.S[1] = w.S[2] = w.S[3] = v3(0.0, 0.0, 0.0);
    w.qk = w.D1 = v3(0.0, 0.0, 0.0); w.mk = 0.0; w.k1 = 2; w.iters = 0; w.n_support = 0; w.fl = 0u;
    // the warp draws seeds from the queue 32 at a time; lane i keeps the small fields of entry bbase + i
    unsigned long long bbase = 0; int bcount = 0, bcur = 0;
    int m_p = 0; unsigned m_fln = 0; long long m_oa = 0, m_ob = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active);
        const int n_idle = __popc(idle);
        if (n_idle >= ta.refill && !(exhausted && bcur >= bcount)) {
            if (bcur >= bcount) {                                   // next batch of 32 seeds
                if (lane == 0) bbase = atomicAdd(ta.q.head, 32ull);
                bbase = __shfl_sync(FULL, bbase, 0);
                exhausted = bbase + 32ull >= count;
                bcount = bbase < count ? (int)((count - bbase) < 32ull ? (count - bbase) : 32ull) : 0;
                bcur = 0;
                if (lane < bcount) {
                    m_p = ta.q.pair[bbase + lane]; m_fln = ta.q.fln[bbase + lane];
                    m_oa = ta.q.oa[bbase + lane];  m_ob = ta.q.ob[bbase + lane];
                }
            }
            const int avail = bcount - bcur;
            if (avail > 0) {
                const int n_take = n_idle < avail ? n_idle : avail;
                const int rank = __popc(idle & lt);
                const bool take = !active && rank < n_take;
                const int src = (bcur + rank) & 31;
                const int sp_ = __shfl_sync(FULL, m_p, src);
                const unsigned sfln = __shfl_sync(FULL, m_fln, src);
                const long long soa = __shfl_sync(FULL, m_oa, src), sob = __shfl_sync(FULL, m_ob, src);
                const int nA = take ? (int)((sfln >> 4) & 0x3fffu) : 0, nB = take ? (int)(sfln >> 18) : 0;
                const int hA = (nA + 1) >> 1, hB = (nB + 1) >> 1;
                const int tmaxA = __reduce_max_sync(FULL, 3 * hA), tmaxB = __reduce_max_sync(FULL, 3 * hB);
                stage_own_hull(col, a.verts + 3 * soa, nA, hA, take, tmaxA);
                stage_own_hull(col + 3 * hA * T2_PITCH, a.verts + 3 * sob, nB, hB, take, tmaxB);
                if (take) {
                    const double *f = ta.q.f64 + (bbase + (unsigned long long)(bcur + rank));
#pragma unroll
                    for (int r = 0; r < 4; ++r) w.S[r] = v3(f[(3 * r) * cap], f[(3 * r + 1) * cap], f[(3 * r + 2) * cap]);
                    w.qk = v3(f[12 * cap], f[13 * cap], f[14 * cap]);
                    w.mk = f[15 * cap];
                    w.D1 = v3(0.0, 0.0, 0.0); w.k1 = 2; w.iters = 0; w.n_support = 4;
                    w.fl = WF_FIRST | ((sfln & SEED_COLD) ? WF_COLD : 0u);
                    p = sp_;
                    A.n = nA; A.h = hA;
                    B.col = col + 3 * hA * T2_PITCH; B.n = nB; B.h = hB;
                    active = true;
                }
                bcur += n_take;
                cp_async_wait_all();
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted && bcur >= bcount) break;
            continue;
        }
        // one round on every lane, finished or not: nothing is skipped, so no branch merges two walk states
        Pick pk;
        const int e1 = walk_test<1>(g, w, pk, active);
        const bool coll = active && e1 != 0 && exit_is_collision(e1);
        if (coll) {                                                 // park the tetrahedron in this lane's (now dead) column
            double2 *d = reinterpret_cast<double2 *>(col);
            d[0 * 32] = make_double2(w.S[0].x, w.S[0].y); d[1 * 32] = make_double2(w.S[0].z, w.S[1].x);
            d[2 * 32] = make_double2(w.S[1].y, w.S[1].z); d[3 * 32] = make_double2(w.S[2].x, w.S[2].y);
            d[4 * 32] = make_double2(w.S[2].z, w.S[3].x); d[5 * 32] = make_double2(w.S[3].y, w.S[3].z);
        }
        const int iters_out = w.iters;
        const int e2 = walk_step<1>(g, A, B, w, pk, active && e1 == 0);
        const int e = e1 ? e1 : e2;
        if (active && e) {
            if (a.collision) a.collision[p] = coll ? 1 : 0;
            if (a.status) a.status[p] = (uint32_t)e | ((uint32_t)iters_out << 8) | ((w.fl & WF_COLD) ? (1u << 30) : 0u);
            active = false;
        }
        __syncwarp();
        {   // EPA work list: one atomic per warp and round; the rows come back from the column
            const unsigned mcol = __ballot_sync(FULL, coll);
            if (mcol && a.wl_pair) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.wl_count, (unsigned long long)__popc(mcol));
                base = __shfl_sync(FULL, base, 0);
                if (coll) {
                    const unsigned long long slot = base + __popc(mcol & lt);
                    a.wl_pair[slot] = (int32_t)p;
                    double2 *dst = reinterpret_cast<double2 *>(a.wl_simplex + 12 * slot);
                    const double2 *srcp = reinterpret_cast<const double2 *>(col);
#pragma unroll
                    for (int r = 0; r < 6; ++r) dst[r] = srcp[r * 32];
                }
            }
        }
    }
}

}  // namespace gjk
