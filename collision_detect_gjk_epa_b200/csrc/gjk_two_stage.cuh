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

__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8_s(unsigned smem_addr, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gmem_src) : "memory");
}

// pull the 128-byte lines of [ptr, ptr + bytes) towards L2 (no register result, never waited for)
__device__ __forceinline__ void prefetch_l2_range(const void *ptr, int bytes)
{
    const char *p = reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(ptr) & ~(uintptr_t)127);
    const char *end = reinterpret_cast<const char *>(ptr) + bytes;
    for (; p < end; p += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Every `on` lane copies its own hull block into its own column, starting at `dst` (lanes run in lockstep,
// each on different data; `tmax` is the warp-wide maximum of 3*h).
__device__ __forceinline__ void stage_own_hull(char *dst, const double *src, int n, int h, bool on, int tmax)
{
    const unsigned FULL = 0xffffffffu;
    const bool lin = ((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);   // one linear 16-byte stream
    const int t_end = on ? 3 * h : 0;
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (__all_sync(FULL, lin || !on)) {
#pragma unroll 4
        for (int t = 0; t < tmax; ++t, d += T2_PITCH, src += 2)
            if (t < t_end) cp_async16_s(d, src);
    } else {
        const int odd = n & 1;
        for (int t = 0; t < tmax; ++t, d += T2_PITCH) {
            if (t < t_end) {
                const int ax = (t >= h) + (t >= 2 * h);
                const int u = t - ax * h;
                const int e0 = 2 * t - ax * odd;                    // = ax*n + 2u
                const int e1 = (2 * u + 1 < n) ? e0 + 1 : e0;       // odd n: repeat the axis' last element
                cp_async8_s(d, src + e0);
                cp_async8_s(d + 8, src + e1);
            }
        }
    }
}

// envelope sphere F:1175-1182 from a T2 column.  The centroid sum keeps the reference order (vertex by
// vertex, padded repeats skipped); the radius is a maximum, for which the repeats are harmless.
__device__ __forceinline__ Sphere hull_sphere_t2(const HullViewT2 &hv)
{
    double sx = 0.0, sy = 0.0, sz = 0.0;
    const int ay = hv.h * T2_PITCH, az = 2 * ay;
    const char *p = hv.col;
    for (int u = 0; u < (hv.n >> 1); ++u, p += T2_PITCH) {
        const double2 X = *reinterpret_cast<const double2 *>(p);
        const double2 Y = *reinterpret_cast<const double2 *>(p + ay);
        const double2 Z = *reinterpret_cast<const double2 *>(p + az);
        sx = add(add(sx, X.x), X.y); sy = add(add(sy, Y.x), Y.y); sz = add(add(sz, Z.x), Z.y);
    }
    if (hv.n & 1) {
        sx = add(sx, *reinterpret_cast<const double *>(p)); sy = add(sy, *reinterpret_cast<const double *>(p + ay));
        sz = add(sz, *reinterpret_cast<const double *>(p + az));
    }
    Sphere s;
    centroid_div(sx, sy, sz, hv.n, s);
    double r2 = -1.0;
    p = hv.col;
    for (int u = 0; u < hv.h; ++u, p += T2_PITCH) {
        const double2 X = *reinterpret_cast<const double2 *>(p);
        const double2 Y = *reinterpret_cast<const double2 *>(p + ay);
        const double2 Z = *reinterpret_cast<const double2 *>(p + az);
        const double dx0 = sub(X.x, s.cx), dy0 = sub(Y.x, s.cy), dz0 = sub(Z.x, s.cz);
        const double dx1 = sub(X.y, s.cx), dy1 = sub(Y.y, s.cy), dz1 = sub(Z.y, s.cz);
        const double t0 = add(add(mul(dx0, dx0), mul(dy0, dy0)), mul(dz0, dz0));
        const double t1 = add(add(mul(dx1, dx1), mul(dy1, dy1)), mul(dz1, dz1));
        if (t0 > r2) r2 = t0;
        if (t1 > r2) r2 = t1;
    }
    s.r = (hv.n > 0) ? sqrt(r2) : -DBL_MAX;
    return s;
}

// ---- seed queue: what K1i hands to K1w (structure of arrays, `cap` entries each) ----
// The four rows of the tetrahedron are A[i] - B[j] for vertex numbers K1i knows, and K1w stages the hulls anyway:
// the queue carries the eight vertex numbers (8 bytes) and K1w rebuilds the rows bit for bit from its staged copy,
// instead of 96 bytes of coordinates written and read back through HBM.  The unit normal of face (1,2,3) travels
// as computed (its straight-line / cold-redo history is not worth repeating): 64 bytes per seed, was 152.
constexpr int SEED_F64 = 4;              // unit normal of face (1,2,3) xyz, its |cross|
constexpr unsigned SEED_COLD = 1u;       // fln = flags (4 bits) | nA << 4 | nB << 18
struct SeedQueue {
    double   *f64;                       // SEED_F64 arrays of cap doubles
    uint2    *idx;                       // vertex numbers of the rows in hull A (.x) and hull B (.y), one byte per row
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
#ifndef GJK_INIT_THREADS
#define GJK_INIT_THREADS 128
#endif
constexpr int INIT_THREADS = GJK_INIT_THREADS;
#ifndef GJK_WALK_MIN_CTAS
#define GJK_WALK_MIN_CTAS 3
#endif

// ------------------------------------------------------------------ K1i
// Index chain of one tile, loaded in two steps so that neither waits inside the tile that issues it: the
// pair's hull numbers two tiles ahead, the hull offsets (which need the numbers) one tile ahead.
struct TileIds  { int ha, hb; int64_t p; bool valid; };
struct TileMeta {                        // raw loads only: the vertex counts are formed when the tile is worked,
    int ha, hb; int64_t p; long long oa, ea, ob, eb; bool valid;   // so nothing waits for these loads in the tile that issues them
    __device__ __forceinline__ int nA() const { return (int)(ea - oa); }
    __device__ __forceinline__ int nB() const { return (int)(eb - ob); }
};
__device__ __forceinline__ TileIds load_tile_ids(const BatchArgs &a, int64_t tile, int64_t n_tiles, int64_t n_items, int lane)
{
    TileIds m; m.ha = m.hb = 0;
    const int64_t q = tile * 32 + lane;
    m.valid = tile < n_tiles && q < n_items;
    m.p = q;
    if (m.valid) {
        if (a.order) m.p = (int64_t)(uint32_t)__ldg(a.order + q);      // visiting order (locality): one more link in a chain fetched two tiles ahead
        m.ha = __ldg(a.pair_a + m.p); m.hb = __ldg(a.pair_b + m.p);
    }
    return m;
}
__device__ __forceinline__ TileMeta load_tile_meta(const BatchArgs &a, const TileIds &id)
{
    TileMeta m; m.ha = id.ha; m.hb = id.hb; m.p = id.p; m.oa = m.ea = m.ob = m.eb = 0; m.valid = id.valid;
    if (id.valid) {
        m.oa = __ldg(a.hull_off + id.ha); m.ea = __ldg(a.hull_off + id.ha + 1);
        m.ob = __ldg(a.hull_off + id.hb); m.eb = __ldg(a.hull_off + id.hb + 1);
    }
    return m;
}

template <bool FUSE>
__global__ void __launch_bounds__(INIT_THREADS, GJK_INIT_MIN_CTAS)
gjk_init_kernel(TwoStageArgs ta)
{
    extern __shared__ __align__(16) char sm2[];
    const BatchArgs &a = ta.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    char *col = sm2 + (size_t)warp * ta.col_units * T2_PITCH + lane * 16;
    const Lanes<1> g;
    const int64_t n_items = (a.order && a.order_count) ? (int64_t)(*a.order_count) : a.n_pairs;      // a filtered batch visits its survivors
    const int64_t n_tiles = (n_items + 31) / 32;
    const int64_t n_warps = (int64_t)gridDim.x * (INIT_THREADS / 32);

    int64_t tile = (int64_t)blockIdx.x * (INIT_THREADS / 32) + warp;
    TileMeta cur = load_tile_meta(a, load_tile_ids(a, tile, n_tiles, n_items, lane));
    TileMeta nx = load_tile_meta(a, load_tile_ids(a, tile + n_warps, n_tiles, n_items, lane));
    TileIds nid = load_tile_ids(a, tile + 2 * n_warps, n_tiles, n_items, lane);
    // issue the staging copies of a tile into this warp's columns (the previous tile must be done reading them)
    auto stage = [&](const TileMeta &t) {
        const int nA = t.nA(), nB = t.nB();
        const int hA = (nA + 1) >> 1, hB = (nB + 1) >> 1;
        const bool fits = t.valid && 3 * (hA + hB) <= ta.col_units;
        const int tmaxA = __reduce_max_sync(FULL, fits ? 3 * hA : 0), tmaxB = __reduce_max_sync(FULL, fits ? 3 * hB : 0);
        stage_own_hull(col, a.verts + 3 * t.oa, nA, hA, fits, tmaxA);
        stage_own_hull(col + 3 * hA * T2_PITCH, a.verts + 3 * t.ob, nB, hB, fits, tmaxB);
    };
    stage(cur);
    // Software pipeline over the warp's tiles: the staging of tile t+1 is issued as soon as tile t has been computed (the
    // column is free then) and BEFORE tile t's results are written out -- the round trips of the queue atomics and the
    // seed stores run while the copies are in flight (ncu, round 2: 11.7 % of this kernel's stall samples sat on the
    // staging wait, 3.4 % on the atomics' round trip, with nothing between issue and wait).
    for (; tile < n_tiles; tile += n_warps) {
        const TileMeta m = cur;
        const int64_t p = m.p;
        const int nA = m.nA(), nB = m.nB();
        const int hA = (nA + 1) >> 1, hB = (nB + 1) >> 1;
        const bool fits = m.valid && 3 * (hA + hB) <= ta.col_units;
        // shared store: this tile's envelope spheres are gathered at random from the table K0 filled (32 B each; with a
        // store of millions of bodies a DRAM round trip) -- issued here so that they travel while the staging copies do
        Sphere sa, sb;
        sa.cx = sa.cy = sa.cz = sa.r = 0.0; sb = sa;
        if (!FUSE && !a.early_out && fits) { sa = a.sphere[m.ha]; sb = a.sphere[m.hb]; }
        if (nx.valid) {                                     // next tile: hull blocks (and spheres) on their way to L2 while this one is worked
            prefetch_l2_range(a.verts + 3 * nx.oa, 24 * nx.nA());
            prefetch_l2_range(a.verts + 3 * nx.ob, 24 * nx.nB());
            if (!FUSE && !a.early_out) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.sphere + nx.ha));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.sphere + nx.hb));
            }
        }
        cp_async_wait_all();
        Walk w;
        int e = -1;                                         // -1: nothing to report for this lane
        if (fits) {
            HullViewT2 A, B;
            A.col = col; A.n = nA; A.h = hA;
            B.col = col + 3 * hA * T2_PITCH; B.n = nB; B.h = hB;
            if (FUSE && !a.early_out) { sa = hull_sphere_t2(A); sb = hull_sphere_t2(B); }     // E1 skips the sphere test (gjk_init)
            e = gjk_init(g, A, B, sa, sb, w, a.early_out != 0);
        }
        __syncwarp();                                       // every lane is done with its column
        cur = nx;
        nx = load_tile_meta(a, nid);                        // the following tiles' index chain overlaps the write-out below
        nid = load_tile_ids(a, tile + 3 * n_warps, n_tiles, n_items, lane);
        stage(cur);                                         // next tile's copies in flight from here on (no-op past the last tile)
        // ---- write-out of tile t ----
        if (e > 0) {
            if (a.collision) a.collision[p] = exit_is_collision(e) ? 1 : 0;
            if (a.status) a.status[p] = pack_status_walk(e, w);
        }
        {   // hulls too large for a column: leave to the sub-warp kernels (one atomic per warp)
            const unsigned big = __ballot_sync(FULL, m.valid && !fits);
            if (big) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.ovf_count, (unsigned long long)__popc(big));
                base = __shfl_sync(FULL, base, 0);
                if (m.valid && !fits) a.ovf_list[base + __popc(big & ((1u << lane) - 1u))] = (int32_t)p;
                if (m.valid && !fits && a.ovf_status) { if (a.status) a.status[p] = a.ovf_status; if (a.collision) a.collision[p] = 0; }
            }
        }
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
                f[0] = w.qk.x; f[cap] = w.qk.y; f[2 * cap] = w.qk.z; f[3 * cap] = w.mk;
                ta.q.idx[s] = make_uint2(w.ia, w.ib);
                ta.q.pair[s] = (int32_t)p;
                ta.q.fln[s] = ((w.fl & WF_COLD) ? SEED_COLD : 0u) | ((unsigned)nA << 4) | ((unsigned)nB << 18);
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
    bool active = false;
    w.S[0] = w.S[1] = w.S[2] = w.S[3] = v3(0.0, 0.0, 0.0);
    w.qk = w.D1 = v3(0.0, 0.0, 0.0); w.mk = 0.0; w.k1 = 2; w.iters = 0; w.n_support = 0; w.fl = 0u;
    // The warp draws seeds from the queue 32 at a time; lane i keeps the small fields of entry base + i.  Two
    // batches are held: `cur` is being handed out, `nxt` was fetched one batch earlier and its hull blocks and
    // seed records have been prefetched to L2, so a re-pack waits for L2, not for HBM behind two index loads.
    struct Batch { unsigned long long base; int n; int p; unsigned fln; long long oa, ob; };
    auto fetch = [&](Batch &bt) {
        if (lane == 0) bt.base = atomicAdd(ta.q.head, 32ull);
        bt.base = __shfl_sync(FULL, bt.base, 0);
        bt.n = bt.base < count ? (int)((count - bt.base) < 32ull ? (count - bt.base) : 32ull) : 0;
        bt.p = 0; bt.fln = 0u; bt.oa = 0; bt.ob = 0;
        if (lane < bt.n) {
            bt.p = ta.q.pair[bt.base + lane]; bt.fln = ta.q.fln[bt.base + lane];
            bt.oa = ta.q.oa[bt.base + lane];  bt.ob = ta.q.ob[bt.base + lane];
        }
    };
    auto prefetch = [&](const Batch &bt) {
        if (lane < bt.n) {
            prefetch_l2_range(a.verts + 3 * bt.oa, 24 * (int)((bt.fln >> 4) & 0x3fffu));
            prefetch_l2_range(a.verts + 3 * bt.ob, 24 * (int)(bt.fln >> 18));
        }
        if (bt.n > 0 && lane < 2 * SEED_F64)                        // seed records: 4 fields x 256 bytes, one line per lane
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ta.q.f64 + (size_t)(lane >> 1) * cap + bt.base + (lane & 1) * 16));
        if (bt.n > 0 && (lane == 8 || lane == 9)) asm volatile("prefetch.global.L2 [%0];" ::"l"(ta.q.idx + bt.base + (lane & 1) * 16));
    };
    Batch cur, nxt;
    fetch(cur); prefetch(cur);
    fetch(nxt);
    bool nxt_pf = false;
    int bcur = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active);
        const int n_idle = __popc(idle);
        if (n_idle >= ta.refill && (bcur < cur.n || nxt.n > 0)) {
            if (!nxt_pf) { prefetch(nxt); nxt_pf = true; }
            if (bcur >= cur.n) { cur = nxt; bcur = 0; fetch(nxt); nxt_pf = false; }
            const int avail = cur.n - bcur;
            if (avail > 0) {
                const int n_take = n_idle < avail ? n_idle : avail;
                const int rank = __popc(idle & lt);
                const bool take = !active && rank < n_take;
                const int src = (bcur + rank) & 31;
                const int sp_ = __shfl_sync(FULL, cur.p, src);
                const unsigned sfln = __shfl_sync(FULL, cur.fln, src);
                const long long soa = __shfl_sync(FULL, cur.oa, src), sob = __shfl_sync(FULL, cur.ob, src);
                const int nA = take ? (int)((sfln >> 4) & 0x3fffu) : 0, nB = take ? (int)(sfln >> 18) : 0;
                const int hA = (nA + 1) >> 1, hB = (nB + 1) >> 1;
                const int tmaxA = __reduce_max_sync(FULL, 3 * hA), tmaxB = __reduce_max_sync(FULL, 3 * hB);
                uint2 t2 = make_uint2(0u, 0u);
                stage_own_hull(col, a.verts + 3 * soa, nA, hA, take, tmaxA);
                stage_own_hull(col + 3 * hA * T2_PITCH, a.verts + 3 * sob, nB, hB, take, tmaxB);
                if (take) {
                    const double *f = ta.q.f64 + (cur.base + (unsigned long long)(bcur + rank));
                    w.qk = v3(f[0], f[cap], f[2 * cap]);
                    w.mk = f[3 * cap];
                    t2 = ta.q.idx[cur.base + (unsigned long long)(bcur + rank)];
                    w.D1 = v3(0.0, 0.0, 0.0); w.k1 = 2; w.iters = 0; w.n_support = 4;
                    w.fl = WF_FIRST | ((sfln & SEED_COLD) ? WF_COLD : 0u) | (a.early_out ? WF_EARLY : 0u);
                    p = sp_;
                    A.n = nA; A.h = hA;
                    B.col = col + 3 * hA * T2_PITCH; B.n = nB; B.h = hB;
                    active = true;
                }
                bcur += n_take;
                cp_async_wait_all();
                if (take) {                                         // the tetrahedron K1i ended with, from the staged hulls
#pragma unroll
                    for (int r = 0; r < 4; ++r) w.S[r] = A.get((int)((t2.x >> (8 * r)) & 0xffu)) - B.get((int)((t2.y >> (8 * r)) & 0xffu));
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (bcur >= cur.n && nxt.n == 0) break;
            continue;
        }
        // one round on every lane, finished or not: nothing is skipped, so no branch merges two walk states
        Pick pk;
        const int e1 = walk_test(g, w, pk, active);
        const bool coll = active && e1 != 0 && exit_is_collision(e1);
        if (coll) {                                                 // park the tetrahedron in this lane's (now dead) column
            double2 *d = reinterpret_cast<double2 *>(col);
            d[0 * 32] = make_double2(w.S[0].x, w.S[0].y); d[1 * 32] = make_double2(w.S[0].z, w.S[1].x);
            d[2 * 32] = make_double2(w.S[1].y, w.S[1].z); d[3 * 32] = make_double2(w.S[2].x, w.S[2].y);
            d[4 * 32] = make_double2(w.S[2].z, w.S[3].x); d[5 * 32] = make_double2(w.S[3].y, w.S[3].z);
        }
        const int iters_out = w.iters;
        const int e2 = walk_step(g, A, B, w, pk, active && e1 == 0);
        const int e = e1 ? e1 : e2;
        if (active && e) {
            if (a.collision) a.collision[p] = coll ? 1 : 0;
            if (a.status) a.status[p] = (uint32_t)e | ((uint32_t)iters_out << 8) | ((w.fl & WF_COLD) ? (1u << 30) : 0u) | (e == EXIT_STOP ? (1u << 25) : 0u);
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
