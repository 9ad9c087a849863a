// gjk_kernels.cuh -- K0 hull-prep, K1a/K1b GJK flag kernels, K2 work-list compaction (sm_100a).
//
// K0  hull_prep_kernel   per-hull envelope sphere of RoughCollisionDetection_SphericalEnvelope
//                        (reference src/GCLIB_GJKEPA.f90:1175-1182).  The centroid is an
//                        order-dependent sequential sum, so it is computed once per hull.  Used when
//                        hulls are shared by many pairs; for pair-private hulls the GJK kernels
//                        compute the spheres themselves from the staged vertices (FUSE = true).
// K1a gjk_tile_kernel    GJKEPA flag path (F:76-236), one THREAD per pair.  A warp stages the
//                        vertices of its 32 pairs into shared memory with coalesced cp.async, one
//                        column per pair (element e of pair j at region[e*33 + j]): every support scan
//                        then reads conflict-free shared memory instead of 32 scattered L1 lines per
//                        load (ncu, profiles/: the un-staged form was L1-wavefront bound).  Pairs
//                        whose hulls do not fit a column go to the overflow list.
// K1b gjk_group_kernel<L> same path, a sub-warp of L lanes per pair (support = (value,index) max-dot
//                        shuffle reduction), vertices read through L1; runs the overflow list (large
//                        hulls) or, when the tile kernel is switched off, the whole batch.
// K2  colliding pairs are appended with their terminating tetrahedron to the EPA work list by both
//     kernels (one atomic per pair, after the pair's walk).
#pragma once
#include "gjk_core.cuh"

namespace gjk {

constexpr int TILE = 128;                // threads per CTA of the tile kernel (4 warps x 32 pairs)
constexpr int TILE_STRIDE = 33;          // odd column pitch of a warp's region: conflict-free transposed stores

struct BatchArgs {
    const double  *verts;       // packed hull blocks
    const int64_t *hull_off;    // n_hulls+1, vertex units
    const Sphere  *sphere;      // per hull (K0 output); unused when FUSE
    const int32_t *pair_a, *pair_b;
    int64_t n_pairs;
    int64_t n_hulls;
    uint8_t  *collision;        // may be null
    uint32_t *status;           // may be null
    // optional indirection: process pair_list[0 .. *pair_list_count) instead of 0..n_pairs
    const int32_t *pair_list;
    const unsigned long long *pair_list_count;
    // overflow list written by the tile kernel
    int32_t *ovf_list;
    unsigned long long *ovf_count;
    int32_t early_out;          // 1 = shortcut E1 (gjk_core.cuh): exit as soon as a support value proves separation
    uint32_t ovf_status;        // 0: a later pass works the overflow list; else no pass follows (the caller's bound on hull
                                // sizes promised there would be none): the pair gets this status word and flag 0
    int32_t col_doubles;        // capacity of one shared-memory column
    // optional visiting order of the thread-per-pair init kernel: entry q names pair (int32)order[q] (the low word of a sorted
    // 64-bit key).  Results are written at the pair's own index, so the order only decides which hulls are in L2 together.
    const unsigned long long *order;
    const unsigned long long *order_count;      // with `order`: how many entries it has (device; null = n_pairs)
    // EPA work list (null when not wanted)
    int32_t *wl_pair;
    double  *wl_simplex;        // 12 doubles per entry, row-major rows 0..3
    unsigned long long *wl_count;
};

__device__ __forceinline__ HullView hull_view(const double *verts, const int64_t *hull_off, int h)
{
    int64_t o = __ldg(hull_off + h);
    HullView v;
    v.n = (int)(__ldg(hull_off + h + 1) - o);
    v.x = verts + 3 * o; v.y = v.x + v.n; v.z = v.y + v.n;
    return v;
}

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// A hull block [x(n) | y(n) | z(n)] behind plain loads (shared memory or global).
struct HullViewPlain {
    const double *x, *y, *z;
    int n;
    __device__ __forceinline__ Vec3 get(int i) const { return v3(x[i], y[i], z[i]); }
};

// K0.  A warp takes 32 consecutive hulls.  Their blocks are one contiguous span of the store, so the warp copies them
// with coalesced loads into shared memory -- hull j shifted by j doubles, which makes the lanes' strides odd and their
// later reads conflict-free for uniform sizes -- and every lane then sums ITS hull in the reference's order (F:1175)
// from there.  (Thread per hull straight from global touched 32 different blocks per load instruction: 14.5 % of the
// DRAM peak, `lg_throttle` 123 per issue, ncu.)  Spans beyond the buffer (average hull > 16 vertices) take that
// direct path.
constexpr int PREP_WARPS = 2;
constexpr int PREP_CAP = 32 * 3 * 16 + 32;          // doubles per warp
__global__ void __launch_bounds__(PREP_WARPS * 32)
hull_prep_kernel(const double *__restrict__ verts, const int64_t *__restrict__ hull_off,
                 int64_t n_hulls, Sphere *__restrict__ out)
{
    __shared__ double sm[PREP_WARPS][PREP_CAP];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t h0 = ((int64_t)blockIdx.x * PREP_WARPS + warp) * 32;
    if (h0 >= n_hulls) return;
    const int64_t h = h0 + lane;
    const bool valid = h < n_hulls;
    const long long o = valid ? __ldg(hull_off + h) : 0, e = valid ? __ldg(hull_off + h + 1) : 0;
    const long long o0 = __shfl_sync(FULL, o, 0);
    const int n = (int)(e - o);
    const int cnt = __popc(__ballot_sync(FULL, valid));
    const long long e_last = __shfl_sync(FULL, e, cnt - 1);
    const long long span = 3 * (e_last - o0);
    if (span + 32 > PREP_CAP) {                          // uniform per warp
        if (valid) out[h] = hull_sphere(hull_view(verts, hull_off, (int)h));
        return;
    }
    double *buf = sm[warp];
    for (int j = 0; j < cnt; ++j) {
        const long long oj = __shfl_sync(FULL, o, j);
        const int nj3 = 3 * __shfl_sync(FULL, n, j);
        const double *src = verts + 3 * oj;
        double *dst = buf + 3 * (oj - o0) + j;
        for (int t = lane; t < nj3; t += 32) cp_async8(dst + t, src + t);       // all in flight at once: nothing waits per hull
    }
    cp_async_wait_all();
    __syncwarp();
    if (valid) {
        HullViewPlain v;
        v.x = buf + 3 * (o - o0) + lane; v.y = v.x + n; v.z = v.y + n; v.n = n;
        out[h] = hull_sphere(v);
    }
}

// envelope sphere by a sub-warp: lane l sums axis l sequentially (the order matters), the radius is a
// max over vertices (order-free)
template <class G, class HV>
__device__ __forceinline__ Sphere group_sphere(const G &g, const HV &h)
{
    constexpr int L = G::LANES;
    double c[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if ((ax % L) == g.lane) {
            double s = 0.0;
            for (int i = 0; i < h.n; ++i) s = add(s, h.comp(ax, i));
            c[ax] = s / (double)h.n;
        }
    }
    Sphere s;
    s.cx = g.bcast(c[0], 0 % L); s.cy = g.bcast(c[1], 1 % L); s.cz = g.bcast(c[2], 2 % L);
    Vec3 cc = v3(s.cx, s.cy, s.cz);
    double r2 = -1.0;
    for (int i = g.lane; i < h.n; i += L) {
        Vec3 d = h.get(i) - cc;
        double t = add(add(mul(d.x, d.x), mul(d.y, d.y)), mul(d.z, d.z));
        if (t > r2) r2 = t;
    }
    if constexpr (L > 1) {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) { double t = __shfl_xor_sync(g.mask, r2, o, L); if (t > r2) r2 = t; }
    }
    s.r = (h.n > 0) ? sqrt(r2) : -DBL_MAX;
    return s;
}

__device__ __forceinline__ uint32_t pack_status(const GjkOut &o)
{
    return (uint32_t)o.exit_path | ((uint32_t)o.iters << 8) | (o.stop_flag ? (1u << 25) : 0u)
         | (o.exit_path == EXIT_TRI_ORIGIN ? (1u << 28) : 0u) | (o.cold ? (1u << 30) : 0u);
}

// ------------------------------------------------------------------ K1a: one thread per pair, staged hulls

// Each WARP owns a tile of 32 pairs and a private shared-memory region of col_doubles x 33 doubles:
// element e of pair j sits at region[e*33 + j].  Staging is asynchronous (cp.async, LDGSTS), lanes
// over consecutive doubles of one hull block (coalesced global reads, conflict-free transposed
// writes thanks to the odd pitch); the walk then reads region[e*33 + lane] (conflict-free).  No
// block-level barrier anywhere: warps never wait for each other's slowest pair.
#ifndef GJK_TILE_MIN_CTAS
#define GJK_TILE_MIN_CTAS 3
#endif
template <bool FUSE>
__global__ void __launch_bounds__(TILE, GJK_TILE_MIN_CTAS)
gjk_tile_kernel(BatchArgs a)
{
    extern __shared__ __align__(16) double sm[];          // (TILE/32) regions of col_doubles * WSTRIDE
    constexpr int WSTRIDE = 33;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    double *region = sm + (size_t)warp * a.col_doubles * WSTRIDE;
    const Lanes<1> g;
    const int64_t n_tiles = (a.n_pairs + 31) / 32;
    const int64_t n_warps = (int64_t)gridDim.x * (TILE / 32);

    for (int64_t tile = (int64_t)blockIdx.x * (TILE / 32) + warp; tile < n_tiles; tile += n_warps) {
        const int64_t p = tile * 32 + lane;
        const bool valid = p < a.n_pairs;
        int ha = 0, hb = 0, nA = 0, nB = 0;
        long long oa = 0, ob = 0;
        if (valid) {
            ha = __ldg(a.pair_a + p); hb = __ldg(a.pair_b + p);
            oa = __ldg(a.hull_off + ha); nA = (int)(__ldg(a.hull_off + ha + 1) - oa);
            ob = __ldg(a.hull_off + hb); nB = (int)(__ldg(a.hull_off + hb + 1) - ob);
        }
        const bool fits = valid && 3 * (nA + nB) <= a.col_doubles;
        const int mA = fits ? 3 * nA : 0, mB = fits ? 3 * nB : 0;
        __syncwarp();                                      // the previous tile's columns are no longer read
        for (int j = 0; j < 32; ++j) {
            const int ja = __shfl_sync(FULL, mA, j), jb = __shfl_sync(FULL, mB, j);
            const long long joa = __shfl_sync(FULL, oa, j), job = __shfl_sync(FULL, ob, j);
            const double *srcA = a.verts + 3 * joa, *srcB = a.verts + 3 * job;
            double *colA = region + j, *colB = region + ja * WSTRIDE + j;
            for (int e = lane; e < ja; e += 32) cp_async8(colA + e * WSTRIDE, srcA + e);
            for (int e = lane; e < jb; e += 32) cp_async8(colB + e * WSTRIDE, srcB + e);
        }
        cp_async_wait_all();
        __syncwarp();
        {   // large hulls: leave to the sub-warp kernels (one atomic per warp)
            const unsigned big = __ballot_sync(FULL, valid && !fits);
            if (big) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.ovf_count, (unsigned long long)__popc(big));
                base = __shfl_sync(FULL, base, 0);
                if (valid && !fits) a.ovf_list[base + __popc(big & ((1u << lane) - 1u))] = (int32_t)p;
                if (valid && !fits && a.ovf_status) { if (a.status) a.status[p] = a.ovf_status; if (a.collision) a.collision[p] = 0; }
            }
        }
        if (fits) {
            HullViewStrided<WSTRIDE> A, B;
            A.base = region + lane; A.n = nA;
            B.base = region + 3 * nA * WSTRIDE + lane; B.n = nB;
            Sphere sa, sb;
            if (FUSE) { sa = hull_sphere(A); sb = hull_sphere(B); }
            else      { sa = a.sphere[ha];   sb = a.sphere[hb]; }
            Vec3 S[4];
            const GjkOut o = gjk_pair(g, A, B, sa, sb, S, a.early_out != 0);
            const bool coll = exit_is_collision(o.exit_path);
            if (a.collision) a.collision[p] = coll ? 1 : 0;
            if (a.status) a.status[p] = pack_status(o);
            if (coll && a.wl_pair) {
                unsigned long long slot = atomicAdd(a.wl_count, 1ull);
                a.wl_pair[slot] = (int32_t)p;
                double *dst = a.wl_simplex + 12 * slot;
#pragma unroll
                for (int r = 0; r < 4; ++r) { dst[3 * r] = S[r].x; dst[3 * r + 1] = S[r].y; dst[3 * r + 2] = S[r].z; }
            }
        }
    }
}

// ------------------------------------------------------------------ K1b: a sub-warp of L lanes per pair
template <int L, bool FUSE>
__global__ void __launch_bounds__(128)
gjk_group_kernel(BatchArgs a)
{
    const Lanes<L> g;
    const int64_t n_groups = (int64_t)gridDim.x * (blockDim.x / L);
    const int64_t n_items = a.pair_list ? (int64_t)(*a.pair_list_count) : a.n_pairs;
    int64_t it = (int64_t)blockIdx.x * (blockDim.x / L) + threadIdx.x / L;
    for (; it < n_items; it += n_groups) {
        const int64_t p = a.pair_list ? (int64_t)a.pair_list[it] : it;
        const int ha = __ldg(a.pair_a + p), hb = __ldg(a.pair_b + p);
        const HullView A = hull_view(a.verts, a.hull_off, ha);
        const HullView B = hull_view(a.verts, a.hull_off, hb);
        Sphere sa, sb;
        if (FUSE) { sa = group_sphere(g, A); sb = group_sphere(g, B); }
        else      { sa = a.sphere[ha];          sb = a.sphere[hb]; }
        Vec3 S[4];
        const GjkOut o = gjk_pair(g, A, B, sa, sb, S, a.early_out != 0);
        const bool coll = exit_is_collision(o.exit_path);
        if (g.lane == 0) {
            if (a.collision) a.collision[p] = coll ? 1 : 0;
            if (a.status) a.status[p] = pack_status(o);
        }
        if (coll && a.wl_pair) {
            unsigned long long slot = 0;
            if (g.lane == 0) slot = atomicAdd(a.wl_count, 1ull);
            slot = ((unsigned long long)(unsigned)g.bcast_i((int)(slot >> 32), 0) << 32)
                 | (unsigned)g.bcast_i((int)(slot & 0xffffffffu), 0);
            if (g.lane == 0) a.wl_pair[slot] = (int32_t)p;
            double *dst = a.wl_simplex + 12 * slot;
#pragma unroll
            for (int r = g.lane; r < 4; r += L) {
                Vec3 row = pick4(S, r);
                dst[3 * r + 0] = row.x; dst[3 * r + 1] = row.y; dst[3 * r + 2] = row.z;
            }
        }
    }
}


// ------------------------------------------------------------------ K1c: sub-warp per pair over TMA-staged hulls
// Shared-memory view of a hull block copied verbatim ([x(n) | y(n) | z(n)]).
struct HullViewShared {
    const double *x, *y, *z; int n;
    __device__ __forceinline__ Vec3 get(int i) const { return v3(x[i], y[i], z[i]); }
    __device__ __forceinline__ double comp(int ax, int i) const { return (ax == 0 ? x : (ax == 1 ? y : z))[i]; }
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async()
{ asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Each warp owns a tile of 32/L pairs; every pair has a slot of slot_doubles in the warp's region.
// Lane 0 of each group issues the pair's two hull blocks as 1-D TMA bulk copies (16-byte aligned
// blocks) or the group copies them element-wise with cp.async (odd vertex counts / offsets); the
// walk then reads shared memory.  Slots are padded by 4 doubles so that the 32/L groups of a warp
// start 32 bytes apart modulo the bank window (conflict-free lane-strided vertex reads).
template <int L, bool FUSE>
__global__ void __launch_bounds__(128)
gjk_group_staged_kernel(BatchArgs a, int slot_doubles)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(8) unsigned long long s_bar[4];
    constexpr int PPW = 32 / L;                          // pairs per warp
    const Lanes<L> g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / L;
    const unsigned FULL = 0xffffffffu;
    double *slot = sm + ((size_t)warp * PPW + grp) * slot_doubles;
    unsigned long long *bar = &s_bar[warp];
    if (lane == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncwarp();
    unsigned parity = 0;

    const int64_t n_items = a.pair_list ? (int64_t)(*a.pair_list_count) : a.n_pairs;
    const int64_t n_tiles = (n_items + PPW - 1) / PPW;
    const int64_t n_warps = (int64_t)gridDim.x * 4;
    for (int64_t tile = (int64_t)blockIdx.x * 4 + warp; tile < n_tiles; tile += n_warps) {
        const int64_t it = tile * PPW + grp;
        const bool valid = it < n_items;
        int64_t p = 0; int ha = 0, hb = 0;
        HullView GA, GB; GA.n = GB.n = 0; GA.x = GB.x = a.verts;
        if (valid) {
            p = a.pair_list ? (int64_t)a.pair_list[it] : it;
            ha = __ldg(a.pair_a + p); hb = __ldg(a.pair_b + p);
            GA = hull_view(a.verts, a.hull_off, ha); GB = hull_view(a.verts, a.hull_off, hb);
        }
        const bool fits = valid && 3 * (GA.n + GB.n) + 4 <= slot_doubles;
        const unsigned bytesA = fits ? 24u * (unsigned)GA.n : 0u, bytesB = fits ? 24u * (unsigned)GB.n : 0u;
        const bool bulkA = bytesA && ((bytesA & 15u) == 0) && ((reinterpret_cast<uintptr_t>(GA.x) & 15u) == 0);
        // B lands right after A: its destination is 16-byte aligned only if A's block is a multiple of 16 bytes
        const bool bulkB = bytesB && ((bytesB & 15u) == 0) && ((bytesA & 15u) == 0) && ((reinterpret_cast<uintptr_t>(GB.x) & 15u) == 0);
        __syncwarp();                                     // the previous tile's slots are no longer read
        fence_proxy_async();                              // order those generic reads before the async writes
        unsigned tx = (g.lane == 0) ? ((bulkA ? bytesA : 0u) + (bulkB ? bytesB : 0u)) : 0u;
        tx = __reduce_add_sync(FULL, tx);
        if (lane == 0) mbar_expect_tx(bar, tx);
        __syncwarp();
        if (g.lane == 0) {
            if (bulkA) bulk_g2s(slot, GA.x, bytesA, bar);
            if (bulkB) bulk_g2s(slot + 3 * GA.n, GB.x, bytesB, bar);
        }
        if (bytesA && !bulkA) for (int e = g.lane; e < 3 * GA.n; e += L) cp_async8(slot + e, GA.x + e);
        if (bytesB && !bulkB) for (int e = g.lane; e < 3 * GB.n; e += L) cp_async8(slot + 3 * GA.n + e, GB.x + e);
        cp_async_wait_all();
        {
            unsigned spins = 0;
            while (!mbar_try_wait(bar, parity)) { if (++spins > 64u) __nanosleep(32); }     // no poll limit (see gjk_two_stage_group.cuh)
            parity ^= 1u;
        }
        __syncwarp();
        if (valid && !fits && g.lane == 0) {             // larger than a slot: plain sub-warp kernel
            unsigned long long s2 = atomicAdd(a.ovf_count, 1ull);
            a.ovf_list[s2] = (int32_t)p;
            if (a.ovf_status) { if (a.status) a.status[p] = a.ovf_status; if (a.collision) a.collision[p] = 0; }
        }
        if (fits) {
            HullViewShared A, B;
            A.x = slot; A.y = slot + GA.n; A.z = slot + 2 * GA.n; A.n = GA.n;
            B.x = slot + 3 * GA.n; B.y = B.x + GB.n; B.z = B.y + GB.n; B.n = GB.n;
            Sphere sa, sb;
            if (FUSE) { sa = group_sphere(g, A); sb = group_sphere(g, B); }
            else      { sa = a.sphere[ha];          sb = a.sphere[hb]; }
            Vec3 S[4];
            const GjkOut o = gjk_pair(g, A, B, sa, sb, S, a.early_out != 0);
            const bool coll = exit_is_collision(o.exit_path);
            if (g.lane == 0) {
                if (a.collision) a.collision[p] = coll ? 1 : 0;
                if (a.status) a.status[p] = pack_status(o);
            }
            if (coll && a.wl_pair) {
                unsigned long long s3 = 0;
                if (g.lane == 0) s3 = atomicAdd(a.wl_count, 1ull);
                s3 = ((unsigned long long)(unsigned)g.bcast_i((int)(s3 >> 32), 0) << 32)
                   | (unsigned)g.bcast_i((int)(s3 & 0xffffffffu), 0);
                if (g.lane == 0) a.wl_pair[s3] = (int32_t)p;
                double *dst = a.wl_simplex + 12 * s3;
#pragma unroll
                for (int r = g.lane; r < 4; r += L) {
                    Vec3 row = pick4(S, r);
                    dst[3 * r + 0] = row.x; dst[3 * r + 1] = row.y; dst[3 * r + 2] = row.z;
                }
            }
        }
    }
}

}  // namespace gjk
