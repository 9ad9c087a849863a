// gjk_f32.cuh -- the GJK flag path in BINARY32 ARITHMETIC (GJKEPA_TUNE_F32_ARITH on gjkepa_batch_f32).
//
// The walk is gjk32::gjk_pair: the text of the binary64 product path (gjk_core_body.inc, reference F:76-236) compiled
// with float operands.  It reads the caller's float hull store as it arrived over PCIe -- no widening, half the bytes,
// twice the lanes of the FP32 pipe, and sqrt / division in a handful of instructions instead of the binary64 sequences.
//
//   K1f  gjk_f32_kernel        one thread per pair; the two hulls of the 32 pairs of a warp staged once in shared memory,
//                              element e of lane l's pair at col[e * 32 + l] (a lane reads its own bank: conflict free),
//                              for hulls of at most `cap` (<= 16) vertices; other pairs go to the overflow list.
//   K1fg gjk_f32_group_kernel  the overflow list by sub-warps of 4 lanes: the pair's blocks copied verbatim into a shared-memory
//                              slot (nA + nB <= 128; larger pairs are walked over global memory), a lane scans vertices l, l+4, ...
// K1f is the one-kernel form on purpose: split into an init kernel and a walk kernel with re-packed warps, as the binary64
// path is, it was SLOWER (cfg2 0.19 + 0.29 ms against 0.38 ms, cfg5 0.35 + 0.49 against 0.62 ms, same bits;
// profiles/r2_ab_f32_two_stage.txt): a float walk round is cheap, and re-staging a lane's hulls at every re-pack costs more
// than the idle lanes it removes.
//
// What leaves the kernels: flag 0 and the binary32 walk's exit path / iteration count (marked GJKEPA_ST_F32_DECIDED) for a
// pair the walk calls separated -- final; a pair it calls colliding is appended to the survivor list.  In GJK mode the
// host layer takes the survivors' flags as they are (flag 1, binary32 status); in the EPA modes the binary64 kernels
// visit the survivors (BatchArgs::order), confirm the flag in binary64 and produce depth and normal as for any batch.
#pragma once
#include "gjk_core_f32.cuh"

namespace gjk32 {

struct F32Args {
    const float *verts;                 // [x(n) | y(n) | z(n)] per hull, as the caller's store
    const int64_t *hull_off;
    const int32_t *pair_a, *pair_b;
    int64_t n_pairs;
    const int32_t *pair_list;           // K1fg: the overflow list of K1f
    const unsigned long long *pair_list_count;
    uint8_t *collision;                 // may be null
    uint32_t *status;                   // may be null
    uint32_t mark;                      // GJKEPA_ST_F32_DECIDED
    int32_t final_flags;                // 1: GJK mode, the survivors' flags are written here too
    unsigned long long *order, *order_count;     // survivors (pairs the binary32 walk calls colliding)
    int32_t *ovf_list;                  // K1f: pairs with a hull above `cap`
    unsigned long long *ovf_count;
    int32_t cap;                        // vertices per hull a column holds
    int32_t early_out;
};

// element e of the staged block [x(n) | y(n) | z(n)] of this lane's pair
struct HullViewCol {
    const float *col;
    int n;
    __device__ __forceinline__ Vec3 get(int i) const { return v3(col[i * 32], col[(n + i) * 32], col[(2 * n + i) * 32]); }
};

__device__ __forceinline__ uint32_t pack_status(const GjkOut &o, uint32_t mark)
{
    return (uint32_t)o.exit_path | ((uint32_t)o.iters << 8) | (o.stop_flag ? (1u << 25) : 0u)
         | (o.exit_path == EXIT_TRI_ORIGIN ? (1u << 28) : 0u) | mark;
}

// flag / status of one decided pair and the warp-aggregated append of the survivors
__device__ __forceinline__ void emit(const F32Args &a, bool valid, bool leader, int64_t p, const GjkOut &o)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool coll = valid && exit_is_collision(o.exit_path);
    if (valid && leader && (!coll || a.final_flags)) {
        if (a.collision) a.collision[p] = coll ? 1 : 0;
        if (a.status) a.status[p] = pack_status(o, a.mark);
    }
    const unsigned m = __ballot_sync(FULL, coll && leader);
    if (m && a.order) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.order_count, (unsigned long long)__popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (coll && leader) a.order[base + __popc(m & ((1u << lane) - 1u))] = (unsigned long long)(uint32_t)p;
    }
}

constexpr int F32_WARPS = 4;

__global__ void __launch_bounds__(F32_WARPS * 32)
gjk_f32_kernel(F32Args a)
{
    extern __shared__ __align__(16) float smf[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *col = smf + (size_t)warp * (6 * a.cap) * 32 + lane;
    const Lanes<1> g;
    const int64_t n_tiles = (a.n_pairs + 31) / 32;
    const int64_t n_warps = (int64_t)gridDim.x * F32_WARPS;
    for (int64_t tile = (int64_t)blockIdx.x * F32_WARPS + warp; tile < n_tiles; tile += n_warps) {
        const int64_t p = tile * 32 + lane;
        const bool valid = p < a.n_pairs;
        int nA = 0, nB = 0;
        const float *srcA = a.verts, *srcB = a.verts;
        if (valid) {
            const int ha = __ldg(a.pair_a + p), hb = __ldg(a.pair_b + p);
            const long long oa = __ldg(a.hull_off + ha), ob = __ldg(a.hull_off + hb);
            nA = (int)(__ldg(a.hull_off + ha + 1) - oa); nB = (int)(__ldg(a.hull_off + hb + 1) - ob);
            srcA = a.verts + 3 * oa; srcB = a.verts + 3 * ob;
        }
        const bool fits = valid && nA <= a.cap && nB <= a.cap;
        const unsigned big = __ballot_sync(FULL, valid && !fits);
        if (big) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.ovf_count, (unsigned long long)__popc(big));
            base = __shfl_sync(FULL, base, 0);
            if (valid && !fits) a.ovf_list[base + __popc(big & ((1u << lane) - 1u))] = (int32_t)p;
        }
        __syncwarp();                                   // the previous tile's columns are no longer read
        if (fits) {
            for (int e = 0; e < 3 * nA; ++e) col[e * 32] = __ldg(srcA + e);
            float *cb = col + 3 * nA * 32;
            for (int e = 0; e < 3 * nB; ++e) cb[e * 32] = __ldg(srcB + e);
        }
        GjkOut o; o.exit_path = 0; o.iters = 0; o.n_support = 0; o.stop_flag = false; o.cold = 0; o.band = 0;
        if (fits) {
            HullViewCol A, B;
            A.col = col; A.n = nA; B.col = col + 3 * nA * 32; B.n = nB;
            Sphere sa, sb;
            if (a.early_out) { sa.cx = sa.cy = sa.cz = sa.r = 0.0f; sb = sa; }
            else { sa = hull_sphere(A); sb = hull_sphere(B); }
            Vec3 S[4];
            o = gjk_pair(g, A, B, sa, sb, S, a.early_out != 0);
        }
        __syncwarp();
        emit(a, fits, true, p, o);
    }
}

// the pair's two blocks [x(n) | y(n) | z(n)] copied verbatim into a shared-memory slot
struct HullViewSlot {
    const float *x, *y, *z;
    int n;
    __device__ __forceinline__ Vec3 get(int i) const { return v3(x[i], y[i], z[i]); }
};

constexpr int F32_SLOT_VERTS = 128;                        // nA + nB a slot holds
constexpr int F32_SLOT_FLOATS = 3 * F32_SLOT_VERTS + 4;    // +4: the 8 slots of a warp start 4 banks apart, a group's 4 lanes read 4 banks

template <class HV>
__device__ __forceinline__ GjkOut group_walk(const Lanes<4> &g, const HV &A, const HV &B, bool early)
{
    Sphere sa, sb;
    if (early) { sa.cx = sa.cy = sa.cz = sa.r = 0.0f; sb = sa; }
    else { sa = hull_sphere(A); sb = hull_sphere(B); }
    Vec3 S[4];
    return gjk_pair(g, A, B, sa, sb, S, early);
}

__global__ void __launch_bounds__(128)
gjk_f32_group_kernel(F32Args a)
{
    extern __shared__ __align__(16) float smf[];
    constexpr int L = 4, GPW = 32 / L;
    const Lanes<L> g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / L;
    float *slot = smf + (size_t)(warp * GPW + grp) * F32_SLOT_FLOATS;
    const int64_t n_items = (int64_t)(*a.pair_list_count);
    const int64_t n_tiles = (n_items + GPW - 1) / GPW;
    const int64_t n_warps = (int64_t)gridDim.x * 4;
    for (int64_t tile = (int64_t)blockIdx.x * 4 + warp; tile < n_tiles; tile += n_warps) {
        const int64_t it = tile * GPW + grp;
        const bool valid = it < n_items;
        int64_t p = 0;
        HullView A, B;
        A.x = A.y = A.z = a.verts; A.n = 0; B = A;
        if (valid) {
            p = a.pair_list[it];
            const int ha = __ldg(a.pair_a + p), hb = __ldg(a.pair_b + p);
            const long long oa = __ldg(a.hull_off + ha), ob = __ldg(a.hull_off + hb);
            A.n = (int)(__ldg(a.hull_off + ha + 1) - oa); B.n = (int)(__ldg(a.hull_off + hb + 1) - ob);
            A.x = a.verts + 3 * oa; A.y = A.x + A.n; A.z = A.y + A.n;
            B.x = a.verts + 3 * ob; B.y = B.x + B.n; B.z = B.y + B.n;
        }
        const bool fits = valid && A.n + B.n <= F32_SLOT_VERTS;
        __syncwarp();                                  // the previous tile's slots are no longer read
        if (fits) {                                    // a group's lanes copy 16-byte runs of its pair's blocks
            for (int e = g.lane; e < 3 * A.n; e += L) slot[e] = __ldg(A.x + e);
            for (int e = g.lane; e < 3 * B.n; e += L) slot[3 * A.n + e] = __ldg(B.x + e);
        }
        __syncwarp();
        GjkOut o; o.exit_path = 0; o.iters = 0; o.n_support = 0; o.stop_flag = false; o.cold = 0; o.band = 0;
        if (fits) {                                    // the group's four lanes take this branch together
            HullViewSlot SA, SB;
            SA.x = slot; SA.y = slot + A.n; SA.z = slot + 2 * A.n; SA.n = A.n;
            SB.x = slot + 3 * A.n; SB.y = SB.x + B.n; SB.z = SB.y + B.n; SB.n = B.n;
            o = group_walk(g, SA, SB, a.early_out != 0);
        } else if (valid) {                            // larger than a slot: over global memory
            o = group_walk(g, A, B, a.early_out != 0);
        }
        __syncwarp();
        emit(a, valid, g.lane == 0, p, o);
    }
}

}  // namespace gjk32
