// gjk_two_stage_group.cuh -- the two-stage GJK flag path (gjk_two_stage.cuh) for hulls too large for the
// thread-per-pair columns (17..64 vertices): a sub-warp of L = 4 lanes per pair, lane l owning tetrahedron face
// l and hull vertices l, l+4, ...; hull blocks copied verbatim ([x(n) | y(n) | z(n)]) into one shared-memory
// slot per pair.
//
//   K1ig gjk_init_group_kernel   runs the pairs the thread-per-pair init kernel declined (its overflow list):
//                                1-D TMA bulk copies of the two hull blocks into the pair's slot, envelope
//                                spheres, initial tetrahedron (gjk_init), seeds appended to the seed queue.
//   K1wg gjk_walk_group_kernel   persistent warps of 8 pairs walking in LOCKSTEP: one instruction stream for
//                                the 8 groups, every shuffle/vote with the constant full mask (LanesW; the
//                                free-running sub-warp kernel spent 7 % of its instructions on member-mask
//                                checks and ran its 8 groups one after the other whenever their control flow
//                                differed: 19 of 32 lanes active on average, ncu), groups re-packed from the
//                                seed queue as they finish.
#pragma once
#include "gjk_two_stage.cuh"
#ifdef GJK_DEBUG_BOUNDS
#include <cstdio>
#endif

namespace gjk {

#ifndef GJK_WALKG_MIN_CTAS
#define GJK_WALKG_MIN_CTAS 4
#endif

// ------------------------------------------------------------------ K1ig
template <int L, bool FUSE>
__global__ void __launch_bounds__(128)
gjk_init_group_kernel(TwoStageArgs ta, int slot_doubles)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(8) unsigned long long s_bar[4];
    const BatchArgs &a = ta.b;
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
            // no poll limit: a bulk copy whose addresses are valid always completes, and under time-slicing, a debugger or a
            // profiler's replay "late" is not "lost" -- a trap here would take the whole CUDA context down with it
            unsigned spins = 0;
            while (!mbar_try_wait(bar, parity)) { if (++spins > 64u) __nanosleep(32); }
            parity ^= 1u;
        }
        __syncwarp();
        if (valid && !fits && g.lane == 0) {             // larger than a slot: plain sub-warp kernel
            unsigned long long s2 = atomicAdd(a.ovf_count, 1ull);
            a.ovf_list[s2] = (int32_t)p;
            if (a.ovf_status) { if (a.status) a.status[p] = a.ovf_status; if (a.collision) a.collision[p] = 0; }
        }
        Walk w;
        int e = -1;
        if (fits) {
            HullViewShared A, B;
            A.x = slot; A.y = slot + GA.n; A.z = slot + 2 * GA.n; A.n = GA.n;
            B.x = slot + 3 * GA.n; B.y = B.x + GB.n; B.z = B.y + GB.n; B.n = GB.n;
            Sphere sa, sb;
            if (a.early_out) { sa.cx = sa.cy = sa.cz = sa.r = 0.0; sb = sa; }     // E1 skips the sphere test (gjk_init)
            else if (FUSE) { sa = group_sphere(g, A); sb = group_sphere(g, B); }
            else           { sa = a.sphere[ha];       sb = a.sphere[hb]; }
            e = gjk_init(g, A, B, sa, sb, w, a.early_out != 0);
            if (e > 0) {
                const bool coll = exit_is_collision(e);
                if (g.lane == 0) {
                    if (a.collision) a.collision[p] = coll ? 1 : 0;
                    if (a.status) a.status[p] = pack_status_walk(e, w);
                }
                if (coll && a.wl_pair) {                 // F:140: the origin on the first triangle
                    unsigned long long s3 = 0;
                    if (g.lane == 0) s3 = atomicAdd(a.wl_count, 1ull);
                    s3 = ((unsigned long long)(unsigned)g.bcast_i((int)(s3 >> 32), 0) << 32)
                       | (unsigned)g.bcast_i((int)(s3 & 0xffffffffu), 0);
                    if (g.lane == 0) a.wl_pair[s3] = (int32_t)p;
                    double *dst = a.wl_simplex + 12 * s3;
#pragma unroll
                    for (int r = g.lane; r < 4; r += L) {
                        Vec3 row = pick4(w.S, r);
                        dst[3 * r + 0] = row.x; dst[3 * r + 1] = row.y; dst[3 * r + 2] = row.z;
                    }
                }
            }
        }
        __syncwarp();
        const unsigned go = __ballot_sync(FULL, e == 0 && g.lane == 0);
        if (go) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(ta.q.count, (unsigned long long)__popc(go));
            base = __shfl_sync(FULL, base, 0);
            if (e == 0 && g.lane == 0) {
                const unsigned long long s = base + __popc(go & ((1u << lane) - 1u));
                double *f = ta.q.f64 + s;
                const int64_t cap = ta.q.cap;
                f[0] = w.qk.x; f[cap] = w.qk.y; f[2 * cap] = w.qk.z; f[3 * cap] = w.mk;
                ta.q.idx[s] = make_uint2(w.ia, w.ib);
                ta.q.pair[s] = (int32_t)p;
                ta.q.fln[s] = ((w.fl & WF_COLD) ? SEED_COLD : 0u) | ((unsigned)GA.n << 4) | ((unsigned)GB.n << 18);
                ta.q.oa[s] = (int64_t)((GA.x - a.verts) / 3); ta.q.ob[s] = (int64_t)((GB.x - a.verts) / 3);
            }
        }
    }
}

// ------------------------------------------------------------------ K1wg
template <int L>
__global__ void __launch_bounds__(128, GJK_WALKG_MIN_CTAS)
gjk_walk_group_kernel(TwoStageArgs ta, int slot_doubles)
{
    extern __shared__ __align__(16) double sm[];
    const BatchArgs &a = ta.b;
    constexpr int GPW = 32 / L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane & (L - 1), grp = lane / L;
    const unsigned FULL = 0xffffffffu;
    const unsigned leaders = (L == 4) ? 0x11111111u : ((L == 2) ? 0x55555555u : ((L == 8) ? 0x01010101u : 0xffffffffu));
    const unsigned before = (1u << (grp * L)) - 1u;             // lanes of the groups before mine
    double *slot = sm + ((size_t)warp * GPW + grp) * slot_doubles;
    const LanesW<L> g;
    const unsigned long long count = *ta.q.count;
    const int64_t cap = ta.q.cap;
    const int refill_groups = ta.refill / L > 0 ? ta.refill / L : 1;

    Walk w;
    HullViewShared A, B;
    A.x = A.y = A.z = slot; A.n = 0; B = A;
    int64_t p = 0;
    bool active = false;
    w.S[0] = w.S[1] = w.S[2] = w.S[3] = v3(0.0, 0.0, 0.0);
    w.qk = w.D1 = v3(0.0, 0.0, 0.0); w.mk = 0.0; w.k1 = 2; w.iters = 0; w.n_support = 0; w.fl = 0u;
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
    // L2 prefetch, a few pairs ahead only: with 1.5 KB per pair a whole batch per warp (x 2400 resident warps)
    // outgrows L2 and the lines are gone again before they are used (ncu: DRAM reads 2x the hull bytes)
    constexpr int AHEAD = 8;
    auto prefetch_records = [&](const Batch &bt) {                // seed records: 4 fields x 256 bytes, one line per lane
        if (bt.n > 0 && lane < 2 * SEED_F64)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ta.q.f64 + (size_t)(lane >> 1) * cap + bt.base + (lane & 1) * 16));
        if (bt.n > 0 && (lane == 8 || lane == 9)) asm volatile("prefetch.global.L2 [%0];" ::"l"(ta.q.idx + bt.base + (lane & 1) * 16));
    };
    auto prefetch_hulls = [&](const Batch &bt, int from) {         // the hull blocks of entries [from, from + AHEAD)
        if (lane >= from && lane < from + AHEAD && lane < bt.n) {
            prefetch_l2_range(a.verts + 3 * bt.oa, 24 * (int)((bt.fln >> 4) & 0x3fffu));
            prefetch_l2_range(a.verts + 3 * bt.ob, 24 * (int)(bt.fln >> 18));
        }
    };
    Batch cur, nxt;
    fetch(cur); prefetch_records(cur); prefetch_hulls(cur, 0);
    fetch(nxt);
    bool nxt_pf = false;
    int bcur = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active) & leaders;
        const int n_idle = __popc(idle);
        if (n_idle >= refill_groups && (bcur < cur.n || nxt.n > 0)) {
            if (!nxt_pf) { prefetch_records(nxt); nxt_pf = true; }
            if (bcur >= cur.n) { cur = nxt; bcur = 0; fetch(nxt); nxt_pf = false; }
            const int avail = cur.n - bcur;
            if (avail > 0) {
                const int n_take = n_idle < avail ? n_idle : avail;
                prefetch_hulls(cur, bcur + n_take);                 // what the next re-pack will take
                if (bcur + n_take + AHEAD > 32) prefetch_hulls(nxt, bcur + n_take - 32);
                const int rank = __popc(idle & before);
                const bool take = !active && rank < n_take;
                const int src = (bcur + rank) & 31;
#ifdef GJK_WALKG_DIRECT_META
                int sp_ = 0; unsigned sfln = 0; long long soa = 0, sob = 0;
                if (take) {
                    const unsigned long long it2 = cur.base + (unsigned long long)(bcur + rank);
                    sp_ = ta.q.pair[it2]; sfln = ta.q.fln[it2]; soa = ta.q.oa[it2]; sob = ta.q.ob[it2];
                }
#else
                const int sp_ = __shfl_sync(FULL, cur.p, src);
                const unsigned sfln = __shfl_sync(FULL, cur.fln, src);
                const long long soa = __shfl_sync(FULL, cur.oa, src), sob = __shfl_sync(FULL, cur.ob, src);
#endif
#ifdef GJK_DEBUG_BOUNDS
                if (take) {
                    const long long nv_total = a.hull_off[a.n_hulls];
                    const int nA_ = (int)((sfln >> 4) & 0x3fffu), nB_ = (int)(sfln >> 18);
                    const unsigned long long it_ = cur.base + (unsigned long long)(bcur + rank);
                    if (soa < 0 || sob < 0 || soa + nA_ > nv_total || sob + nB_ > nv_total || sp_ < 0 || sp_ >= a.n_pairs || it_ >= count ||
                        3 * (nA_ + nB_) + 4 > slot_doubles)
                        if (gl == 0) printf("BAD seed item=%llu count=%llu base=%llu bcur=%d rank=%d n_take=%d cur.n=%d src=%d p=%d oa=%lld ob=%lld nA=%d nB=%d\n",
                                            it_, count, cur.base, bcur, rank, n_take, cur.n, src, sp_, soa, sob, nA_, nB_);
                }
#endif
                uint2 t2 = make_uint2(0u, 0u);
                if (take) {
                    const int nA = (int)((sfln >> 4) & 0x3fffu), nB = (int)(sfln >> 18);
                    const double *srcA = a.verts + 3 * soa, *srcB = a.verts + 3 * sob;
                    for (int e = gl; e < 3 * nA; e += L) cp_async8(slot + e, srcA + e);
                    for (int e = gl; e < 3 * nB; e += L) cp_async8(slot + 3 * nA + e, srcB + e);
                    const double *f = ta.q.f64 + (cur.base + (unsigned long long)(bcur + rank));
                    w.qk = v3(f[0], f[cap], f[2 * cap]);
                    w.mk = f[3 * cap];
                    t2 = ta.q.idx[cur.base + (unsigned long long)(bcur + rank)];
                    w.D1 = v3(0.0, 0.0, 0.0); w.k1 = 2; w.iters = 0; w.n_support = 4;
                    w.fl = WF_FIRST | ((sfln & SEED_COLD) ? WF_COLD : 0u) | (a.early_out ? WF_EARLY : 0u);
                    p = sp_;
                    A.x = slot; A.y = slot + nA; A.z = slot + 2 * nA; A.n = nA;
                    B.x = slot + 3 * nA; B.y = B.x + nB; B.z = B.y + nB; B.n = nB;
                    active = true;
                }
                bcur += n_take;
                cp_async_wait_all();
                __syncwarp();                                   // the group's lanes read each other's copies
                if (take) {                                     // the tetrahedron K1ig ended with, from the staged hulls
#pragma unroll
                    for (int r = 0; r < 4; ++r) w.S[r] = A.get((int)((t2.x >> (8 * r)) & 0xffu)) - B.get((int)((t2.y >> (8 * r)) & 0xffu));
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (bcur >= cur.n && nxt.n == 0) break;
            continue;
        }
        // one round for all groups in lockstep, finished or not (see walk_test / walk_step)
        Pick pk;
        const int e1 = walk_test(g, w, pk, active);
        const bool coll = active && e1 != 0 && exit_is_collision(e1);
        {   // EPA work list, before the step overwrites the tetrahedron: one atomic per warp and round
            const unsigned mcol = __ballot_sync(FULL, coll) & leaders;
            if (mcol && a.wl_pair) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.wl_count, (unsigned long long)__popc(mcol));
                base = __shfl_sync(FULL, base, 0);
                if (coll) {
                    const unsigned long long s = base + __popc(mcol & before);
                    if (gl == 0) a.wl_pair[s] = (int32_t)p;
                    double *dst = a.wl_simplex + 12 * s;
#pragma unroll
                    for (int r = gl; r < 4; r += L) {
                        const Vec3 row = pick4(w.S, r);
                        dst[3 * r + 0] = row.x; dst[3 * r + 1] = row.y; dst[3 * r + 2] = row.z;
                    }
                }
            }
        }
        const int iters_out = w.iters;
        const int e2 = walk_step(g, A, B, w, pk, active && e1 == 0);
        const int e = e1 ? e1 : e2;
        if (active && e) {
            if (gl == 0) {
                if (a.collision) a.collision[p] = coll ? 1 : 0;
                if (a.status) a.status[p] = (uint32_t)e | ((uint32_t)iters_out << 8) | ((w.fl & WF_COLD) ? (1u << 30) : 0u) | (e == EXIT_STOP ? (1u << 25) : 0u);
            }
            active = false;
        }
    }
}

}  // namespace gjk
