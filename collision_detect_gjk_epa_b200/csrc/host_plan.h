// host_plan.h -- how a host batch is cut up: one contiguous pair range per device (shard), each shard into chunks for the
// copy-in / compute / copy-out pipeline, and per chunk the range of hulls its pairs name.  Plain C++ (no CUDA), shared by
// host_batch.inl and the CPU tests (tests/host_sim/host_plan_sim.cpp, tests/test_host_plan.py).
//
// Pairs are independent (reference src/GCLIB_GJKEPA.f90:55-60: the caller's loop over pairs may run on any number of
// threads), so any partition of the pair list gives the same results; this one keeps ranges contiguous so that every
// device / chunk copies one slice of each of the caller's per-pair arrays.
#pragma once
#include <cstdint>
#include <vector>

namespace hostplan {

constexpr int MAX_CHUNKS = 64;

// shard s of n: sizes differ by at most one, shards in index order
inline void shard_range(int64_t n_pairs, int s, int n, int64_t &lo, int64_t &hi)
{
    const int64_t base = n_pairs / n, rem = n_pairs % n;
    lo = s * base + (s < rem ? s : rem);
    hi = lo + base + (s < rem ? 1 : 0);
}

// devices a batch of n_pairs is spread over: a shard below ~4k pairs does not fill a GPU
inline int shards_for(int64_t n_pairs, int n_devices)
{
    if (n_pairs >= (int64_t)n_devices * 4096) return n_devices;
    const int64_t n = n_pairs / 4096;
    return n < 1 ? 1 : (int)n;
}

struct Chunk { int64_t a, b; int64_t hlo, hhi; };      // pairs [a,b) of the shard (shard-relative), hulls [hlo,hhi] they name

// lowest and highest value of v[0..n): the scan is serial host time in front of the first copy of a batch (8 MB of pair lists per
// 1M pairs), so it is written for the vectoriser (independent accumulators, no branches) and cloned for AVX2 where the CPU has it
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__CUDACC__) && !defined(__clang__)
#define HOSTPLAN_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define HOSTPLAN_CLONES
#endif
HOSTPLAN_CLONES inline void minmax_i32(const int32_t *v, int64_t n, int32_t &mn_out, int32_t &mx_out)
{
    int32_t mn[16], mx[16];
    for (int k = 0; k < 16; ++k) { mn[k] = 0x7fffffff; mx[k] = -0x7fffffff - 1; }
    int64_t i = 0;
    for (; i + 16 <= n; i += 16)
        for (int k = 0; k < 16; ++k) { const int32_t x = v[i + k]; mn[k] = x < mn[k] ? x : mn[k]; mx[k] = x > mx[k] ? x : mx[k]; }
    for (; i < n; ++i) { const int32_t x = v[i]; mn[0] = x < mn[0] ? x : mn[0]; mx[0] = x > mx[0] ? x : mx[0]; }
    for (int k = 1; k < 16; ++k) { mn[0] = mn[k] < mn[0] ? mn[k] : mn[0]; mx[0] = mx[k] > mx[0] ? mx[k] : mx[0]; }
    mn_out = mn[0]; mx_out = mx[0];
}

// lowest and highest hull number named by pairs [a,b)
inline void hull_range(const int32_t *pa, const int32_t *pb, int64_t a, int64_t b, int64_t &lo, int64_t &hi)
{
    int32_t mn = 0x7fffffff, mx = -1, m0, m1;
    if (b > a) {
        minmax_i32(pa + a, b - a, m0, m1); mn = m0 < mn ? m0 : mn; mx = m1 > mx ? m1 : mx;
        minmax_i32(pb + a, b - a, m0, m1); mn = m0 < mn ? m0 : mn; mx = m1 > mx ? m1 : mx;
    }
    lo = mn; hi = mx;
}

// Chunks pay when the copies are a large part of the step: about one per 32 MB that crosses PCIe (bytes_per_pair = the
// pair-proportional traffic of the call: hull bytes of a pair-private store, pair lists, per-pair results), at most 8, at
// least 64k pairs each; `forced` > 0 overrides (GJKEPA_TUNE_HOST_CHUNKS).  Chunk sizes are multiples of 1024 pairs.
inline void chunk_sizes(int64_t np, int forced, double bytes_per_pair, int &nchunk, int64_t &cp)
{
    nchunk = forced;
    if (nchunk <= 0) {
        const double mb = bytes_per_pair * (double)np / (32.0 * 1024 * 1024);
        nchunk = mb < 1.0 ? 1 : (mb > 8.0 ? 8 : (int)mb);
        if ((int64_t)nchunk > np / 65536) nchunk = (int)(np / 65536 < 1 ? 1 : np / 65536);
    }
    if (nchunk > MAX_CHUNKS) nchunk = MAX_CHUNKS;
    cp = (np + nchunk - 1) / nchunk;
    cp = (cp + 1023) & ~(int64_t)1023;
    nchunk = (int)((np + cp - 1) / cp);
}

struct Plan {
    int nchunk = 0; int64_t cp = 0;
    bool ranged = false;                 // every chunk brings its own hull range (pair-private stores)
    bool lazy = false;                   // ranged, and the chunks' exact ranges are still to be scanned (scan_chunk, just in time)
    int64_t Hlo = 0, Hhi = -1;           // hull range uploaded whole when !ranged
    std::vector<Chunk> ch;
    bool bad_index = false;              // a pair names a hull outside [0, n_hulls)
};

// exact hull range of chunk k (lazy plans); false if a pair names a hull outside the store
inline bool scan_chunk(Plan &P, int k, int64_t n_hulls, const int32_t *pa, const int32_t *pb)
{
    hull_range(pa, pb, P.ch[k].a, P.ch[k].b, P.ch[k].hlo, P.ch[k].hhi);
    return P.ch[k].hlo >= 0 && P.ch[k].hhi < n_hulls;
}

// pa / pb: the SHARD's pair lists (np entries).  scan_store: the store is a host store that may be pair-private (about as
// many hulls as pairs); otherwise it is shared or already on the device and no pair is looked at.
// lazy (batches of several chunks): the pair lists are only SAMPLED here (one pair in 509) to tell an ordered list from a random
// one; the exact range of a chunk is scanned just before its copy is queued, while the copy of the chunk before it is in flight --
// the full scan (8 MB per 1M pairs, memory-bound) would otherwise be serial host time in front of the first byte on the bus.
inline Plan plan_shard(int64_t np, int forced_chunks, double bytes_per_pair, bool scan_store, int64_t n_hulls, const int64_t *hull_off,
                       const int32_t *pa, const int32_t *pb, bool lazy = false)
{
    Plan P;
    chunk_sizes(np, forced_chunks, bytes_per_pair, P.nchunk, P.cp);
    P.ch.resize((size_t)P.nchunk);
    for (int k = 0; k < P.nchunk; ++k) {
        P.ch[k].a = (int64_t)k * P.cp; P.ch[k].b = P.ch[k].a + P.cp < np ? P.ch[k].a + P.cp : np; P.ch[k].hlo = 0; P.ch[k].hhi = -1;
    }
    P.Hlo = 0; P.Hhi = n_hulls - 1;
    if (scan_store && lazy && P.nchunk > 1) {
        int64_t sum_v = 0, lo_all = n_hulls, hi_all = -1;
        for (int k = 0; k < P.nchunk; ++k) {
            int64_t lo = n_hulls, hi = -1;
            for (int64_t i = P.ch[k].a; i < P.ch[k].b; i += 509) {
                const int64_t x = pa[i], y = pb[i];
                if (x < 0 || x >= n_hulls || y < 0 || y >= n_hulls) { P.bad_index = true; return P; }
                lo = x < lo ? x : lo; lo = y < lo ? y : lo; hi = x > hi ? x : hi; hi = y > hi ? y : hi;
            }
            sum_v += hull_off[hi + 1] - hull_off[lo];
            lo_all = lo < lo_all ? lo : lo_all; hi_all = hi > hi_all ? hi : hi_all;
        }
        const int64_t whole_v = hull_off[hi_all + 1] - hull_off[lo_all];
        P.ranged = sum_v <= whole_v + whole_v / 4;
        P.lazy = P.ranged;
        if (P.ranged) return P;          // exact ranges: scan_chunk, chunk by chunk
    }
    if (scan_store) {
        // look at what the chunks name: when pair i uses hulls near 2i every chunk needs a narrow range and the ranges
        // together are about the store; when pairs name hulls at random every chunk would need all of it
        int64_t sum_v = 0;
        P.Hlo = n_hulls; P.Hhi = -1;
        for (int k = 0; k < P.nchunk; ++k) {
            hull_range(pa, pb, P.ch[k].a, P.ch[k].b, P.ch[k].hlo, P.ch[k].hhi);
            if (P.ch[k].hlo < 0 || P.ch[k].hhi >= n_hulls) { P.bad_index = true; return P; }
            sum_v += hull_off[P.ch[k].hhi + 1] - hull_off[P.ch[k].hlo];
            P.Hlo = P.ch[k].hlo < P.Hlo ? P.ch[k].hlo : P.Hlo; P.Hhi = P.ch[k].hhi > P.Hhi ? P.ch[k].hhi : P.Hhi;
        }
        const int64_t whole_v = hull_off[P.Hhi + 1] - hull_off[P.Hlo];
        P.ranged = sum_v <= whole_v + whole_v / 4;
    }
    return P;
}

}  // namespace hostplan
