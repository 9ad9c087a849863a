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

// lowest and highest hull number named by pairs [a,b)
inline void hull_range(const int32_t *pa, const int32_t *pb, int64_t a, int64_t b, int64_t &lo, int64_t &hi)
{
    int32_t mn = 0x7fffffff, mx = -1;
    for (int64_t i = a; i < b; ++i) {
        const int32_t x = pa[i], y = pb[i];
        mn = x < mn ? x : mn; mn = y < mn ? y : mn;
        mx = x > mx ? x : mx; mx = y > mx ? y : mx;
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
    int64_t Hlo = 0, Hhi = -1;           // hull range uploaded whole when !ranged
    std::vector<Chunk> ch;
    bool bad_index = false;              // a pair names a hull outside [0, n_hulls)
};

// pa / pb: the SHARD's pair lists (np entries).  scan_store: the store is a host store that may be pair-private (about as
// many hulls as pairs); otherwise it is shared or already on the device and no pair is looked at.
inline Plan plan_shard(int64_t np, int forced_chunks, double bytes_per_pair, bool scan_store, int64_t n_hulls, const int64_t *hull_off,
                       const int32_t *pa, const int32_t *pb)
{
    Plan P;
    chunk_sizes(np, forced_chunks, bytes_per_pair, P.nchunk, P.cp);
    P.ch.resize((size_t)P.nchunk);
    for (int k = 0; k < P.nchunk; ++k) {
        P.ch[k].a = (int64_t)k * P.cp; P.ch[k].b = P.ch[k].a + P.cp < np ? P.ch[k].a + P.cp : np; P.ch[k].hlo = 0; P.ch[k].hhi = -1;
    }
    P.Hlo = 0; P.Hhi = n_hulls - 1;
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
