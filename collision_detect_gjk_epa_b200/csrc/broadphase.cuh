// broadphase.cuh -- Kb: candidate pairs on the device, upstream of the reference path.
//
// The reference decides per pair; its first step is RoughCollisionDetection_SphericalEnvelope (reference
// src/GCLIB_GJKEPA.f90:1165-1188): centroid + largest vertex distance per body, a pair passes iff
//     NORM2(mp1 - mp2) <= r1 + r2 + 1.0.
// This stage emits exactly the pairs (i < j) that pass THAT test (same arithmetic as gjk::sphere_hit, so the set is
// bit-exact with a host that tests all pairs), found through a uniform grid instead of all N^2 pairs:
//   spheres (K0)  ->  bounds: lower corner and largest radius (block reductions + atomics on ordered keys)
//   ->  cell of each centre, cell size h = (2 r_max + 1)(1 + 1e-9): two passing bodies are at most h apart, so their
//       cells differ by at most one step per axis  ->  sort bodies by cell key (cub radix sort)
//   ->  per body: binary-search its neighbour cells in the sorted list -- the three cells of a row along x have
//       consecutive keys, so 9 searches cover the 27 cells (27 single-cell searches when the scene is more than 2^21
//       cells across and keys alias) -- and test the candidates; a counting pass, an exclusive scan and a filling pass give a deterministic pair order.
#pragma once
#include "gjk_kernels.cuh"

namespace bp {

using gjk::Sphere;

struct Bounds { unsigned long long min_key[3]; unsigned long long max_r_key; unsigned long long max_key[3]; double slack; };   // order-preserving keys of doubles

__device__ __forceinline__ unsigned long long dkey(double v)
{   // monotone map double -> uint64 (-0 canonicalised)
    long long b = __double_as_longlong(v + 0.0);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__host__ __device__ __forceinline__ double dkey_inv(unsigned long long k)
{
    long long b = (k & 0x8000000000000000ull) ? (long long)(k ^ 0x8000000000000000ull) : (long long)~k;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double d; memcpy(&d, &b, sizeof d); return d;
#endif
}

// A body whose envelope sphere is not finite (a NaN or Inf vertex) takes no part: it neither moves the grid's bounds --
// one NaN would put every body into cell 0 (an all-pairs scan) or drop whole rows of cells -- nor gets a cell, and it
// is emitted in no pair.  The narrow phase answers FALSE for every pair that has such a body (NaN: every comparison of
// F:1185 fails; Inf: the walk ends on the reference's STOP path, tests/test_gpu_nonfinite.py), so the CONTACT set is
// what an all-pairs host loop gets; only candidates that could never collide are left out.
__device__ __forceinline__ bool finite_sphere(const Sphere &v)
{
    return fabs(v.cx) <= DBL_MAX && fabs(v.cy) <= DBL_MAX && fabs(v.cz) <= DBL_MAX && fabs(v.r) <= DBL_MAX;
}
constexpr unsigned long long KEY_NONE = 0x7fffffffffffffffull;      // sorts after every cell of an exact grid

// slack: the absolute term of the sphere test, 1.0 in the reference (F:1185)
__global__ void bounds_init_kernel(Bounds *b, double slack)
{
    b->slack = slack;
    b->min_key[0] = b->min_key[1] = b->min_key[2] = ~0ull; b->max_r_key = 0ull;
    b->max_key[0] = b->max_key[1] = b->max_key[2] = 0ull;
}

__global__ void __launch_bounds__(256)
bounds_kernel(const Sphere *__restrict__ s, int64_t n, Bounds *b)
{
    unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull, kr = 0ull, m0 = 0ull, m1 = 0ull, m2 = 0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Sphere v = s[i];
        if (!finite_sphere(v)) continue;
        const unsigned long long a0 = dkey(v.cx), a1 = dkey(v.cy), a2 = dkey(v.cz), ar = dkey(v.r);
        k0 = a0 < k0 ? a0 : k0; k1 = a1 < k1 ? a1 : k1; k2 = a2 < k2 ? a2 : k2; kr = ar > kr ? ar : kr;
        m0 = a0 > m0 ? a0 : m0; m1 = a1 > m1 ? a1 : m1; m2 = a2 > m2 ? a2 : m2;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t;
        t = __shfl_xor_sync(0xffffffffu, k0, o); k0 = t < k0 ? t : k0;
        t = __shfl_xor_sync(0xffffffffu, k1, o); k1 = t < k1 ? t : k1;
        t = __shfl_xor_sync(0xffffffffu, k2, o); k2 = t < k2 ? t : k2;
        t = __shfl_xor_sync(0xffffffffu, kr, o); kr = t > kr ? t : kr;
        t = __shfl_xor_sync(0xffffffffu, m0, o); m0 = t > m0 ? t : m0;
        t = __shfl_xor_sync(0xffffffffu, m1, o); m1 = t > m1 ? t : m1;
        t = __shfl_xor_sync(0xffffffffu, m2, o); m2 = t > m2 ? t : m2;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&b->min_key[0], k0); atomicMin(&b->min_key[1], k1); atomicMin(&b->min_key[2], k2); atomicMax(&b->max_r_key, kr);
        atomicMax(&b->max_key[0], m0); atomicMax(&b->max_key[1], m1); atomicMax(&b->max_key[2], m2);
    }
}

struct Grid { double ox, oy, oz, h, slack; long long nx, ny, nz; bool exact; };   // n*: largest cell index per axis; exact: keys never alias
__device__ __forceinline__ Grid grid_of(const Bounds *b)
{
    Grid g;
    g.ox = dkey_inv(b->min_key[0]); g.oy = dkey_inv(b->min_key[1]); g.oz = dkey_inv(b->min_key[2]);
    const double rmax = dkey_inv(b->max_r_key);
    g.slack = b->slack;
    g.h = (2.0 * rmax + g.slack) * (1.0 + 1.0e-9);
    g.nx = (long long)floor((dkey_inv(b->max_key[0]) - g.ox) / g.h);
    g.ny = (long long)floor((dkey_inv(b->max_key[1]) - g.oy) / g.h);
    g.nz = (long long)floor((dkey_inv(b->max_key[2]) - g.oz) / g.h);
    g.exact = g.nx < 0x1ffffe && g.ny < 0x1ffffe && g.nz < 0x1ffffe && g.nx >= 0 && g.ny >= 0 && g.nz >= 0;
    return g;
}
// 21 bits per axis: scenes up to 2M cells across; beyond that cells alias (still correct: aliasing only adds candidates
// that the sphere test rejects ... as long as the 27-neighbourhood is searched by aliased key too, which it is)
__device__ __forceinline__ unsigned long long cell_key(long long ix, long long iy, long long iz)
{
    return ((unsigned long long)(ix & 0x1fffff)) | ((unsigned long long)(iy & 0x1fffff) << 21) | ((unsigned long long)(iz & 0x1fffff) << 42);
}
__device__ __forceinline__ void cell_of(const Grid &g, const Sphere &s, long long &ix, long long &iy, long long &iz)
{
    ix = (long long)floor((s.cx - g.ox) / g.h); iy = (long long)floor((s.cy - g.oy) / g.h); iz = (long long)floor((s.cz - g.oz) / g.h);
}

__global__ void __launch_bounds__(256)
keys_kernel(const Sphere *__restrict__ s, int64_t n, const Bounds *__restrict__ b, unsigned long long *__restrict__ keys, int32_t *__restrict__ ids)
{
    const Grid g = grid_of(b);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        long long ix, iy, iz;
        const Sphere v = s[i];
        cell_of(g, v, ix, iy, iz);
        keys[i] = finite_sphere(v) ? cell_key(ix, iy, iz) : KEY_NONE;
        ids[i] = (int32_t)i;
    }
}

// F:1185 with the absolute term as a parameter: NORM2(mp1 - mp2) <= r1 + r2 + slack (slack = 1.0: gjk::sphere_hit, bit for bit)
__device__ __forceinline__ bool sphere_pass(const Sphere &a, const Sphere &b, double slack)
{
    return gjk::norm3(gjk::v3(a.cx, a.cy, a.cz) - gjk::v3(b.cx, b.cy, b.cz)) <= gjk::add(gjk::add(a.r, b.r), slack);
}

__device__ __forceinline__ int64_t lower_bound(const unsigned long long *k, int64_t n, unsigned long long v)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (k[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// FILL = false: counts[t] = pairs whose smaller body is the t-th body of the sorted list; FILL = true: write them at
// offsets[t]..  A pair (i, j), i < j, is emitted by body i's thread.
template <bool FILL>
__global__ void __launch_bounds__(128)
pairs_kernel(const Sphere *__restrict__ s, int64_t n, const Bounds *__restrict__ b,
             const unsigned long long *__restrict__ skeys, const int32_t *__restrict__ sids,
             int64_t *__restrict__ counts, const int64_t *__restrict__ offsets, int64_t cap,
             int32_t *__restrict__ pair_a, int32_t *__restrict__ pair_b)
{
    const Grid g = grid_of(b);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = sids[t];
        const Sphere si = s[i];
        long long ix, iy, iz;
        cell_of(g, si, ix, iy, iz);
        int64_t cnt = 0;
        const int64_t out0 = FILL ? offsets[t] : 0;
        if (!finite_sphere(si)) { if (!FILL) counts[t] = 0; continue; }
        if (g.exact) {
            // every body's cell index lies in [0, n*] (floor is monotone), so rows outside are empty and the three cells
            // of a row along x are one run of consecutive keys
            const long long xlo = ix > 0 ? ix - 1 : 0, xhi = ix < g.nx ? ix + 1 : g.nx;
            for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) {
                const long long cy = iy + dy, cz = iz + dz;
                if (cy < 0 || cy > g.ny || cz < 0 || cz > g.nz) continue;
                const unsigned long long klo = cell_key(xlo, cy, cz), khi = cell_key(xhi, cy, cz);
                for (int64_t u = lower_bound(skeys, n, klo); u < n && skeys[u] <= khi; ++u) {
                    const int32_t j = sids[u];
                    if (j <= i) continue;
                    if (sphere_pass(si, s[j], g.slack)) {             // F:1185, the reference's own test
                        if (FILL) { const int64_t o = out0 + cnt; if (o < cap) { pair_a[o] = i; pair_b[o] = j; } }
                        ++cnt;
                    }
                }
            }
        } else {
        unsigned long long seen[27]; int n_seen = 0;             // aliased neighbour keys may repeat: search each key once
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            const unsigned long long key = cell_key(ix + dx, iy + dy, iz + dz);
            bool dup = false;
            for (int q = 0; q < n_seen; ++q) dup = dup || seen[q] == key;
            if (dup) continue;
            seen[n_seen++] = key;
            for (int64_t u = lower_bound(skeys, n, key); u < n && skeys[u] == key; ++u) {
                const int32_t j = sids[u];
                if (j <= i) continue;
                const Sphere sj = s[j];
                if (finite_sphere(sj) && sphere_pass(si, sj, g.slack)) {   // F:1185, the reference's own test
                    if (FILL) { const int64_t o = out0 + cnt; if (o < cap) { pair_a[o] = i; pair_b[o] = j; } }
                    ++cnt;
                }
            }
        }
        }
        if (!FILL) counts[t] = cnt;
    }
}

}  // namespace bp
