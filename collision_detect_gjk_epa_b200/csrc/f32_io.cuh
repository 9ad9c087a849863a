// f32_io.cuh -- Kc: the binary32 boundary of gjkepa_batch_f32 ("fp32 mode").
//
// The reference computes in REAL*8 and its decisions hang on thresholds below binary32's epsilon (1.2e-7): the
// 1e-8 on-face bands of isPointInSimplex / the coplanar exits (reference src/GCLIB_GJKEPA.f90:157,203,1248) and the
// 1e-12 degeneracy tests of UTZVEC / UNINML (F:1349,1388).  A binary32 walk cannot evaluate them, so the fp32 mode is
// binary32 at the BOUNDARY -- vertices, tolerance and results stored and moved as float, half the PCIe and HBM
// bytes of the step that PCIe bounds -- while the device widens the hull store once (exact: every float is a
// double) and runs the same binary64 kernels.  Flags are therefore the reference's own on the float-rounded hulls.
#pragma once
#include <cstdint>

namespace f32io {

// dst[i] = (double)src[i]; two elements per thread where alignment allows
__global__ void __launch_bounds__(256)
widen_kernel(const float *__restrict__ src, double *__restrict__ dst, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4 *>(src)[i];
        reinterpret_cast<double2 *>(dst)[2 * i] = make_double2((double)v.x, (double)v.y);
        reinterpret_cast<double2 *>(dst)[2 * i + 1] = make_double2((double)v.z, (double)v.w);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (double)src[i];
}

// dst[i] = (float)src[i], round to nearest
__global__ void __launch_bounds__(256)
narrow_kernel(const double *__restrict__ src, float *__restrict__ dst, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n2 = n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 v = reinterpret_cast<const double2 *>(src)[i];
        reinterpret_cast<float2 *>(dst)[i] = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) dst[n - 1] = __double2float_rn(src[n - 1]);
}

// ---- binary32 ARITHMETIC, as a filter (GJKEPA_TUNE_F32_FILTER, fp32 entry only) -----------------------------------
// A binary32 walk cannot evaluate the reference's 1e-8 / 1e-12 thresholds (header), but binary32 can PROVE that two hulls are
// apart: if along some direction d every vertex of B projects beyond every vertex of A by more than the rounding error of the
// projections plus a margin, the hulls are separated by more than 1e-6, and the reference answers FALSE for such a pair (all its
// failure exits are FALSE; TRUE needs the origin inside, or within its 1e-8 band of, a tetrahedron of Minkowski points: the
// argument of shortcut E1, gjk_core.cuh, with the same caveat).  The filter can only answer "separated"; every other pair goes
// on to the binary64 kernels, which see exactly these float vertices widened.  Eight lanes per pair, the float store read twice (L1):
//   pass 1: centroids, bounding boxes and the largest |coordinate| M of both hulls (3 axis tests: boxes disjoint by the margin);
//   pass 2: projections on d = cB - cA, normalised (1 more test).
// Error bound: a 3-term binary32 dot product of |d| <= 1 with coordinates <= M is off by at most 3.5e-7 M, the difference of two
// of them by 1e-6 M at most; the margin is 1e-5 (1 + M), so a pair that passes has a true gap above 1e-6.
// Survivors are appended to the visiting order of the init kernel (low word of a 64-bit entry = pair number), one atomic per warp.
// FLP lanes per pair (lane l takes vertices l, l + FLP, ...; consecutive lanes read consecutive floats, so a warp's load touches
// 8 sectors instead of 32 -- one thread per pair was bound by L1 sector throughput: 0.18-0.6 ms per 1M pairs, more than the filter
// saved; eight lanes per pair by shuffle throughput: 0.16-0.22 ms), butterflies over the group for the sums, boxes and projections.
constexpr int FLP = 4;          // lanes per pair: 8 pairs per warp, two butterfly steps per reduction
__device__ __forceinline__ float bfly8_add(float v) { for (int o = FLP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, FLP); return v; }
__device__ __forceinline__ float bfly8_min(float v) { for (int o = FLP / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o, FLP)); return v; }
__device__ __forceinline__ float bfly8_max(float v) { for (int o = FLP / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, FLP)); return v; }

__global__ void __launch_bounds__(256)
separation_filter_kernel(const float *__restrict__ v, const int64_t *__restrict__ hull_off, const int32_t *__restrict__ pair_a,
                         const int32_t *__restrict__ pair_b, int64_t n_pairs, uint8_t *__restrict__ collision, uint32_t *__restrict__ status,
                         uint32_t status_filtered, unsigned long long *__restrict__ order, unsigned long long *__restrict__ order_count)
{
    const unsigned FULL = 0xffffffffu;
    constexpr int PPW = 32 / FLP;
    const int lane = threadIdx.x & 31, l = lane & (FLP - 1), grp = lane / FLP;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // the hull numbers of the NEXT round are loaded while this one is worked (the index chain pair -> hull -> offsets is three
    // dependent round trips)
    int nha = 0, nhb = 0;
    { const int64_t p1 = warp0 * PPW + grp; if (p1 < n_pairs) { nha = pair_a[p1]; nhb = pair_b[p1]; } }
    for (int64_t q = warp0 * PPW; q < n_pairs; q += n_warps * PPW) {
        const int64_t p = q + grp;
        const bool valid = p < n_pairs;
        int na = 0, nb = 0;
        const float *ax = v, *bx = v;
        const int ha = nha, hb = nhb;
        { const int64_t p1 = p + n_warps * PPW; if (p1 < n_pairs) { nha = pair_a[p1]; nhb = pair_b[p1]; } }
        if (valid) {
            const int64_t oa = hull_off[ha], ob = hull_off[hb];
            na = (int)(hull_off[ha + 1] - oa); nb = (int)(hull_off[hb + 1] - ob);
            ax = v + 3 * oa; bx = v + 3 * ob;
        }
        const float *ay = ax + na, *az = ay + na, *by = bx + nb, *bz = by + nb;
        float sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f}, alo[3], ahi[3], blo[3], bhi[3], M = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { alo[k] = blo[k] = 3.0e38f; ahi[k] = bhi[k] = -3.0e38f; }
        for (int i = l; i < na; i += FLP) {
            const float c[3] = {ax[i], ay[i], az[i]};
#pragma unroll
            for (int k = 0; k < 3; ++k) { sa[k] += c[k]; alo[k] = fminf(alo[k], c[k]); ahi[k] = fmaxf(ahi[k], c[k]); M = fmaxf(M, fabsf(c[k])); }
        }
        for (int i = l; i < nb; i += FLP) {
            const float c[3] = {bx[i], by[i], bz[i]};
#pragma unroll
            for (int k = 0; k < 3; ++k) { sb[k] += c[k]; blo[k] = fminf(blo[k], c[k]); bhi[k] = fmaxf(bhi[k], c[k]); M = fmaxf(M, fabsf(c[k])); }
        }
        M = bfly8_max(M);
        // a NaN anywhere is dropped by fminf / fmaxf: such a pair may be called apart, which is the reference's answer too
        // (a NaN centroid fails its sphere test, F:1185); an Inf makes the margin infinite and keeps the pair
        const float margin = 1.0e-5f * (1.0f + M);
        bool apart = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float al = bfly8_min(alo[k]), ah = bfly8_max(ahi[k]), bl = bfly8_min(blo[k]), bh = bfly8_max(bhi[k]);
            apart = apart || (bl - ah > margin) || (al - bh > margin);
            sa[k] = bfly8_add(sa[k]); sb[k] = bfly8_add(sb[k]);
        }
        float d[3] = {0.f, 0.f, 0.f};
        bool have_d = false;
        if (na > 0 && nb > 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) d[k] = sb[k] / (float)nb - sa[k] / (float)na;
            const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (len > 1.0e-6f) { const float inv = 1.0f / len; d[0] *= inv; d[1] *= inv; d[2] *= inv; have_d = true; }
        }
        float hA = -3.0e38f, lB = 3.0e38f;
        for (int i = l; i < na; i += FLP) hA = fmaxf(hA, d[0] * ax[i] + d[1] * ay[i] + d[2] * az[i]);
        for (int i = l; i < nb; i += FLP) lB = fminf(lB, d[0] * bx[i] + d[1] * by[i] + d[2] * bz[i]);
        hA = bfly8_max(hA); lB = bfly8_min(lB);
        if (have_d && (lB - hA) > margin) apart = true;
        const bool keep = valid && !apart;
        if (valid && apart && l == 0) {
            if (collision) collision[p] = 0;
            if (status) status[p] = status_filtered;
        }
        const unsigned m = __ballot_sync(FULL, keep && l == 0);
        if (m) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(order_count, (unsigned long long)__popc(m));
            base = __shfl_sync(FULL, base, 0);
            if (keep && l == 0) order[base + __popc(m & ((1u << lane) - 1u))] = (unsigned long long)(uint32_t)p;
        }
    }
}

}  // namespace f32io
