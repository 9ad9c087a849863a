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
// on to the binary64 kernels, which see exactly these float vertices widened.  One thread per pair, the float store read once:
//   pass 1: centroids, bounding boxes and the largest |coordinate| M of both hulls (3 axis tests: boxes disjoint by the margin);
//   pass 2: projections on d = cB - cA, normalised (1 more test).
// Error bound: a 3-term binary32 dot product of |d| <= 1 with coordinates <= M is off by at most 3.5e-7 M, the difference of two
// of them by 1e-6 M at most; the margin is 1e-5 (1 + M), so a pair that passes has a true gap above 1e-6.
// Survivors are appended to the visiting order of the init kernel (low word of a 64-bit entry = pair number), one atomic per warp.
__global__ void __launch_bounds__(256)
separation_filter_kernel(const float *__restrict__ v, const int64_t *__restrict__ hull_off, const int32_t *__restrict__ pair_a,
                         const int32_t *__restrict__ pair_b, int64_t n_pairs, uint8_t *__restrict__ collision, uint32_t *__restrict__ status,
                         uint32_t status_filtered, unsigned long long *__restrict__ order, unsigned long long *__restrict__ order_count)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(int64_t)31; p0 < n_pairs; p0 += stride) {
        const int64_t p = p0 + lane;
        bool keep = false, valid = p < n_pairs;
        if (valid) {
            const int ha = pair_a[p], hb = pair_b[p];
            const int64_t oa = hull_off[ha], ob = hull_off[hb];
            const int na = (int)(hull_off[ha + 1] - oa), nb = (int)(hull_off[hb + 1] - ob);
            const float *ax = v + 3 * oa, *ay = ax + na, *az = ay + na, *bx = v + 3 * ob, *by = bx + nb, *bz = by + nb;
            float sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f}, alo[3], ahi[3], blo[3], bhi[3], M = 0.f;
            for (int k = 0; k < 3; ++k) { alo[k] = blo[k] = 3.0e38f; ahi[k] = bhi[k] = -3.0e38f; }
            for (int i = 0; i < na; ++i) {
                const float c[3] = {ax[i], ay[i], az[i]};
                for (int k = 0; k < 3; ++k) { sa[k] += c[k]; alo[k] = fminf(alo[k], c[k]); ahi[k] = fmaxf(ahi[k], c[k]); M = fmaxf(M, fabsf(c[k])); }
            }
            for (int i = 0; i < nb; ++i) {
                const float c[3] = {bx[i], by[i], bz[i]};
                for (int k = 0; k < 3; ++k) { sb[k] += c[k]; blo[k] = fminf(blo[k], c[k]); bhi[k] = fmaxf(bhi[k], c[k]); M = fmaxf(M, fabsf(c[k])); }
            }
            const float margin = 1.0e-5f * (1.0f + M);
            // NaN / Inf anywhere: every comparison below fails and the pair is kept for the binary64 kernels
            bool apart = false;
            for (int k = 0; k < 3; ++k) apart = apart || (blo[k] - ahi[k] > margin) || (alo[k] - bhi[k] > margin);
            if (!apart && na > 0 && nb > 0) {
                float d[3] = {sb[0] / (float)nb - sa[0] / (float)na, sb[1] / (float)nb - sa[1] / (float)na, sb[2] / (float)nb - sa[2] / (float)na};
                const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                if (len > 1.0e-6f) {
                    const float inv = 1.0f / len;
                    d[0] *= inv; d[1] *= inv; d[2] *= inv;
                    float hA = -3.0e38f, lB = 3.0e38f;
                    for (int i = 0; i < na; ++i) hA = fmaxf(hA, d[0] * ax[i] + d[1] * ay[i] + d[2] * az[i]);
                    for (int i = 0; i < nb; ++i) lB = fminf(lB, d[0] * bx[i] + d[1] * by[i] + d[2] * bz[i]);
                    apart = (lB - hA) > margin;
                }
            }
            keep = !apart;
            if (apart) {
                if (collision) collision[p] = 0;
                if (status) status[p] = status_filtered;
            }
        }
        const unsigned m = __ballot_sync(FULL, keep);
        if (m) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(order_count, (unsigned long long)__popc(m));
            base = __shfl_sync(FULL, base, 0);
            if (keep) order[base + __popc(m & ((1u << lane) - 1u))] = (unsigned long long)(uint32_t)p;
        }
    }
}

}  // namespace f32io
