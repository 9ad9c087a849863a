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

}  // namespace f32io
