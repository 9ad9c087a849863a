// pose_kernel.cuh -- Kp: world-space hull store from body-space shapes and per-body poses, on the device.
//
// Upstream of the reference path: a caller of GJKEPA (reference src/GCLIB_GJKEPA.f90:39-52)
// hands over world-space vertex arrays p1_, p2_, i.e. it has applied each body's pose on the host.  When the
// shapes stay resident on the GPU only the poses cross PCIe (96 B per body instead of 24 B per vertex).
// Arithmetic contract (so that flags stay bit-exact with a host that does the same): per vertex and axis
//     x' = ((r00*x + r01*y) + r02*z) + tx          R row-major, every operation rounded, no FMA.
// A sub-warp of 8 lanes per body; lane l transforms vertices l, l+8, ...; the block layout [x(n) | y(n) | z(n)]
// is kept, so consecutive lanes read and write consecutive doubles.
#pragma once
#include "gjk_core.cuh"

namespace pose {

// sizes[b] = vertices of body b's shape (sizes[n_bodies] = 0 so that an exclusive scan yields n_bodies+1 offsets)
__global__ void __launch_bounds__(256)
body_sizes_kernel(const int32_t *__restrict__ body_shape, const int64_t *__restrict__ shape_off, int64_t n_bodies,
                  int64_t n_shapes, int64_t *__restrict__ sizes)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_bodies) return;
    int64_t n = 0;
    if (b < n_bodies) {
        const int64_t s = body_shape[b];
        if (s >= 0 && s < n_shapes) n = shape_off[s + 1] - shape_off[s];
    }
    sizes[b] = n;
}

__global__ void __launch_bounds__(256)
apply_pose_kernel(const double *__restrict__ shape_verts, const int64_t *__restrict__ shape_off,
                  const int32_t *__restrict__ body_shape, const double *__restrict__ pose,
                  const int64_t *__restrict__ world_off, int64_t n_bodies, int64_t n_shapes, double *__restrict__ world)
{
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int l = threadIdx.x & 7;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
    for (int64_t b = g0; b < n_bodies; b += stride) {
        const int64_t s = body_shape[b];
        if (s < 0 || s >= n_shapes) continue;
        const int64_t so = shape_off[s];
        const int n = (int)(shape_off[s + 1] - so);
        const double *sx = shape_verts + 3 * so, *sy = sx + n, *sz = sy + n;
        double *wx = world + 3 * world_off[b], *wy = wx + n, *wz = wy + n;
        const double *p = pose + 12 * b;
        const double r00 = p[0], r01 = p[1], r02 = p[2], r10 = p[3], r11 = p[4], r12 = p[5],
                     r20 = p[6], r21 = p[7], r22 = p[8], tx = p[9], ty = p[10], tz = p[11];
        for (int i = l; i < n; i += 8) {
            const double x = sx[i], y = sy[i], z = sz[i];
            wx[i] = gjk::add(gjk::add(gjk::add(gjk::mul(r00, x), gjk::mul(r01, y)), gjk::mul(r02, z)), tx);
            wy[i] = gjk::add(gjk::add(gjk::add(gjk::mul(r10, x), gjk::mul(r11, y)), gjk::mul(r12, z)), ty);
            wz[i] = gjk::add(gjk::add(gjk::add(gjk::mul(r20, x), gjk::mul(r21, y)), gjk::mul(r22, z)), tz);
        }
    }
}

}  // namespace pose
