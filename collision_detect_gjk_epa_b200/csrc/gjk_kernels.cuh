// gjk_kernels.cuh -- K0 hull-prep and K1 GJK flag kernel (sm_100a).
//
// K0  hull_prep_kernel : per-hull envelope sphere of RoughCollisionDetection_SphericalEnvelope
//                        (reference src/GCLIB_GJKEPA.f90:1175-1182).  The centroid is an
//                        order-dependent sequential sum, so it is computed once per hull (one thread
//                        per hull) instead of once per pair.
// K1  gjk_kernel<L>    : GJKEPA flag path (F:76-236) for a batch of pairs, L lanes per pair
//                        (gjk_core.cuh).  Colliding pairs are appended, with their terminating
//                        tetrahedron, to the EPA work list (K2 compaction, warp-aggregated).
#pragma once
#include "gjk_core.cuh"

namespace gjk {

struct BatchArgs {
    const double  *verts;       // packed hull blocks
    const int64_t *hull_off;    // n_hulls+1, vertex units
    const Sphere  *sphere;      // per hull (K0 output)
    const int32_t *pair_a, *pair_b;
    int64_t n_pairs;
    int64_t n_hulls;
    uint8_t  *collision;        // may be null
    uint32_t *status;           // may be null
    // EPA work list (null when mode == GJK)
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

__global__ void __launch_bounds__(256)
hull_prep_kernel(const double *__restrict__ verts, const int64_t *__restrict__ hull_off,
                 int64_t n_hulls, Sphere *__restrict__ out)
{
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hulls) return;
    HullView v = hull_view(verts, hull_off, (int)h);
    out[h] = hull_sphere(v);
}

template <int L>
__global__ void __launch_bounds__(128)
gjk_kernel(BatchArgs a)
{
    const Lanes<L> g;
    const int64_t n_groups = (int64_t)gridDim.x * (blockDim.x / L);
    int64_t p = (int64_t)blockIdx.x * (blockDim.x / L) + threadIdx.x / L;
    for (; p < a.n_pairs; p += n_groups) {
        const int ha = __ldg(a.pair_a + p), hb = __ldg(a.pair_b + p);
        const HullView A = hull_view(a.verts, a.hull_off, ha);
        const HullView B = hull_view(a.verts, a.hull_off, hb);
        const Sphere sa = a.sphere[ha], sb = a.sphere[hb];
        Vec3 S[4];
        const GjkOut o = gjk_pair<L>(g, A, B, sa, sb, S);
        const bool coll = exit_is_collision(o.exit_path);
        if (g.lane == 0) {
            if (a.collision) a.collision[p] = coll ? 1 : 0;
            if (a.status)
                a.status[p] = (uint32_t)o.exit_path | ((uint32_t)o.iters << 8)
                            | (o.stop_flag ? (1u << 25) : 0u)
                            | (o.exit_path == EXIT_TRI_ORIGIN ? (1u << 28) : 0u);
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

}  // namespace gjk
