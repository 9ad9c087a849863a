// scene.cuh -- Ks: the compaction at the end of a scene step (gjkepa_scene_posed_f64, SURVEY.md §8f.4: pose
// application + candidate pairs + the reference path in one call).  The narrow phase writes its results at the
// candidate's index and leaves the indices of the colliding candidates in a device list (unordered: warps append as
// they finish); after a sort of that list this kernel gathers the contacts' rows into dense arrays, so only contacts
// cross PCIe.  One thread per contact; the reads are scattered (one candidate in ~4 collides), the writes coalesced.
#pragma once
#include <cstdint>

namespace scene {

struct GatherArgs {
    const int32_t *order;                // colliding candidates, in (body a, body b) order
    int64_t n;
    const int32_t *cand_a, *cand_b;
    const double *depth, *normal, *point, *nearest;   // per candidate (normal/point 3, nearest 6 doubles); may be null
    const int32_t *colli_type;
    int32_t *out_a, *out_b;
    double *o_depth, *o_normal, *o_point, *o_nearest;
    int32_t *o_type;
};

// sort key of a contact: (body a, body b) -- the contact list comes out in lexicographic order whatever order the broad
// phase found the candidates in
__global__ void __launch_bounds__(256)
contact_keys_kernel(const int32_t *__restrict__ idx, int64_t n, const int32_t *__restrict__ cand_a, const int32_t *__restrict__ cand_b,
                    unsigned long long *__restrict__ keys)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx[k];
        keys[k] = ((unsigned long long)(uint32_t)cand_a[i] << 32) | (unsigned long long)(uint32_t)cand_b[i];
    }
}

__global__ void __launch_bounds__(256)
gather_kernel(GatherArgs g)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < g.n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g.order[k];
        g.out_a[k] = g.cand_a[i];
        g.out_b[k] = g.cand_b[i];
        if (g.depth) g.o_depth[k] = g.depth[i];
        if (g.normal) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g.o_normal[3 * k + c] = g.normal[3 * i + c];
        }
        if (g.point) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g.o_point[3 * k + c] = g.point[3 * i + c];
        }
        if (g.nearest) {
#pragma unroll
            for (int c = 0; c < 6; ++c) g.o_nearest[6 * k + c] = g.nearest[6 * i + c];
        }
        if (g.colli_type) g.o_type[k] = g.colli_type[i];
    }
}

}  // namespace scene
