// contact_kernel.cuh -- K4: the contact epilogue of GJKEPA (reference src/GCLIB_GJKEPA.f90:326-343) as a
// kernel of its own: get_nearest_points (F:813-855), get_collisionPoint_01 (F:700-806) / _03 (F:426-452) and
// get_info_collisionType (F:353-413) for every pair the EPA kernels retired with a depth and a normal.
//
// Everything here is a scan over ONE hull along +-normal, and the scans that pick the contact point are
// order-dependent (quirk Q8: "> max - 1e-8, the last near-maximum wins, the previous winner becomes second"),
// so the work is split by hull, not by vertex: two lanes per contact, lane 0 walks p1 along +n, lane 1 walks
// p2 along -n, and they swap their few results with one shuffle each.  No shared memory, no reductions.
#pragma once
#include "epa_kernel.cuh"

namespace contact {

using gjk::Vec3; using gjk::v3; using gjk::dot3; using gjk::neg;
using epa::EpaArgs;

__device__ __forceinline__ Vec3 shfl_xor3(Vec3 v)
{ return v3(__shfl_xor_sync(0xffffffffu, v.x, 1), __shfl_xor_sync(0xffffffffu, v.y, 1), __shfl_xor_sync(0xffffffffu, v.z, 1)); }

__global__ void __launch_bounds__(256)
contact_kernel(EpaArgs a)
{
    const unsigned FULL = 0xffffffffu;
    const long long n_items = (long long)*a.wl_count;
    const long long n_slots = ((2 * n_items + 31) / 32) * 32;            // whole warps, so the shuffles stay convergent
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_slots; t += (long long)gridDim.x * blockDim.x) {
        const long long item = t >> 1;
        const int role = (int)(t & 1);                                    // 0: p1 along +n, 1: p2 along -n
        bool on = item < n_items;
        int p = 0; Vec3 nrm = v3(0, 0, 0);
        gjk::HullView H; H.n = 0; H.x = H.y = H.z = a.verts;
        if (on) {
            const double *r = a.result + 12 * item;
            on = r[3] >= 0.0;                                             // retired with a result
            nrm = v3(r[0], r[1], r[2]);
            p = a.wl_pair[item];
            H = gjk::hull_view(a.verts, a.hull_off, role ? a.pair_b[p] : a.pair_a[p]);
        }
        const int n = on ? H.n : 0;
        const Vec3 dir = role ? neg(nrm) : nrm;
        // pass 1: support vertex (strict >, first wins), near-maximum tracker (F:722-747), p1's z sum (F:446)
        double best = -DBL_MAX, mx = -DBL_MAX, zsum = 0.0; int ib = 0, id0 = 0, id1 = 0;
        for (int i = 0; i < n; ++i) {
            const Vec3 v = H.get(i);
            const double tt = dot3(dir, v);
            if (tt > best) { best = tt; ib = i; }
            if (tt > mx - 1.0e-8) { mx = tt; id1 = id0; id0 = i + 1; }
            zsum = gjk::add(zsum, v.z);
        }
        if (id1 == 0) id1 = id0;
        const Vec3 mine = on ? H.get(ib) : v3(0, 0, 0);                    // get_nearest_points, F:813-855
        const Vec3 theirs = shfl_xor3(mine);
        if (on && role == 0 && a.nearest) {
            double *o = a.nearest + 6 * p;
            o[0] = mine.x; o[1] = theirs.x; o[2] = mine.y; o[3] = theirs.y; o[4] = mine.z; o[5] = theirs.z;
        }
        const bool multi = id0 != id1;                                     // two near-maximal vertices on this hull
        const bool multi_other = __shfl_xor_sync(FULL, multi ? 1 : 0, 1) != 0;
        const Vec3 near_mine = (on && id0 > 0) ? H.get(id0 - 1) : v3(0, 0, 0);
        const Vec3 near_theirs = shfl_xor3(near_mine);
        const double zsum_other = __shfl_xor_sync(FULL, zsum, 1);
        const int n_other = __shfl_xor_sync(FULL, n, 1);
        Vec3 nrm_out = nrm;
        uint32_t st = 0;
        if (a.version == 1) {                                              // get_collisionPoint_01, F:700-806
            if (on && role == 0 && a.point) {
                const bool m1 = multi, m2 = multi_other;
                Vec3 r;
                if (!m1 && !m2) r = v3(gjk::add(near_mine.x, near_theirs.x) / 2.0, gjk::add(near_mine.y, near_theirs.y) / 2.0,
                                       gjk::add(near_mine.z, near_theirs.z) / 2.0);
                else if (m1 && !m2) r = near_theirs;
                else if (!m1 && m2) r = near_mine;
                else {                                                     // F:766-804: mean of p1 supports within 0.1 of the maximum
                    double sx = 0, sy = 0, sz = 0; int c = 0;
                    for (int i = 0; i < n; ++i) {
                        const Vec3 v = H.get(i);
                        if (dot3(dir, v) > best - 1.0e-1) { ++c; sx = gjk::add(sx, v.x); sy = gjk::add(sy, v.y); sz = gjk::add(sz, v.z); }
                    }
                    r = v3(sx / (double)c, sy / (double)c, sz / (double)c);
                }
                double *o = a.point + 3 * p;
                o[0] = r.x; o[1] = r.y; o[2] = r.z;
            }
        } else if (a.version == 3) {                                       // get_collisionPoint_03, F:426-452
            if (on && role == 1 && a.point) {
                double *o = a.point + 3 * p;
                o[0] = near_mine.x; o[1] = near_mine.y; o[2] = zsum_other / (double)(float)n_other;
            }
            const Vec3 nn = v3(nrm.x, nrm.y, 0.0);
            const double md = gjk::norm3(nn);
            nrm_out = v3(nn.x / md, nn.y / md, nn.z / md);                 // NaN when the normal is parallel to z, as the reference
            if (on && role == 0 && a.normal) { double *o = a.normal + 3 * p; o[0] = nrm_out.x; o[1] = nrm_out.y; o[2] = nrm_out.z; }
        } else {
            st |= epa::ST_BAD_VERSION;                                     // version 2 needs the un-vendored GCLIB_List; others STOP (F:336)
        }
        // get_info_collisionType, F:353-413, on the possibly re-projected normal (F:335-343)
        const Vec3 dir2 = role ? neg(nrm_out) : nrm_out;
        double m2v = best;
        if (a.version == 3) {
            m2v = -DBL_MAX;
            for (int i = 0; i < n; ++i) { const double tt = dot3(dir2, H.get(i)); if (tt > m2v) m2v = tt; }
        }
        int cnt = 0;
        const double thr = gjk::sub(m2v, a.tol_ff);
        for (int i = 0; i < n; ++i) cnt += dot3(dir2, H.get(i)) > thr;
        const int cnt_other = __shfl_xor_sync(FULL, cnt, 1);
        if (on && role == 0) {
            if (a.colli_type) a.colli_type[p] = (cnt >= 3 && cnt_other >= 3) ? 2 : 1;
            if (st && a.status) a.status[p] |= st;
        }
    }
}

inline int launch(const EpaArgs &a, int sm_count, cudaStream_t st)
{
    contact_kernel<<<sm_count * 8, 256, 0, st>>>(a);
    return 0;
}

}  // namespace contact
