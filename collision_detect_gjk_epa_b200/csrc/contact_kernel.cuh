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

// ---- get_collisionPoint_02 (F:457-696) pieces; GCLIB_List's List_Array3d is an append-only array here ----
constexpr int CP2_CAP = 128;            // support vertices kept per hull (a hull with more is flagged, not truncated silently)

__device__ __forceinline__ Vec3 utz(Vec3 v)
{   // UTZVEC, F:1343-1351
    const double md = gjk::norm3(v);
    return md < 1.0e-12 ? v3(0, 0, 0) : v3(v.x / md, v.y / md, v.z / md);
}
__device__ __forceinline__ Vec3 foot_pl(Vec3 p, Vec3 v1, Vec3 v2)
{   // FOOT_PL, F:1496-1510
    const Vec3 u = utz(v2 - v1);
    const double t = dot3(p - v1, u);
    return v3(gjk::add(v1.x, gjk::mul(t, u.x)), gjk::add(v1.y, gjk::mul(t, u.y)), gjk::add(v1.z, gjk::mul(t, u.z)));
}
__device__ __forceinline__ void foot_ll(Vec3 p1, Vec3 q1, Vec3 p2, Vec3 q2, Vec3 &f1, Vec3 &f2)
{   // FOOT_LL, F:1445-1491
    const Vec3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    const double a = dot3(d1, d1), b = dot3(d1, d2), c = dot3(d1, r), e = dot3(d2, d2), f = dot3(d2, r);
    const double d = gjk::sub(gjk::mul(a, e), gjk::mul(b, b));
    if (fabs(d) < 1.0e-12) {
        f1 = v3(gjk::add(p1.x, q1.x) / 2.0, gjk::add(p1.y, q1.y) / 2.0, gjk::add(p1.z, q1.z) / 2.0);
        f2 = foot_pl(f1, p2, q2);
    } else {
        const double s = gjk::sub(gjk::mul(b, f), gjk::mul(c, e)) / d, t = gjk::sub(gjk::mul(a, f), gjk::mul(b, c)) / d;
        f1 = v3(gjk::add(p1.x, gjk::mul(s, d1.x)), gjk::add(p1.y, gjk::mul(s, d1.y)), gjk::add(p1.z, gjk::mul(s, d1.z)));
        f2 = v3(gjk::add(p2.x, gjk::mul(t, d2.x)), gjk::add(p2.y, gjk::mul(t, d2.y)), gjk::add(p2.z, gjk::mul(t, d2.z)));
    }
}
// SORT_CLOCK, F:1517-1575, on the support vertices idx[0..m) of hull H; ord[] receives positions into idx[].
// Same definitions as the oracle for the reference's undefined corners (all points coincident: input order;
// "already placed" compares coordinates with the rows written so far).
__device__ void sort_clock(const gjk::HullView &H, const unsigned short *idx, int m, unsigned char *ord)
{
    bool overlap = true;
    for (int i = 0; i < m - 1 && overlap; ++i)
        for (int j = i + 1; j < m && overlap; ++j) {
            const Vec3 d = H.get(idx[i]) - H.get(idx[j]);
            if (fabs(d.x) > 1.0e-12 || fabs(d.y) > 1.0e-12 || fabs(d.z) > 1.0e-12) overlap = false;
        }
    for (int i = 0; i < m; ++i) ord[i] = (unsigned char)i;
    if (overlap) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = 0; i < m; ++i) { const Vec3 v = H.get(idx[i]); sx = gjk::add(sx, v.x); sy = gjk::add(sy, v.y); sz = gjk::add(sz, v.z); }
    const Vec3 cen = v3(sx / (double)m, sy / (double)m, sz / (double)m);
    const Vec3 p0 = H.get(idx[0]);
    const Vec3 nrm = gjk::cross3(H.get(idx[1]) - p0, H.get(idx[2]) - p0);
    const double two_pi = (double)(2.0f * acosf(-1.0f));             // single precision, F:1547
    for (int i = 1; i < m; ++i) {
        double min_angle = DBL_MAX; int index = -1;
        const Vec3 v2 = H.get(idx[ord[i - 1]]) - cen;
        for (int j = 0; j < m; ++j) {
            const Vec3 pj = H.get(idx[j]);
            bool used = false;
            for (int q = 0; q < i && !used; ++q) { const Vec3 o = H.get(idx[ord[q]]); used = (pj.x == o.x && pj.y == o.y && pj.z == o.z); }
            if (used) continue;
            const Vec3 v1 = pj - cen;
            double ang = atan2(dot3(nrm, gjk::cross3(v2, v1)), dot3(v1, v2));
            ang = fmod(ang + two_pi, two_pi);
            if (ang < 0.0) ang += two_pi;
            if (ang < min_angle) { min_angle = ang; index = j; }
        }
        ord[i] = (unsigned char)(index < 0 ? 0 : index);
    }
}
// IS_INSIDE_PF (F:1271-1337) for the polygon H[idx[ord[0..m)]] and point p
__device__ bool inside_polygon(const gjk::HullView &H, const unsigned short *idx, const unsigned char *ord, int m, Vec3 p)
{
    double c0 = 0.0; bool any_pos = false;
    for (int i = 0; i < m; ++i) {
        const Vec3 vi = H.get(idx[ord[i]]), vj = H.get(idx[ord[i == m - 1 ? 0 : i + 1]]);
        double c = gjk::sub(gjk::mul(gjk::sub(vj.x, vi.x), gjk::sub(p.y, vi.y)), gjk::mul(gjk::sub(vj.y, vi.y), gjk::sub(p.x, vi.x)));
        if (fabs(c) < 1.0e-12) c = 0.0;
        if (c > 1.0e-15) any_pos = true;
        if (i == 0) c0 = c;
    }
    const bool xz = !any_pos;                                         // F:1310: re-project on XZ (no re-zeroing)
    bool inside = true;
    for (int i = 0; i < m; ++i) {
        const Vec3 vi = H.get(idx[ord[i]]), vj = H.get(idx[ord[i == m - 1 ? 0 : i + 1]]);
        double c;
        if (xz) c = gjk::sub(gjk::mul(gjk::sub(vj.x, vi.x), gjk::sub(p.z, vi.z)), gjk::mul(gjk::sub(vj.z, vi.z), gjk::sub(p.x, vi.x)));
        else { c = gjk::sub(gjk::mul(gjk::sub(vj.x, vi.x), gjk::sub(p.y, vi.y)), gjk::mul(gjk::sub(vj.y, vi.y), gjk::sub(p.x, vi.x))); if (fabs(c) < 1.0e-12) c = 0.0; }
        if (i == 0) c0 = c;
        if (gjk::mul(c0, c) < 0.0) inside = false;
    }
    return inside;
}

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
        } else if (a.version == 2) {                                       // get_collisionPoint_02, F:457-696
            unsigned short idx[CP2_CAP]; unsigned char ord[CP2_CAP];
            int m = 0;                                                     // AddAllSupports F:510-530: within 0.1 of the maximum
            for (int i = 0; i < n; ++i)
                if (dot3(dir, H.get(i)) > best - 1.0e-1) { if (m < CP2_CAP) idx[m] = (unsigned short)i; ++m; }
            if (m > CP2_CAP) { m = CP2_CAP; st |= epa::ST_HULL_FAIL; }
            const int m_other = __shfl_xor_sync(FULL, m, 1);
            const Vec3 s0 = (on && m > 0) ? H.get(idx[0]) : v3(0, 0, 0), s1 = (on && m > 1) ? H.get(idx[1]) : v3(0, 0, 0);
            const Vec3 o0 = shfl_xor3(s0), o1 = shfl_xor3(s1);
            const int n1 = role == 0 ? m : m_other, n2 = role == 0 ? m_other : m;
            // the lane that holds what the case needs computes and writes
            Vec3 r = v3(0, 0, 0); bool mine_writes = false;
            if (n1 == 1 && n2 == 1) { mine_writes = role == 0; r = v3(gjk::add(s0.x, o0.x) / 2.0, gjk::add(s0.y, o0.y) / 2.0, gjk::add(s0.z, o0.z) / 2.0); }
            else if (n1 == 1 && n2 >= 2) { mine_writes = role == 0; r = s0; }                        // case_02 on p1's vertex
            else if (n1 >= 2 && n2 == 1) { mine_writes = role == 1; r = s0; }                        // case_02 on p2's vertex
            else if (n1 == 2 && n2 == 2) {                                                            // case_03
                mine_writes = role == 0;
                Vec3 f1, f2;
                foot_ll(s0, s1, o0, o1, f1, f2);
                r = v3(gjk::add(f1.x, f2.x) / 2.0, gjk::add(f1.y, f2.y) / 2.0, gjk::add(f1.z, f2.z) / 2.0);
            } else if ((n1 == 2 && n2 >= 3) || (n1 >= 3 && n2 == 2)) {                                // case_04: polygon vs segment
                mine_writes = m >= 3;
                if (on && mine_writes) {
                    sort_clock(H, idx, m, ord);
                    const int cin = (inside_polygon(H, idx, ord, m, o0) ? 1 : 0) + (inside_polygon(H, idx, ord, m, o1) ? 1 : 0);
                    if (cin == 0) {
                        double sx = 0, sy = 0, sz = 0;
                        for (int i = 0; i < m; ++i) { const Vec3 v = H.get(idx[i]); sx = gjk::add(sx, v.x); sy = gjk::add(sy, v.y); sz = gjk::add(sz, v.z); }
                        r = foot_pl(v3(sx / (double)m, sy / (double)m, sz / (double)m), o0, o1);
                    } else r = v3(gjk::mul(gjk::add(o0.x, o1.x), 0.5), gjk::mul(gjk::add(o0.y, o1.y), 0.5), gjk::mul(gjk::add(o0.z, o1.z), 0.5));
                }
            } else if (n1 >= 3 && n2 >= 3) {                                                          // case_05: mean of p1's supports
                mine_writes = role == 0;
                if (on && mine_writes) {
                    double sx = 0, sy = 0, sz = 0;
                    for (int i = 0; i < m; ++i) { const Vec3 v = H.get(idx[i]); sx = gjk::add(sx, v.x); sy = gjk::add(sy, v.y); sz = gjk::add(sz, v.z); }
                    r = v3(sx / (double)m, sy / (double)m, sz / (double)m);
                }
            }
            if (on && mine_writes && a.point) { double *o = a.point + 3 * p; o[0] = r.x; o[1] = r.y; o[2] = r.z; }
        } else {
            st |= epa::ST_BAD_VERSION;                                     // the reference PAUSEs and STOPs, F:336-339
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
        const uint32_t st_all = st | __shfl_xor_sync(FULL, st, 1);
        if (on && role == 0) {
            if (a.colli_type) a.colli_type[p] = (cnt >= 3 && cnt_other >= 3) ? 2 : 1;
            if (st_all && a.status) a.status[p] |= st_all;
        }
    }
}

inline int launch(const EpaArgs &a, int sm_count, cudaStream_t st)
{
    contact_kernel<<<sm_count * 8, 256, 0, st>>>(a);
    return 0;
}

}  // namespace contact
