// Host build of the product's GJK core (L = 1 lane instantiation of csrc/gjk_core.cuh) so that the
// logic shared with the CUDA kernels can be compared against the oracle on a machine with no GPU.
// Test-only: the shipped library never runs this path.
#define GJK_SHADOW_HISTORY 1
#include "../../collision_detect_gjk_epa_b200/csrc/gjk_core.cuh"
#include "../../collision_detect_gjk_epa_b200/csrc/gjk_core_f32.cuh"
#include <cstring>

static void gjk_batch_impl(const double *verts, const int64_t *hull_off, int64_t n_pairs,
                           const int32_t *pa, const int32_t *pb,
                           uint8_t *collision, uint32_t *status, int32_t *n_support, double *simplex, bool early)
{
    using namespace gjk;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n_pairs; ++i) {
        HullView A, B;
        int64_t oa = hull_off[pa[i]], ob = hull_off[pb[i]];
        A.n = (int)(hull_off[pa[i] + 1] - oa); A.x = verts + 3 * oa; A.y = A.x + A.n; A.z = A.y + A.n;
        B.n = (int)(hull_off[pb[i] + 1] - ob); B.x = verts + 3 * ob; B.y = B.x + B.n; B.z = B.y + B.n;
        Sphere sa = hull_sphere(A), sb = hull_sphere(B);
        Lanes<1> g;
        Vec3 S[4];
        GjkOut o = gjk_pair(g, A, B, sa, sb, S, early);
        collision[i] = exit_is_collision(o.exit_path) ? 1 : 0;
        status[i] = (uint32_t)o.exit_path | ((uint32_t)o.iters << 8) | (o.stop_flag ? (1u << 25) : 0u)
                  | (o.exit_path == EXIT_TRI_ORIGIN ? (1u << 28) : 0u);
        n_support[i] = o.n_support;
        if (simplex) std::memcpy(simplex + 12 * i, S, 12 * sizeof(double));
    }
}

extern "C" void hostsim_gjk_batch(const double *verts, const int64_t *hull_off, int64_t n_pairs,
                                  const int32_t *pa, const int32_t *pb,
                                  uint8_t *collision, uint32_t *status, int32_t *n_support, double *simplex)
{ gjk_batch_impl(verts, hull_off, n_pairs, pa, pb, collision, status, n_support, simplex, false); }

// the same walk with the opt-in shortcut E1 (GJKEPA_TUNE_EARLY_OUT): flags must not change, exit paths and counts may
extern "C" void hostsim_gjk_batch_early(const double *verts, const int64_t *hull_off, int64_t n_pairs,
                                        const int32_t *pa, const int32_t *pb,
                                        uint8_t *collision, uint32_t *status, int32_t *n_support, double *simplex)
{ gjk_batch_impl(verts, hull_off, n_pairs, pa, pb, collision, status, n_support, simplex, true); }

// Unit hook for the stall test F:219-234: the product keeps only the vertex the previous update dropped
// and the face it kept; the test harness passes explicit simplex_last_1 / simplex_last_2 arrays as the
// reference holds them and compares the two answers.
extern "C" int hostsim_stall(const double *last1, const double *last2, int have2, int k1, int k, const double *sp_)
{
    using namespace gjk;
    Vec3 p[4], l2[4];
    for (int i = 0; i < 4; ++i) { p[i] = v3(last1[3 * i], last1[3 * i + 1], last1[3 * i + 2]); l2[i] = v3(last2[3 * i], last2[3 * i + 1], last2[3 * i + 2]); }
    int a1, b1, c1; face_rows(k1, a1, b1, c1);
    const int d1 = 6 - a1 - b1 - c1;
    int a, b, c; face_rows(k, a, b, c);
    Vec3 sp = v3(sp_[0], sp_[1], sp_[2]);
    (void)a; (void)b; (void)c;
    return stall_test(p[0], p[1], p[2], p[3], k, sp, l2[d1], k1, have2 != 0) ? 1 : 0;
}

// walk iterations whose history bookkeeping was checked against explicit simplex_last_1/_2 arrays, and how
// many disagreed (must be 0)
extern "C" void hostsim_shadow_counts(long long *checks, long long *mismatch)
{
    *checks = gjk::g_shadow_checks; *mismatch = gjk::g_shadow_mismatch;
}

// The binary32 instantiation of the same text (csrc/gjk_core_f32.cuh, GJKEPA_TUNE_F32_ARITH): float vertices in, flags and the
// walk's exit path / iteration count out.
extern "C" void hostsim_gjk32_batch(const float *verts, const int64_t *hull_off, int64_t n_pairs,
                                    const int32_t *pa, const int32_t *pb, uint8_t *collision, uint32_t *status)
{
    using namespace gjk32;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n_pairs; ++i) {
        HullView A, B;
        int64_t oa = hull_off[pa[i]], ob = hull_off[pb[i]];
        A.n = (int)(hull_off[pa[i] + 1] - oa); A.x = verts + 3 * oa; A.y = A.x + A.n; A.z = A.y + A.n;
        B.n = (int)(hull_off[pb[i] + 1] - ob); B.x = verts + 3 * ob; B.y = B.x + B.n; B.z = B.y + B.n;
        Sphere sa = hull_sphere(A), sb = hull_sphere(B);
        Lanes<1> g;
        Vec3 S[4];
        GjkOut o = gjk_pair(g, A, B, sa, sb, S, false);
        collision[i] = exit_is_collision(o.exit_path) ? 1 : 0;
        status[i] = (uint32_t)o.exit_path | ((uint32_t)o.iters << 8) | (o.stop_flag ? (1u << 25) : 0u);
    }
}
