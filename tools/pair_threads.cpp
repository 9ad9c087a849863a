// pair_threads.cpp -- the reference's own calling pattern against the drop-in: an OpenMP PARALLEL DO over pairs with one
// gjkepa_pair_f64 (= CALL GJKEPA, reference src/GCLIB_GJKEPA.f90:39-52, callable from OpenMP threads: F:55-60) per
// iteration.  Reports pairs/s for the given number of host threads and checks every result against ONE gjkepa_batch_f64
// call over the same pairs (flags and colli_type equal, depth/normal/point bit-identical: same kernels either way).
//   g++ -O2 -fopenmp tools/pair_threads.cpp -Iinclude -Lcollision_detect_gjk_epa_b200 -lgjkepa_b200 -Wl,-rpath,$PWD/collision_detect_gjk_epa_b200 -o build/pair_threads
//   build/pair_threads <threads> <pairs> [n_vert]
#include "../include/gjkepa.h"
#include <omp.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

int main(int argc, char **argv)
{
    const int threads = argc > 1 ? atoi(argv[1]) : 32;
    const int n_pairs = argc > 2 ? atoi(argv[2]) : 20000;
    const int nv = argc > 3 ? atoi(argv[3]) : 8;
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> N(0.0, 1.0);
    std::uniform_real_distribution<double> U(0.5, 1.0);
    // hull h: nv points on an ellipsoid, column-major p(nv,3) = [x | y | z]; the second hull of a pair is shifted by N(0, 0.85)
    std::vector<double> verts((size_t)2 * n_pairs * nv * 3);
    for (int h = 0; h < 2 * n_pairs; ++h) {
        double *p = &verts[(size_t)h * nv * 3];
        const double ax = U(rng), ay = U(rng), az = U(rng);
        const double ox = (h & 1) ? 0.85 * N(rng) : 0.0, oy = (h & 1) ? 0.85 * N(rng) : 0.0, oz = (h & 1) ? 0.85 * N(rng) : 0.0;
        for (int v = 0; v < nv; ++v) {
            double x = N(rng), y = N(rng), z = N(rng), r = std::sqrt(x * x + y * y + z * z);
            p[v] = ax * x / r + ox; p[nv + v] = ay * y / r + oy; p[2 * nv + v] = az * z / r + oz;
        }
    }
    std::vector<uint8_t> col(n_pairs), col_b(n_pairs);
    std::vector<int32_t> ty(n_pairs), ty_b(n_pairs);
    std::vector<double> nr(6 * (size_t)n_pairs), nm(3 * (size_t)n_pairs), pt(3 * (size_t)n_pairs), dp(n_pairs);
    std::vector<double> nr_b(nr.size()), nm_b(nm.size()), pt_b(pt.size()), dp_b(dp.size());
    const int32_t version = 1, n32 = nv;
    const double tol = 1.0;
    omp_set_num_threads(threads);
    int bad_rc = 0;
    auto run = [&](int count) {
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad_rc)
        for (int i = 0; i < count; ++i) {
            const double *p1 = &verts[(size_t)(2 * i) * nv * 3], *p2 = &verts[(size_t)(2 * i + 1) * nv * 3];
            bad_rc += gjkepa_pair_f64(&version, &tol, p1, &n32, p2, &n32, &col[i], &ty[i], &nr[6 * (size_t)i], &nm[3 * (size_t)i], &pt[3 * (size_t)i], &dp[i]) != 0;
        }
    };
    run(n_pairs < 2000 ? n_pairs : 2000);                     // warm-up: context creation, module load
    const auto t0 = std::chrono::steady_clock::now();
    run(n_pairs);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // the same pairs as one batch
    gjkepa_ctx *ctx = nullptr;
    if (gjkepa_create(&ctx, 0)) { fprintf(stderr, "gjkepa_create failed\n"); return 2; }
    std::vector<int64_t> off(2 * (size_t)n_pairs + 1);
    std::vector<int32_t> pa(n_pairs), pb(n_pairs);
    for (size_t h = 0; h < off.size(); ++h) off[h] = (int64_t)h * nv;
    for (int i = 0; i < n_pairs; ++i) { pa[i] = 2 * i; pb[i] = 2 * i + 1; }
    const int rc = gjkepa_batch_f64(ctx, version, tol, 2 * (int64_t)n_pairs, verts.data(), off.data(), n_pairs, pa.data(), pb.data(), GJKEPA_MODE_FULL,
                                    col_b.data(), ty_b.data(), nr_b.data(), nm_b.data(), pt_b.data(), dp_b.data(), nullptr, nullptr, nullptr);
    long long flag_diff = 0, type_diff = 0, real_diff = 0, hits = 0;
    for (int i = 0; i < n_pairs; ++i) {
        flag_diff += col[i] != col_b[i]; type_diff += ty[i] != ty_b[i]; hits += col[i] != 0;
        real_diff += dp[i] != dp_b[i] || std::memcmp(&nm[3 * (size_t)i], &nm_b[3 * (size_t)i], 24) != 0;
    }
    gjkepa_destroy(ctx);
    printf("{\"threads\": %d, \"pairs\": %d, \"n_vert\": %d, \"seconds\": %.6f, \"pairs_per_s\": %.1f, \"us_per_call_per_thread\": %.2f, "
           "\"colliding\": %lld, \"call_errors\": %d, \"batch_rc\": %d, \"flag_mismatches_vs_batch\": %lld, \"colli_type_mismatches\": %lld, "
           "\"depth_or_normal_not_bit_identical\": %lld}\n",
           threads, n_pairs, nv, s, n_pairs / s, s / n_pairs * threads * 1e6, hits, bad_rc, rc, flag_diff, type_diff, real_diff);
    return (bad_rc || rc || flag_diff) ? 1 : 0;
}
