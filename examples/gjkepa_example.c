/* examples/gjkepa_example.c -- the library from plain C (C99), the way a cgo / JNI / ctypes / ISO_C_BINDING stub would call it.
 *
 *   gcc -std=c99 -Iinclude examples/gjkepa_example.c -Lcollision_detect_gjk_epa_b200 -lgjkepa_b200 \
 *       -Wl,-rpath,$PWD/collision_detect_gjk_epa_b200 -lm -o build/gjkepa_example && build/gjkepa_example
 *
 * 1. the reference's own call, CALL GJKEPA(version_, TOL_FF_, p1_, p2_, ...) (F:39-52), on its documented test case: two unit
 *    cubes offset by 0.5 along x -> collision, depth 0.5, normal (1,0,0);
 * 2. the same pair and a separated one as a batch out of one hull store, with the contact list;
 * 3. the store left resident on the GPU(s), only pair lists in and compact contact rows out.
 * Prints one line per step and exits non-zero if a result is not the expected one. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "gjkepa.h"

/* p(n,3) column-major = [x(n) | y(n) | z(n)]: a unit cube centred at (cx, 0, 0) */
static void cube(double *p, double cx)
{
    int i;
    for (i = 0; i < 8; ++i) {
        p[i]      = cx + ((i & 1) ? 0.5 : -0.5);
        p[8 + i]  = (i & 2) ? 0.5 : -0.5;
        p[16 + i] = (i & 4) ? 0.5 : -0.5;
    }
}

static int fail(const char *what, int rc)
{
    fprintf(stderr, "%s failed: rc = %d\n", what, rc);
    return 1;
}

int main(void)
{
    const int32_t version = 1, n = 8;
    const double tol = 1.0;
    double store[3 * 24];                       /* three hulls: cube at 0, at 0.5, at 3.0 */
    const int64_t hull_off[4] = {0, 8, 16, 24};
    const int32_t pair_a[2] = {0, 0}, pair_b[2] = {1, 2};
    uint8_t flag = 0, flags[2];
    int32_t type = 0, types[2], index[2];
    double nearest[12], normal[6], point[6], depth[2];
    uint32_t status[2];
    int64_t n_contacts = -1;
    gjkepa_ctx *ctx = NULL;
    int rc;

    cube(store, 0.0); cube(store + 24, 0.5); cube(store + 48, 3.0);

    rc = gjkepa_pair_f64(&version, &tol, store, &n, store + 24, &n, &flag, &type, nearest, normal, point, depth);
    if (rc) return fail("gjkepa_pair_f64", rc);
    printf("pair:     collision %d  type %d  depth %.17g  normal (%g, %g, %g)\n", flag, type, depth[0], normal[0], normal[1], normal[2]);
    if (!flag || fabs(depth[0] - 0.5) > 1e-12 || fabs(normal[0] - 1.0) > 1e-12) return 2;

    rc = gjkepa_create_multi(&ctx, NULL, 0);    /* every visible GPU */
    if (rc) return fail("gjkepa_create_multi", rc);
    rc = gjkepa_batch_f64(ctx, version, tol, 3, store, hull_off, 2, pair_a, pair_b, GJKEPA_MODE_FULL,
                          flags, types, nearest, normal, point, depth, status, &n_contacts, index);
    if (rc) { fprintf(stderr, "%s\n", gjkepa_last_error(ctx)); return fail("gjkepa_batch_f64", rc); }
    printf("batch:    flags %d %d  contacts %lld (pair %d)  depth %.17g  exit paths %u %u on %d GPU(s)\n", flags[0], flags[1],
           (long long)n_contacts, index[0], depth[0], status[0] & 0xffu, status[1] & 0xffu, (int)gjkepa_device_count(ctx));
    if (flags[0] != 1 || flags[1] != 0 || n_contacts != 1 || index[0] != 0 || fabs(depth[0] - 0.5) > 1e-12) return 3;
    if ((status[1] & 0xffu) != GJKEPA_EXIT_SPHERE) return 4;       /* 3.0 apart: the envelope spheres already miss (F:77) */

    rc = gjkepa_store_upload(ctx, 3, store, hull_off);
    if (rc) return fail("gjkepa_store_upload", rc);
    depth[0] = depth[1] = -1.0;
    rc = gjkepa_batch_resident_f64(ctx, version, tol, 2, pair_a, pair_b, GJKEPA_MODE_EPA | GJKEPA_MODE_COMPACT,
                                   flags, NULL, NULL, normal, NULL, depth, NULL, &n_contacts, index);
    if (rc) { fprintf(stderr, "%s\n", gjkepa_last_error(ctx)); return fail("gjkepa_batch_resident_f64", rc); }
    printf("resident: contacts %lld  row 0 = pair %d, depth %.17g, normal (%g, %g, %g)\n", (long long)n_contacts, index[0], depth[0],
           normal[0], normal[1], normal[2]);
    if (n_contacts != 1 || index[0] != 0 || fabs(depth[0] - 0.5) > 1e-12 || fabs(normal[0] - 1.0) > 1e-12) return 5;
    gjkepa_destroy(ctx);
    printf("ok\n");
    return 0;
}
