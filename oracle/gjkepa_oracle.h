/* gjkepa_oracle.h -- CPU restatement of GCLIB_GJKEPA (reference src/GCLIB_GJKEPA.f90, "F:<line>").
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker / the timed CPU baseline.  The CUDA library never links or calls it.
 *
 * PARITY UNPINNED (EPA hull only): the reference repository has no tests, no golden vectors and
 * cannot be compiled here (no Fortran compiler; it USEs three un-vendored modules).  The GJK flag
 * path (F:39-239 + helpers) is fully specified in-repo and restated statement by statement; the
 * EPA loop (F:242-323, F:863-1022) calls an external `QuickHull`/`getHullMeshesVertex` whose source
 * is absent, so this file supplies its own triangulating hull (qh_* below) and the results are
 * pinned only by analytic cube cases and by scipy/Qhull ground truth (tests/).
 */
#ifndef GJKEPA_ORACLE_H
#define GJKEPA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* exit paths of GJKEPA (F:39-239); values are shared with the CUDA library's status word */
enum {
    ORC_EXIT_NONE          = 0,
    ORC_EXIT_SPHERE        = 1,  /* F:77   envelope spheres disjoint            -> FALSE */
    ORC_EXIT_INIT_CAP      = 2,  /* F:86   100 table rows exhausted             -> FALSE */
    ORC_EXIT_DUP3          = 3,  /* F:123  3rd support repeats 1st/2nd          -> FALSE */
    ORC_EXIT_TRI_ORIGIN    = 4,  /* F:140  origin on the first triangle         -> TRUE, EPA (hazard H1) */
    ORC_EXIT_COPLANAR4     = 5,  /* F:157  4th support coplanar                 -> FALSE */
    ORC_EXIT_INIT_CONTAIN  = 6,  /* F:164  initial tetra holds the origin       -> TRUE, EPA */
    ORC_EXIT_LOOP_CAP      = 7,  /* F:186  50 walk steps                        -> FALSE */
    ORC_EXIT_LOOP_COLLINEAR= 8,  /* F:199                                       -> FALSE */
    ORC_EXIT_LOOP_COPLANAR = 9,  /* F:203                                       -> FALSE */
    ORC_EXIT_LOOP_CONTAIN  = 10, /* F:210                                       -> TRUE, EPA */
    ORC_EXIT_LOOP_STALL    = 11, /* F:219-234 simplex repeats last two          -> FALSE */
    ORC_EXIT_STOP          = 12  /* F:1369 DIST_PF_SIGN degenerate plane: the reference blocks on
                                    stdin and STOPs; defined here as FALSE + ORC_ST_STOP */
};

/* status word: bits 0-7 exit path, 8-15 GJK walk iterations, 16-23 EPA iterations, 24-31 flags */
#define ORC_ST_EPA_CAP      (1u << 24) /* F:299 EPA iteration cap hit (reference PAUSEs) */
#define ORC_ST_STOP         (1u << 25) /* a reference STOP path was reached */
#define ORC_ST_HULL_FAIL    (1u << 26) /* hull stand-in could not build a tetrahedron */
#define ORC_ST_BAD_VERSION  (1u << 27) /* version_ not in {1,3}: reference STOPs (F:336) or needs GCLIB_List (version 2) */
#define ORC_ST_H1           (1u << 28) /* F:140 path taken with simplex row 4 := 0 */

enum { ORC_MODE_GJK = 0, ORC_MODE_EPA = 1, ORC_MODE_FULL = 2 };

typedef struct {
    int32_t collision;      /* collision_        (F:47) 0/1 */
    int32_t colli_type;     /* colliType_        (F:48) */
    double  nearest[6];     /* nearest_points_(2,3) column-major: p1x,p2x,p1y,p2y,p1z,p2z (F:49) */
    double  normal[3];      /* collision_normal_ (F:50) */
    double  point[3];       /* collision_point_  (F:51) */
    double  depth;          /* penetration_depth_(F:52) */
    uint32_t status;
    int32_t n_support;      /* support_mapping calls made */
    double  simplex[12];    /* final simplex(4,3), row-major here: row i = simplex[3*i..] */
} orc_result;

/* One pair.  p1, p2: Fortran p(n,3) column-major blocks = [x(n) | y(n) | z(n)].
 * mode: ORC_MODE_GJK (flag only; EPA skipped, flag identical), ORC_MODE_EPA (+depth, normal),
 * ORC_MODE_FULL (+nearest points, collision point by `version`, collision type).           */
void orc_gjkepa_pair(int32_t version, double tol_ff,
                     const double *p1, int32_t n1, const double *p2, int32_t n2,
                     int32_t mode, orc_result *out);

/* Batch over a packed hull store: hull h = verts[3*hull_off[h] ...], n = hull_off[h+1]-hull_off[h],
 * stored as the Fortran block above.  OpenMP parallel-for over pairs (schedule(dynamic,1024)),
 * the structure the reference's THREADPRIVATE design is meant for.  Any output pointer may be NULL.
 * Returns the number of threads used. */
int orc_gjkepa_batch(int32_t version, double tol_ff,
                     const double *verts, const int64_t *hull_off,
                     int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b,
                     int32_t mode, int32_t n_threads,
                     uint8_t *collision, int32_t *colli_type, double *nearest /*6*np*/,
                     double *normal /*3*np*/, double *point /*3*np*/, double *depth,
                     uint32_t *status, int32_t *n_support);

/* exposed pieces, for unit tests */
void orc_epilogue(int32_t version, double tol_ff, const double *p1, int32_t n1, const double *p2, int32_t n2,
                  const double normal_in[3], double nearest[6], double point[3], double normal_out[3], int32_t *colli_type);
void orc_support(const double *p1, int32_t n1, const double *p2, int32_t n2,
                 const double dir[3], double res[3], int32_t *i1, int32_t *i2);
int  orc_is_inside_pf(const double *v /*nnode*3 row-major*/, int32_t nnode, const double p[3]);
int  orc_point_in_simplex(const double p[3], const double s[12]);
int  orc_sphere_test(const double *p1, int32_t n1, const double *p2, int32_t n2);
/* hull stand-in: pts (npts,3) row-major -> tris (ntri,9) row-major; returns 0 on success */
int  orc_quickhull(const double *pts, int32_t npts, double *tris, int32_t *ntri, int32_t cap);
void orc_dir_table(int32_t row1based, double d[3]);
/* all pairs (a < b) passing the F:1185 sphere test, brute force, lexicographic; returns the count */
int64_t orc_sphere_pairs(const double *verts, const int64_t *hull_off, int64_t n_hulls,
                         int64_t cap, int32_t *pair_a, int32_t *pair_b);
int  orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
