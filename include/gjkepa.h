/* gjkepa.h -- C ABI of the B200-native batched GJK+EPA narrow phase (libgjkepa_b200.so).
 *
 * Drop-in boundary for ONE path of xiejihong0306/collision-detect-GJK-EPA: the public subroutine
 *     SUBROUTINE GJKEPA(version_, TOL_FF_, p1_, p2_, collision_, colliType_, nearest_points_,
 *                       collision_normal_, collision_point_, penetration_depth_)
 * of MODULE GCLIB_GJKEPA (reference src/GCLIB_GJKEPA.f90:39-52, cited below as F:<line>).
 * Plain pointers and sizes only; no C++/torch types.  All compute runs in hand-written sm_100a CUDA
 * kernels; there is NO CPU fallback -- every entry point returns GJKEPA_ERR_CUDA when no device
 * (or no loadable kernel image) is available.
 *
 * Array conventions (Fortran, column-major), identical to the reference dummy arguments:
 *   hull block   p(n,3)            -> [x(1..n) | y(1..n) | z(1..n)]            (F:46)
 *   nearest      nearest_points_(2,3) -> [p1x,p2x,p1y,p2y,p1z,p2z]             (F:49)
 *   normal/point (3)                                                            (F:50-51)
 * A hull store packs blocks back to back: hull h starts at double index 3*hull_off[h] and has
 * n_h = hull_off[h+1]-hull_off[h] vertices (hull_off in VERTEX units, n_hulls+1 entries).
 */
#ifndef GJKEPA_H
#define GJKEPA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gjkepa_ctx gjkepa_ctx;

/* return codes: never exit(), never block on stdin (the reference PAUSEs / READs / STOPs at
 * F:300-302, 337-339, 1370-1372; those paths become status bits below) */
#define GJKEPA_OK            0
#define GJKEPA_ERR_ARG       1
#define GJKEPA_ERR_CUDA      2
#define GJKEPA_ERR_NOMEM     3

/* how far along the reference call chain a batch goes */
#define GJKEPA_MODE_GJK      0   /* collision_ only                       (F:76-236)            */
#define GJKEPA_MODE_EPA      1   /* + penetration_depth_, collision_normal_ (F:263-323, 863-1022) */
#define GJKEPA_MODE_FULL     2   /* + nearest_points_, collision_point_, colliType_ (F:326-343)  */
/* OR-ed into the mode of a HOST-buffer batch call: the real-valued results, colli_type and status come back as one row
 * per CONTACT instead of one per pair -- row k belongs to pair contact_index[k] (ascending), *n_contacts rows in all
 * (both arguments are then required; every array needs room for n_pairs rows at worst).  collision stays one byte per
 * pair.  The rows are compacted on the device, so only contacts cross PCIe (45-75 % of the pairs of a broad-phase
 * batch are misses whose rows would be the zeros of F:65-71). */
#define GJKEPA_MODE_COMPACT  16

/* per-pair status word: bits 0-7 GJK exit path, 8-15 GJK walk iterations (F:182-236),
 * 16-23 EPA iterations (F:296-323), 24-31 flags */
#define GJKEPA_EXIT_SPHERE          1   /* F:77  */
#define GJKEPA_EXIT_INIT_CAP        2   /* F:86  */
#define GJKEPA_EXIT_DUP3            3   /* F:123 */
#define GJKEPA_EXIT_TRI_ORIGIN      4   /* F:140 (collision) */
#define GJKEPA_EXIT_COPLANAR4       5   /* F:157 */
#define GJKEPA_EXIT_INIT_CONTAIN    6   /* F:164 (collision) */
#define GJKEPA_EXIT_LOOP_CAP        7   /* F:186 */
#define GJKEPA_EXIT_LOOP_COLLINEAR  8   /* F:199 */
#define GJKEPA_EXIT_LOOP_COPLANAR   9   /* F:203 */
#define GJKEPA_EXIT_LOOP_CONTAIN    10  /* F:210 (collision) */
#define GJKEPA_EXIT_LOOP_STALL      11  /* F:219-234 */
#define GJKEPA_EXIT_STOP            12  /* F:1369 degenerate plane: reference STOPs; here FALSE + flag */
#define GJKEPA_ST_EPA_CAP      (1u << 24)  /* F:299 */
#define GJKEPA_ST_STOP         (1u << 25)
#define GJKEPA_ST_HULL_FAIL    (1u << 26)  /* EPA polytope degenerate or over capacity */
#define GJKEPA_ST_BAD_VERSION  (1u << 27)  /* version_ not 1, 2 or 3 (F:329-340: the reference PAUSEs and STOPs) */
#define GJKEPA_ST_H1           (1u << 28)  /* F:140 path (simplex row 4 defined as zero) */
#define GJKEPA_ST_TOPO_WARN    (1u << 29)  /* diagnostic: an EPA visible region was not a disk */
#define GJKEPA_ST_COLD_MATH    (1u << 30)  /* diagnostic: an operand left the straight-line sqrt/div range and
                                              the block was redone with the IEEE operators (same result) */

#define GJKEPA_ST_BAD_HINT     (1u << 31)  /* the pair has a hull larger than the bound given to gjkepa_set_max_hull_vertices:
                                              NOT processed (flag 0, outputs zero); fix the bound or set it to 0 */

/* ---- context: one GPU (gjkepa_create) or several GPUs of one box (gjkepa_create_multi) ----
 * Threading: a context owns its streams and its workspace and is used by ONE host thread at a time (several contexts on
 * one device may run from several threads).  gjkepa_pair_f64 is safe to call from any number of threads, OpenMP
 * threads included -- the reference is callable from an !$OMP PARALLEL DO (THREADPRIVATE scratch, F:55-60): concurrent
 * callers are coalesced into one small batch per round trip (see gjkepa_pair_f64).
 * Every entry point leaves the calling thread's current CUDA device as it found it. */
int  gjkepa_create(gjkepa_ctx **ctx, int32_t device);
/* One context over n_devices GPUs (devices == NULL or n_devices == 0: every visible GPU).  The reference's model is
 * "the caller runs the loop over pairs on all its cores" (F:55-60); here the HOST-buffer batch calls (gjkepa_batch_f64,
 * _f32, _posed_f64, _resident_f64) split the pair list into one contiguous range per device, one host thread and one
 * set of streams per device, and EVERY device copies its results straight into the caller's arrays at the pair's global
 * index (compact rows: at the row offset the earlier ranges determine) -- N PCIe links in parallel, no collective, no
 * gather through one GPU.  Results are bit-identical to a single-device context.  gjkepa_shapes_upload and
 * gjkepa_store_upload replicate to every device.  Calls that are one chain of passes over a whole scene
 * (gjkepa_broadphase_*, gjkepa_scene_posed_f64) run on the first device; gjkepa_batch_device_f64 and gjkepa_set_stream
 * need a single-device context (GJKEPA_ERR_ARG). */
int  gjkepa_create_multi(gjkepa_ctx **ctx, const int32_t *devices, int32_t n_devices);
int32_t gjkepa_device_count(const gjkepa_ctx *ctx);
int  gjkepa_f32_filter_stats(const gjkepa_ctx *ctx, int64_t *out /*[2]*/);
/* wall clock (ms) of the last HOST-buffer batch call: ms[0] = the shards (copy-in, kernels, per-pair results copied out;
 * devices side by side), ms[1] = the contact rows / contact_index copied to the caller's arrays at their row offsets */
int  gjkepa_host_phase_ms(const gjkepa_ctx *ctx, double *ms /*[2]*/);

/* ---- checking mode: GJKEPA_GUARD=1 in the environment before the library's first call ----
 * Every device allocation the library makes is then bracketed by two 256-byte bands holding a pattern and handed out filled
 * with 0xFF bytes (NaN / -1) instead of the zeros a fresh cudaMalloc usually holds; gjkepa_guard_check idles the devices and
 * reports the live allocations and how many bands no longer hold the pattern (bands found changed when an allocation was freed
 * are included).  The reference has nothing like it (its scratch is ALLOCATE / DEALLOCATE per call, F:274-277): this is the
 * library's own stand-in for a memory checker, used by tests/test_gpu_guard.py.  Outside checking mode both return
 * GJKEPA_ERR_ARG.  gjkepa_guard_selftest plants a write of overrun_bytes past the end (negative: before the start) of a
 * scratch allocation so that a test can see the detector fire. */
int  gjkepa_guard_check(int64_t *n_allocations, int64_t *n_bands_changed);
int  gjkepa_guard_selftest(int64_t overrun_bytes);
int  gjkepa_destroy(gjkepa_ctx *ctx);
/* Optional: page-lock an array the caller passes to the batch calls again and again (cudaHostRegister, portable).  The
 * arrays a Fortran host ALLOCATEs are pageable; copies from pageable memory go through the driver's staging buffer at a
 * fraction of PCIe speed and cannot overlap the kernels. */
int  gjkepa_host_register(void *p, int64_t bytes);
int  gjkepa_host_unregister(void *p);
/* A new context launches on its own non-blocking stream.  gjkepa_set_stream switches to the
 * caller's stream (cudaStream_t cast to void*, taken literally: NULL is CUDA's legacy default
 * stream); gjkepa_use_own_stream switches back. */
int  gjkepa_set_stream(gjkepa_ctx *ctx, void *cuda_stream);
int  gjkepa_use_own_stream(gjkepa_ctx *ctx);
const char *gjkepa_last_error(const gjkepa_ctx *ctx);
/* kernels launched by this context since creation (or since the last reset); for bench accounting */
int64_t gjkepa_launch_count(const gjkepa_ctx *ctx, int32_t reset);
/* optional per-kernel timing: when on, every batch records CUDA events around K0 (hull prep),
 * K1 (GJK) and K3 (EPA) on the launching stream; gjkepa_kernel_ms waits for the last batch and
 * returns their durations in milliseconds in ms[0..2] (0 for a kernel that did not run). */
int  gjkepa_set_profiling(gjkepa_ctx *ctx, int32_t on);
int  gjkepa_kernel_ms(gjkepa_ctx *ctx, double *ms /*[3]*/);
/* tuning knobs.  lanes: 0 = automatic (one thread per pair over shared-memory-staged hulls, a 4-lane
 * sub-warp per pair for hulls too large to stage); 1, 2 or 4 = sub-warp kernel only, that many lanes.
 * max_hull_vertices: caller's bound on vertices per hull (0 = unknown); sizes the shared-memory
 * columns and lets the large-hull pass be skipped. */
int  gjkepa_set_lanes(gjkepa_ctx *ctx, int32_t lanes);
int  gjkepa_set_max_hull_vertices(gjkepa_ctx *ctx, int32_t n);
/* further knobs, by key.  WALK_REFILL: the walk kernel re-packs a warp with new pairs when this many of
 * its 32 lanes are idle (1..32, default 8).  SINGLE_KERNEL: 1 = run the flag path as one kernel per tile
 * (no iteration compaction; kept for A/B measurements), 0 = init kernel + compacting walk kernel (default).
 * EPA_GENERAL: 1 = every colliding pair through the general EPA kernel (in-place face list), 0 = the linked-
 * mesh kernels first, the general kernel for what they decline (default). */
#define GJKEPA_TUNE_WALK_REFILL    1
#define GJKEPA_TUNE_SINGLE_KERNEL  2
#define GJKEPA_TUNE_EPA_GENERAL    3
#define GJKEPA_TUNE_TILE_MAX_VERTICES 4   /* 4..16 (default 16): largest hull run one thread per pair; larger ones get 4 lanes */
/* EARLY_OUT (default 0 = the parity build): 1 = shortcut E1.  The walk leaves as soon as one of its own support points s
 * along a unit direction d has s.d < -1e-6, which proves the hulls are apart; the reference answers FALSE for every
 * such pair (all its failure exits are FALSE, TRUE needs the origin inside or within 1e-8 of a tetrahedron of
 * Minkowski points), so the FLAGS are unchanged -- tests/test_gpu_parity.py checks that on every workload -- while the
 * exit path of those pairs reads GJKEPA_EXIT_EARLY_OUT and the walk is ~2x shorter for separated pairs.  The envelope-
 * sphere test (F:76-77) is dropped as well under this knob -- a sphere miss implies separation, which the walk detects --
 * so GJKEPA_EXIT_SPHERE is not reported either. */
#define GJKEPA_TUNE_EARLY_OUT      5
/* HOST_CHUNKS: a host-buffer batch runs as this many chunks per device through a copy-in / compute / copy-out pipeline
 * (0 = automatic: up to 8 chunks of at least 64k pairs; 1 = copy everything, compute, copy everything back). */
#define GJKEPA_TUNE_HOST_CHUNKS    6
/* LOCALITY_SORT (host-buffer batches against a shared or resident store): visit the pairs in the order of their first hull, so
 * that pairs sharing a hull find it in L2 -- what a broad phase's output order gives for free.  0 = never, 1 = always, 2 =
 * automatic (default): when the store is larger than L2 and a sample of the pair list is not already ordered.  Results are
 * written at the pair's own index and do not depend on this. */
#define GJKEPA_TUNE_LOCALITY_SORT  7
/* F32_FILTER (gjkepa_batch_f32 only; default 0): binary32 ARITHMETIC as a filter in front of the binary64 kernels.  A kernel
 * in float tests every pair for separation along the coordinate axes and the line of centroids with a margin above its own
 * rounding error; a pair it proves apart gets flag 0 and exit path GJKEPA_EXIT_EARLY_OUT without a binary64 walk, every other
 * pair runs the binary64 kernels as before.  The filter can only answer "separated" (by more than 1e-6: the argument of
 * EARLY_OUT above), so flags are unchanged -- tests/test_gpu_f32_mode.py: 0 mismatches -- while status differs for the filtered
 * pairs.  gjkepa_f32_filter_stats: out[0] = pairs of the last fp32 batch, out[1] = pairs the filter proved apart. */
#define GJKEPA_TUNE_F32_FILTER     8
/* F32_ARITH (gjkepa_batch_f32 only; default 0; takes precedence over F32_FILTER): the GJK walk itself in BINARY32 ARITHMETIC --
 * the reference's flag path (F:76-236) compiled with float operands from the same source text as the binary64 kernels
 * (csrc/gjk_core_body.inc), reading the float store as it arrived, no widening.  A pair that walk calls separated gets flag 0
 * and the walk's exit path and iteration count, marked GJKEPA_ST_F32_DECIDED.  GJKEPA_MODE_GJK: the pairs it calls colliding get
 * flag 1 the same way and no binary64 kernel runs at all.  GJKEPA_MODE_EPA / _FULL: those pairs are handed to the binary64
 * kernels, which confirm the flag in binary64 and produce depth, normal and the contact epilogue as for any batch.
 * The reference's absolute thresholds (1e-8 bands, 1e-12 / 1e-15 degeneracy tests) lie below binary32's rounding noise, so
 * this build scales them to its precision (1e-5, 1e-7, 1e-9; csrc/gjk_core_f32.cuh says why).  Flags can then differ from the
 * binary64 answer only for pairs that touch to within ~1e-5 of the hull scale: tests/test_host_sim.py (CPU) and
 * tests/test_gpu_f32_mode.py bound the rate at 1e-4 (measured: 0 of 1.1M random pairs) and check every differing pair's gap
 * with Qhull; bench.py reports rate and timing. */
#define GJKEPA_TUNE_F32_ARITH      9
#define GJKEPA_ST_F32_DECIDED  (32u << 16) /* status of a pair decided by the binary32 walk: 32 in the EPA-iterations field,
                                            * which is otherwise 0 for every pair without an EPA run */
#define GJKEPA_EXIT_EARLY_OUT      13
int  gjkepa_set_tuning(gjkepa_ctx *ctx, int32_t key, int32_t value);

/* ---- replaces CALL GJKEPA(...) (F:39-52): one pair, every scalar by reference, HOST arrays ----
 * Uses a process-wide default context on device 0 (created on first call).  Safe to call from any number of threads at once:
 * a caller that finds no batch in flight becomes the leader, takes every request queued so far (its own included), runs them as
 * ONE small batch and hands each caller its results; callers arriving meanwhile queue up for the next leader.  Throughput is
 * (calling threads) / (one GPU round trip, ~50-110 us), every result the bits the same pair gets in any batch call. */
int gjkepa_pair_f64(const int32_t *version, const double *tol_ff,
                    const double *p1, const int32_t *n1, const double *p2, const int32_t *n2,
                    uint8_t *collision, int32_t *colli_type,
                    double *nearest /*[6]*/, double *normal /*[3]*/, double *point /*[3]*/,
                    double *depth);

/* ---- batch form of the same call: n_pairs independent (p1_, p2_) pairs out of one hull store ----
 * HOST buffers; H2D / D2H copies are part of the call.  Output pointers may be NULL (not wanted).
 * Outputs are written at the pair's index and are zero for non-colliding pairs, as F:65-71.
 * contact_index (optional, n_pairs entries) receives the indices of colliding pairs in ascending
 * order, *n_contacts their count (both built on the device from the flags).  Large batches run in chunks: the copy-in
 * of a chunk overlaps the kernels of the one before and the copy-out of the one before that; when the pairs name
 * hulls in roughly ascending order (pair-private stores) a chunk brings along only the hull range it needs. */
int gjkepa_batch_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                     int64_t n_hulls, const double *verts, const int64_t *hull_off,
                     int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                     uint8_t *collision, int32_t *colli_type, double *nearest /*6*np*/,
                     double *normal /*3*np*/, double *point /*3*np*/, double *depth,
                     uint32_t *status, int64_t *n_contacts, int32_t *contact_index);

/* ---- hull store resident on the device(s): only pair lists go in, only results come out ----
 * A Fortran host cannot hold device pointers, so the store is named by the context.  gjkepa_store_upload copies a
 * world-space hull store (layout as above, HOST buffers) to every device of the context -- for a multi-device context in
 * parallel over their PCIe links -- and computes the envelope spheres of F:1175-1182 once; it replaces any earlier store.
 * gjkepa_batch_resident_f64 is gjkepa_batch_f64 against that store: per pair 8 B go in and 1 B (flags), 33 B
 * (+depth, normal) or 109 B (everything) come out, or only the contacts' rows with GJKEPA_MODE_COMPACT.  Results are
 * bit-identical to gjkepa_batch_f64 on the same store. */
int gjkepa_store_upload(gjkepa_ctx *ctx, int64_t n_hulls, const double *verts, const int64_t *hull_off);
int gjkepa_batch_resident_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                              int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                              uint8_t *collision, int32_t *colli_type, double *nearest /*6*np*/,
                              double *normal /*3*np*/, double *point /*3*np*/, double *depth,
                              uint32_t *status, int64_t *n_contacts, int32_t *contact_index);

/* ---- fp32 mode: binary32 at the boundary (north_star: "1e-5 if an fp32 mode is offered") ----
 * Same call with the hull store, the tolerance and the real-valued results as float: half the bytes over PCIe, which
 * bounds the host-buffer step.  By default the device widens the store once (exact) and runs the SAME binary64 kernels,
 * because the reference's decision thresholds (1e-8 bands F:157,203,1248; 1e-12 degeneracy tests F:1349,1388) lie below
 * binary32's epsilon and a binary32 walk cannot evaluate them as written (GJKEPA_TUNE_F32_ARITH above is the opt-in walk in
 * binary32 arithmetic, with thresholds scaled to it).  Hence, by default: flags, colli_type and status are exactly
 * what gjkepa_batch_f64 returns for the float-rounded vertices; depth / normal / nearest / point are that call's
 * doubles rounded once to float.  Against the unrounded binary64 hulls: depth and normal within 1e-5 (relative to
 * the hull scale), flags equal except for pairs whose gap or overlap is below the rounding of the vertices. */
int gjkepa_batch_f32(gjkepa_ctx *ctx, int32_t version, float tol_ff,
                     int64_t n_hulls, const float *verts, const int64_t *hull_off,
                     int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                     uint8_t *collision, int32_t *colli_type, float *nearest /*6*np*/,
                     float *normal /*3*np*/, float *point /*3*np*/, float *depth,
                     uint32_t *status, int64_t *n_contacts, int32_t *contact_index);

/* ---- same, DEVICE buffers already resident in HBM (no copies; asynchronous on the stream) ----
 * n_contacts_dev: device int64 (optional).  contact_index_dev: optional, UNORDERED on this path. */
int gjkepa_batch_device_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                            int64_t n_hulls, const double *verts_dev, const int64_t *hull_off_dev,
                            int64_t n_pairs, const int32_t *pair_a_dev, const int32_t *pair_b_dev,
                            int32_t mode,
                            uint8_t *collision_dev, int32_t *colli_type_dev, double *nearest_dev,
                            double *normal_dev, double *point_dev, double *depth_dev,
                            uint32_t *status_dev, int64_t *n_contacts_dev, int32_t *contact_index_dev);

/* ---- shapes resident on the device, poses per call (upstream of the reference path) ----
 * A caller of GJKEPA passes world-space vertices, i.e. it applies each body's pose on the host first.  With the
 * body-space shapes uploaded once, only poses cross PCIe per call (96 B per body instead of 24 B per vertex) and
 * the pose is applied on the device with this arithmetic (IEEE, every operation rounded, no FMA), R row-major:
 *     x' = ((r00*x + r01*y) + r02*z) + tx     (likewise y', z')
 * so a host that transforms the same way and calls gjkepa_batch_f64 gets bit-identical results.
 * shape_verts/shape_off: a hull store as above (HOST buffers, copied; replaces any earlier upload).
 * body_shape[b]: shape number of body b; pose: 12 doubles per body (r00 r01 r02 r10 .. r22 tx ty tz);
 * pair_a/pair_b: BODY numbers.  Outputs as gjkepa_batch_f64. */
int gjkepa_shapes_upload(gjkepa_ctx *ctx, int64_t n_shapes, const double *shape_verts, const int64_t *shape_off);
int gjkepa_batch_posed_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                           int64_t n_bodies, const int32_t *body_shape, const double *pose /*12*n_bodies*/,
                           int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                           uint8_t *collision, int32_t *colli_type, double *nearest /*6*np*/,
                           double *normal /*3*np*/, double *point /*3*np*/, double *depth,
                           uint32_t *status, int64_t *n_contacts, int32_t *contact_index);

/* ---- candidate pairs on the device (upstream of the reference path) ----
 * Emits every pair of bodies (a < b) that passes the reference's own rough test, RoughCollisionDetection_
 * SphericalEnvelope (F:1165-1188: NORM2(mp1 - mp2) <= r1 + r2 + 1.0, same arithmetic, so the SET is bit-exact with a host
 * that tests all pairs), found through a uniform grid.  pair_a/pair_b: HOST buffers of `cap` entries; *n_pairs
 * receives the number of passing pairs, which may exceed cap (then only the first cap were written: call again
 * with larger buffers).  The order is deterministic.  The _posed form takes bodies as gjkepa_batch_posed_f64 does. */
int gjkepa_broadphase_f64(gjkepa_ctx *ctx, int64_t n_hulls, const double *verts, const int64_t *hull_off,
                          int64_t cap, int32_t *pair_a, int32_t *pair_b, int64_t *n_pairs);
int gjkepa_broadphase_posed_f64(gjkepa_ctx *ctx, int64_t n_bodies, const int32_t *body_shape, const double *pose,
                                int64_t cap, int32_t *pair_a, int32_t *pair_b, int64_t *n_pairs);

/* ---- scene step: poses in, contacts out (pose application + candidate pairs + the reference path in one call) ----
 * For a simulation whose shapes are resident (gjkepa_shapes_upload): per step 96 B per body go to the device and only
 * the CONTACTS come back -- the candidate list never crosses PCIe.  Equivalent, result for result, to
 *     gjkepa_broadphase_posed_f64 -> gjkepa_batch_posed_f64 on its pairs -> keep the pairs whose collision_ is TRUE:
 * body_a/body_b (a < b) name the colliding bodies, rows sorted by (body_a, body_b); depth/normal (mode >= EPA) and
 * point/nearest/colli_type (mode FULL) are GJKEPA's outputs for that pair (F:47-52); any of them may be NULL.
 * All output arrays have `cap` rows; *n_contacts may exceed cap (then the first cap were written). */
int gjkepa_scene_posed_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                           int64_t n_bodies, const int32_t *body_shape, const double *pose /*12*n_bodies*/, int32_t mode,
                           int64_t cap, int32_t *body_a, int32_t *body_b, double *depth, double *normal /*3*cap*/,
                           double *point /*3*cap*/, double *nearest /*6*cap*/, int32_t *colli_type,
                           int64_t *n_candidates, int64_t *n_contacts);

/* self-test hook (host buffers, n elements each): the library's straight-line IEEE sqrt / division
 * sequences next to the compiler's operators; in_range bit 0 = division fast path valid, bit 1 = sqrt */
int gjkepa_debug_fp64(gjkepa_ctx *ctx, int64_t n, const double *a, const double *b, double *q_fast,
                      double *q_ref, double *s_fast, double *s_ref, uint8_t *in_range);

/* block until everything queued by this context has finished */
int gjkepa_sync(gjkepa_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
