// gjkepa_capi.cu -- C ABI (include/gjkepa.h) and host batch scheduler of libgjkepa_b200.so.
//
// Replaces, for batches, the call  CALL GJKEPA(version_, TOL_FF_, p1_, p2_, collision_, ...)
// (reference src/GCLIB_GJKEPA.f90:39-52).  gjkepa_batch_device_f64 is the device pipeline, all on one CUDA stream:
//   K0 hull_prep (shared stores) -> K1 init + compacting walk kernels (flags + EPA work list) -> EPA routing by hull size ->
//   linked-mesh / general EPA kernels on the colliding pairs only -> K4 contact epilogue.
// The host-buffer entries (gjkepa_batch_f64 / _f32 / _posed_f64 / _resident_f64) sit on top of it in host_batch.inl: chunked
// copy / compute / copy pipeline, device-side contact lists, one shard per device of a multi-device context.
// There is no CPU compute path: without a usable device every entry point fails.
#include "../../include/gjkepa.h"
#include "gjk_kernels.cuh"
#include "gjk_two_stage.cuh"
#include "gjk_two_stage_group.cuh"
#include "epa_kernel.cuh"
#include "epa_v3.cuh"
#include "contact_kernel.cuh"
#include "pose_kernel.cuh"
#include "broadphase.cuh"
#include "f32_io.cuh"
#include "gjk_f32.cuh"
#include "scene.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

// ---- checking mode (GJKEPA_GUARD=1 in the environment, read once) ----------------------------------------------------
// compute-sanitizer is not available where this library is developed, so the library carries its own detector for the two
// faults a workspace scheduler can hide: every device allocation it makes gets a 256-byte band on either side filled with
// a pattern (a write past either end lands in a band; gjkepa_guard_check counts the bands that changed), and the payload is
// filled with 0xFF bytes (NaN as a double, -1 as an index) instead of whatever cudaMalloc returns -- usually zeros, which
// hides a kernel that reads workspace it never wrote.  tests/test_gpu_guard.py runs the whole pipeline in this mode.
// Off (the default) the wrappers are the plain runtime calls.
namespace guard {
constexpr size_t BAND = 256;
constexpr unsigned char FILL = 0xA5;
struct Rec { char *base; size_t bytes; int dev; };
static std::mutex g_mu;
static std::vector<Rec> g_live;
static std::atomic<long long> g_bad_at_free{0};
static bool enabled()
{
    static const bool on = [] { const char *e = std::getenv("GJKEPA_GUARD"); return e && *e && *e != '0'; }();
    return on;
}
static int bands_changed(const Rec &r)           // the caller has made r.dev current and idle
{
    unsigned char h[2 * BAND];
    if (cudaMemcpy(h, r.base, BAND, cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
    if (cudaMemcpy(h + BAND, r.base + BAND + r.bytes, BAND, cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
    int bad = 0;
    for (size_t k = 0; k < 2; ++k)
        for (size_t i = 0; i < BAND; ++i) if (h[k * BAND + i] != FILL) { ++bad; break; }
    return bad;
}
static cudaError_t gmalloc(void **p, size_t bytes)
{
    if (!enabled()) return ::cudaMalloc(p, bytes);
    char *base = nullptr;
    cudaError_t e = ::cudaMalloc((void **)&base, bytes + 2 * BAND);
    if (e != cudaSuccess) { *p = nullptr; return e; }
    e = cudaMemset(base, FILL, BAND);
    if (e == cudaSuccess && bytes) e = cudaMemset(base + BAND, 0xFF, bytes);
    if (e == cudaSuccess) e = cudaMemset(base + BAND + bytes, FILL, BAND);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { ::cudaFree(base); *p = nullptr; return e; }
    Rec r{base, bytes, 0};
    cudaGetDevice(&r.dev);
    { std::lock_guard<std::mutex> lk(g_mu); g_live.push_back(r); }
    *p = base + BAND;
    return cudaSuccess;
}
static cudaError_t gfree(void *p)
{
    if (!enabled() || !p) return ::cudaFree(p);
    Rec r{nullptr, 0, 0};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (size_t i = 0; i < g_live.size(); ++i)
            if (g_live[i].base + BAND == (char *)p) { r = g_live[i]; g_live[i] = g_live.back(); g_live.pop_back(); break; }
    }
    if (!r.base) return ::cudaFree(p);
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(r.dev);
    cudaDeviceSynchronize();
    g_bad_at_free += bands_changed(r);
    const cudaError_t e = ::cudaFree(r.base);
    cudaSetDevice(cur);
    return e;
}
}  // namespace guard
#define cudaMalloc(p, n) guard::gmalloc((void **)(p), (n))
#define cudaFree(p) guard::gfree((void *)(p))

constexpr int hostb_max_chunks = 64;
constexpr int N_COUNTERS = 64;           // device counter block of a context (work lists, queue heads, EPA routing)
struct HostPipe;       // host_batch.inl: chunked copy/compute pipeline of the host-buffer entries
struct MultiDev;       // host_batch.inl: one child context + worker thread per device

struct gjkepa_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int lanes = 0;                 // 0 = automatic (tile kernel + 4-lane kernel for large hulls)
    int max_hull_hint = 0;         // caller's bound on vertices per hull (0 = unknown)
    int tile_ctas_per_sm = 2;      // persistent grid of the tile kernel
    int walk_refill = 8;           // K1w re-packs a warp when this many lanes are idle
    bool single_kernel = false;    // monolithic tile kernel instead of K1i + K1w
    bool epa_general = false;      // general EPA kernel only (no linked-mesh fast path)
    int early_out = 0;             // shortcut E1 (opt-in): see gjk_core.cuh
    int tile_max_vertices = 16;    // largest hull the thread-per-pair kernels stage (larger ones: 4-lane kernels)
    double *d_seed_f64 = nullptr; int32_t *d_seed_pair = nullptr; uint32_t *d_seed_fl = nullptr; int64_t *d_seed_off = nullptr; int64_t cap_seed = 0;
    uint2 *d_seed_idx = nullptr;
    int32_t *d_ovf = nullptr; int64_t cap_ovf = 0;
    int32_t *d_ovf2 = nullptr; int64_t cap_ovf2 = 0;
    int32_t *d_route = nullptr; int64_t cap_route = 0;        // EPA routing lists (small | big), 2 x n_pairs
    // workspace (device)
    gjk::Sphere *d_sphere = nullptr; int64_t cap_sphere = 0;
    int32_t *d_wl_pair = nullptr; double *d_wl_simplex = nullptr; int64_t cap_wl = 0;
    unsigned long long *d_counters = nullptr;      // [0] work-list length, [1] EPA queue head, [2] GJK overflow, [3] EPA pass-2 head, [4] EPA overflow
    // staging for the host-buffer entry (device)
    void *d_stage = nullptr; size_t cap_stage = 0;
    // resident shapes (gjkepa_shapes_upload) and the world-space store rebuilt from poses per call
    double *d_shape_verts = nullptr; int64_t *d_shape_off = nullptr; int64_t n_shapes = 0, max_shape_verts = 0;
    double *d_world = nullptr; int64_t cap_world = 0;            // doubles
    int64_t *d_world_off = nullptr; int64_t cap_world_off = 0;   // sizes, then offsets (n_bodies + 1 each)
    void *d_scan_tmp = nullptr; size_t cap_scan_tmp = 0;
    void *d_stage2 = nullptr; size_t cap_stage2 = 0;
    void *d_bp = nullptr; size_t cap_bp = 0;                     // broad-phase workspace
    int64_t bp_layout_n = -1; size_t bp_sort_tmp = 0, bp_scan_tmp = 0;
    int32_t *d_cand_a = nullptr, *d_cand_b = nullptr; int64_t cap_cand = 0;   // scene step: candidate pairs
    void *d_scene = nullptr; size_t cap_scene = 0;               // scene step: per-candidate results, compacted contacts
    unsigned long long *h_counters = nullptr;      // pinned
    char *h_small = nullptr, *d_small = nullptr;   // packed staging for tiny batches (pinned host mirror + device block)
    // optional per-kernel timing (CUDA events on the launching stream)
    bool profiling = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid[3] = {false, false, false};      // K0, K1, K3 recorded by the last batch
    // envelope spheres computed ahead of the batch (resident store, or once per host batch that is run in chunks): K0 is skipped
    const gjk::Sphere *ext_sphere = nullptr;
    const unsigned long long *ext_order = nullptr;      // visiting order of the next batch's pairs (host_batch.inl: locality sort, fp32 filter), or null
    const unsigned long long *ext_order_count = nullptr;   // its length on the device (fp32 filter: the survivors), or null = n_pairs
    int f32_filter = 0;                                  // GJKEPA_TUNE_F32_FILTER
    int f32_arith = 0;                                   // GJKEPA_TUNE_F32_ARITH: the GJK walk of the fp32 entry in binary32 arithmetic
    int64_t filter_stats[2] = {0, 0};                    // pairs seen / pairs proven apart by the fp32 filter in the last fp32 batch
    int locality_sort = 2;                               // 0 off, 1 always, 2 automatic (GJKEPA_TUNE_LOCALITY_SORT)
    // resident world-space hull store (gjkepa_store_upload)
    double *d_store_verts = nullptr; int64_t *d_store_off = nullptr; gjk::Sphere *d_store_sphere = nullptr;
    int64_t n_store_hulls = 0; int store_max_verts = 0;
    int host_chunks = 0;           // chunks per host batch (0 = automatic)
    int epa_classes = 3;           // size classes present in the next small batch (bit 0: nA+nB <= 64, bit 1: larger), when the host knows
    double phase_ms[2] = {0.0, 0.0};        // wall clock of the last host batch: shards (copies + kernels), contact rows to the host
    double host_ms[3] = {0.0, 0.0, 0.0};    // profiling: K0/K1/K3 durations summed over the chunks of the last host batch
    bool host_ms_valid = false;
    HostPipe *pipe = nullptr;      // created on first use
    MultiDev *multi = nullptr;     // gjkepa_create_multi: this context only fans out to its children
};

namespace {

constexpr int32_t GJKEPA_MODE_NO_CLEAR_ = 1 << 8;      // internal: the output block was cleared by the caller

int fail(gjkepa_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

// developer aid: GJKEPA_DEBUG_SYNC=1 in the environment synchronises after every kernel and names the one that failed
static bool debug_sync_enabled() { static const bool on = [] { const char *e = getenv("GJKEPA_DEBUG_SYNC"); return e && e[0] == '1'; }(); return on; }
#define DBG_SYNC(name) do { if (debug_sync_enabled()) { cudaError_t e_ = cudaStreamSynchronize(st); if (e_ == cudaSuccess) e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) { fprintf(stderr, "gjkepa: kernel %s failed: %s\n", name, cudaGetErrorString(e_)); return fail(ctx, GJKEPA_ERR_CUDA, name, e_); } } } while (0)

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, GJKEPA_ERR_CUDA, #call, e_); } while (0)

// every entry point runs on its context's device and hands the calling thread's current device back on return
struct DeviceKeep {
    int prev = -1;
    DeviceKeep() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceKeep() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(c) DeviceKeep keep_device_; CU(cudaSetDevice((c)->device))

// grow-on-demand device block (contents are not kept): 12.5 % slack so that slowly growing batches do not reallocate
static int grow_bytes(gjkepa_ctx *ctx, void *&p, size_t &cap, size_t need, const char *what)
{
    if (need <= cap) return GJKEPA_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    const size_t want = need + need / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, what, e);
    cap = want;
    return GJKEPA_OK;
}

template <typename T>
int ensure(gjkepa_ctx *ctx, T *&p, int64_t &cap, int64_t need)
{
    if (need <= cap) return GJKEPA_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    int64_t ncap = need + need / 8 + 1024;
    cudaError_t e = cudaMalloc((void **)&p, (size_t)ncap * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc workspace", e);
    cap = ncap;
    return GJKEPA_OK;
}

void pipe_free(gjkepa_ctx *ctx);        // host_batch.inl
void multi_free(gjkepa_ctx *ctx);

}  // namespace

// a multi-device context forwards a setter to every child
#define FORWARD_MULTI(call_on_k) do { if (ctx->multi) { int rc_ = GJKEPA_OK; for (gjkepa_ctx *k : multi_kids(ctx)) { int q_ = (call_on_k); if (q_ && !rc_) rc_ = q_; } return rc_; } } while (0)
static const std::vector<gjkepa_ctx *> &multi_kids(const gjkepa_ctx *ctx);
static gjkepa_ctx *first_kid(const gjkepa_ctx *ctx);

// seed queue between the two GJK stages (gjk_two_stage.cuh; the binary32 walk of gjk_f32.cuh reuses the same arrays)
static int ensure_seed_queue(gjkepa_ctx *ctx, int64_t n_pairs)
{
    if (n_pairs <= ctx->cap_seed) return GJKEPA_OK;
    if (ctx->d_seed_f64) cudaFree(ctx->d_seed_f64);
    if (ctx->d_seed_pair) cudaFree(ctx->d_seed_pair);
    if (ctx->d_seed_fl) cudaFree(ctx->d_seed_fl);
    if (ctx->d_seed_off) cudaFree(ctx->d_seed_off);
    if (ctx->d_seed_idx) cudaFree(ctx->d_seed_idx);
    ctx->d_seed_f64 = nullptr; ctx->d_seed_pair = nullptr; ctx->d_seed_fl = nullptr; ctx->d_seed_off = nullptr; ctx->d_seed_idx = nullptr; ctx->cap_seed = 0;
    const int64_t ncap = ((n_pairs + n_pairs / 8 + 1024) + 31) & ~(int64_t)31;
    cudaError_t e = cudaMalloc((void **)&ctx->d_seed_f64, (size_t)ncap * gjk::SEED_F64 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_seed_pair, (size_t)ncap * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_seed_fl, (size_t)ncap * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_seed_off, (size_t)ncap * 2 * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_seed_idx, (size_t)ncap * sizeof(uint2));
    if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc seed queue", e);
    ctx->cap_seed = ncap;
    return GJKEPA_OK;
}

extern "C" {

int gjkepa_create(gjkepa_ctx **out, int32_t device)
{
    if (!out) return GJKEPA_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return GJKEPA_ERR_CUDA;
    gjkepa_ctx *ctx = new gjkepa_ctx();
    ctx->device = device;
    DeviceKeep keep_device_;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_counters, N_COUNTERS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&ctx->h_counters, 4 * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) {
        // the kernels are built for sm_100a only: fail loudly when the image cannot load here
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, gjk::gjk_group_kernel<4, false>);
    }
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e != cudaSuccess) {
        fprintf(stderr, "gjkepa_create: %s\n", cudaGetErrorString(e));
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
        if (ctx->d_counters) cudaFree(ctx->d_counters);
        if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
        delete ctx;
        return GJKEPA_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return GJKEPA_OK;
}

int gjkepa_destroy(gjkepa_ctx *ctx)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (ctx->multi) { multi_free(ctx); delete ctx; return GJKEPA_OK; }
    DeviceKeep keep_device_;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pipe_free(ctx);
    if (ctx->d_store_verts) cudaFree(ctx->d_store_verts);
    if (ctx->d_store_off) cudaFree(ctx->d_store_off);
    if (ctx->d_store_sphere) cudaFree(ctx->d_store_sphere);
    if (ctx->d_sphere) cudaFree(ctx->d_sphere);
    if (ctx->d_wl_pair) cudaFree(ctx->d_wl_pair);
    if (ctx->d_wl_simplex) cudaFree(ctx->d_wl_simplex);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_ovf) cudaFree(ctx->d_ovf);
    if (ctx->d_ovf2) cudaFree(ctx->d_ovf2);
    if (ctx->d_route) cudaFree(ctx->d_route);
    if (ctx->d_seed_f64) cudaFree(ctx->d_seed_f64);
    if (ctx->d_seed_pair) cudaFree(ctx->d_seed_pair);
    if (ctx->d_seed_fl) cudaFree(ctx->d_seed_fl);
    if (ctx->d_seed_off) cudaFree(ctx->d_seed_off);
    if (ctx->d_seed_idx) cudaFree(ctx->d_seed_idx);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->d_shape_verts) cudaFree(ctx->d_shape_verts);
    if (ctx->d_shape_off) cudaFree(ctx->d_shape_off);
    if (ctx->d_world) cudaFree(ctx->d_world);
    if (ctx->d_world_off) cudaFree(ctx->d_world_off);
    if (ctx->d_scan_tmp) cudaFree(ctx->d_scan_tmp);
    if (ctx->d_stage2) cudaFree(ctx->d_stage2);
    if (ctx->d_bp) cudaFree(ctx->d_bp);
    if (ctx->d_cand_a) cudaFree(ctx->d_cand_a);
    if (ctx->d_cand_b) cudaFree(ctx->d_cand_b);
    if (ctx->d_scene) cudaFree(ctx->d_scene);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->d_small) cudaFree(ctx->d_small);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    delete ctx;
    return GJKEPA_OK;
}

int gjkepa_set_stream(gjkepa_ctx *ctx, void *s)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (ctx->multi) return fail(ctx, GJKEPA_ERR_ARG, "a multi-device context runs on its own streams");
    ctx->stream = (cudaStream_t)s;          // taken literally: NULL is CUDA's legacy default stream
    return GJKEPA_OK;
}

int gjkepa_use_own_stream(gjkepa_ctx *ctx)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (ctx->multi) return GJKEPA_OK;
    ctx->stream = ctx->own_stream;
    return GJKEPA_OK;
}

const char *gjkepa_last_error(const gjkepa_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int64_t gjkepa_launch_count(const gjkepa_ctx *ctx, int32_t reset)
{
    if (!ctx) return -1;
    if (ctx->multi) { int64_t t = 0; for (gjkepa_ctx *k : multi_kids(ctx)) t += gjkepa_launch_count(k, reset); return t; }
    int64_t n = ctx->launches;
    if (reset) const_cast<gjkepa_ctx *>(ctx)->launches = 0;
    return n;
}

int gjkepa_set_lanes(gjkepa_ctx *ctx, int32_t lanes)
{
    if (!ctx || !(lanes == 0 || lanes == 1 || lanes == 2 || lanes == 4)) return GJKEPA_ERR_ARG;
    FORWARD_MULTI(gjkepa_set_lanes(k, lanes));
    ctx->lanes = lanes;
    return GJKEPA_OK;
}

int gjkepa_set_max_hull_vertices(gjkepa_ctx *ctx, int32_t n)
{
    if (!ctx || n < 0) return GJKEPA_ERR_ARG;
    FORWARD_MULTI(gjkepa_set_max_hull_vertices(k, n));
    ctx->max_hull_hint = n;
    return GJKEPA_OK;
}

int gjkepa_set_tuning(gjkepa_ctx *ctx, int32_t key, int32_t value)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    FORWARD_MULTI(gjkepa_set_tuning(k, key, value));
    if (key == GJKEPA_TUNE_HOST_CHUNKS) {
        if (value < 0 || value > hostb_max_chunks) return GJKEPA_ERR_ARG;
        ctx->host_chunks = value;
        return GJKEPA_OK;
    }
    if (key == GJKEPA_TUNE_WALK_REFILL) {
        if (value < 1 || value > 32) return GJKEPA_ERR_ARG;
        ctx->walk_refill = value;
        return GJKEPA_OK;
    }
    if (key == GJKEPA_TUNE_SINGLE_KERNEL) { ctx->single_kernel = value != 0; return GJKEPA_OK; }
    if (key == GJKEPA_TUNE_EPA_GENERAL) { ctx->epa_general = value != 0; return GJKEPA_OK; }
    if (key == GJKEPA_TUNE_TILE_MAX_VERTICES) {
        if (value < 4 || value > 16) return GJKEPA_ERR_ARG;
        ctx->tile_max_vertices = value;
        return GJKEPA_OK;
    }
    if (key == GJKEPA_TUNE_EARLY_OUT) { ctx->early_out = value != 0; return GJKEPA_OK; }
    if (key == GJKEPA_TUNE_F32_FILTER) { ctx->f32_filter = value != 0; return GJKEPA_OK; }
    if (key == GJKEPA_TUNE_F32_ARITH) { ctx->f32_arith = value != 0; return GJKEPA_OK; }
    if (key == GJKEPA_TUNE_LOCALITY_SORT) {
        if (value < 0 || value > 2) return GJKEPA_ERR_ARG;
        ctx->locality_sort = value;
        return GJKEPA_OK;
    }
    return GJKEPA_ERR_ARG;
}

int gjkepa_set_profiling(gjkepa_ctx *ctx, int32_t on)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    FORWARD_MULTI(gjkepa_set_profiling(k, on));
    ctx->profiling = on != 0;
    return GJKEPA_OK;
}

int gjkepa_kernel_ms(gjkepa_ctx *ctx, double *ms)
{
    if (!ctx || !ms) return GJKEPA_ERR_ARG;
    if (ctx->multi) {               // devices run side by side: the slowest one per kernel
        ms[0] = ms[1] = ms[2] = 0.0;
        for (gjkepa_ctx *k : multi_kids(ctx)) {
            double t[3];
            if (int rc = gjkepa_kernel_ms(k, t)) return rc;
            for (int i = 0; i < 3; ++i) ms[i] = t[i] > ms[i] ? t[i] : ms[i];
        }
        return GJKEPA_OK;
    }
    if (ctx->host_ms_valid) { for (int k = 0; k < 3; ++k) ms[k] = ctx->host_ms[k]; return GJKEPA_OK; }
    ON_DEVICE(ctx);
    for (int k = 0; k < 3; ++k) {
        ms[k] = 0.0;
        if (!ctx->ev_valid[k]) continue;
        float t = 0.f;
        CU(cudaEventSynchronize(ctx->ev[2 * k + 1]));
        CU(cudaEventElapsedTime(&t, ctx->ev[2 * k], ctx->ev[2 * k + 1]));
        ms[k] = (double)t;
    }
    return GJKEPA_OK;
}

int gjkepa_sync(gjkepa_ctx *ctx)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    FORWARD_MULTI(gjkepa_sync(k));
    ON_DEVICE(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    return GJKEPA_OK;
}

int gjkepa_batch_device_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                            int64_t n_hulls, const double *verts, const int64_t *hull_off,
                            int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b,
                            int32_t mode,
                            uint8_t *collision, int32_t *colli_type, double *nearest,
                            double *normal, double *point, double *depth,
                            uint32_t *status, int64_t *n_contacts, int32_t *contact_index)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (ctx->multi) return fail(ctx, GJKEPA_ERR_ARG, "device buffers belong to one GPU: use a single-device context");
    ctx->host_ms_valid = false;
    const bool no_clear = (mode & GJKEPA_MODE_NO_CLEAR_) != 0;
    mode &= ~GJKEPA_MODE_NO_CLEAR_;
    if (n_pairs < 0 || n_hulls < 0 || mode < GJKEPA_MODE_GJK || mode > GJKEPA_MODE_FULL)
        return fail(ctx, GJKEPA_ERR_ARG, "bad sizes or mode");
    if (n_pairs > 0x7fffffffLL || n_hulls > 0x7fffffffLL)      // pair and hull numbers travel as int32 (pair lists, work lists)
        return fail(ctx, GJKEPA_ERR_ARG, "more than 2^31-1 pairs or hulls in one call: split the batch");
    if (n_pairs > 0 && (!verts || !hull_off || !pair_a || !pair_b)) return fail(ctx, GJKEPA_ERR_ARG, "null input");
    ON_DEVICE(ctx);
    cudaStream_t st = ctx->stream;

    // outputs of non-colliding pairs are zero (F:65-71)
    if (n_pairs > 0 && !no_clear) {
        if (colli_type) CU(cudaMemsetAsync(colli_type, 0, (size_t)n_pairs * sizeof(int32_t), st));
        if (nearest)    CU(cudaMemsetAsync(nearest, 0, (size_t)n_pairs * 6 * sizeof(double), st));
        if (normal)     CU(cudaMemsetAsync(normal, 0, (size_t)n_pairs * 3 * sizeof(double), st));
        if (point)      CU(cudaMemsetAsync(point, 0, (size_t)n_pairs * 3 * sizeof(double), st));
        if (depth)      CU(cudaMemsetAsync(depth, 0, (size_t)n_pairs * sizeof(double), st));
    }
    CU(cudaMemsetAsync(ctx->d_counters, 0, N_COUNTERS * sizeof(unsigned long long), st));
    if (n_pairs == 0) {
        if (n_contacts) CU(cudaMemsetAsync(n_contacts, 0, sizeof(int64_t), st));
        return GJKEPA_OK;
    }

    // pair-private hulls (every hull used by ~one pair): the GJK kernels compute the envelope spheres
    // themselves from the vertices they stage anyway; shared stores get one K0 pass per batch
    const gjk::Sphere *ext = ctx->ext_sphere;          // spheres of this store are already on the device
    const bool fuse = !ext && n_hulls >= n_pairs;
    int rc = GJKEPA_OK;
    if (!fuse && !ext) { rc = ensure(ctx, ctx->d_sphere, ctx->cap_sphere, n_hulls); if (rc) return rc; }
    const bool want_list = (mode >= GJKEPA_MODE_EPA) || n_contacts || contact_index;
    if (want_list) {
        if (n_pairs > ctx->cap_wl) {
            if (ctx->d_wl_pair) cudaFree(ctx->d_wl_pair);
            if (ctx->d_wl_simplex) cudaFree(ctx->d_wl_simplex);
            ctx->d_wl_pair = nullptr; ctx->d_wl_simplex = nullptr; ctx->cap_wl = 0;
            int64_t ncap = n_pairs + n_pairs / 8 + 1024;
            cudaError_t e = cudaMalloc((void **)&ctx->d_wl_pair, (size_t)ncap * sizeof(int32_t));
            if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_wl_simplex, (size_t)ncap * 12 * sizeof(double));
            if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc EPA work list", e);
            ctx->cap_wl = ncap;
        }
    }
    // tiny batches (the drop-in single-pair call above all) are launch-latency bound: one free-running 4-lane GJK
    // kernel and the general EPA kernel instead of the staged pipelines (4 launches instead of 9)
    const bool tiny = ctx->lanes == 0 && n_pairs <= 256;
    const bool use_tile = ctx->lanes == 0 && !tiny;
    if (use_tile || mode >= GJKEPA_MODE_EPA) { rc = ensure(ctx, ctx->d_ovf, ctx->cap_ovf, n_pairs); if (rc) return rc; }
    if (use_tile || mode >= GJKEPA_MODE_EPA) { rc = ensure(ctx, ctx->d_ovf2, ctx->cap_ovf2, n_pairs); if (rc) return rc; }
    const bool epa_unrouted = n_pairs <= 8192;       // two routing launches cost more than they save; the size classes filter in-kernel
    if (mode >= GJKEPA_MODE_EPA && !epa_unrouted && !ctx->epa_general) { rc = ensure(ctx, ctx->d_route, ctx->cap_route, 2 * n_pairs + n_pairs / 4 + 4); if (rc) return rc; }

    ctx->ev_valid[0] = ctx->ev_valid[1] = ctx->ev_valid[2] = false;
    const bool prof = ctx->profiling;
    // K0 (not needed under shortcut E1, which drops the sphere test)
    if (!fuse && !ext && !(ctx->early_out && use_tile && !ctx->single_kernel)) {
        const int threads = gjk::PREP_WARPS * 32;
        const int64_t blocks = (n_hulls + gjk::PREP_WARPS * 32 - 1) / (gjk::PREP_WARPS * 32);
        if (prof) CU(cudaEventRecord(ctx->ev[0], st));
        gjk::hull_prep_kernel<<<(unsigned)blocks, threads, 0, st>>>(verts, hull_off, n_hulls, ctx->d_sphere);
        if (prof) { CU(cudaEventRecord(ctx->ev[1], st)); ctx->ev_valid[0] = true; }
        ++ctx->launches;
    }
    // K1
    gjk::BatchArgs a;
    a.verts = verts; a.hull_off = hull_off; a.sphere = ext ? ext : ctx->d_sphere;
    a.pair_a = pair_a; a.pair_b = pair_b; a.n_pairs = n_pairs; a.n_hulls = n_hulls;
    a.collision = collision; a.status = status;
    a.pair_list = nullptr; a.pair_list_count = nullptr;
    a.ovf_list = ctx->d_ovf; a.ovf_count = ctx->d_counters + 2; a.col_doubles = 0;
    a.early_out = ctx->early_out;
    a.order = (use_tile && !ctx->single_kernel) ? ctx->ext_order : nullptr;
    a.order_count = a.order ? ctx->ext_order_count : nullptr;
    a.ovf_status = (ctx->max_hull_hint != 0 && ctx->max_hull_hint <= ctx->tile_max_vertices) ? GJKEPA_ST_BAD_HINT : 0u;
    a.wl_pair = want_list ? ctx->d_wl_pair : nullptr;
    a.wl_simplex = want_list ? ctx->d_wl_simplex : nullptr;
    a.wl_count = ctx->d_counters;
    if (prof) CU(cudaEventRecord(ctx->ev[2], st));
    if (use_tile && !ctx->single_kernel) {
        // K1i + K1w: one thread per pair, init in lockstep, walk with iteration compaction (gjk_two_stage.cuh)
        const int tmax_v = ctx->tile_max_vertices;
        const int cap_vertices = ctx->max_hull_hint > 0 ? (ctx->max_hull_hint < tmax_v ? ctx->max_hull_hint : tmax_v) : tmax_v;
        gjk::TwoStageArgs ta;
        ta.b = a;
        ta.col_units = 2 * 3 * ((cap_vertices + 1) / 2);
        ta.refill = ctx->walk_refill;
        rc = ensure_seed_queue(ctx, n_pairs); if (rc) return rc;
        ta.q.f64 = ctx->d_seed_f64; ta.q.pair = ctx->d_seed_pair; ta.q.fln = ctx->d_seed_fl; ta.q.idx = ctx->d_seed_idx;
        ta.q.oa = ctx->d_seed_off; ta.q.ob = ctx->d_seed_off + ctx->cap_seed;
        ta.q.count = ctx->d_counters + 6; ta.q.head = ctx->d_counters + 7; ta.q.cap = ctx->cap_seed;
        const size_t smem = (size_t)ta.col_units * gjk::T2_PITCH * (gjk::TILE / 32);
        const size_t smem_i = (size_t)ta.col_units * gjk::T2_PITCH * (gjk::INIT_THREADS / 32);
        const int64_t n_ctas = (n_pairs + gjk::TILE - 1) / gjk::TILE;
        const int64_t n_ctas_i = (n_pairs + gjk::INIT_THREADS - 1) / gjk::INIT_THREADS;
        int per_sm = 0;
        if (fuse) {
            CU(cudaFuncSetAttribute(gjk::gjk_init_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_init_kernel<true>, gjk::INIT_THREADS, smem_i));
            const int64_t res = (int64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
            gjk::gjk_init_kernel<true><<<(unsigned)(n_ctas_i < res ? n_ctas_i : res), gjk::INIT_THREADS, smem_i, st>>>(ta);
        } else {
            CU(cudaFuncSetAttribute(gjk::gjk_init_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_init_kernel<false>, gjk::INIT_THREADS, smem_i));
            const int64_t res = (int64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
            gjk::gjk_init_kernel<false><<<(unsigned)(n_ctas_i < res ? n_ctas_i : res), gjk::INIT_THREADS, smem_i, st>>>(ta);
        }
        ++ctx->launches;
        CU(cudaFuncSetAttribute(gjk::gjk_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_walk_kernel, gjk::TILE, smem));
        {
            const int64_t res = (int64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
            DBG_SYNC("gjk_init_kernel");
            gjk::gjk_walk_kernel<<<(unsigned)(n_ctas < res ? n_ctas : res), gjk::TILE, smem, st>>>(ta);
            DBG_SYNC("gjk_walk_kernel");
        }
        ++ctx->launches;
    } else if (use_tile) {
        // K1a: one thread per pair over shared-memory columns; pairs that do not fit overflow to K1b
        const int cap_vertices = ctx->max_hull_hint > 0 ? (ctx->max_hull_hint < 16 ? ctx->max_hull_hint : 16) : 16;
        a.col_doubles = 3 * 2 * cap_vertices;
        const size_t smem = (size_t)a.col_doubles * gjk::TILE_STRIDE * (gjk::TILE / 32) * sizeof(double);
        const int64_t n_tiles = (n_pairs + 31) / 32 / (gjk::TILE / 32) + 1;
        const int64_t resident = (int64_t)ctx->sm_count * ctx->tile_ctas_per_sm;
        int64_t blocks = n_tiles < resident ? n_tiles : resident;
        if (fuse) {
            CU(cudaFuncSetAttribute(gjk::gjk_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_tile_kernel<true>, gjk::TILE, smem));
            blocks = n_tiles < (int64_t)ctx->sm_count * per_sm ? n_tiles : (int64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
            gjk::gjk_tile_kernel<true><<<(unsigned)blocks, gjk::TILE, smem, st>>>(a);
        } else {
            CU(cudaFuncSetAttribute(gjk::gjk_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_tile_kernel<false>, gjk::TILE, smem));
            blocks = n_tiles < (int64_t)ctx->sm_count * per_sm ? n_tiles : (int64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
            gjk::gjk_tile_kernel<false><<<(unsigned)blocks, gjk::TILE, smem, st>>>(a);
        }
        ++ctx->launches;
    }
    if (use_tile) {
        // K1c: hulls too large for a column -> 4-lane sub-warps over TMA-staged slots; K1b: anything larger still
        const int hint = ctx->max_hull_hint;
        if (hint == 0 || hint > ctx->tile_max_vertices) {
            const int cap_v = (hint > 0 && hint < 64) ? hint : 64;
            const int slot_doubles = ((3 * 2 * cap_v + 15) / 16) * 16 + 4;
            const size_t smem_c = (size_t)slot_doubles * 8 /*pairs per warp*/ * 4 /*warps*/ * sizeof(double);
            gjk::BatchArgs b = a;
            b.pair_list = ctx->d_ovf; b.pair_list_count = ctx->d_counters + 2;
            b.ovf_list = ctx->d_ovf2; b.ovf_count = ctx->d_counters + 5;
            b.ovf_status = (hint != 0 && hint <= 64) ? GJKEPA_ST_BAD_HINT : 0u;
            int per_sm = 1;
            if (ctx->single_kernel) {
                if (fuse) {
                    CU(cudaFuncSetAttribute(gjk::gjk_group_staged_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_group_staged_kernel<4, true>, 128, smem_c));
                    gjk::gjk_group_staged_kernel<4, true><<<ctx->sm_count * (per_sm > 0 ? per_sm : 1), 128, smem_c, st>>>(b, slot_doubles);
                } else {
                    CU(cudaFuncSetAttribute(gjk::gjk_group_staged_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_group_staged_kernel<4, false>, 128, smem_c));
                    gjk::gjk_group_staged_kernel<4, false><<<ctx->sm_count * (per_sm > 0 ? per_sm : 1), 128, smem_c, st>>>(b, slot_doubles);
                }
                ++ctx->launches;
            } else {
                // K1ig + K1wg: init by free-running 4-lane groups, walk by 8 groups per warp in lockstep; the seed
                // buffers of the thread-per-pair stage are free again by now (same stream), only the counters differ
                gjk::TwoStageArgs tg;
                tg.b = b; tg.col_units = 0; tg.refill = ctx->walk_refill;
                tg.q.f64 = ctx->d_seed_f64; tg.q.pair = ctx->d_seed_pair; tg.q.fln = ctx->d_seed_fl; tg.q.idx = ctx->d_seed_idx;
                tg.q.oa = ctx->d_seed_off; tg.q.ob = ctx->d_seed_off + ctx->cap_seed;
                tg.q.count = ctx->d_counters + 10; tg.q.head = ctx->d_counters + 11; tg.q.cap = ctx->cap_seed;
                if (fuse) {
                    CU(cudaFuncSetAttribute(gjk::gjk_init_group_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_init_group_kernel<4, true>, 128, smem_c));
                    gjk::gjk_init_group_kernel<4, true><<<ctx->sm_count * (per_sm > 0 ? per_sm : 1), 128, smem_c, st>>>(tg, slot_doubles);
                } else {
                    CU(cudaFuncSetAttribute(gjk::gjk_init_group_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_init_group_kernel<4, false>, 128, smem_c));
                    gjk::gjk_init_group_kernel<4, false><<<ctx->sm_count * (per_sm > 0 ? per_sm : 1), 128, smem_c, st>>>(tg, slot_doubles);
                }
                DBG_SYNC("gjk_init_group_kernel");
                CU(cudaFuncSetAttribute(gjk::gjk_walk_group_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk::gjk_walk_group_kernel<4>, 128, smem_c));
                gjk::gjk_walk_group_kernel<4><<<ctx->sm_count * (per_sm > 0 ? per_sm : 1), 128, smem_c, st>>>(tg, slot_doubles);
                DBG_SYNC("gjk_walk_group_kernel");
                ctx->launches += 2;
            }
            if (hint == 0 || hint > 64) {
                gjk::BatchArgs c = a;
                c.pair_list = ctx->d_ovf2; c.pair_list_count = ctx->d_counters + 5;
                const int blocks_c = ctx->sm_count * 4;
                if (fuse) gjk::gjk_group_kernel<4, true><<<blocks_c, 128, 0, st>>>(c);
                else      gjk::gjk_group_kernel<4, false><<<blocks_c, 128, 0, st>>>(c);
                ++ctx->launches;
            }
        }
    } else {
        const int L = tiny ? 4 : ctx->lanes;
        const int threads = 128;
        const int64_t groups_per_block = threads / L;
        int64_t blocks = (n_pairs + groups_per_block - 1) / groups_per_block;
        const int64_t max_blocks = (int64_t)ctx->sm_count * 64;
        if (blocks > max_blocks) blocks = max_blocks;
        if (fuse) {
            if (L == 1)      gjk::gjk_group_kernel<1, true><<<(unsigned)blocks, threads, 0, st>>>(a);
            else if (L == 2) gjk::gjk_group_kernel<2, true><<<(unsigned)blocks, threads, 0, st>>>(a);
            else             gjk::gjk_group_kernel<4, true><<<(unsigned)blocks, threads, 0, st>>>(a);
        } else {
            if (L == 1)      gjk::gjk_group_kernel<1, false><<<(unsigned)blocks, threads, 0, st>>>(a);
            else if (L == 2) gjk::gjk_group_kernel<2, false><<<(unsigned)blocks, threads, 0, st>>>(a);
            else             gjk::gjk_group_kernel<4, false><<<(unsigned)blocks, threads, 0, st>>>(a);
        }
        ++ctx->launches;
    }
    if (prof) { CU(cudaEventRecord(ctx->ev[3], st)); ctx->ev_valid[1] = true; }
    CU(cudaGetLastError());

    // K3 on the colliding pairs
    if (mode >= GJKEPA_MODE_EPA) {
        epa::EpaArgs e;
        e.verts = verts; e.hull_off = hull_off; e.pair_a = pair_a; e.pair_b = pair_b;
        e.wl_pair = ctx->d_wl_pair; e.wl_simplex = ctx->d_wl_simplex;
        e.wl_count = ctx->d_counters; e.queue_head = ctx->d_counters + 1;
        e.version = version; e.tol_ff = tol_ff; e.full = (mode == GJKEPA_MODE_FULL);
        e.colli_type = colli_type; e.nearest = nearest; e.normal = normal; e.point = point;
        e.depth = depth; e.status = status; e.result = nullptr;
        // contact epilogue as its own kernel (always for version 2, which the in-kernel epilogue does not have)
        const bool k4 = mode == GJKEPA_MODE_FULL && (!ctx->epa_general || version == 2);
        if (k4) { e.full = false; e.result = ctx->d_wl_simplex; }
        int nl = 0;
        if (prof) CU(cudaEventRecord(ctx->ev[4], st));
        // every pair meets the EPA kernel its own hull sizes choose, whatever the batch: tiny batches included
        if (ctx->epa_general) rc = epa::launch(e, ctx->sm_count, st, ctx->d_ovf, ctx->d_counters, &nl);
        else                  rc = epa3::launch(e, ctx->sm_count, st, ctx->d_ovf, ctx->d_ovf2, ctx->d_route, n_pairs, ctx->d_counters,
                                                ctx->max_hull_hint > 0 && ctx->max_hull_hint <= 32, epa_unrouted, ctx->epa_classes, &nl);
        if (rc) return fail(ctx, GJKEPA_ERR_CUDA, "epa launch", (cudaError_t)rc);
        DBG_SYNC("epa kernels");
        if (k4) { contact::launch(e, ctx->sm_count, st); ++nl; }
        DBG_SYNC("contact_kernel");
        if (prof) { CU(cudaEventRecord(ctx->ev[5], st)); ctx->ev_valid[2] = true; }
        ctx->launches += nl;
        CU(cudaGetLastError());
    }
    if (n_contacts)
        CU(cudaMemcpyAsync(n_contacts, ctx->d_counters, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    if (contact_index)   // the list length lives on the device: copy n_pairs slots, entries beyond *n_contacts are unspecified
        CU(cudaMemcpyAsync(contact_index, ctx->d_wl_pair, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return GJKEPA_OK;
}

// Kp: world-space hull store of n_bodies posed shapes (device inputs), left in ctx->d_world; *d_off_out = its offsets
static int posed_world(gjkepa_ctx *ctx, int64_t n_bodies, const int32_t *d_body_shape, const double *d_pose, int64_t **d_off_out)
{
    cudaStream_t st = ctx->stream;
    // world store sized by the largest shape (no device->host round trip for the exact total)
    const int64_t need_world = n_bodies * ctx->max_shape_verts * 3;
    if (need_world > ctx->cap_world) {
        if (ctx->d_world) { cudaFree(ctx->d_world); ctx->d_world = nullptr; ctx->cap_world = 0; }
        cudaError_t e = cudaMalloc((void **)&ctx->d_world, (size_t)need_world * sizeof(double));
        if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc world store", e);
        ctx->cap_world = need_world;
    }
    if (2 * (n_bodies + 1) > ctx->cap_world_off) {
        if (ctx->d_world_off) { cudaFree(ctx->d_world_off); ctx->d_world_off = nullptr; ctx->cap_world_off = 0; }
        cudaError_t e = cudaMalloc((void **)&ctx->d_world_off, (size_t)2 * (n_bodies + 1) * sizeof(int64_t));
        if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc world offsets", e);
        ctx->cap_world_off = 2 * (n_bodies + 1);
    }
    int64_t *d_sizes = ctx->d_world_off, *d_off = ctx->d_world_off + (n_bodies + 1);
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_sizes, d_off, (int)(n_bodies + 1), st));
    if (tmp_bytes > ctx->cap_scan_tmp) {
        if (ctx->d_scan_tmp) { cudaFree(ctx->d_scan_tmp); ctx->d_scan_tmp = nullptr; ctx->cap_scan_tmp = 0; }
        cudaError_t e = cudaMalloc(&ctx->d_scan_tmp, tmp_bytes + 256);
        if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc scan scratch", e);
        ctx->cap_scan_tmp = tmp_bytes + 256;
    }
    const int threads = 256;
    pose::body_sizes_kernel<<<(unsigned)((n_bodies + 1 + threads - 1) / threads), threads, 0, st>>>(
        d_body_shape, ctx->d_shape_off, n_bodies, ctx->n_shapes, d_sizes);
    size_t tb = ctx->cap_scan_tmp;
    CU(cub::DeviceScan::ExclusiveSum(ctx->d_scan_tmp, tb, d_sizes, d_off, (int)(n_bodies + 1), st));
    int64_t blocks = (n_bodies * 8 + threads - 1) / threads;
    const int64_t max_blocks = (int64_t)ctx->sm_count * 32;
    if (blocks > max_blocks) blocks = max_blocks;
    pose::apply_pose_kernel<<<(unsigned)blocks, threads, 0, st>>>(ctx->d_shape_verts, ctx->d_shape_off,
        d_body_shape, d_pose, d_off, n_bodies, ctx->n_shapes, ctx->d_world);
    ctx->launches += 2;
    CU(cudaGetLastError());
    *d_off_out = d_off;
    return GJKEPA_OK;
}

// Kb, first half: spheres -> grid -> sorted cell keys -> per-body counts -> offsets; *h_total = number of passing pairs
// (synchronises).  The workspace (ctx->d_bp) then holds what broadphase_fill needs.
struct BpLayout { size_t o_bounds, o_k0, o_k1, o_i0, o_i1, o_cnt, o_off, o_tmp, total, sort_tmp, scan_tmp; };
static BpLayout bp_layout(int64_t n, size_t sort_tmp, size_t scan_tmp)
{
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    BpLayout L;
    L.sort_tmp = sort_tmp; L.scan_tmp = scan_tmp;
    L.o_bounds = 0; L.o_k0 = al(sizeof(bp::Bounds)); L.o_k1 = L.o_k0 + al((size_t)n * 8); L.o_i0 = L.o_k1 + al((size_t)n * 8);
    L.o_i1 = L.o_i0 + al((size_t)n * 4); L.o_cnt = L.o_i1 + al((size_t)n * 4); L.o_off = L.o_cnt + al((size_t)(n + 1) * 8);
    L.o_tmp = L.o_off + al((size_t)(n + 1) * 8); L.total = L.o_tmp + al(sort_tmp > scan_tmp ? sort_tmp : scan_tmp);
    return L;
}
static int broadphase_count(gjkepa_ctx *ctx, int64_t n, const double *d_verts, const int64_t *d_hull_off, int64_t *h_total, double slack = 1.0)
{
    cudaStream_t st = ctx->stream;
    int rc = ensure(ctx, ctx->d_sphere, ctx->cap_sphere, n);
    if (rc) return rc;
    size_t sort_tmp = 0, scan_tmp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                       (const int32_t *)nullptr, (int32_t *)nullptr, (int)n, 0, 63, st));
    CU(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int64_t *)nullptr, (int64_t *)nullptr, (int)(n + 1), st));
    const BpLayout L = bp_layout(n, sort_tmp, scan_tmp);
    if (L.total > ctx->cap_bp) {
        if (ctx->d_bp) { cudaFree(ctx->d_bp); ctx->d_bp = nullptr; ctx->cap_bp = 0; }
        cudaError_t e = cudaMalloc(&ctx->d_bp, L.total + L.total / 8);
        if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc broad-phase workspace", e);
        ctx->cap_bp = L.total + L.total / 8;
    }
    ctx->bp_layout_n = n; ctx->bp_sort_tmp = sort_tmp; ctx->bp_scan_tmp = scan_tmp;
    char *w = (char *)ctx->d_bp;
    bp::Bounds *d_b = (bp::Bounds *)(w + L.o_bounds);
    unsigned long long *k0 = (unsigned long long *)(w + L.o_k0), *k1 = (unsigned long long *)(w + L.o_k1);
    int32_t *i0 = (int32_t *)(w + L.o_i0), *i1 = (int32_t *)(w + L.o_i1);
    int64_t *d_cnt = (int64_t *)(w + L.o_cnt), *d_offs = (int64_t *)(w + L.o_off);
    const int threads = 256;
    const unsigned blocks = (unsigned)(((n + threads - 1) / threads) < (int64_t)ctx->sm_count * 16 ? ((n + threads - 1) / threads) : (int64_t)ctx->sm_count * 16);
    gjk::hull_prep_kernel<<<(unsigned)((n + gjk::PREP_WARPS * 32 - 1) / (gjk::PREP_WARPS * 32)), gjk::PREP_WARPS * 32, 0, st>>>(d_verts, d_hull_off, n, ctx->d_sphere);
    bp::bounds_init_kernel<<<1, 1, 0, st>>>(d_b, slack);
    bp::bounds_kernel<<<blocks, threads, 0, st>>>(ctx->d_sphere, n, d_b);
    bp::keys_kernel<<<blocks, threads, 0, st>>>(ctx->d_sphere, n, d_b, k0, i0);
    size_t tb = sort_tmp;
    CU(cub::DeviceRadixSort::SortPairs(w + L.o_tmp, tb, k0, k1, i0, i1, (int)n, 0, 63, st));
    const unsigned pblocks = (unsigned)(((n + 127) / 128) < (int64_t)ctx->sm_count * 32 ? ((n + 127) / 128) : (int64_t)ctx->sm_count * 32);
    CU(cudaMemsetAsync(d_cnt + n, 0, sizeof(int64_t), st));
    bp::pairs_kernel<false><<<pblocks, 128, 0, st>>>(ctx->d_sphere, n, d_b, k1, i1, d_cnt, nullptr, 0, nullptr, nullptr);
    tb = scan_tmp;
    CU(cub::DeviceScan::ExclusiveSum(w + L.o_tmp, tb, d_cnt, d_offs, (int)(n + 1), st));
    ctx->launches += 5;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_total, d_offs + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return GJKEPA_OK;
}
// Kb, second half: the passing pairs into device buffers of `cap` entries (asynchronous)
static int broadphase_fill(gjkepa_ctx *ctx, int64_t n, int64_t cap, int32_t *d_pa, int32_t *d_pb)
{
    if (n != ctx->bp_layout_n) return fail(ctx, GJKEPA_ERR_ARG, "broad-phase fill without its counting pass");
    cudaStream_t st = ctx->stream;
    const BpLayout L = bp_layout(n, ctx->bp_sort_tmp, ctx->bp_scan_tmp);
    char *w = (char *)ctx->d_bp;
    const unsigned pblocks = (unsigned)(((n + 127) / 128) < (int64_t)ctx->sm_count * 32 ? ((n + 127) / 128) : (int64_t)ctx->sm_count * 32);
    bp::pairs_kernel<true><<<pblocks, 128, 0, st>>>(ctx->d_sphere, n, (const bp::Bounds *)(w + L.o_bounds),
        (const unsigned long long *)(w + L.o_k1), (const int32_t *)(w + L.o_i1), (int64_t *)(w + L.o_cnt), (const int64_t *)(w + L.o_off), cap, d_pa, d_pb);
    ++ctx->launches;
    CU(cudaGetLastError());
    return GJKEPA_OK;
}
static int broadphase_device(gjkepa_ctx *ctx, int64_t n, const double *d_verts, const int64_t *d_hull_off,
                             int64_t cap, int32_t *d_pa, int32_t *d_pb, int64_t *h_total)
{
    int rc = broadphase_count(ctx, n, d_verts, d_hull_off, h_total);
    if (rc) return rc;
    if (cap > 0 && *h_total > 0) { rc = broadphase_fill(ctx, n, cap, d_pa, d_pb); if (rc) return rc; }
    CU(cudaStreamSynchronize(ctx->stream));
    return GJKEPA_OK;
}

// ---- shapes resident on the device, poses per call ----
int gjkepa_shapes_upload(gjkepa_ctx *ctx, int64_t n_shapes, const double *shape_verts, const int64_t *shape_off)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (n_shapes <= 0 || !shape_verts || !shape_off || shape_off[0] != 0) return fail(ctx, GJKEPA_ERR_ARG, "bad shape store");
    FORWARD_MULTI(gjkepa_shapes_upload(k, n_shapes, shape_verts, shape_off));
    int64_t mx = 0;
    for (int64_t i = 0; i < n_shapes; ++i) {
        const int64_t n = shape_off[i + 1] - shape_off[i];
        if (n <= 0) return fail(ctx, GJKEPA_ERR_ARG, "empty shape");
        if (n > mx) mx = n;
    }
    ON_DEVICE(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_shape_verts) { cudaFree(ctx->d_shape_verts); ctx->d_shape_verts = nullptr; }
    if (ctx->d_shape_off) { cudaFree(ctx->d_shape_off); ctx->d_shape_off = nullptr; }
    ctx->n_shapes = 0;
    const int64_t nv = shape_off[n_shapes];
    cudaError_t e = cudaMalloc((void **)&ctx->d_shape_verts, (size_t)nv * 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_shape_off, (size_t)(n_shapes + 1) * sizeof(int64_t));
    if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc shape store", e);
    CU(cudaMemcpy(ctx->d_shape_verts, shape_verts, (size_t)nv * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_shape_off, shape_off, (size_t)(n_shapes + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    ctx->n_shapes = n_shapes; ctx->max_shape_verts = mx;
    return GJKEPA_OK;
}

}  // extern "C"

#include "host_batch.inl"

static const std::vector<gjkepa_ctx *> &multi_kids(const gjkepa_ctx *ctx) { return ctx->multi->kids; }
static gjkepa_ctx *first_kid(const gjkepa_ctx *ctx) { return ctx->multi->kids[0]; }

extern "C" {

int gjkepa_batch_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                     int64_t n_hulls, const double *verts, const int64_t *hull_off,
                     int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                     uint8_t *collision, int32_t *colli_type, double *nearest,
                     double *normal, double *point, double *depth,
                     uint32_t *status, int64_t *n_contacts, int32_t *contact_index)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    const bool compact = (mode & GJKEPA_MODE_COMPACT) != 0;
    mode &= ~GJKEPA_MODE_COMPACT;
    if (n_pairs < 0 || n_hulls < 0 || mode < GJKEPA_MODE_GJK || mode > GJKEPA_MODE_FULL) return fail(ctx, GJKEPA_ERR_ARG, "bad sizes or mode");
    if (n_pairs > 0x7fffffffLL || n_hulls > 0x7fffffffLL) return fail(ctx, GJKEPA_ERR_ARG, "more than 2^31-1 pairs or hulls in one call: split the batch");
    if (n_pairs > 0 && (!verts || !hull_off || !pair_a || !pair_b)) return fail(ctx, GJKEPA_ERR_ARG, "null input");
    if (n_pairs == 0) { if (n_contacts) *n_contacts = 0; return GJKEPA_OK; }
    if (!ctx->multi && !compact) {
        ON_DEVICE(ctx);
        cudaStream_t st = ctx->stream;
        const int64_t n_vert = hull_off[n_hulls];
        {   // Tiny batches (the drop-in single-pair call): every copy or memset is ~10 us of driver time, so inputs travel
            // as ONE block from a pinned mirror, all outputs come back as ONE block, one memset clears them.
            constexpr size_t SMALL = 96 * 1024;
            auto a16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
            const size_t i_verts = 0, i_off = a16((size_t)n_vert * 24), i_pa = i_off + a16((size_t)(n_hulls + 1) * 8),
                         i_pb = i_pa + a16((size_t)n_pairs * 4), in_bytes = i_pb + a16((size_t)n_pairs * 4);
            const size_t q_col = in_bytes, q_type = q_col + a16((size_t)n_pairs), q_near = q_type + a16((size_t)n_pairs * 4),
                         q_nrm = q_near + a16((size_t)n_pairs * 48), q_pt = q_nrm + a16((size_t)n_pairs * 24),
                         q_dep = q_pt + a16((size_t)n_pairs * 24), q_st = q_dep + a16((size_t)n_pairs * 8),
                         all_bytes = q_st + a16((size_t)n_pairs * 4);
            if (n_pairs <= 256 && all_bytes <= SMALL) {
                if (!ctx->h_small) {
                    cudaError_t e = cudaHostAlloc((void **)&ctx->h_small, SMALL, cudaHostAllocDefault);
                    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_small, SMALL);
                    if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "small staging", e);
                }
                char *h = ctx->h_small, *d = ctx->d_small;
                std::memcpy(h + i_verts, verts, (size_t)n_vert * 24);
                std::memcpy(h + i_off, hull_off, (size_t)(n_hulls + 1) * 8);
                std::memcpy(h + i_pa, pair_a, (size_t)n_pairs * 4);
                std::memcpy(h + i_pb, pair_b, (size_t)n_pairs * 4);
                CU(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st));
                CU(cudaMemsetAsync(d + q_col, 0, all_bytes - q_col, st));           // F:65-71 for every output at once
                int classes = 0;                                                    // the hull sizes are in host memory: skip EPA passes that would find nothing
                for (int64_t i = 0; i < n_pairs; ++i) {
                    const int32_t a_ = pair_a[i], b_ = pair_b[i];
                    if (a_ < 0 || a_ >= n_hulls || b_ < 0 || b_ >= n_hulls) return fail(ctx, GJKEPA_ERR_ARG, "pair list names a hull outside the store");
                    classes |= ((hull_off[a_ + 1] - hull_off[a_]) + (hull_off[b_ + 1] - hull_off[b_]) <= 64) ? 1 : 2;
                }
                struct ClassesBack { gjkepa_ctx *c; ~ClassesBack() { c->epa_classes = 3; } } classes_back{ctx};
                ctx->epa_classes = classes;
                int rc = gjkepa_batch_device_f64(ctx, version, tol_ff, n_hulls, (const double *)(d + i_verts), (const int64_t *)(d + i_off),
                    n_pairs, (const int32_t *)(d + i_pa), (const int32_t *)(d + i_pb), mode | GJKEPA_MODE_NO_CLEAR_,
                    (uint8_t *)(d + q_col), colli_type ? (int32_t *)(d + q_type) : nullptr, nearest ? (double *)(d + q_near) : nullptr,
                    normal ? (double *)(d + q_nrm) : nullptr, point ? (double *)(d + q_pt) : nullptr, depth ? (double *)(d + q_dep) : nullptr,
                    status ? (uint32_t *)(d + q_st) : nullptr, nullptr, nullptr);
                if (rc) return rc;
                CU(cudaMemcpyAsync(h + q_col, d + q_col, all_bytes - q_col, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                const uint8_t *hf = (const uint8_t *)(h + q_col);
                if (collision) std::memcpy(collision, hf, (size_t)n_pairs);
                if (colli_type) std::memcpy(colli_type, h + q_type, (size_t)n_pairs * 4);
                if (nearest) std::memcpy(nearest, h + q_near, (size_t)n_pairs * 48);
                if (normal) std::memcpy(normal, h + q_nrm, (size_t)n_pairs * 24);
                if (point) std::memcpy(point, h + q_pt, (size_t)n_pairs * 24);
                if (depth) std::memcpy(depth, h + q_dep, (size_t)n_pairs * 8);
                if (status) std::memcpy(status, h + q_st, (size_t)n_pairs * 4);
                if (n_contacts || contact_index) {
                    int64_t cnt = 0;
                    for (int64_t i = 0; i < n_pairs; ++i) if (hf[i]) { if (contact_index) contact_index[cnt] = (int32_t)i; ++cnt; }
                    if (n_contacts) *n_contacts = cnt;
                }
                return GJKEPA_OK;
            }
        }
    }
    Req r;
    r.version = version; r.tol_ff = tol_ff; r.mode = mode; r.compact = compact; r.src = SRC_HOST;
    r.n_hulls = n_hulls; r.verts = verts; r.hull_off = hull_off; r.n_pairs = n_pairs; r.pair_a = pair_a; r.pair_b = pair_b;
    r.collision = collision; r.colli_type = colli_type; r.nearest = nearest; r.normal = normal; r.point = point; r.depth = depth;
    r.status = status; r.n_contacts = n_contacts; r.contact_index = contact_index;
    return run_request(ctx, r);
}

// ---- binary32 at the boundary: float hull store in, float results out; the device widens and runs the binary64 kernels ----
int gjkepa_batch_f32(gjkepa_ctx *ctx, int32_t version, float tol_ff,
                     int64_t n_hulls, const float *verts, const int64_t *hull_off,
                     int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                     uint8_t *collision, int32_t *colli_type, float *nearest,
                     float *normal, float *point, float *depth,
                     uint32_t *status, int64_t *n_contacts, int32_t *contact_index)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (n_pairs < 0 || n_hulls < 0 || mode < GJKEPA_MODE_GJK || mode > GJKEPA_MODE_FULL) return fail(ctx, GJKEPA_ERR_ARG, "bad sizes or mode");
    if (n_pairs > 0x7fffffffLL || n_hulls > 0x7fffffffLL) return fail(ctx, GJKEPA_ERR_ARG, "more than 2^31-1 pairs or hulls in one call: split the batch");
    if (n_pairs > 0 && (!verts || !hull_off || !pair_a || !pair_b)) return fail(ctx, GJKEPA_ERR_ARG, "null input");
    Req r;
    r.version = version; r.tol_ff = (double)tol_ff; r.mode = mode; r.src = SRC_HOST_F32; r.f32 = true;
    r.n_hulls = n_hulls; r.verts = verts; r.hull_off = hull_off; r.n_pairs = n_pairs; r.pair_a = pair_a; r.pair_b = pair_b;
    r.collision = collision; r.colli_type = colli_type; r.nearest = nearest; r.normal = normal; r.point = point; r.depth = depth;
    r.status = status; r.n_contacts = n_contacts; r.contact_index = contact_index;
    return run_request(ctx, r);
}

// a body_shape entry outside the shape store would become an empty hull (NaN sphere, poisoned broad-phase bounds): refuse it
static int check_body_shapes(gjkepa_ctx *ctx, int64_t n_bodies, const int32_t *body_shape)
{
    const int64_t ns = ctx->multi ? ctx->multi->kids[0]->n_shapes : ctx->n_shapes;
    if (ns <= 0) return fail(ctx, GJKEPA_ERR_ARG, "no shapes uploaded (gjkepa_shapes_upload)");
    int bad = 0;
    for (int64_t b = 0; b < n_bodies; ++b) bad |= (body_shape[b] < 0) | (body_shape[b] >= ns);
    if (bad) return fail(ctx, GJKEPA_ERR_ARG, "body_shape names a shape outside the uploaded store");
    return GJKEPA_OK;
}

int gjkepa_batch_posed_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                           int64_t n_bodies, const int32_t *body_shape, const double *pose,
                           int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                           uint8_t *collision, int32_t *colli_type, double *nearest,
                           double *normal, double *point, double *depth,
                           uint32_t *status, int64_t *n_contacts, int32_t *contact_index)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    const bool compact = (mode & GJKEPA_MODE_COMPACT) != 0;
    mode &= ~GJKEPA_MODE_COMPACT;
    if (n_pairs < 0 || n_bodies <= 0 || !body_shape || !pose || mode < GJKEPA_MODE_GJK || mode > GJKEPA_MODE_FULL || n_bodies > 0x7fffffffLL || n_pairs > 0x7fffffffLL)
        return fail(ctx, GJKEPA_ERR_ARG, "bad sizes, mode or null input");
    if (n_pairs > 0 && (!pair_a || !pair_b)) return fail(ctx, GJKEPA_ERR_ARG, "null input");
    if (int rc = check_body_shapes(ctx, n_bodies, body_shape)) return rc;
    Req r;
    r.version = version; r.tol_ff = tol_ff; r.mode = mode; r.compact = compact; r.src = SRC_POSED;
    r.n_bodies = n_bodies; r.body_shape = body_shape; r.pose = pose; r.n_pairs = n_pairs; r.pair_a = pair_a; r.pair_b = pair_b;
    r.collision = collision; r.colli_type = colli_type; r.nearest = nearest; r.normal = normal; r.point = point; r.depth = depth;
    r.status = status; r.n_contacts = n_contacts; r.contact_index = contact_index;
    return run_request(ctx, r);
}

// ---- world-space hull store resident on the device(s): pair lists in, results out ----
static int store_upload_one(gjkepa_ctx *ctx, int64_t n_hulls, const double *verts, const int64_t *hull_off, int max_verts)
{
    ON_DEVICE(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_store_verts) { cudaFree(ctx->d_store_verts); ctx->d_store_verts = nullptr; }
    if (ctx->d_store_off) { cudaFree(ctx->d_store_off); ctx->d_store_off = nullptr; }
    if (ctx->d_store_sphere) { cudaFree(ctx->d_store_sphere); ctx->d_store_sphere = nullptr; }
    ctx->n_store_hulls = 0;
    const int64_t v0 = hull_off[0], nv = hull_off[n_hulls] - v0;
    cudaError_t e = cudaMalloc((void **)&ctx->d_store_verts, (size_t)nv * 24 + 256);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_store_off, (size_t)(n_hulls + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_store_sphere, (size_t)n_hulls * sizeof(gjk::Sphere));
    if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc hull store", e);
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->d_store_verts, verts + 3 * v0, (size_t)nv * 24, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_store_off, hull_off, (size_t)(n_hulls + 1) * 8, cudaMemcpyHostToDevice, st));
    if (v0 != 0) {
        hostb::rebase_kernel<<<(unsigned)((n_hulls + 256) / 256 < 4096 ? (n_hulls + 256) / 256 : 4096), 256, 0, st>>>(nullptr, nullptr, 0, 0, 0, ctx->d_store_off, n_hulls + 1, v0, nullptr);
        ++ctx->launches;
    }
    if (int rc = spheres_for(ctx, ctx->d_store_verts, ctx->d_store_off, n_hulls, ctx->d_store_sphere, st)) return rc;
    CU(cudaStreamSynchronize(st));
    ctx->n_store_hulls = n_hulls; ctx->store_max_verts = max_verts;
    return GJKEPA_OK;
}

int gjkepa_store_upload(gjkepa_ctx *ctx, int64_t n_hulls, const double *verts, const int64_t *hull_off)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (n_hulls <= 0 || n_hulls > 0x7fffffffLL || !verts || !hull_off) return fail(ctx, GJKEPA_ERR_ARG, "bad hull store");
    int64_t mx = 0;
    for (int64_t h = 0; h < n_hulls; ++h) {
        const int64_t n = hull_off[h + 1] - hull_off[h];
        if (n <= 0) return fail(ctx, GJKEPA_ERR_ARG, "empty hull in the store");
        mx = n > mx ? n : mx;
    }
    if (mx > 0x7fffffff) return fail(ctx, GJKEPA_ERR_ARG, "hull too large");
    const int mxi = (int)mx;
    if (ctx->multi)      // replicated: every device copies the store over its own PCIe link, in parallel
        return ctx->multi->all(ctx, [=](int, gjkepa_ctx *k) { return store_upload_one(k, n_hulls, verts, hull_off, mxi); });
    return store_upload_one(ctx, n_hulls, verts, hull_off, mxi);
}

int gjkepa_batch_resident_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                              int64_t n_pairs, const int32_t *pair_a, const int32_t *pair_b, int32_t mode,
                              uint8_t *collision, int32_t *colli_type, double *nearest,
                              double *normal, double *point, double *depth,
                              uint32_t *status, int64_t *n_contacts, int32_t *contact_index)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    const bool compact = (mode & GJKEPA_MODE_COMPACT) != 0;
    mode &= ~GJKEPA_MODE_COMPACT;
    if (n_pairs < 0 || n_pairs > 0x7fffffffLL || mode < GJKEPA_MODE_GJK || mode > GJKEPA_MODE_FULL) return fail(ctx, GJKEPA_ERR_ARG, "bad sizes or mode");
    if (n_pairs > 0 && (!pair_a || !pair_b)) return fail(ctx, GJKEPA_ERR_ARG, "null input");
    const gjkepa_ctx *c0 = ctx->multi ? ctx->multi->kids[0] : ctx;
    if (c0->n_store_hulls <= 0) return fail(ctx, GJKEPA_ERR_ARG, "no hull store uploaded (gjkepa_store_upload)");
    Req r;
    r.version = version; r.tol_ff = tol_ff; r.mode = mode; r.compact = compact; r.src = SRC_RESIDENT;
    r.n_pairs = n_pairs; r.pair_a = pair_a; r.pair_b = pair_b;
    r.collision = collision; r.colli_type = colli_type; r.nearest = nearest; r.normal = normal; r.point = point; r.depth = depth;
    r.status = status; r.n_contacts = n_contacts; r.contact_index = contact_index;
    return run_request(ctx, r);
}

// ---- one context over several devices ----
int gjkepa_create_multi(gjkepa_ctx **out, const int32_t *devices, int32_t n_devices)
{
    if (!out) return GJKEPA_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return GJKEPA_ERR_CUDA;
    if (n_devices < 0 || n_devices > 64) return GJKEPA_ERR_ARG;
    std::vector<int32_t> devs;
    if (n_devices == 0 || !devices) { for (int i = 0; i < ndev; ++i) devs.push_back(i); }      // every visible device
    else devs.assign(devices, devices + n_devices);
    gjkepa_ctx *ctx = new gjkepa_ctx();
    ctx->multi = new MultiDev();
    ctx->device = devs[0];
    DeviceKeep keep_device_;
    for (int32_t d : devs) {
        gjkepa_ctx *k = nullptr;
        const int rc = gjkepa_create(&k, d);
        if (rc) { multi_free(ctx); delete ctx; return rc; }
        ctx->multi->kids.push_back(k);
    }
    ctx->sm_count = ctx->multi->kids[0]->sm_count;
    for (size_t i = 0; i < devs.size(); ++i) {
        MultiDev::Worker *w = new MultiDev::Worker();
        ctx->multi->workers.push_back(w);
        w->th = std::thread(MultiDev::loop, w, (int)devs[i]);
    }
    *out = ctx;
    return GJKEPA_OK;
}

int gjkepa_host_phase_ms(const gjkepa_ctx *ctx, double *ms)
{
    if (!ctx || !ms) return GJKEPA_ERR_ARG;
    ms[0] = ctx->phase_ms[0]; ms[1] = ctx->phase_ms[1];
    return GJKEPA_OK;
}

int gjkepa_f32_filter_stats(const gjkepa_ctx *ctx, int64_t *out)
{
    if (!ctx || !out) return GJKEPA_ERR_ARG;
    out[0] = out[1] = 0;
    if (ctx->multi) { for (gjkepa_ctx *k : multi_kids(ctx)) { out[0] += k->filter_stats[0]; out[1] += k->filter_stats[1]; } }
    else { out[0] = ctx->filter_stats[0]; out[1] = ctx->filter_stats[1]; }
    return GJKEPA_OK;
}

int32_t gjkepa_device_count(const gjkepa_ctx *ctx)
{
    if (!ctx) return 0;
    return ctx->multi ? (int32_t)ctx->multi->kids.size() : 1;
}

// pin a caller's array once so that the copies of every later batch run at full PCIe speed and overlap the kernels
// (arrays a Fortran host ALLOCATEs are pageable: the driver then stages every copy through its own small pinned buffer)
int gjkepa_host_register(void *p, int64_t bytes)
{
    if (!p || bytes <= 0) return GJKEPA_ERR_ARG;
    return cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable) == cudaSuccess ? GJKEPA_OK : GJKEPA_ERR_CUDA;
}
int gjkepa_host_unregister(void *p)
{
    if (!p) return GJKEPA_ERR_ARG;
    return cudaHostUnregister(p) == cudaSuccess ? GJKEPA_OK : GJKEPA_ERR_CUDA;
}

// ---- candidate pairs on the device ----
static int broadphase_out(gjkepa_ctx *ctx, int64_t n, const double *d_verts, const int64_t *d_off,
                          int64_t cap, int32_t *pair_a, int32_t *pair_b, int64_t *n_pairs)
{
    cudaStream_t st = ctx->stream;
    int32_t *d_pa = nullptr, *d_pb = nullptr;
    if (cap > 0) {
        int rc = ensure(ctx, ctx->d_ovf, ctx->cap_ovf, cap); if (rc) return rc;       // borrowed: free between batches
        rc = ensure(ctx, ctx->d_ovf2, ctx->cap_ovf2, cap); if (rc) return rc;
        d_pa = ctx->d_ovf; d_pb = ctx->d_ovf2;
    }
    int64_t total = 0;
    int rc = broadphase_device(ctx, n, d_verts, d_off, cap, d_pa, d_pb, &total);
    if (rc) return rc;
    const int64_t m = total < cap ? total : cap;
    if (m > 0) {
        CU(cudaMemcpyAsync(pair_a, d_pa, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(pair_b, d_pb, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    *n_pairs = total;
    return GJKEPA_OK;
}

int gjkepa_broadphase_f64(gjkepa_ctx *ctx, int64_t n_hulls, const double *verts, const int64_t *hull_off,
                          int64_t cap, int32_t *pair_a, int32_t *pair_b, int64_t *n_pairs)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (ctx->multi) {               // the broad phase of one scene runs on the context's first device
        const int rc = gjkepa_broadphase_f64(first_kid(ctx), n_hulls, verts, hull_off, cap, pair_a, pair_b, n_pairs);
        if (rc) ctx->err = first_kid(ctx)->err;
        return rc;
    }
    if (n_hulls <= 0 || !verts || !hull_off || cap < 0 || !n_pairs || (cap > 0 && (!pair_a || !pair_b)) || n_hulls > 0x7fffffff)
        return fail(ctx, GJKEPA_ERR_ARG, "bad sizes or null input");
    ON_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const int64_t n_vert = hull_off[n_hulls];
    const size_t o_off = al((size_t)n_vert * 24), total = o_off + al((size_t)(n_hulls + 1) * 8);
    if (int rcg = grow_bytes(ctx, ctx->d_stage, ctx->cap_stage, total, "cudaMalloc staging")) return rcg;
    char *base = (char *)ctx->d_stage;
    CU(cudaMemcpyAsync(base, verts, (size_t)n_vert * 24, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_off, hull_off, (size_t)(n_hulls + 1) * 8, cudaMemcpyHostToDevice, st));
    return broadphase_out(ctx, n_hulls, (const double *)base, (const int64_t *)(base + o_off), cap, pair_a, pair_b, n_pairs);
}

int gjkepa_broadphase_posed_f64(gjkepa_ctx *ctx, int64_t n_bodies, const int32_t *body_shape, const double *pose,
                                int64_t cap, int32_t *pair_a, int32_t *pair_b, int64_t *n_pairs)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (ctx->multi) {
        const int rc = gjkepa_broadphase_posed_f64(first_kid(ctx), n_bodies, body_shape, pose, cap, pair_a, pair_b, n_pairs);
        if (rc) ctx->err = first_kid(ctx)->err;
        return rc;
    }
    if (ctx->n_shapes <= 0) return fail(ctx, GJKEPA_ERR_ARG, "no shapes uploaded (gjkepa_shapes_upload)");
    if (n_bodies <= 0 || !body_shape || !pose || cap < 0 || !n_pairs || (cap > 0 && (!pair_a || !pair_b)) || n_bodies > 0x7fffffff)
        return fail(ctx, GJKEPA_ERR_ARG, "bad sizes or null input");
    if (int rcs = check_body_shapes(ctx, n_bodies, body_shape)) return rcs;
    ON_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_pose = al((size_t)n_bodies * sizeof(int32_t)), total = o_pose + al((size_t)n_bodies * 12 * sizeof(double));
    if (int rcg = grow_bytes(ctx, ctx->d_stage2, ctx->cap_stage2, total, "cudaMalloc staging")) return rcg;
    char *base = (char *)ctx->d_stage2;
    CU(cudaMemcpyAsync(base, body_shape, (size_t)n_bodies * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_pose, pose, (size_t)n_bodies * 12 * sizeof(double), cudaMemcpyHostToDevice, st));
    int64_t *d_off = nullptr;
    int rc = posed_world(ctx, n_bodies, (const int32_t *)base, (const double *)(base + o_pose), &d_off);
    if (rc) return rc;
    return broadphase_out(ctx, n_bodies, ctx->d_world, d_off, cap, pair_a, pair_b, n_pairs);
}

// ---- scene step: poses in, contacts out ----
int gjkepa_scene_posed_f64(gjkepa_ctx *ctx, int32_t version, double tol_ff,
                           int64_t n_bodies, const int32_t *body_shape, const double *pose, int32_t mode,
                           int64_t cap, int32_t *body_a, int32_t *body_b, double *depth, double *normal,
                           double *point, double *nearest, int32_t *colli_type,
                           int64_t *n_candidates, int64_t *n_contacts)
{
    if (!ctx) return GJKEPA_ERR_ARG;
    if (ctx->multi) {               // one scene, one device: the step is a chain of dependent passes over all bodies
        const int rc = gjkepa_scene_posed_f64(first_kid(ctx), version, tol_ff, n_bodies, body_shape, pose, mode, cap, body_a, body_b, depth, normal,
                                              point, nearest, colli_type, n_candidates, n_contacts);
        if (rc) ctx->err = first_kid(ctx)->err;
        return rc;
    }
    if (ctx->n_shapes <= 0) return fail(ctx, GJKEPA_ERR_ARG, "no shapes uploaded (gjkepa_shapes_upload)");
    if (n_bodies <= 0 || !body_shape || !pose || cap < 0 || !n_contacts || (cap > 0 && (!body_a || !body_b)) || n_bodies > 0x7fffffff
        || mode < GJKEPA_MODE_GJK || mode > GJKEPA_MODE_FULL)
        return fail(ctx, GJKEPA_ERR_ARG, "bad sizes, mode or null input");
    if (int rcs = check_body_shapes(ctx, n_bodies, body_shape)) return rcs;
    ON_DEVICE(ctx);
    cudaStream_t st = ctx->stream;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // 1. poses to the device, world-space hull store (Kp)
    const size_t o_pose = al((size_t)n_bodies * sizeof(int32_t)), in_total = o_pose + al((size_t)n_bodies * 12 * sizeof(double));
    if (int rcg = grow_bytes(ctx, ctx->d_stage2, ctx->cap_stage2, in_total, "cudaMalloc staging")) return rcg;
    char *sb = (char *)ctx->d_stage2;
    CU(cudaMemcpyAsync(sb, body_shape, (size_t)n_bodies * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sb + o_pose, pose, (size_t)n_bodies * 12 * sizeof(double), cudaMemcpyHostToDevice, st));
    int64_t *d_off = nullptr;
    int rc = posed_world(ctx, n_bodies, (const int32_t *)sb, (const double *)(sb + o_pose), &d_off);
    if (rc) return rc;
    // 2. candidate pairs (Kb): count, size the buffers, fill
    int64_t total = 0;
    // Under shortcut E1 (flags of separated pairs are FALSE whichever way they are found) the candidates are the pairs whose
    // envelope spheres actually touch (absolute term 1e-6 instead of the reference's 1.0): the contact list is the same, the
    // candidate list ~4x shorter
    rc = broadphase_count(ctx, n_bodies, ctx->d_world, d_off, &total, ctx->early_out ? gjk::EARLY_OUT_MARGIN : 1.0);
    if (rc) return rc;
    if (n_candidates) *n_candidates = total;
    *n_contacts = 0;
    if (total == 0) return GJKEPA_OK;
    if (total > 0x7fffffff) return fail(ctx, GJKEPA_ERR_ARG, "more than 2^31 candidate pairs: split the scene");
    if (total > ctx->cap_cand) {
        if (ctx->d_cand_a) cudaFree(ctx->d_cand_a);
        if (ctx->d_cand_b) cudaFree(ctx->d_cand_b);
        ctx->d_cand_a = ctx->d_cand_b = nullptr; ctx->cap_cand = 0;
        const int64_t ncap = total + total / 8 + 1024;
        cudaError_t e = cudaMalloc((void **)&ctx->d_cand_a, (size_t)ncap * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_cand_b, (size_t)ncap * sizeof(int32_t));
        if (e != cudaSuccess) return fail(ctx, GJKEPA_ERR_NOMEM, "cudaMalloc candidate pairs", e);
        ctx->cap_cand = ncap;
    }
    rc = broadphase_fill(ctx, n_bodies, total, ctx->d_cand_a, ctx->d_cand_b);
    if (rc) return rc;
    // 3. the reference path over the candidates, results at the candidate's index in device scratch
    const bool want_d = depth && mode >= GJKEPA_MODE_EPA, want_n = normal && mode >= GJKEPA_MODE_EPA;
    const bool want_p = point && mode == GJKEPA_MODE_FULL, want_q = nearest && mode == GJKEPA_MODE_FULL, want_t = colli_type && mode == GJKEPA_MODE_FULL;
    size_t sort_tmp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                       (const int32_t *)nullptr, (int32_t *)nullptr, (int)total, 0, 64, st));
    const size_t o_cnt = 0, o_idx = al(sizeof(int64_t)), o_ord = o_idx + al((size_t)total * 4), o_k0 = o_ord + al((size_t)total * 4),
                 o_k1 = o_k0 + al((size_t)total * 8), o_tmp = o_k1 + al((size_t)total * 8),
                 o_dep = o_tmp + al(sort_tmp), o_nrm = o_dep + al(want_d ? (size_t)total * 8 : 0), o_pt = o_nrm + al(want_n ? (size_t)total * 24 : 0),
                 o_nr = o_pt + al(want_p ? (size_t)total * 24 : 0), o_ty = o_nr + al(want_q ? (size_t)total * 48 : 0),
                 sc_total = o_ty + al(want_t ? (size_t)total * 4 : 0);
    if (int rcg = grow_bytes(ctx, ctx->d_scene, ctx->cap_scene, sc_total, "cudaMalloc scene scratch")) return rcg;
    char *w = (char *)ctx->d_scene;
    const int save_hint = ctx->max_hull_hint;
    ctx->max_hull_hint = (int)ctx->max_shape_verts;      // the shape store bounds the hull sizes
    rc = gjkepa_batch_device_f64(ctx, version, tol_ff, n_bodies, ctx->d_world, d_off, total, ctx->d_cand_a, ctx->d_cand_b, mode,
        nullptr, want_t ? (int32_t *)(w + o_ty) : nullptr, want_q ? (double *)(w + o_nr) : nullptr, want_n ? (double *)(w + o_nrm) : nullptr,
        want_p ? (double *)(w + o_pt) : nullptr, want_d ? (double *)(w + o_dep) : nullptr, nullptr,
        (int64_t *)(w + o_cnt), (int32_t *)(w + o_idx));
    ctx->max_hull_hint = save_hint;
    if (rc) return rc;
    int64_t nc = 0;
    CU(cudaMemcpyAsync(&nc, w + o_cnt, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_contacts = nc;
    const int64_t m = nc < cap ? nc : cap;
    if (nc == 0 || m == 0) return GJKEPA_OK;
    // 4. contacts in lexicographic (body a, body b) order, gathered densely (Ks)
    {
        const int threads = 256;
        int64_t blocks = (nc + threads - 1) / threads;
        const int64_t mx = (int64_t)ctx->sm_count * 32;
        scene::contact_keys_kernel<<<(unsigned)(blocks < mx ? blocks : mx), threads, 0, st>>>((const int32_t *)(w + o_idx), nc, ctx->d_cand_a,
                                                                                            ctx->d_cand_b, (unsigned long long *)(w + o_k0));
        ++ctx->launches;
    }
    size_t tb = sort_tmp;
    CU(cub::DeviceRadixSort::SortPairs(w + o_tmp, tb, (const unsigned long long *)(w + o_k0), (unsigned long long *)(w + o_k1),
                                       (const int32_t *)(w + o_idx), (int32_t *)(w + o_ord), (int)nc, 0, 64, st));
    // dense outputs: the staging block is free again (poses are consumed)
    const size_t c_a = 0, c_b = c_a + al((size_t)nc * 4), c_d = c_b + al((size_t)nc * 4), c_n = c_d + al(want_d ? (size_t)nc * 8 : 0),
                 c_p = c_n + al(want_n ? (size_t)nc * 24 : 0), c_q = c_p + al(want_p ? (size_t)nc * 24 : 0), c_t = c_q + al(want_q ? (size_t)nc * 48 : 0),
                 c_total = c_t + al(want_t ? (size_t)nc * 4 : 0);
    if (int rcg = grow_bytes(ctx, ctx->d_stage, ctx->cap_stage, c_total, "cudaMalloc staging")) return rcg;
    char *c = (char *)ctx->d_stage;
    scene::GatherArgs g;
    g.order = (const int32_t *)(w + o_ord); g.n = nc; g.cand_a = ctx->d_cand_a; g.cand_b = ctx->d_cand_b;
    g.depth = want_d ? (const double *)(w + o_dep) : nullptr; g.normal = want_n ? (const double *)(w + o_nrm) : nullptr;
    g.point = want_p ? (const double *)(w + o_pt) : nullptr; g.nearest = want_q ? (const double *)(w + o_nr) : nullptr;
    g.colli_type = want_t ? (const int32_t *)(w + o_ty) : nullptr;
    g.out_a = (int32_t *)(c + c_a); g.out_b = (int32_t *)(c + c_b); g.o_depth = (double *)(c + c_d); g.o_normal = (double *)(c + c_n);
    g.o_point = (double *)(c + c_p); g.o_nearest = (double *)(c + c_q); g.o_type = (int32_t *)(c + c_t);
    {
        const int threads = 256;
        int64_t blocks = (nc + threads - 1) / threads;
        const int64_t mx = (int64_t)ctx->sm_count * 32;
        scene::gather_kernel<<<(unsigned)(blocks < mx ? blocks : mx), threads, 0, st>>>(g);
        ++ctx->launches;
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(body_a, c + c_a, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(body_b, c + c_b, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (want_d) CU(cudaMemcpyAsync(depth, c + c_d, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    if (want_n) CU(cudaMemcpyAsync(normal, c + c_n, (size_t)m * 24, cudaMemcpyDeviceToHost, st));
    if (want_p) CU(cudaMemcpyAsync(point, c + c_p, (size_t)m * 24, cudaMemcpyDeviceToHost, st));
    if (want_q) CU(cudaMemcpyAsync(nearest, c + c_q, (size_t)m * 48, cudaMemcpyDeviceToHost, st));
    if (want_t) CU(cudaMemcpyAsync(colli_type, c + c_t, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return GJKEPA_OK;
}

// ---- self-test hook: the straight-line sqrt / division sequences of fp64_exact.cuh against the
// compiler's IEEE operators, element by element (used by tests/test_gpu_fp64_exact.py) ----
__global__ void fp64_exact_probe(int64_t n, const double *a, const double *b, double *q_fast, double *q_ref,
                                 double *s_fast, double *s_ref, uint8_t *in_range)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool okd = true, oks = true;
    double r = gjk::rcp_div_seed(b[i]);
    q_fast[i] = gjk::div_with_seed(a[i], b[i], r, okd);
    q_ref[i] = a[i] / b[i];
    s_fast[i] = gjk::sqrt_fast(b[i], oks);
    s_ref[i] = sqrt(b[i]);
    in_range[i] = (okd ? 1 : 0) | (oks ? 2 : 0);
}

int gjkepa_debug_fp64(gjkepa_ctx *ctx, int64_t n, const double *a, const double *b, double *q_fast,
                      double *q_ref, double *s_fast, double *s_ref, uint8_t *in_range)
{
    if (!ctx || n <= 0) return GJKEPA_ERR_ARG;
    if (ctx->multi) return gjkepa_debug_fp64(first_kid(ctx), n, a, b, q_fast, q_ref, s_fast, s_ref, in_range);
    ON_DEVICE(ctx);
    double *d = nullptr; uint8_t *f = nullptr;
    CU(cudaMalloc((void **)&d, (size_t)n * 6 * sizeof(double)));
    CU(cudaMalloc((void **)&f, (size_t)n));
    CU(cudaMemcpy(d, a, (size_t)n * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + n, b, (size_t)n * 8, cudaMemcpyHostToDevice));
    fp64_exact_probe<<<(unsigned)((n + 255) / 256), 256>>>(n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, d + 5 * n, f);
    ++ctx->launches;
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(q_fast, d + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(q_ref, d + 3 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(s_fast, d + 4 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(s_ref, d + 5 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(in_range, f, (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(f);
    return GJKEPA_OK;
}

// ---- drop-in single pair: same dummy-argument list as GJKEPA (F:39-52), everything by reference ----
// The reference is called from inside an !$OMP PARALLEL DO over pairs (THREADPRIVATE scratch, F:55-60): many host threads,
// one pair each, all at once.  A GPU round trip costs tens of microseconds whatever it carries, so concurrent callers are
// COMBINED: a caller that finds no batch in flight becomes the leader, takes every request queued so far (its own included),
// runs them as ONE small batch on the process-wide context and hands each caller its results; callers that arrive while
// that batch is on the GPU queue up for the next one, which the next leader (the first of them to notice) runs.
// Throughput is then (calling threads) / (one round trip) instead of 1 / (one round trip).
} // extern "C"
namespace {
struct PairJob {
    int32_t version; double tol; const double *p1; int32_t n1; const double *p2; int32_t n2;
    uint8_t *collision; int32_t *colli_type; double *nearest, *normal, *point, *depth;
    int rc = GJKEPA_OK;
    std::atomic<int> done{0};
};
std::mutex g_pair_mu;                          // guards the queue and the leader flag
std::vector<PairJob *> g_pair_q;
std::atomic<bool> g_pair_leader{false};       // written under the mutex; waiting callers peek at it without
gjkepa_ctx *g_default_ctx = nullptr;           // created by the first leader, used by one leader at a time

// one batch: jobs that share version_ and TOL_FF_ go through one gjkepa_batch_f64 call (tiny-batch path: one block in, one block out)
void run_pair_jobs(std::vector<PairJob *> &jobs)
{
    static std::vector<double> verts, nr, nm, pt, dp;      // leader-only scratch
    static std::vector<int64_t> off;
    static std::vector<int32_t> pa, pb, ty;
    static std::vector<uint8_t> col;
    static std::vector<PairJob *> grp;
    if (!g_default_ctx) {
        const int rc = gjkepa_create(&g_default_ctx, 0);
        if (rc) { for (PairJob *j : jobs) { j->rc = rc; j->done.store(1, std::memory_order_release); } return; }
    }
    size_t next = 0;
    std::vector<char> taken(jobs.size(), 0);
    while (next < jobs.size()) {
        while (next < jobs.size() && taken[next]) ++next;
        if (next >= jobs.size()) break;
        const int32_t ver = jobs[next]->version; const double tol = jobs[next]->tol;
        grp.clear(); verts.clear(); off.assign(1, 0); pa.clear(); pb.clear();
        size_t bytes = 0;
        for (size_t i = next; i < jobs.size(); ++i) {
            PairJob *j = jobs[i];
            if (taken[i] || j->version != ver || !(j->tol == tol)) continue;
            const size_t add = (size_t)24 * (j->n1 + j->n2) + 160;
            if (!grp.empty() && (bytes + add > 80 * 1024 || grp.size() >= 256)) break;      // stays on the tiny-batch path
            bytes += add; taken[i] = 1; grp.push_back(j);
            verts.insert(verts.end(), j->p1, j->p1 + 3 * (size_t)j->n1);
            off.push_back(off.back() + j->n1);
            verts.insert(verts.end(), j->p2, j->p2 + 3 * (size_t)j->n2);
            off.push_back(off.back() + j->n2);
            pa.push_back((int32_t)(2 * (grp.size() - 1))); pb.push_back((int32_t)(2 * (grp.size() - 1) + 1));
        }
        const size_t m = grp.size();
        col.assign(m, 0); ty.assign(m, 0); nr.assign(6 * m, 0.0); nm.assign(3 * m, 0.0); pt.assign(3 * m, 0.0); dp.assign(m, 0.0);
        const int rc = gjkepa_batch_f64(g_default_ctx, ver, tol, (int64_t)(2 * m), verts.data(), off.data(), (int64_t)m, pa.data(), pb.data(),
                                        GJKEPA_MODE_FULL, col.data(), ty.data(), nr.data(), nm.data(), pt.data(), dp.data(), nullptr, nullptr, nullptr);
        for (size_t k = 0; k < m; ++k) {
            PairJob *j = grp[k];
            j->rc = rc;
            if (!rc) {
                if (j->collision) *j->collision = col[k];          /* LOGICAL*1: write 1, read "!= 0" */
                if (j->colli_type) *j->colli_type = ty[k];
                if (j->nearest) std::memcpy(j->nearest, &nr[6 * k], 48);
                if (j->normal) std::memcpy(j->normal, &nm[3 * k], 24);
                if (j->point) std::memcpy(j->point, &pt[3 * k], 24);
                if (j->depth) *j->depth = dp[k];
            }
            j->done.store(1, std::memory_order_release);           // j may leave its caller's stack frame from here on
        }
    }
}
}  // namespace
extern "C" {

int gjkepa_guard_check(int64_t *n_allocations, int64_t *n_bands_changed)
{
    if (n_allocations) *n_allocations = 0;
    if (n_bands_changed) *n_bands_changed = 0;
    if (!guard::enabled()) return GJKEPA_ERR_ARG;          // not a checking run: nothing to report
    int cur = 0;
    cudaGetDevice(&cur);
    std::vector<guard::Rec> live;
    { std::lock_guard<std::mutex> lk(guard::g_mu); live = guard::g_live; }
    long long bad = guard::g_bad_at_free.load();
    for (const guard::Rec &r : live) {
        cudaSetDevice(r.dev);
        cudaDeviceSynchronize();
        bad += guard::bands_changed(r);
    }
    cudaSetDevice(cur);
    if (n_allocations) *n_allocations = (int64_t)live.size();
    if (n_bands_changed) *n_bands_changed = bad;
    return GJKEPA_OK;
}

int gjkepa_guard_selftest(int64_t overrun_bytes)
{
    // plants a write of `overrun_bytes` past the end of a 1000-byte allocation made the way the library makes all of its own:
    // a following gjkepa_guard_check must count it (the allocation stays live until the process ends)
    if (!guard::enabled()) return GJKEPA_ERR_ARG;
    char *p = nullptr;
    if (cudaMalloc((void **)&p, 1000) != cudaSuccess) return GJKEPA_ERR_NOMEM;
    if (overrun_bytes > 0 && cudaMemset(p + 1000, 0, (size_t)(overrun_bytes < 256 ? overrun_bytes : 256)) != cudaSuccess) return GJKEPA_ERR_CUDA;
    if (overrun_bytes < -256) overrun_bytes = -256;
    if (overrun_bytes < 0 && cudaMemset(p + overrun_bytes, 0, (size_t)(-overrun_bytes)) != cudaSuccess) return GJKEPA_ERR_CUDA;
    return cudaDeviceSynchronize() == cudaSuccess ? GJKEPA_OK : GJKEPA_ERR_CUDA;
}

int gjkepa_pair_f64(const int32_t *version, const double *tol_ff,
                    const double *p1, const int32_t *n1, const double *p2, const int32_t *n2,
                    uint8_t *collision, int32_t *colli_type,
                    double *nearest, double *normal, double *point, double *depth)
{
    if (!version || !tol_ff || !p1 || !n1 || !p2 || !n2 || *n1 <= 0 || *n2 <= 0) return GJKEPA_ERR_ARG;
    PairJob job;
    job.version = *version; job.tol = *tol_ff; job.p1 = p1; job.n1 = *n1; job.p2 = p2; job.n2 = *n2;
    job.collision = collision; job.colli_type = colli_type; job.nearest = nearest; job.normal = normal; job.point = point; job.depth = depth;
    bool lead = false;
    {
        std::lock_guard<std::mutex> lk(g_pair_mu);
        g_pair_q.push_back(&job);
        if (!g_pair_leader.load(std::memory_order_relaxed)) { g_pair_leader.store(true, std::memory_order_relaxed); lead = true; }
    }
    std::vector<PairJob *> mine;
    unsigned spins = 0;
    for (;;) {
        if (lead) {
            // run batches until this caller's own request is served, then pass the lead on.  Callers released by the previous
            // batch come back within a microsecond or two: a short gathering window (until the queue stops growing, a few
            // microseconds at most, against a round trip of ~50) keeps them in THIS batch instead of the one after
            {
                size_t seen = 0;
                for (int w = 0; w < 6; ++w) {
                    size_t now;
                    { std::lock_guard<std::mutex> lk(g_pair_mu); now = g_pair_q.size(); }
                    if (w > 0 && now == seen) break;
                    seen = now;
                    const auto t0 = std::chrono::steady_clock::now();
                    while (std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() < 1.5) {
#if defined(__x86_64__) || defined(__i386__)
                        __builtin_ia32_pause();
#endif
                    }
                }
                std::lock_guard<std::mutex> lk(g_pair_mu);
                mine.swap(g_pair_q);
            }
            if (!mine.empty()) run_pair_jobs(mine);
            mine.clear();
            if (job.done.load(std::memory_order_acquire)) {
                std::lock_guard<std::mutex> lk(g_pair_mu);
                g_pair_leader.store(false, std::memory_order_release);       // a waiting caller picks the lead up (below)
                return job.rc;
            }
            continue;
        }
        if (job.done.load(std::memory_order_acquire)) return job.rc;
        if ((++spins & 15u) == 0u && !g_pair_leader.load(std::memory_order_acquire)) {       // no leader: take the lead (lock only then)
            std::lock_guard<std::mutex> lk(g_pair_mu);
            if (!g_pair_leader.load(std::memory_order_relaxed) && !job.done.load(std::memory_order_acquire)) {
                g_pair_leader.store(true, std::memory_order_relaxed); lead = true;
            }
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if (spins > 4000u) std::this_thread::yield();
    }
}

}  // extern "C"
