// host_batch.inl -- the HOST-buffer side of the C ABI (included by gjkepa_capi.cu after gjkepa_batch_device_f64).
//
// What a Fortran caller of GJKEPA (reference src/GCLIB_GJKEPA.f90:39-52) holds is host memory, and its model of
// parallelism is "the caller runs the loop over pairs on every core" (THREADPRIVATE scratch, F:55-60).  Here the
// loop over pairs is the batch, and this file is what carries a host batch to the GPUs and back:
//
//   * one SHARD per device: a contiguous range of the pair list (pairs are independent: no data-path collective,
//     SURVEY.md §8e).  A context made by gjkepa_create_multi owns one child context and one worker thread per
//     device; every device copies its own results straight into the caller's arrays at the pair's global index,
//     so N PCIe links work in parallel and nothing funnels through one GPU.
//   * every shard runs in CHUNKS through a three-stream pipeline (copy-in, compute, copy-out; two slots each):
//     H2D(k+1) overlaps the kernels of chunk k and the D2H of chunk k-1.  When the hull store is pair-private
//     (cfg2-cfg4: pair i uses hulls near 2i) a chunk brings only the hull range its pairs name and a small kernel
//     makes the pair list relative to it; a shared store (cfg5) is uploaded whole, once per shard -- or not at all
//     when it is resident (gjkepa_store_upload).
//   * contact_index / n_contacts / GJKEPA_MODE_COMPACT rows are built ON THE DEVICE (cub::DeviceSelect over the
//     flags, then a row gather into per-shard accumulators): no host loop over pairs, and only contacts cross PCIe.
#pragma once
#include "host_plan.h"

namespace hostb {

constexpr int MAX_CHUNKS = hostplan::MAX_CHUNKS;
constexpr int N_CNT = 3 * MAX_CHUNKS + 2;      // HostPipe::d_cnt

// Pair list and hull offsets of a chunk made relative to the first hull / first vertex of the range it brought along, and
// every hull number checked against the store the kernels will index (n_hulls entries): a number outside it is replaced by 0
// -- so that no kernel reads out of bounds -- and reported through *bad, which fails the call with GJKEPA_ERR_ARG.
__global__ void __launch_bounds__(256)
rebase_kernel(int32_t *__restrict__ pa, int32_t *__restrict__ pb, int64_t n_pairs, int32_t hull_lo, int64_t n_hulls,
              int64_t *__restrict__ off, int64_t n_off, int64_t vert_lo, int64_t *__restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n = n_pairs > n_off ? n_pairs : n_off;
    bool wrong = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (i < n_pairs) {
            int32_t a = pa[i] - hull_lo, b = pb[i] - hull_lo;
            if (a < 0 || a >= n_hulls) { a = 0; wrong = true; }
            if (b < 0 || b >= n_hulls) { b = 0; wrong = true; }
            pa[i] = a; pb[i] = b;
        }
        if (i < n_off) off[i] -= vert_lo;
    }
    if (bad && __any_sync(0xffffffffu, wrong) && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)bad, 1ull);
}

struct CompactArgs {
    const int32_t *sel;          // colliding pairs of this chunk: GLOBAL pair numbers, ascending
    const int64_t *count;        // how many
    const int64_t *base;         // rows already in the accumulators
    int64_t pair0;               // global number of the chunk's first pair
    const double *depth, *normal, *point, *nearest;      // chunk-dense results (null = not wanted)
    const int32_t *colli_type; const uint32_t *status;
    int32_t *o_idx; double *o_depth, *o_normal, *o_point, *o_nearest; int32_t *o_type; uint32_t *o_status;
};

// rows of the chunk's contacts appended to the shard's accumulators (ascending pair number: chunks are processed in order)
__global__ void __launch_bounds__(256)
compact_rows_kernel(CompactArgs g)
{
    const int64_t n = *g.count, base = *g.base;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t gp = g.sel[k];
        const int64_t i = (int64_t)gp - g.pair0, o = base + k;
        g.o_idx[o] = gp;
        if (g.depth) g.o_depth[o] = g.depth[i];
        if (g.normal) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g.o_normal[3 * o + c] = g.normal[3 * i + c];
        }
        if (g.point) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g.o_point[3 * o + c] = g.point[3 * i + c];
        }
        if (g.nearest) {
#pragma unroll
            for (int c = 0; c < 6; ++c) g.o_nearest[6 * o + c] = g.nearest[6 * i + c];
        }
        if (g.colli_type) g.o_type[o] = g.colli_type[i];
        if (g.status) g.o_status[o] = g.status[i];
    }
}

__global__ void advance_kernel(int64_t *base, const int64_t *count) { *base += *count; }

// sort keys of the locality order: first hull of the pair in the high word, the pair's number in the chunk in the low word
__global__ void __launch_bounds__(256)
order_keys_kernel(const int32_t *__restrict__ pa, int64_t n, unsigned long long *__restrict__ keys)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = ((unsigned long long)(uint32_t)pa[i] << 32) | (unsigned long long)(uint32_t)i;
}

}  // namespace hostb

struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_store = nullptr;
    void *d_in[2] = {nullptr, nullptr};  size_t cap_in[2] = {0, 0};
    void *d_out[2] = {nullptr, nullptr}; size_t cap_out[2] = {0, 0};
    void *d_whole = nullptr; size_t cap_whole = 0;      // a host batch's hull store uploaded whole (shared stores)
    void *d_acc = nullptr;   size_t cap_acc = 0;        // compact accumulators of the shard
    void *d_sel_tmp = nullptr; size_t cap_sel_tmp = 0;
    void *d_order = nullptr; size_t cap_order = 0;      // locality order of a chunk: keys in | keys out | cub scratch
    int64_t *d_cnt = nullptr;                           // [0..MAX_CHUNKS) contacts per chunk, [MAX_CHUNKS] rows accumulated, [MAX_CHUNKS+1] bad hull numbers,
                                                        // [MAX_CHUNKS+2 ..) survivors of the fp32 filter / binary32 walk per chunk,
                                                        // [2*MAX_CHUNKS+2 ..) pairs the binary32 thread-per-pair kernel left to its group kernel
    int64_t *h_cnt = nullptr;                           // pinned mirror
    // accumulator layout of the last shard (run_shard -> finish_shard)
    size_t a_idx = 0, a_dep = 0, a_nrm = 0, a_pt = 0, a_near = 0, a_type = 0, a_st = 0;
    bool acc_valid = false;
};

namespace {

enum Src { SRC_HOST = 0, SRC_HOST_F32, SRC_RESIDENT, SRC_POSED };

// one call of a host-buffer entry point
struct Req {
    int32_t version = 1; double tol_ff = 0.0; int32_t mode = 0; bool compact = false; Src src = SRC_HOST;
    int64_t n_hulls = 0; const void *verts = nullptr; const int64_t *hull_off = nullptr;          // SRC_HOST / SRC_HOST_F32
    int64_t n_bodies = 0; const int32_t *body_shape = nullptr; const double *pose = nullptr;      // SRC_POSED
    int64_t n_pairs = 0; const int32_t *pair_a = nullptr, *pair_b = nullptr;
    uint8_t *collision = nullptr; int32_t *colli_type = nullptr; void *nearest = nullptr, *normal = nullptr, *point = nullptr, *depth = nullptr;
    uint32_t *status = nullptr; int64_t *n_contacts = nullptr; int32_t *contact_index = nullptr;
    bool f32 = false;            // the caller's real-valued arrays (store and results) are float
};

size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

void pipe_free(gjkepa_ctx *ctx);

int pipe_get(gjkepa_ctx *ctx, HostPipe **out)
{
    if (ctx->pipe) { *out = ctx->pipe; return GJKEPA_OK; }
    HostPipe *p = new HostPipe();
    cudaError_t e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_store, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_cnt, hostb::N_CNT * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&p->h_cnt, hostb::N_CNT * sizeof(int64_t), cudaHostAllocDefault);
    ctx->pipe = p;
    if (e != cudaSuccess) { pipe_free(ctx); return fail(ctx, GJKEPA_ERR_CUDA, "host pipeline streams/events", e); }
    *out = p;
    return GJKEPA_OK;
}

void pipe_free(gjkepa_ctx *ctx)
{
    HostPipe *p = ctx->pipe;
    if (!p) return;
    for (int i = 0; i < 2; ++i) {
        if (p->d_in[i]) cudaFree(p->d_in[i]);
        if (p->d_out[i]) cudaFree(p->d_out[i]);
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
        if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
    }
    if (p->ev_store) cudaEventDestroy(p->ev_store);
    if (p->d_whole) cudaFree(p->d_whole);
    if (p->d_acc) cudaFree(p->d_acc);
    if (p->d_sel_tmp) cudaFree(p->d_sel_tmp);
    if (p->d_order) cudaFree(p->d_order);
    if (p->d_cnt) cudaFree(p->d_cnt);
    if (p->h_cnt) cudaFreeHost(p->h_cnt);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
    ctx->pipe = nullptr;
}

using hostplan::Chunk;

// K0 once for a store that several chunks share (skipped under shortcut E1, which drops the sphere test)
int spheres_for(gjkepa_ctx *ctx, const double *d_verts, const int64_t *d_off, int64_t n_hulls, gjk::Sphere *out, cudaStream_t st)
{
    const int64_t blocks = (n_hulls + gjk::PREP_WARPS * 32 - 1) / (gjk::PREP_WARPS * 32);
    gjk::hull_prep_kernel<<<(unsigned)blocks, gjk::PREP_WARPS * 32, 0, st>>>(d_verts, d_off, n_hulls, out);
    ++ctx->launches;
    CU(cudaGetLastError());
    return GJKEPA_OK;
}

// Runs pairs [P0,P1) of the request on ctx's device.  Dense outputs are copied to the caller's arrays at the pair's
// global index as the chunks finish; compact rows / contact indices stay in the device accumulators until
// finish_shard places them (their row offset depends on the shards before this one).  *count = contacts of the shard
// (valid when contact_index, n_contacts or compact rows were asked for).  Synchronises before returning.
int run_shard(gjkepa_ctx *ctx, const Req &r, int64_t P0, int64_t P1, int64_t *count)
{
    *count = 0;
    const int64_t np = P1 - P0;
    if (np <= 0) return GJKEPA_OK;
    DeviceKeep keep_device_;
    CU(cudaSetDevice(ctx->device));
    HostPipe *pp = nullptr;
    if (int rc = pipe_get(ctx, &pp)) return rc;
    HostPipe &p = *pp;
    p.acc_valid = false;
    cudaStream_t sc = ctx->stream, si = p.s_in, so = p.s_out;
    const bool epa = r.mode >= GJKEPA_MODE_EPA, full = r.mode == GJKEPA_MODE_FULL;
    const bool w_dep = r.depth && epa, w_nrm = r.normal && epa, w_pt = r.point && full, w_near = r.nearest && full, w_type = r.colli_type && full;
    const bool w_st = r.status != nullptr;
    const bool want_idx = r.contact_index || r.compact;
    const bool want_cnt = r.n_contacts || want_idx;
    const bool need_flags = r.collision || want_cnt;
    const size_t RS = r.f32 ? 4 : 8;                    // bytes per real in the caller's arrays
    const int32_t *pa = r.pair_a + P0, *pb = r.pair_b + P0;

    // ---- chunk plan (host_plan.h) ----
    const bool host_store = r.src == SRC_HOST || r.src == SRC_HOST_F32;
    const bool private_store = host_store && r.n_hulls >= r.n_pairs / 2;     // else shared by the pairs (cfg5): it goes up whole
    double per_pair = 8.0 + (r.collision ? 1.0 : 0.0) + (r.status && !r.compact ? 4.0 : 0.0);
    if (!r.compact) per_pair += (r.f32 ? 0.5 : 1.0) * ((w_dep ? 8 : 0) + (w_nrm ? 24 : 0) + (w_pt ? 24 : 0) + (w_near ? 48 : 0)) + (w_type ? 4 : 0);
    if (private_store) per_pair += (r.f32 ? 12.0 : 24.0) * (double)(r.hull_off[r.n_hulls] - r.hull_off[0]) / (double)r.n_pairs;
    hostplan::Plan plan = hostplan::plan_shard(np, ctx->host_chunks, per_pair, private_store, r.n_hulls, r.hull_off, pa, pb, /*lazy=*/true);
    if (plan.bad_index) return fail(ctx, GJKEPA_ERR_ARG, "pair list names a hull outside the store");
    const int nchunk = plan.nchunk;
    const int64_t cp = plan.cp;
    // fp32 entry with GJKEPA_TUNE_F32_ARITH: the GJK walk itself runs in binary32 (gjk_f32.cuh); the binary64 kernels only see the
    // pairs it calls colliding, and in GJK mode not even those
    const bool use_arith = r.src == SRC_HOST_F32 && ctx->f32_arith && ctx->lanes == 0 && !ctx->single_kernel && cp > 256;
    const bool arith_only = use_arith && !epa;
    std::vector<Chunk> &ch = plan.ch;
    const bool ranged = plan.ranged;
    const int64_t Hlo = plan.Hlo, Hhi = plan.Hhi;               // hull range uploaded whole when !ranged
    const int64_t whole_nh = Hhi - Hlo + 1;
    const int64_t whole_v0 = host_store ? r.hull_off[Hlo] : 0, whole_nv = host_store ? r.hull_off[Hhi + 1] - whole_v0 : 0;

    // ---- slot sizes ----
    size_t need_in = 0, need_out = 0;
    struct InL { size_t v32, v64, off, pa, pb, total; };
    auto in_layout = [&](int64_t nv, int64_t nh, int64_t n) {
        InL L;
        L.v32 = 0;
        L.v64 = L.v32 + al256(r.src == SRC_HOST_F32 ? (size_t)nv * 12 : 0);
        L.off = L.v64 + al256((size_t)nv * 24);
        L.pa = L.off + al256(nh > 0 ? (size_t)(nh + 1) * 8 : 0);
        L.pb = L.pa + al256((size_t)n * 4);
        L.total = L.pb + al256((size_t)n * 4);
        return L;
    };
    for (int k = 0; k < nchunk; ++k) {
        const int64_t n = ch[k].b - ch[k].a;
        if (plan.lazy && k > 0) break;                   // lazy plan: sized by the first chunk here, re-checked per chunk below
        if (plan.lazy && !hostplan::scan_chunk(plan, 0, r.n_hulls, pa, pb)) return fail(ctx, GJKEPA_ERR_ARG, "pair list names a hull outside the store");
        const InL L = ranged ? in_layout(r.hull_off[ch[k].hhi + 1] - r.hull_off[ch[k].hlo], ch[k].hhi - ch[k].hlo + 1, n) : in_layout(0, 0, n);
        need_in = L.total > need_in ? L.total : need_in;
    }
    if (plan.lazy) need_in += need_in / 16;              // the other chunks of an ordered list are about this size
    // results of a chunk, dense: flags | type | status | reals (nearest, normal, point, depth back to back) | float copy of the reals | selected pairs
    const size_t n_real = (w_near ? 6 : 0) + (w_nrm ? 3 : 0) + (w_pt ? 3 : 0) + (w_dep ? 1 : 0);        // per pair
    const size_t o_col = 0, o_type = o_col + al256(need_flags ? (size_t)cp : 0), o_st = o_type + al256(w_type ? (size_t)cp * 4 : 0),
                 o_r64 = o_st + al256(w_st ? (size_t)cp * 4 : 0), o_r32 = o_r64 + al256((size_t)cp * n_real * 8),
                 o_sel = o_r32 + al256(r.f32 && !r.compact ? (size_t)cp * n_real * 4 : 0);
    need_out = o_sel + al256(want_cnt ? (size_t)cp * 4 : 0);
    const size_t r_near = 0, r_nrm = r_near + (w_near ? 6 : 0), r_pt = r_nrm + (w_nrm ? 3 : 0), r_dep = r_pt + (w_pt ? 3 : 0);   // x cp elements
    for (int s = 0; s < (nchunk > 1 ? 2 : 1); ++s) {
        if (int rc = grow_bytes(ctx, p.d_in[s], p.cap_in[s], need_in, "cudaMalloc input slot")) return rc;
        if (int rc = grow_bytes(ctx, p.d_out[s], p.cap_out[s], need_out, "cudaMalloc output slot")) return rc;
    }
    size_t sel_tmp = 0;
    if (want_cnt) {
        CU(cub::DeviceSelect::Flagged(nullptr, sel_tmp, thrust::counting_iterator<int32_t>(0), (const uint8_t *)nullptr, (int32_t *)nullptr,
                                      (int64_t *)nullptr, (int)cp, sc));
        if (int rc = grow_bytes(ctx, p.d_sel_tmp, p.cap_sel_tmp, sel_tmp + 256, "cudaMalloc select scratch")) return rc;
    }
    if (want_idx) {
        // accumulators: worst case every pair of the shard collides
        p.a_idx = 0;
        p.a_dep = p.a_idx + al256((size_t)np * 4);
        p.a_nrm = p.a_dep + al256(r.compact && w_dep ? (size_t)np * 8 : 0);
        p.a_pt = p.a_nrm + al256(r.compact && w_nrm ? (size_t)np * 24 : 0);
        p.a_near = p.a_pt + al256(r.compact && w_pt ? (size_t)np * 24 : 0);
        p.a_type = p.a_near + al256(r.compact && w_near ? (size_t)np * 48 : 0);
        p.a_st = p.a_type + al256(r.compact && w_type ? (size_t)np * 4 : 0);
        const size_t a_total = p.a_st + al256(r.compact && w_st ? (size_t)np * 4 : 0);
        if (int rc = grow_bytes(ctx, p.d_acc, p.cap_acc, a_total, "cudaMalloc contact accumulators")) return rc;
    }
    CU(cudaMemsetAsync(p.d_cnt, 0, hostb::N_CNT * sizeof(int64_t), sc));

    // ---- the store the chunks run against ----
    const double *d_verts = nullptr; const int64_t *d_off = nullptr; int64_t nh_dev = 0;
    const gjk::Sphere *sph = nullptr;
    const int save_hint = ctx->max_hull_hint;
    const gjk::Sphere *save_ext = ctx->ext_sphere;
    struct Restore { gjkepa_ctx *c; int hint; const gjk::Sphere *ext; ~Restore() { c->max_hull_hint = hint; c->ext_sphere = ext; } } restore{ctx, save_hint, save_ext};
    const bool e1_skips_spheres = ctx->early_out && ctx->lanes == 0 && !ctx->single_kernel;
    if (r.src == SRC_RESIDENT) {
        if (ctx->n_store_hulls <= 0) return fail(ctx, GJKEPA_ERR_ARG, "no hull store uploaded (gjkepa_store_upload)");
        d_verts = ctx->d_store_verts; d_off = ctx->d_store_off; nh_dev = ctx->n_store_hulls; sph = ctx->d_store_sphere;
        ctx->max_hull_hint = ctx->store_max_verts;
    } else if (r.src == SRC_POSED) {
        // poses up, world-space store rebuilt on the device (Kp); the chunks then carry pair lists only
        const size_t o_pose = al256((size_t)r.n_bodies * 4), tot = o_pose + al256((size_t)r.n_bodies * 96);
        if (int rc = grow_bytes(ctx, ctx->d_stage2, ctx->cap_stage2, tot, "cudaMalloc staging")) return rc;
        char *b = (char *)ctx->d_stage2;
        CU(cudaMemcpyAsync(b, r.body_shape, (size_t)r.n_bodies * 4, cudaMemcpyHostToDevice, sc));
        CU(cudaMemcpyAsync(b + o_pose, r.pose, (size_t)r.n_bodies * 96, cudaMemcpyHostToDevice, sc));
        int64_t *doff = nullptr;
        if (int rc = posed_world(ctx, r.n_bodies, (const int32_t *)b, (const double *)(b + o_pose), &doff)) return rc;
        d_verts = ctx->d_world; d_off = doff; nh_dev = r.n_bodies;
        ctx->max_hull_hint = (int)ctx->max_shape_verts;
        if (nchunk > 1 && !e1_skips_spheres) {
            if (int rc = ensure(ctx, ctx->d_sphere, ctx->cap_sphere, nh_dev)) return rc;
            if (int rc = spheres_for(ctx, d_verts, d_off, nh_dev, ctx->d_sphere, sc)) return rc;
            sph = ctx->d_sphere;
        }
    } else if (!ranged) {
        // hull range [Hlo,Hhi] of a shared host store, once for the shard
        const size_t o_v64 = al256(r.f32 ? (size_t)whole_nv * 12 : 0), o_off = o_v64 + al256((size_t)whole_nv * 24),
                     tot = o_off + al256((size_t)(whole_nh + 1) * 8);
        if (int rc = grow_bytes(ctx, p.d_whole, p.cap_whole, tot, "cudaMalloc hull store")) return rc;
        char *b = (char *)p.d_whole;
        if (r.f32) CU(cudaMemcpyAsync(b, (const float *)r.verts + 3 * whole_v0, (size_t)whole_nv * 12, cudaMemcpyHostToDevice, si));
        else       CU(cudaMemcpyAsync(b + o_v64, (const double *)r.verts + 3 * whole_v0, (size_t)whole_nv * 24, cudaMemcpyHostToDevice, si));
        CU(cudaMemcpyAsync(b + o_off, r.hull_off + Hlo, (size_t)(whole_nh + 1) * 8, cudaMemcpyHostToDevice, si));
        CU(cudaEventRecord(p.ev_store, si));
        CU(cudaStreamWaitEvent(sc, p.ev_store, 0));
        if (r.f32 && !arith_only) {
            const int64_t ne = whole_nv * 3;
            int64_t gb = (ne / 4 + 256) / 256; const int64_t gm = (int64_t)ctx->sm_count * 16;
            f32io::widen_kernel<<<(unsigned)(gb < gm ? gb : gm), 256, 0, sc>>>((const float *)b, (double *)(b + o_v64), ne);
            ++ctx->launches;
        }
        d_verts = (const double *)(b + o_v64); d_off = (const int64_t *)(b + o_off); nh_dev = whole_nh;
        if (Hlo != 0 || whole_v0 != 0) {            // offsets relative to the first vertex brought along (pair lists: per chunk)
            hostb::rebase_kernel<<<(unsigned)((whole_nh + 1 + 255) / 256 < 2048 ? (whole_nh + 1 + 255) / 256 : 2048), 256, 0, sc>>>(
                nullptr, nullptr, 0, 0, 0, (int64_t *)(b + o_off), whole_nh + 1, whole_v0, nullptr);
            ++ctx->launches;
        }
        if (nchunk > 1 && !e1_skips_spheres && !arith_only) {
            if (int rc = ensure(ctx, ctx->d_sphere, ctx->cap_sphere, nh_dev)) return rc;
            if (int rc = spheres_for(ctx, d_verts, d_off, nh_dev, ctx->d_sphere, sc)) return rc;
            sph = ctx->d_sphere;
        }
    }

    // ---- locality order (GJKEPA_TUNE_LOCALITY_SORT) ----
    // Against a store larger than L2 the init kernel's gathers hit DRAM unless neighbouring pairs share hulls (ncu, round 2: the
    // same 4M pairs over 2M bodies cost 5.15 ms in random order and 3.49 ms grouped by first hull).  A broad phase emits its pairs
    // grouped; a list that is not (a sample of it has many descents) is VISITED in the order of its first hull -- one key pass
    // and one radix sort per chunk on the device, ~0.1 ms per 1M pairs.
    bool use_order = false;
    int order_bits = 1;
    if (!ranged && ctx->lanes == 0 && !ctx->single_kernel && ctx->locality_sort != 0 && cp > 256) {
        use_order = ctx->locality_sort == 1;
        if (ctx->locality_sort == 2 && np >= 262144) {
            const int64_t store_bytes = r.src == SRC_RESIDENT ? (int64_t)ctx->n_store_hulls * 24 * ctx->store_max_verts
                                      : r.src == SRC_POSED ? r.n_bodies * 24 * ctx->max_shape_verts : whole_nv * 24;
            if (store_bytes > (int64_t)96 << 20) {
                int64_t desc = 0; const int64_t ns = np < 8192 ? np : 8192, s0 = (np - ns) / 2;
                for (int64_t i = s0 + 1; i < s0 + ns; ++i) desc += pa[i] < pa[i - 1];
                use_order = desc * 8 > ns;                   // more than one descent in eight steps: not a broad phase's order
            }
        }
        const int64_t nh_all = r.src == SRC_RESIDENT ? ctx->n_store_hulls : (r.src == SRC_POSED ? r.n_bodies : whole_nh);
        while (order_bits < 32 && ((int64_t)1 << order_bits) < nh_all) ++order_bits;
    }
    // fp32 entry with GJKEPA_TUNE_F32_FILTER: the binary32 separation filter runs first and the init kernel visits its survivors
    const bool use_filter = !use_arith && r.src == SRC_HOST_F32 && ctx->f32_filter && ctx->lanes == 0 && !ctx->single_kernel && cp > 256;
    if (use_filter || use_arith) use_order = false;
    ctx->filter_stats[0] = ctx->filter_stats[1] = 0;
    size_t order_tmp = 0, o_keys1 = 0, o_otmp = 0;
    if (use_filter) { if (int rc = grow_bytes(ctx, p.d_order, p.cap_order, al256((size_t)cp * 8), "cudaMalloc survivor list")) return rc; }
    if (use_arith) { if (int rc = grow_bytes(ctx, p.d_order, p.cap_order, al256((size_t)cp * 8) + al256((size_t)cp * 4), "cudaMalloc survivor list")) return rc; }
    if (use_order) {
        CU(cub::DeviceRadixSort::SortKeys(nullptr, order_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)cp, 32, 32 + order_bits, sc));
        o_keys1 = al256((size_t)cp * 8); o_otmp = 2 * o_keys1;
        if (int rc = grow_bytes(ctx, p.d_order, p.cap_order, o_otmp + al256(order_tmp), "cudaMalloc locality order")) return rc;
    }
    struct OrderBack { gjkepa_ctx *c; ~OrderBack() { c->ext_order = nullptr; c->ext_order_count = nullptr; } } order_back{ctx};

    // ---- the pipeline ----
    double acc_ms[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < nchunk; ++k) {
        const int s = k & 1;
        const int64_t a = ch[k].a, n = ch[k].b - ch[k].a;
        if (plan.lazy) {
            // this chunk's exact hull range, scanned while the copies of the chunk before are on the bus
            if (k > 0 && !hostplan::scan_chunk(plan, k, r.n_hulls, pa, pb)) return fail(ctx, GJKEPA_ERR_ARG, "pair list names a hull outside the store");
            const InL Lk = in_layout(r.hull_off[ch[k].hhi + 1] - r.hull_off[ch[k].hlo], ch[k].hhi - ch[k].hlo + 1, n);
            if (Lk.total > p.cap_in[s]) {                  // rare: a chunk that names a much wider range than the first one
                CU(cudaStreamSynchronize(sc)); CU(cudaStreamSynchronize(si));
                if (int rcg = grow_bytes(ctx, p.d_in[s], p.cap_in[s], Lk.total, "cudaMalloc input slot")) return rcg;
            }
        }
        char *in = (char *)p.d_in[s], *out = (char *)p.d_out[s];
        const int64_t hlo = ranged ? ch[k].hlo : Hlo;
        const int64_t nh = ranged ? ch[k].hhi - ch[k].hlo + 1 : 0;
        const int64_t v0 = ranged ? r.hull_off[hlo] : 0, nv = ranged ? r.hull_off[ch[k].hhi + 1] - v0 : 0;
        const InL L = in_layout(nv, nh, n);
        // copy-in: the slot is free once the kernels of chunk k-2 are done
        if (k >= 2) CU(cudaStreamWaitEvent(si, p.ev_done[s], 0));
        if (ranged) {
            if (r.f32) CU(cudaMemcpyAsync(in + L.v32, (const float *)r.verts + 3 * v0, (size_t)nv * 12, cudaMemcpyHostToDevice, si));
            else       CU(cudaMemcpyAsync(in + L.v64, (const double *)r.verts + 3 * v0, (size_t)nv * 24, cudaMemcpyHostToDevice, si));
            CU(cudaMemcpyAsync(in + L.off, r.hull_off + hlo, (size_t)(nh + 1) * 8, cudaMemcpyHostToDevice, si));
        }
        CU(cudaMemcpyAsync(in + L.pa, pa + a, (size_t)n * 4, cudaMemcpyHostToDevice, si));
        CU(cudaMemcpyAsync(in + L.pb, pb + a, (size_t)n * 4, cudaMemcpyHostToDevice, si));
        CU(cudaEventRecord(p.ev_in[s], si));
        // compute: inputs landed, and the D2H of chunk k-2 has drained this output slot
        CU(cudaStreamWaitEvent(sc, p.ev_in[s], 0));
        if (k >= 2) CU(cudaStreamWaitEvent(sc, p.ev_out[s], 0));
        const int64_t gm = (int64_t)ctx->sm_count * 16;
        if (ranged && r.f32 && !arith_only) {
            const int64_t ne = nv * 3; const int64_t gb = (ne / 4 + 256) / 256;
            f32io::widen_kernel<<<(unsigned)(gb < gm ? gb : gm), 256, 0, sc>>>((const float *)(in + L.v32), (double *)(in + L.v64), ne);
            ++ctx->launches;
        }
        {   // always: it is also the bounds check of the pair list (a wrong hull number must not become a wild device read)
            const int64_t w = ranged ? (n > nh + 1 ? n : nh + 1) : n; const int64_t gb = (w + 255) / 256;
            hostb::rebase_kernel<<<(unsigned)(gb < gm ? gb : gm), 256, 0, sc>>>((int32_t *)(in + L.pa), (int32_t *)(in + L.pb), n, (int32_t)hlo,
                ranged ? nh : nh_dev, ranged ? (int64_t *)(in + L.off) : nullptr, ranged ? nh + 1 : 0, v0, p.d_cnt + hostb::MAX_CHUNKS + 1);
            ++ctx->launches;
        }
        if (use_order) {
            unsigned long long *k0 = (unsigned long long *)p.d_order, *k1 = (unsigned long long *)((char *)p.d_order + o_keys1);
            const int64_t gb = (n + 255) / 256;
            hostb::order_keys_kernel<<<(unsigned)(gb < gm ? gb : gm), 256, 0, sc>>>((const int32_t *)(in + L.pa), n, k0);
            size_t tb = order_tmp;
            CU(cub::DeviceRadixSort::SortKeys((char *)p.d_order + o_otmp, tb, k0, k1, (int)n, 32, 32 + order_bits, sc));
            ctx->launches += 2;
            ctx->ext_order = k1;
        }
        if (use_filter) {
            unsigned long long *cnt = (unsigned long long *)(p.d_cnt + hostb::MAX_CHUNKS + 2 + k);      // zeroed with the counter block
            const float *v32 = ranged ? (const float *)(in + L.v32) : (const float *)p.d_whole;
            const int64_t gb = (n * f32io::FLP + 255) / 256;         // FLP lanes per pair
            f32io::separation_filter_kernel<<<(unsigned)(gb < gm ? gb : gm), 256, 0, sc>>>(v32, ranged ? (const int64_t *)(in + L.off) : d_off,
                (const int32_t *)(in + L.pa), (const int32_t *)(in + L.pb), n, need_flags ? (uint8_t *)(out + o_col) : nullptr,
                w_st ? (uint32_t *)(out + o_st) : nullptr, (uint32_t)GJKEPA_EXIT_EARLY_OUT, (unsigned long long *)p.d_order, cnt);
            ++ctx->launches;
            ctx->ext_order = (const unsigned long long *)p.d_order; ctx->ext_order_count = cnt;
        }
        if (use_arith) {
            gjk32::F32Args fa;
            fa.verts = ranged ? (const float *)(in + L.v32) : (const float *)p.d_whole;
            fa.hull_off = ranged ? (const int64_t *)(in + L.off) : d_off;
            fa.pair_a = (const int32_t *)(in + L.pa); fa.pair_b = (const int32_t *)(in + L.pb); fa.n_pairs = n;
            fa.collision = need_flags ? (uint8_t *)(out + o_col) : nullptr;
            fa.status = w_st ? (uint32_t *)(out + o_st) : nullptr;
            fa.mark = GJKEPA_ST_F32_DECIDED; fa.final_flags = arith_only ? 1 : 0;
            fa.order = (unsigned long long *)p.d_order;
            fa.order_count = (unsigned long long *)(p.d_cnt + hostb::MAX_CHUNKS + 2 + k);
            fa.ovf_list = (int32_t *)((char *)p.d_order + al256((size_t)cp * 8));
            fa.ovf_count = (unsigned long long *)(p.d_cnt + 2 * hostb::MAX_CHUNKS + 2 + k);
            fa.pair_list = fa.ovf_list; fa.pair_list_count = fa.ovf_count;
            const int hint = ctx->max_hull_hint;
            fa.cap = (hint > 0 && hint < 16) ? hint : 16;
            fa.early_out = ctx->early_out;
            const size_t smem_f = (size_t)gjk32::F32_WARPS * 6 * fa.cap * 32 * sizeof(float);
            const int64_t want = (n + gjk32::F32_WARPS * 32 - 1) / (gjk32::F32_WARPS * 32);
            auto launch_f = [&](void (*kernel)(gjk32::F32Args), bool persistent) -> int {
                int per_sm = 1;
                CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, gjk32::F32_WARPS * 32, smem_f));
                const int64_t res = (int64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
                kernel<<<(unsigned)(persistent || res < want ? res : want), gjk32::F32_WARPS * 32, smem_f, sc>>>(fa);
                ++ctx->launches;
                return GJKEPA_OK;
            };
            if (int rcl = launch_f(gjk32::gjk_f32_kernel, false)) return rcl;
            // always: a hull above the caller's bound must still get its answer (the bound only sizes K1f's columns here)
            {
                const size_t smem_g = (size_t)4 * 8 * gjk32::F32_SLOT_FLOATS * sizeof(float);
                int per_sm = 1;
                CU(cudaFuncSetAttribute(gjk32::gjk_f32_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjk32::gjk_f32_group_kernel, 128, smem_g));
                gjk32::gjk_f32_group_kernel<<<(unsigned)(ctx->sm_count * (per_sm > 0 ? per_sm : 1)), 128, smem_g, sc>>>(fa);
                ++ctx->launches;
            }
            ctx->ext_order = (const unsigned long long *)p.d_order; ctx->ext_order_count = fa.order_count;
        }
        double *r64 = (double *)(out + o_r64);
        ctx->ext_sphere = ranged ? nullptr : sph;
        int rc = arith_only ? GJKEPA_OK : gjkepa_batch_device_f64(ctx, r.version, r.tol_ff, ranged ? nh : nh_dev,
            ranged ? (const double *)(in + L.v64) : d_verts, ranged ? (const int64_t *)(in + L.off) : d_off, n,
            (const int32_t *)(in + L.pa), (const int32_t *)(in + L.pb), r.mode,
            need_flags ? (uint8_t *)(out + o_col) : nullptr, w_type ? (int32_t *)(out + o_type) : nullptr,
            w_near ? r64 + r_near * cp : nullptr, w_nrm ? r64 + r_nrm * cp : nullptr, w_pt ? r64 + r_pt * cp : nullptr, w_dep ? r64 + r_dep * cp : nullptr,
            w_st ? (uint32_t *)(out + o_st) : nullptr, nullptr, nullptr);
        if (rc) return rc;
        if (ctx->profiling && !arith_only) {    // a profiled run gives up the overlap: per-kernel durations summed over the chunks
            double t[3];
            if (int rq = gjkepa_kernel_ms(ctx, t)) return rq;
            for (int q = 0; q < 3; ++q) acc_ms[q] += t[q];
        }
        if (want_cnt) {          // colliding pairs of the chunk, ascending, and their number -- on the device, from the flags
            size_t tb = p.cap_sel_tmp;
            CU(cub::DeviceSelect::Flagged(p.d_sel_tmp, tb, thrust::counting_iterator<int32_t>((int32_t)(P0 + a)), (const uint8_t *)(out + o_col),
                                          (int32_t *)(out + o_sel), p.d_cnt + k, (int)n, sc));
            ++ctx->launches;
        }
        if (want_idx) {
            hostb::CompactArgs g;
            g.sel = (const int32_t *)(out + o_sel); g.count = p.d_cnt + k; g.base = p.d_cnt + hostb::MAX_CHUNKS; g.pair0 = P0 + a;
            const bool c = r.compact;
            g.depth = c && w_dep ? r64 + r_dep * cp : nullptr; g.normal = c && w_nrm ? r64 + r_nrm * cp : nullptr;
            g.point = c && w_pt ? r64 + r_pt * cp : nullptr; g.nearest = c && w_near ? r64 + r_near * cp : nullptr;
            g.colli_type = c && w_type ? (const int32_t *)(out + o_type) : nullptr; g.status = c && w_st ? (const uint32_t *)(out + o_st) : nullptr;
            char *acc = (char *)p.d_acc;
            g.o_idx = (int32_t *)(acc + p.a_idx); g.o_depth = (double *)(acc + p.a_dep); g.o_normal = (double *)(acc + p.a_nrm);
            g.o_point = (double *)(acc + p.a_pt); g.o_nearest = (double *)(acc + p.a_near); g.o_type = (int32_t *)(acc + p.a_type);
            g.o_status = (uint32_t *)(acc + p.a_st);
            const int64_t gb = (n + 255) / 256;
            hostb::compact_rows_kernel<<<(unsigned)(gb < gm ? gb : gm), 256, 0, sc>>>(g);
            hostb::advance_kernel<<<1, 1, 0, sc>>>(p.d_cnt + hostb::MAX_CHUNKS, p.d_cnt + k);
            ctx->launches += 2;
        }
        const bool dense_reals = !r.compact && n_real > 0;
        if (dense_reals && r.f32) {
            const int64_t ne = (int64_t)cp * (int64_t)n_real; const int64_t gb = (ne / 2 + 256) / 256;
            f32io::narrow_kernel<<<(unsigned)(gb < gm ? gb : gm), 256, 0, sc>>>(r64, (float *)(out + o_r32), ne);
            ++ctx->launches;
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(p.ev_done[s], sc));
        // copy-out, straight into the caller's arrays at the pair's global index
        CU(cudaStreamWaitEvent(so, p.ev_done[s], 0));
        const int64_t g0 = P0 + a;
        if (r.collision) CU(cudaMemcpyAsync(r.collision + g0, out + o_col, (size_t)n, cudaMemcpyDeviceToHost, so));
        if (!r.compact) {
            if (w_type) CU(cudaMemcpyAsync(r.colli_type + g0, out + o_type, (size_t)n * 4, cudaMemcpyDeviceToHost, so));
            if (w_st)   CU(cudaMemcpyAsync(r.status + g0, out + o_st, (size_t)n * 4, cudaMemcpyDeviceToHost, so));
            const char *rb = r.f32 ? out + o_r32 : out + o_r64;
            if (w_near) CU(cudaMemcpyAsync((char *)r.nearest + (size_t)g0 * 6 * RS, rb + r_near * cp * RS, (size_t)n * 6 * RS, cudaMemcpyDeviceToHost, so));
            if (w_nrm)  CU(cudaMemcpyAsync((char *)r.normal + (size_t)g0 * 3 * RS, rb + r_nrm * cp * RS, (size_t)n * 3 * RS, cudaMemcpyDeviceToHost, so));
            if (w_pt)   CU(cudaMemcpyAsync((char *)r.point + (size_t)g0 * 3 * RS, rb + r_pt * cp * RS, (size_t)n * 3 * RS, cudaMemcpyDeviceToHost, so));
            if (w_dep)  CU(cudaMemcpyAsync((char *)r.depth + (size_t)g0 * RS, rb + r_dep * cp * RS, (size_t)n * RS, cudaMemcpyDeviceToHost, so));
        }
        CU(cudaEventRecord(p.ev_out[s], so));
    }
    CU(cudaMemcpyAsync(p.h_cnt, p.d_cnt, hostb::N_CNT * sizeof(int64_t), cudaMemcpyDeviceToHost, sc));
    CU(cudaStreamSynchronize(sc));
    CU(cudaStreamSynchronize(so));
    CU(cudaStreamSynchronize(si));
    if (use_filter || use_arith) {
        int64_t kept = 0;
        for (int k = 0; k < nchunk; ++k) kept += p.h_cnt[hostb::MAX_CHUNKS + 2 + k];
        ctx->filter_stats[0] = np; ctx->filter_stats[1] = np - kept;
    }
    if (p.h_cnt[hostb::MAX_CHUNKS + 1] != 0) return fail(ctx, GJKEPA_ERR_ARG, "pair list names a hull outside the store (results of those pairs are meaningless)");
    if (want_cnt) {
        int64_t c = 0;
        for (int k = 0; k < nchunk; ++k) c += p.h_cnt[k];
        *count = c;
    }
    p.acc_valid = want_idx;
    if (ctx->profiling) { for (int q = 0; q < 3; ++q) ctx->host_ms[q] = acc_ms[q]; ctx->host_ms_valid = true; }
    return GJKEPA_OK;
}

// Contact indices / compact rows of the shard into the caller's arrays from row `row0` on.  Synchronises.
int finish_shard(gjkepa_ctx *ctx, const Req &r, int64_t row0, int64_t count)
{
    if (count <= 0 || !ctx->pipe || !ctx->pipe->acc_valid) return GJKEPA_OK;
    DeviceKeep keep_device_;
    CU(cudaSetDevice(ctx->device));
    HostPipe &p = *ctx->pipe;
    cudaStream_t so = p.s_out;
    const char *acc = (const char *)p.d_acc;
    const bool epa = r.mode >= GJKEPA_MODE_EPA, full = r.mode == GJKEPA_MODE_FULL;
    if (r.contact_index) CU(cudaMemcpyAsync(r.contact_index + row0, acc + p.a_idx, (size_t)count * 4, cudaMemcpyDeviceToHost, so));
    if (r.compact) {
        if (r.depth && epa)       CU(cudaMemcpyAsync((double *)r.depth + row0, acc + p.a_dep, (size_t)count * 8, cudaMemcpyDeviceToHost, so));
        if (r.normal && epa)      CU(cudaMemcpyAsync((double *)r.normal + 3 * row0, acc + p.a_nrm, (size_t)count * 24, cudaMemcpyDeviceToHost, so));
        if (r.point && full)      CU(cudaMemcpyAsync((double *)r.point + 3 * row0, acc + p.a_pt, (size_t)count * 24, cudaMemcpyDeviceToHost, so));
        if (r.nearest && full)    CU(cudaMemcpyAsync((double *)r.nearest + 6 * row0, acc + p.a_near, (size_t)count * 48, cudaMemcpyDeviceToHost, so));
        if (r.colli_type && full) CU(cudaMemcpyAsync(r.colli_type + row0, acc + p.a_type, (size_t)count * 4, cudaMemcpyDeviceToHost, so));
        if (r.status)             CU(cudaMemcpyAsync(r.status + row0, acc + p.a_st, (size_t)count * 4, cudaMemcpyDeviceToHost, so));
    }
    CU(cudaStreamSynchronize(so));
    return GJKEPA_OK;
}

}  // namespace

// ---- several devices behind one context ----
struct MultiDev {
    struct Worker {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::function<int()> job;
        bool has = false, done = true, quit = false;
        int rc = 0;
    };
    std::vector<gjkepa_ctx *> kids;
    std::vector<Worker *> workers;

    static void loop(Worker *w, int device)
    {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(w->mu);
        for (;;) {
            w->cv.wait(lk, [w] { return w->has || w->quit; });
            if (w->quit) return;
            std::function<int()> j = std::move(w->job);
            w->has = false;
            lk.unlock();
            const int rc = j();
            lk.lock();
            w->rc = rc; w->done = true;
            w->cv.notify_all();
        }
    }
    void post(int i, std::function<int()> j)
    {
        Worker *w = workers[(size_t)i];
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = std::move(j); w->has = true; w->done = false;
        w->cv.notify_all();
    }
    int wait(int i)
    {
        Worker *w = workers[(size_t)i];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [w] { return w->done; });
        return w->rc;
    }
    // the same job on every device, in parallel; first error wins (its text is copied to the parent)
    int all(gjkepa_ctx *parent, const std::function<int(int, gjkepa_ctx *)> &f)
    {
        const int n = (int)kids.size();
        for (int i = 0; i < n; ++i) { gjkepa_ctx *k = kids[(size_t)i]; post(i, [f, i, k] { return f(i, k); }); }
        int rc = GJKEPA_OK;
        for (int i = 0; i < n; ++i) { const int r = wait(i); if (r && !rc) { rc = r; parent->err = kids[(size_t)i]->err; } }
        return rc;
    }
};

namespace {

void multi_free(gjkepa_ctx *ctx)
{
    MultiDev *m = ctx->multi;
    if (!m) return;
    for (MultiDev::Worker *w : m->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (gjkepa_ctx *k : m->kids) gjkepa_destroy(k);
    delete m;
    ctx->multi = nullptr;
}

using hostplan::shard_range;

// one host-buffer batch on whatever the context stands for: its own device, or every device of a multi context
int run_request(gjkepa_ctx *ctx, const Req &r)
{
    const bool want_idx = r.contact_index || r.compact;
    if (r.compact && !r.n_contacts) return fail(ctx, GJKEPA_ERR_ARG, "GJKEPA_MODE_COMPACT needs n_contacts");
    if (r.compact && r.f32) return fail(ctx, GJKEPA_ERR_ARG, "GJKEPA_MODE_COMPACT is not offered by the fp32 entry");
    if (r.n_pairs == 0) { if (r.n_contacts) *r.n_contacts = 0; return GJKEPA_OK; }
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
    const clk::time_point t_start = clk::now();
    ctx->phase_ms[0] = ctx->phase_ms[1] = 0.0;
    if (!ctx->multi) {
        int64_t cnt = 0;
        if (int rc = run_shard(ctx, r, 0, r.n_pairs, &cnt)) return rc;
        ctx->phase_ms[0] = ms_since(t_start);
        const clk::time_point t1 = clk::now();
        if (want_idx) { if (int rc = finish_shard(ctx, r, 0, cnt)) return rc; }
        ctx->phase_ms[1] = ms_since(t1);
        if (r.n_contacts) *r.n_contacts = cnt;
        return GJKEPA_OK;
    }
    MultiDev &m = *ctx->multi;
    const int n = hostplan::shards_for(r.n_pairs, (int)m.kids.size());      // few pairs: fewer devices
    std::vector<int64_t> cnt((size_t)n, 0);
    for (int s = 0; s < n; ++s) {
        gjkepa_ctx *k = m.kids[(size_t)s];
        int64_t lo, hi; shard_range(r.n_pairs, s, n, lo, hi);
        int64_t *c = &cnt[(size_t)s];
        const Req *rq = &r;
        m.post(s, [k, rq, lo, hi, c] { return run_shard(k, *rq, lo, hi, c); });
    }
    int rc = GJKEPA_OK;
    for (int s = 0; s < n; ++s) { const int q = m.wait(s); if (q && !rc) { rc = q; ctx->err = m.kids[(size_t)s]->err; } }
    if (rc) return rc;
    ctx->phase_ms[0] = ms_since(t_start);
    const clk::time_point t_fin = clk::now();
    int64_t total = 0;
    if (want_idx) {
        for (int s = 0; s < n; ++s) {
            gjkepa_ctx *k = m.kids[(size_t)s];
            const int64_t row0 = total, c = cnt[(size_t)s];
            const Req *rq = &r;
            m.post(s, [k, rq, row0, c] { return finish_shard(k, *rq, row0, c); });
            total += c;
        }
        for (int s = 0; s < n; ++s) { const int q = m.wait(s); if (q && !rc) { rc = q; ctx->err = m.kids[(size_t)s]->err; } }
    } else {
        for (int s = 0; s < n; ++s) total += cnt[(size_t)s];
    }
    ctx->phase_ms[1] = ms_since(t_fin);
    if (r.n_contacts) *r.n_contacts = total;
    return rc;
}

}  // namespace
