// C entry points over csrc/host_plan.h (the shard / chunk planning of the host-buffer batch calls) for the CPU tests.
// Test-only.
#include "../../collision_detect_gjk_epa_b200/csrc/host_plan.h"

extern "C" {

void plan_shard_range(int64_t n_pairs, int32_t s, int32_t n, int64_t *lo, int64_t *hi) { hostplan::shard_range(n_pairs, s, n, *lo, *hi); }
int32_t plan_shards_for(int64_t n_pairs, int32_t n_devices) { return hostplan::shards_for(n_pairs, n_devices); }

// out: [nchunk, cp, ranged, Hlo, Hhi, bad_index]; chunks: up to 64 rows of (a, b, hlo, hhi).  scan_store = 2: the lazy plan
// (sampled decision, then the chunks' exact ranges one by one, as host_batch.inl does while the copies are in flight)
void plan_shard_c(int64_t np, int32_t forced, double bytes_per_pair, int32_t scan_store, int64_t n_hulls, const int64_t *hull_off,
                  const int32_t *pa, const int32_t *pb, int64_t *out, int64_t *chunks)
{
    hostplan::Plan P = hostplan::plan_shard(np, forced, bytes_per_pair, scan_store != 0, n_hulls, hull_off, pa, pb, scan_store == 2);
    if (P.lazy) for (int k = 0; k < P.nchunk; ++k) if (!hostplan::scan_chunk(P, k, n_hulls, pa, pb)) P.bad_index = true;
    out[0] = P.nchunk; out[1] = P.cp; out[2] = P.ranged; out[3] = P.Hlo; out[4] = P.Hhi; out[5] = P.bad_index;
    for (int k = 0; k < P.nchunk && !P.bad_index; ++k) {
        chunks[4 * k] = P.ch[k].a; chunks[4 * k + 1] = P.ch[k].b; chunks[4 * k + 2] = P.ch[k].hlo; chunks[4 * k + 3] = P.ch[k].hhi;
    }
}

}
