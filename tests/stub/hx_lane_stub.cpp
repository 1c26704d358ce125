// hx_lane_stub.cpp — TEST INFRASTRUCTURE (CPU suite only, never part of the product library).
//
// Stands in for the device runtime (hx_api.cu) behind the internal lane interface of hx_host.h, so that the host
// pipeline of hx_get_hypervar_regions — chunker, parallel parsers, lanes finishing out of order, ordered reservation
// of the output byte ranges, ordered log lines, pwrite pool, error paths — can be exercised without a GPU.  The
// "lanes" compute their batch with the CPU oracle (oracle/hx_oracle.h: the checker), sleep for random times so that
// batches complete out of order, and hand the text to the sink in small chunks.
#include "../../hyperex_b200/csrc/hx_host.h"
#include "../../include/hyperex_b200.h"
#include "../../oracle/hx_oracle.h"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

struct hx_ctx {
    std::vector<std::string> fwd, rev;
    uint8_t k = 0;
    std::string err;
    std::mutex mu;
    void *host_cache = nullptr;
    void (*host_cache_free)(void *) = nullptr;
    int n_lanes = 4;
    int fail_at_batch = -1;  // the n-th lane call fails (error-path test)
    int calls = 0;
};

extern "C" {
hx_ctx *stub_ctx_new(int n_lanes, int fail_at_batch) {
    hx_ctx *c = new hx_ctx();
    c->n_lanes = n_lanes;
    c->fail_at_batch = fail_at_batch;
    return c;
}
void stub_ctx_free(hx_ctx *c) {
    if (c && c->host_cache && c->host_cache_free) c->host_cache_free(c->host_cache);
    delete c;
}
int hx_set_primers(hx_ctx *ctx, uint32_t n_pairs, const char *const *fwd, const uint32_t *, const char *const *rev,
                   const uint32_t *, uint8_t k) {
    ctx->fwd.assign(fwd, fwd + n_pairs);
    ctx->rev.assign(rev, rev + n_pairs);
    ctx->k = k;
    for (uint32_t i = 0; i < n_pairs; ++i)
        if (ctx->fwd[i].empty() || ctx->fwd[i].size() > 64 || ctx->rev[i].empty() || ctx->rev[i].size() > 64) return HX_ERR_ARG;
    return HX_OK;
}
void *hx_alloc_pinned(size_t n) { return malloc(n ? n : 1); }
void hx_free_pinned(void *p) { free(p); }
const char *hx_last_error(const hx_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
}

int hxh_device_count(const hx_ctx *) { return 1; }
int hxh_lane_count(const hx_ctx *ctx) { return ctx->n_lanes; }
void **hxh_host_cache(hx_ctx *ctx, void (*free_fn)(void *)) {
    ctx->host_cache_free = free_fn;
    return &ctx->host_cache;
}
void hxh_set_error(hx_ctx *ctx, const char *msg) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err = msg ? msg : "";
}

static int compute(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n, std::vector<hxo_result> &res) {
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (ctx->calls++ == ctx->fail_at_batch) {
            ctx->err = "stub: injected device failure";
            return HX_ERR_CUDA;
        }
    }
    std::vector<const char *> f, r;
    for (size_t i = 0; i < ctx->fwd.size(); ++i) {
        f.push_back(ctx->fwd[i].c_str());
        r.push_back(ctx->rev[i].c_str());
    }
    res.resize((size_t)n * f.size());
    // hxo_run_batch indexes the arena with the offsets as given
    if (hxo_run_batch(arena, offsets, n, (uint32_t)f.size(), f.data(), r.data(), ctx->k, res.data(), 1, 0) != 0) return HX_ERR_ARG;
    static thread_local std::mt19937 rng(std::random_device{}());
    std::this_thread::sleep_for(std::chrono::microseconds(rng() % 3000));
    return HX_OK;
}

int hxh_run_batch_lane(hx_ctx *ctx, int, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out) {
    std::vector<hxo_result> res;
    const int rc = compute(ctx, arena, offsets, n_records, res);
    if (rc != HX_OK) return rc;
    memcpy(out, res.data(), res.size() * sizeof(hx_result));
    return HX_OK;
}

int hxh_run_text_lane(hx_ctx *ctx, int, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out,
                      const char *ids, const uint64_t *id_off, const char *fa_tails, const uint32_t *fa_tail_off,
                      const char *gff_tails, const uint32_t *gff_tail_off, const HxhTextSink *sink) {
    std::vector<hxo_result> res;
    const int rc = compute(ctx, arena, offsets, n_records, res);
    if (rc != HX_OK) return rc;
    if (out) memcpy(out, res.data(), res.size() * sizeof(hx_result));
    // the text the device kernels assemble: id + per-pair tail + slice, in record x pair order
    const uint32_t np = (uint32_t)ctx->fwd.size();
    std::string text[2];
    for (uint32_t r = 0; r < n_records; ++r)
        for (uint32_t p = 0; p < np; ++p) {
            const hxo_result &x = res[(size_t)r * np + p];
            if ((x.flags & 7u) != 7u) continue;
            const std::string id(ids + id_off[r], ids + id_off[r + 1]);
            text[0] += ">" + id + std::string(fa_tails + fa_tail_off[p], fa_tails + fa_tail_off[p + 1]);
            text[0].append((const char *)arena + offsets[r] + x.f_start, (size_t)x.r_end + 1 - x.f_start);
            text[0] += "\n";
            text[1] += id + "\thyperex\tregion\t" + std::to_string(x.f_start + 1) + "\t" + std::to_string(x.r_end + 1) +
                       std::string(gff_tails + gff_tail_off[p], gff_tails + gff_tail_off[p + 1]);
        }
    if (sink->begin && sink->begin(sink->user, text[0].size(), text[1].size()) != 0) return HX_ERR_IO;
    static thread_local std::mt19937 rng(12345);
    for (int kind = 0; kind < 2; ++kind)
        for (size_t off = 0; off < text[kind].size();) {
            const size_t n = std::min<size_t>(text[kind].size() - off, 1 + rng() % 5000);
            if (sink->chunk(sink->user, kind, off, text[kind].data() + off, n) != 0) return HX_ERR_IO;
            off += n;
        }
    return HX_OK;
}
