// hx_api.cu — C ABI runtime of the B200-native HyperEx hot path (include/hyperex_b200.h).
//
// Host side of the device path: primer set -> pattern groups + IUPAC Peq tables
// (replaces MyersBuilder, utils.rs:251-268,299-302), and the batch pipeline that shards a
// host arena into slabs, copies them host->device on alternating streams, runs
// tile build -> classify -> Myers+pairing scan -> combine, and copies the result table back
// in input order (replaces the record loop utils.rs:270-351).  No CPU fallback exists: all
// compute entry points fail with HX_ERR_CUDA when CUDA is unusable.
#include "../../include/hyperex_b200.h"
#include "hx_device.h"
#include "hx_host.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

static_assert(sizeof(hx_result) == 12, "hx_result must be 12 bytes");

namespace {

constexpr int HX_NSLOT = 4;  // pipeline depth per device (H2D / scan / D2H overlap; the mixed path of hx_run_batch fills two at once)
constexpr int HX_NLANE = 2;  // text lanes per device (see TextLane)

std::string g_create_error;

// One device allocation per user (a slab slot, the extraction scratch, the output text), carved afresh for every
// slab: cudaMalloc / cudaFree cost tens of milliseconds each on a virtualised B200, so the number of calls must not
// scale with the number of buffers or with how often a size changes.
struct DevPool {
    uint8_t *base = nullptr;
    size_t cap = 0, used = 0;
    void reset() { used = 0; }
    size_t bump(size_t bytes) {
        const size_t off = (used + 255) & ~(size_t)255;
        used = off + bytes;
        return off;
    }
    bool fits() const { return used <= cap; }
    cudaError_t grow() {  // callers make sure no work that uses the old block is in flight
        if (base) {
            cudaError_t e = cudaFree(base);
            base = nullptr;
            cap = 0;
            if (e != cudaSuccess) return e;
        }
        const size_t want = used + used / 4 + ((size_t)1 << 20);
        cudaError_t e = cudaMalloc((void **)&base, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (base) cudaFree(base);
        base = nullptr;
        cap = used = 0;
    }
};

template <typename T>
struct DevBuf {  // a typed view into a DevPool; p is only meaningful once the pool fits()
    T *p = nullptr;
    void carve(DevPool &pool, size_t n) { p = (T *)(pool.base + pool.bump(n * sizeof(T))); }
};

struct Slot {
    cudaStream_t stream = nullptr;
    DevPool pool;
    DevBuf<uint8_t> arena;
    DevBuf<uint64_t> offsets;
    DevBuf<uint32_t> rec_u32;  // rec_raw | tile_first | tile_cnt | long_list   (4 x rec_cap)
    DevBuf<uint32_t> scratch;  // ntiles, overflow, n_long, n_plain, hist[NB], cursor[NB]
    DevBuf<HxTile> tiles;
    DevBuf<uint32_t> order;
    DevBuf<uint64_t> sum_f, sum_r, pair_hi;
    DevBuf<uint32_t> pair_lo;
    DevBuf<hx_result> out;
    DevBuf<uint8_t> flags;         // packed text: per-record alphabet evidence gathered by the host packer
    uint32_t *h_status = nullptr;  // pinned: [0] ntiles, [1] overflow
    cudaEvent_t done_ev = nullptr; // blocking-sync event after the slab's last copy (pack-on-the-way path: a spinning wait
                                   // would take a core from the packer)
    void release() {
        if (done_ev) cudaEventDestroy(done_ev);
        done_ev = nullptr;
        pool.release();
        if (h_status) cudaFreeHost(h_status);
        if (stream) cudaStreamDestroy(stream);
        h_status = nullptr;
        stream = nullptr;
    }
};

// A text lane = everything ONE in-flight batch of the text path (scan + on-device FASTA/GFF3 assembly) owns on a device:
// a slot, the extraction scratch, the output text block and the pinned staging of the streamed copy.  Two lanes per
// device let the kernels of one batch run under the device->host text copy of the previous one; a lane is driven by
// one host thread at a time (hx_get_hypervar_regions runs one thread per lane).
struct TextLane {
    Slot slot;
    DevPool ex_pool, ex_text_pool;
    DevBuf<uint8_t> ex_ids, ex_tails, ex_text, ex_gtails, ex_gtext;
    DevBuf<uint64_t> ex_id_off, ex_lens, ex_sums, ex_glens, ex_gsums;
    DevBuf<uint32_t> ex_tail_off, ex_gtail_off;
    uint8_t *ex_host = nullptr, *ex_ghost = nullptr;  // whole-text pinned buffers of hx_run_batch_text
    size_t ex_host_cap = 0, ex_ghost_cap = 0;
    uint8_t *ex_stage[2] = {nullptr, nullptr};  // pinned staging of the streamed text path
    cudaEvent_t ex_stage_ev[2] = {nullptr, nullptr};
    uint64_t *h_totals = nullptr;  // pinned: [0] FASTA bytes, [1] GFF3 bytes of the batch
    void release() {
        slot.release();
        ex_pool.release();
        ex_text_pool.release();
        if (ex_host) cudaFreeHost(ex_host);
        if (ex_ghost) cudaFreeHost(ex_ghost);
        if (h_totals) cudaFreeHost(h_totals);
        for (int i = 0; i < 2; ++i) {
            if (ex_stage[i]) cudaFreeHost(ex_stage[i]);
            if (ex_stage_ev[i]) cudaEventDestroy(ex_stage_ev[i]);
            ex_stage[i] = nullptr;
            ex_stage_ev[i] = nullptr;
        }
        ex_host = ex_ghost = nullptr;
        h_totals = nullptr;
        ex_host_cap = ex_ghost_cap = 0;
    }
};

// Pinned staging of hx_run_batch with pack_h2d.  The packed text goes through a RING of small piece buffers (a few MiB
// each): small enough to stay in the host's last-level cache between the packer's stores and the DMA engine's reads, so
// the packed bytes cost no DRAM traffic (whole-slab staging buffers were measured: the packer drops from 91 to 73
// Gbases/s while the copies of the slabs before it read host memory, profiles/r02_pack_h2d_auto.md, r02_pack_h2d_ab_ring.json).  The ring belongs to the
// context (portable pinned memory, any device copies from it); every device has one event per ring buffer, which marks
// the end of the host->device copy that read it.
constexpr int HX_NRING = 8;  // piece buffers per device of the context
struct PackRing {
    uint8_t *p = nullptr;       // n_buf buffers of buf_bytes
    size_t buf_bytes = 0;
    int n_buf = 0;
    uint8_t *flags = nullptr;   // rec_flags of the whole batch (final slab by slab)
    size_t flags_cap = 0;
    void release() {
        if (p) cudaFreeHost(p);
        if (flags) cudaFreeHost(flags);
        p = flags = nullptr;
        buf_bytes = flags_cap = 0;
        n_buf = 0;
    }
};

struct Device {
    int dev = 0;
    Slot slots[HX_NSLOT];
    std::vector<cudaEvent_t> ring_ev;  // per ring buffer of the context: end of this device's copy out of it
    std::vector<cudaEvent_t> link_ev;  // mixed path: ends of the ASCII chunk copies in flight
    TextLane lanes[HX_NLANE];
    uint8_t *d_tables = nullptr;
    uint8_t *d_tables_wm = nullptr;
    uint8_t *d_tables_pk = nullptr;  // -m 0 packed tables (two 16-bit automata per word)
    uint64_t *d_vtables = nullptr;  // 64-bit verification tables of the suffix-filtered primers
    HxGroup *d_groups = nullptr;
    HxPairRef *d_pairs = nullptr;
};

struct TimedLaunch {
    cudaEvent_t a, b;
    int dev;
};

}  // namespace

struct hx_ctx {
    std::vector<Device> devs;
    std::string err;
    // primer set
    uint32_t n_pairs = 0;
    uint8_t k = 0;
    uint32_t m_max = 0;
    uint32_t n_slots = 0;  // per-pattern summary slots over all groups
    std::vector<HxGroup> groups;
    std::vector<HxPairRef> pair_refs;
    std::vector<uint8_t> tables;
    bool primers_set = false;
    bool packable = true;  // every primer symbol is one of ACGTURYSWKMBDHVN (or lower case, which matches nothing)
    // options
    int64_t opt_tile_len = 0, opt_single_max = 0, opt_slab_bytes = 128ll << 20, opt_time = 0, opt_text_tma = 0, opt_fast_small_k = 1,
            opt_long_filter = 1, opt_group_np_max = 0, opt_pack16 = 1, opt_pack8 = 1, opt_pack_h2d = -1, opt_pack_threads = 0, opt_pack_piece_bytes = 4ll << 20;
    std::atomic<uint64_t> launches{0};
    std::mutex mu;  // guards err and timed: text lanes of one context run on several host threads
    DevPool find_pool;
    PackRing pack_ring;
    uint64_t batch_stats[7] = {0, 0, 0, 0, 0, 0, 0};  // hx_batch_stats
    // pack_h2d = automatic: the first six large batches alternate ASCII slabs / the mixed path, the last four of them
    // are timed, and the faster mode serves every later call (run_batch_impl)
    struct {
        int calls = 0, decided = -1;
        double rate[2] = {0, 0};  // bases per second of the timed ASCII / mixed call
    } auto_pack;
    // host-side state the file-level driver (hx_host.cpp) keeps across calls: pinned record batches etc.
    void *host_cache = nullptr;
    void (*host_cache_free)(void *) = nullptr;
    std::vector<TimedLaunch> timed;
    double myers_ms = 0;
    uint64_t myers_launches = 0;
};

namespace {

// The entry points select the context's devices on the calling thread; the caller's current device is put back on the
// way out (a host that also drives CUDA itself — torch, another library — must not find it changed).
// Only a device whose primary context is alive is restored: cudaSetDevice would otherwise CREATE a context (0.4 s and a
// few hundred MB) on a device the caller never used, e.g. device 0 for a context made on device 3 alone.
bool primary_ctx_active(int dev) {
    typedef int (*fn_t)(int, unsigned int *, int *);  // CUresult cuDevicePrimaryCtxGetState(CUdevice, unsigned *, int *)
    static const fn_t fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (fn_t)p;
    }();
    unsigned int flags = 0;
    int active = 0;
    return fn && fn(dev, &flags, &active) == 0 && active;
}
struct DeviceGuard {
    int dev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (dev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != dev && primary_ctx_active(dev)) cudaSetDevice(dev);
    }
};

int fail(hx_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->err = buf;
    } else {
        g_create_error = buf;
    }
    return code;
}

#define HX_CUDA(ctx, call)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? HX_ERR_NOMEM : HX_ERR_CUDA, "%s: %s",  \
                        #call, cudaGetErrorString(e__));                                                \
    } while (0)

// Top-aligned Peq word of `pat` for text byte b (IUPAC table of utils.rs:251-263 through
// hxh_symbol_matches; build_64 semantics: bit i set iff pattern[i] matches b).
uint64_t peq_word(const std::string &pat, uint8_t b, int bits) {
    uint64_t w = 0;
    const int m = (int)pat.size();
    for (int i = 0; i < m; ++i)
        if (hxh_symbol_matches((uint8_t)pat[i], b)) w |= 1ull << (i + bits - m);
    return w;
}

// All 256 rows of one pattern at once: rows[b] = Peq word of the text byte b AFTER case folding (utils.rs:295), row 0
// (the null byte) all-zero.  Walks the pattern once and sets the bit in the rows of the few bytes each symbol accepts
// (itself + its IUPAC expansion) instead of testing 256 bytes against every symbol: 65 535 pairs plan in seconds.
void peq_rows(const std::string &pat, int bits, uint64_t rows[256]) {
    struct Accepts {  // text bytes a pattern symbol accepts; built once (thread-safe function-local static)
        std::vector<uint8_t> v[256];
        Accepts() {
            for (int c = 0; c < 256; ++c)
                for (int b = 1; b < 256; ++b)
                    if (hxh_symbol_matches((uint8_t)c, (uint8_t)b)) v[c].push_back((uint8_t)b);
        }
    };
    static const Accepts acc;
    const std::vector<uint8_t> *accepts = acc.v;
    uint64_t up[256];
    memset(up, 0, sizeof up);
    const int m = (int)pat.size();
    for (int i = 0; i < m; ++i)
        for (uint8_t b : accepts[(uint8_t)pat[i]]) up[b] |= 1ull << (i + bits - m);
    rows[0] = 0;
    for (int b = 1; b < 256; ++b) rows[b] = up[(b >= 'a' && b <= 'z') ? b - 32 : b];
}

inline void hx_cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

// Like HX_CUDA, for loops that must not return while earlier asynchronous copies still use caller memory: records the
// error in `rc` and leaves the loop; the caller's epilogue then drains every stream it touched.
#define HX_CUDA_BREAK(ctx, rc, call)                                                                          \
    {                                                                                                         \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            rc = fail(ctx, e__ == cudaErrorMemoryAllocation ? HX_ERR_NOMEM : HX_ERR_CUDA, "%s: %s", #call,    \
                      cudaGetErrorString(e__));                                                               \
            break;                                                                                            \
        }                                                                                                     \
    }

struct PatternKey {
    std::string dna, rna;
    bool operator==(const PatternKey &o) const { return dna == o.dna && rna == o.rna; }
};

void sync_timed(hx_ctx *ctx) {
    std::vector<TimedLaunch> timed;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        timed.swap(ctx->timed);
    }
    double ms_sum = 0;
    uint64_t n_sum = 0;
    for (auto &t : timed) {
        cudaSetDevice(t.dev);
        cudaEventSynchronize(t.b);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            ms_sum += ms;
            n_sum += 1;
        }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->myers_ms += ms_sum;
    ctx->myers_launches += n_sum;
}

struct SlabPlan {
    uint32_t tile_len, single_max, warm, gran, tile_cap;
};

SlabPlan plan_slab(const hx_ctx *ctx, uint32_t n_records, uint64_t bytes) {
    SlabPlan p;
    const uint64_t target = 148ull * 1024;  // ~2 resident waves of scan threads
    uint64_t tile_len, single_max;
    if (n_records >= target / 4) {
        tile_len = 2048;
        single_max = 4096;
    } else {
        tile_len = ((bytes / target + 15) / 16) * 16;
        tile_len = std::min<uint64_t>(std::max<uint64_t>(tile_len, 256), 2048);
        single_max = tile_len + tile_len / 2;
    }
    if (ctx->opt_tile_len > 0) tile_len = (uint64_t)((ctx->opt_tile_len + 15) / 16 * 16);
    if (ctx->opt_single_max > 0) single_max = (uint64_t)ctx->opt_single_max;
    if (single_max < tile_len) single_max = tile_len;
    p.tile_len = (uint32_t)tile_len;
    p.single_max = (uint32_t)single_max;
    p.warm = ctx->m_max + ctx->k - 1;  // a hit with dist <= k spans <= m + k text bytes
    const uint64_t maxspan = std::max<uint64_t>(single_max, tile_len + 16) + p.warm;
    uint64_t gran = (maxspan + HX_NBUCKETS - 1) / HX_NBUCKETS;
    p.gran = (uint32_t)std::max<uint64_t>(16, (gran + 15) / 16 * 16);
    p.tile_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFF0ull, (uint64_t)n_records + bytes / tile_len + 1);
    return p;
}

// Carves the slot's device block for one slab (one cudaMalloc at most, and only when the slab needs more than any
// before it).  io_bytes > 0: the slot also holds the arena copy, the offsets and the result table (host-buffer
// entry points); 0: those live in caller-owned device memory (hx_run_batch_device).  The slot must be idle.
int prepare_slab(hx_ctx *ctx, Slot &S, const SlabPlan &pl, uint32_t n_records, uint64_t io_bytes, bool with_flags = false) {
    for (int pass = 0;; ++pass) {
        S.pool.reset();
        if (io_bytes) {
            S.arena.carve(S.pool, io_bytes);
            S.offsets.carve(S.pool, (size_t)n_records + 1);
            S.out.carve(S.pool, (size_t)n_records * ctx->n_pairs);
        }
        if (with_flags) S.flags.carve(S.pool, (size_t)n_records + 16);
        S.rec_u32.carve(S.pool, (size_t)n_records * 4);
        S.scratch.carve(S.pool, 4 + 2 * HX_NBUCKETS + 2);
        S.tiles.carve(S.pool, pl.tile_cap);
        S.order.carve(S.pool, pl.tile_cap);
        S.sum_f.carve(S.pool, (size_t)ctx->n_slots * pl.tile_cap);
        S.sum_r.carve(S.pool, (size_t)ctx->n_slots * pl.tile_cap);
        S.pair_hi.carve(S.pool, (size_t)ctx->n_pairs * pl.tile_cap);
        S.pair_lo.carve(S.pool, (size_t)ctx->n_pairs * pl.tile_cap);
        if (S.pool.fits()) return HX_OK;
        if (pass) return fail(ctx, HX_ERR_STATE, "prepare_slab: pool did not fit after growing");
        HX_CUDA(ctx, S.pool.grow());
    }
}

// Enqueue the whole device path for one slab on `st` (after prepare_slab).  d_out receives n_records x n_pairs results.
// d_flags != NULL: packed text (4-bit codes; offsets and off_base count symbols) with the host packer's per-record evidence.
int enqueue_slab(hx_ctx *ctx, Device &D, Slot &S, const SlabPlan &pl, const uint8_t *d_arena, const uint64_t *d_offsets,
                 uint64_t off_base, uint32_t n_records, hx_result *d_out, cudaStream_t st, const uint8_t *d_flags = nullptr) {
    if (n_records == 0) return HX_OK;
    const bool nib = d_flags != nullptr;

    uint32_t *rec_raw = S.rec_u32.p, *tile_first = rec_raw + n_records, *tile_cnt = tile_first + n_records;
    uint32_t *long_list = tile_cnt + n_records;
    uint32_t *ntiles = S.scratch.p, *overflow = ntiles + 1, *n_long = ntiles + 2, *hist = ntiles + 4, *cursor = hist + HX_NBUCKETS;
    HX_CUDA(ctx, cudaMemsetAsync(S.scratch.p, 0, (4 + 2 * HX_NBUCKETS + 2) * sizeof(uint32_t), st));
    HX_CUDA(ctx, cudaMemsetAsync(rec_raw, 0, (size_t)n_records * sizeof(uint32_t), st));

    HxBuildArgs b;
    b.offsets = d_offsets;
    b.n_records = n_records;
    b.single_max = pl.single_max;
    b.tile_len = pl.tile_len;
    b.warm = pl.warm;
    b.bucket_gran = pl.gran;
    b.tile_cap = pl.tile_cap;
    b.tiles = S.tiles.p;
    b.tile_first = tile_first;
    b.tile_cnt = tile_cnt;
    b.ntiles = ntiles;
    b.hist = hist;
    b.overflow = overflow;
    b.long_list = long_list;
    b.n_long = n_long;
    ctx->launches += hx_launch_tile_build(b, cursor, S.order.p, st);
    if (nib)  // the packer classified the records while it touched every byte anyway: no device pass over the text
        ctx->launches += hx_launch_flags(d_flags, rec_raw, n_records, st);
    else
        ctx->launches += hx_launch_classify(d_arena, d_offsets, off_base, S.tiles.p, ntiles, pl.tile_cap, rec_raw, tile_cnt,
                                            n_records, ntiles + 3, st);  // scratch[3] = n_plain (zeroed with the scratch block)

    for (size_t gi = 0; gi < ctx->groups.size(); ++gi) {
        const HxGroup &hg = ctx->groups[gi];
        HxScanArgs a;
        // the scan kernels fetch whole 32-byte sectors: hand them a sector-aligned base.  d_arena is 16-byte aligned, so
        // the bytes in front of it belong to the same sector (same allocation); off_base moves with it (wrapping
        // arithmetic: rec0 = offsets[r] - off_base), counted in symbols for packed text
        const uint64_t mis = (uint64_t)((uintptr_t)d_arena & 31u);
        a.arena = d_arena - mis;
        a.offsets = d_offsets;
        a.off_base = off_base - (nib ? 2 * mis : mis);
        a.tiles = S.tiles.p;
        a.order = S.order.p;
        a.ntiles = ntiles;
        a.rec_raw = rec_raw;
        a.tile_cnt = tile_cnt;
        a.out = d_out;
        a.n_pairs_total = ctx->n_pairs;
        a.tables = D.d_tables + hg.table_off;
        a.tables_wm = D.d_tables_wm + hg.table_off;
        a.vtables = D.d_vtables;
        // small-k fast path: identical hits, fewer ALU operations; 32-bit groups, vector-load text path
        // (measured: at -m 2 the fast path only pays for narrow groups; wide groups run out of registers)
        // ... and two <= 16-symbol automata per word when the planner found the group packable
        const bool fast_ok = ctx->opt_fast_small_k && !hg.word64 && (!ctx->opt_text_tma || nib) && ctx->k <= 2;
        a.tables_pk = D.d_tables_pk ? D.d_tables_pk + hg.pk_table_off : nullptr;
        a.packed = (fast_ok && ctx->opt_pack16 && hg.packed && D.d_tables_pk) ? 1 : 0;
        a.wm_k = (fast_ok && (a.packed || ctx->k <= 1 || hg.np <= 8)) ? (int)ctx->k : -1;
        a.group = D.d_groups + gi;
        a.sum_f = S.sum_f.p;
        a.sum_r = S.sum_r.p;
        a.pair_hi = S.pair_hi.p;
        a.pair_lo = S.pair_lo.p;
        a.tile_stride = pl.tile_cap;
        a.k = ctx->k;
        a.nib = nib ? 1 : 0;
        a.variant = (hg.longmask || a.packed || nib) ? 0 : (int)ctx->opt_text_tma;  // suffix-filtered groups: vector-load text path only
        a.two = 2;
        a.sched = cursor + HX_NBUCKETS;  // batch counters of the persistent scan grid (zeroed with the scratch block)
        TimedLaunch tl;
        if (ctx->opt_time) {
            tl.dev = D.dev;
            cudaEventCreate(&tl.a);
            cudaEventCreate(&tl.b);
            cudaEventRecord(tl.a, st);
        }
        int nl = hx_launch_scan(a, hg, pl.tile_cap, st);
        if (nl < 0) return fail(ctx, HX_ERR_STATE, "no scan kernel for group of %d patterns", hg.np);
        ctx->launches += nl;
        if (ctx->opt_time) {
            cudaEventRecord(tl.b, st);
            std::lock_guard<std::mutex> lk(ctx->mu);
            ctx->timed.push_back(tl);
        }
    }
    ctx->launches += hx_launch_combine(tile_first, tile_cnt, rec_raw, D.d_pairs, ctx->n_pairs, n_records, S.sum_f.p,
                                       S.sum_r.p, S.pair_hi.p, S.pair_lo.p, pl.tile_cap, long_list, n_long, d_out, st);
    HX_CUDA(ctx, cudaMemcpyAsync(S.h_status, S.scratch.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    HX_CUDA(ctx, cudaGetLastError());
    return HX_OK;
}

int check_status(hx_ctx *ctx, const Slot &S) {
    if (S.h_status[1] == 2) return fail(ctx, HX_ERR_ARG, "a record is longer than 2^32-257 bytes");
    if (S.h_status[1]) return fail(ctx, HX_ERR_CAPACITY, "internal tile capacity exceeded");
    return HX_OK;
}

// Host half of hx_set_primers (no CUDA): primer strings -> pattern groups, Peq tables, verification tables.
struct PlanOpts {
    int64_t fast_small_k, long_filter, group_np_max, pack16, pack8;
};
struct PrimerPlan {
    std::vector<HxGroup> groups;
    std::vector<HxPairRef> refs;
    std::vector<uint8_t> blob, blob_wm, blob_pk;
    std::vector<uint64_t> vblob;  // [vslot][alphabet][256]
    uint32_t m_max = 0, n_slots = 0;
};

int pfail(std::string &err, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
}

int plan_primers(const PlanOpts &o, uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len,
                 const char *const *rev, const uint32_t *rev_len, uint8_t k, PrimerPlan &P, std::string &err) {
    if (n_pairs == 0 || n_pairs > 65535 || !fwd || !rev)
        return pfail(err, HX_ERR_ARG, "hx_set_primers: need 1..65535 primer pairs");
    std::vector<std::string> F(n_pairs), R(n_pairs);
    for (uint32_t i = 0; i < n_pairs; ++i) {
        if (!fwd[i] || !rev[i]) return pfail(err, HX_ERR_ARG, "hx_set_primers: NULL primer in pair %u", i);
        const size_t lf = fwd_len ? fwd_len[i] : strlen(fwd[i]);
        const size_t lr = rev_len ? rev_len[i] : strlen(rev[i]);
        // rust-bio build_64 asserts 0 < m <= 64 (call sites utils.rs:301-302)
        if (lf == 0 || lf > 64 || lr == 0 || lr > 64)
            return pfail(err, HX_ERR_ARG, "hx_set_primers: pair %u: primer length must be 1..64 (got %zu, %zu)", i, lf,
                        lr);
        F[i].assign(fwd[i], lf);
        R[i].assign(rev[i], lr);
        if (memchr(F[i].data(), 0, lf) || memchr(R[i].data(), 0, lr))
            return pfail(err, HX_ERR_ARG, "hx_set_primers: pair %u: NUL byte inside a primer", i);
    }
    // unique patterns per pair: forward as given; reverse = reverse complement per alphabet (utils.rs:299)
    struct PairPat { PatternKey f, r; bool is_long; int cls; };  // cls: 0 packable 32-bit, 1 other 32-bit, 2 64-bit
    std::vector<PairPat> pp(n_pairs);
    uint32_t m_max = 0;
    // Suffix filter (hx_kernels.cu: hx_verify_long): primers of 33..64 symbols are scanned as the 32-bit automaton of
    // their last 32 symbols and every candidate is verified against the whole primer, so no 64-bit group is needed.
    // Only for small k: with k close to 32 every position becomes a candidate.
    const bool filter = o.long_filter && k <= 8;
    // ... and only for a suffix that is specific enough: the expected candidate rate on random ACGT text,
    // 2^-bits(suffix) x (number of k-edit neighbourhoods of a 32-mer), must stay below 1e-3 per position, or the
    // verification (m + k 64-bit steps per candidate) would cost more than stepping the 64-bit automaton.  A primer
    // whose last 32 symbols are mostly N keeps its 64-bit group.  long_filter = 2 skips this test (test-suite).
    auto suffix_ok = [&](const std::string &pat) {
        if (pat.size() <= 32 || o.long_filter >= 2) return true;
        double bits = 0;
        for (size_t i = pat.size() - 32; i < pat.size(); ++i) {
            const char c = pat[i];
            const int deg = strchr("RYSWKM", c) ? 2 : strchr("BDHV", c) ? 3 : c == 'N' ? 4 : 1;
            bits += std::log2(4.0 / deg);
        }
        double nb = 20.0;  // indels on top of substitutions
        for (int i = 0; i < (int)k; ++i) nb *= (32.0 - i) / (i + 1) * 3.0;
        return bits >= std::log2(nb / 1e-3);
    };
    for (uint32_t i = 0; i < n_pairs; ++i) {
        pp[i].f = PatternKey{F[i], F[i]};
        std::string rd(R[i].size(), '\0'), rr(R[i].size(), '\0');
        hx_to_reverse_complement(R[i].data(), R[i].size(), 0, &rd[0]);
        hx_to_reverse_complement(R[i].data(), R[i].size(), 1, &rr[0]);
        pp[i].r = PatternKey{rd, rr};
        pp[i].is_long = (F[i].size() > 32 || R[i].size() > 32) && !(filter && suffix_ok(F[i]) && suffix_ok(rd));
        m_max = std::max<uint32_t>(m_max, (uint32_t)std::max(F[i].size(), R[i].size()));
    }
    // Small-k pattern packing: on the Shift-Or / Wu-Manber path (-m 0..2) two automata of <= 16 symbols
    // share a 32-bit word.  A primer longer than 16 symbols is scanned as its last 16 symbols and every candidate end is
    // verified against the whole primer in the drain (exact match of the rest at -m 0, hx_verify_long otherwise), so the
    // suffix must be specific enough: the expected rate of chance candidates on random ACGT text, 2^-bits(suffix) x
    // (number of k-edit neighbourhoods of a 16-mer), stays below 1e-3 per position — every built-in primer qualifies at
    // -m 0..2 (pack16 = 2 skips the test: test-suite).
    auto suffix16_ok = [&](const std::string &pat) {
        if (pat.size() <= 16 || o.pack16 >= 2) return true;
        double bits = 0;
        for (size_t i = pat.size() - 16; i < pat.size(); ++i) {
            const char c = pat[i];
            const int deg = strchr("RYSWKM", c) ? 2 : strchr("BDHV", c) ? 3 : c == 'N' ? 4 : 1;
            bits += std::log2(4.0 / deg);
        }
        double nb = k ? 20.0 : 1.0;  // indels on top of substitutions
        for (int i = 0; i < (int)k; ++i) nb *= (16.0 - i) / (i + 1) * 3.0;
        return bits >= std::log2(nb / 1e-3);
    };
    // Pairs whose two patterns qualify form their own grouping class, so one vague primer in a CSV does not unpack the rest.
    const bool packing = k <= 2 && o.fast_small_k && o.pack16;
    for (uint32_t i = 0; i < n_pairs; ++i) {
        const bool ok16 = packing && suffix16_ok(pp[i].f.dna) && suffix16_ok(pp[i].f.rna) && suffix16_ok(pp[i].r.dna) && suffix16_ok(pp[i].r.rna);
        pp[i].cls = pp[i].is_long ? 2 : (ok16 ? 0 : 1);
    }
    // Grouping, short (32-bit words) and long (64-bit words) pairs separately.  Pairs that share a pattern (27F and
    // 1492Rmod serve three built-in regions each) should land in the same group, or the shared automaton is stepped
    // once per group: pairs are first joined into the connected components of the "shares a pattern" relation, whole
    // components are packed first-fit-decreasing into groups of <= np_max patterns and <= HX_MAX_GROUP_PAIRS pairs,
    // and only a component that does not fit a group by itself is cut (in pair order, duplicating what it must).
    std::vector<HxGroup> &groups = P.groups;
    std::vector<std::vector<PatternKey>> gpats;
    std::vector<HxPairRef> &refs = P.refs;
    groups.clear();
    refs.assign(n_pairs, HxPairRef{0, 0, 0, 0});
    std::map<std::pair<std::string, std::string>, int> pat_id;
    std::vector<int> fid(n_pairs), rid(n_pairs);
    for (uint32_t i = 0; i < n_pairs; ++i) {
        fid[i] = pat_id.emplace(std::make_pair(pp[i].f.dna, pp[i].f.rna), (int)pat_id.size()).first->second;
        rid[i] = pat_id.emplace(std::make_pair(pp[i].r.dna, pp[i].r.rna), (int)pat_id.size()).first->second;
    }
    for (int pass = 0; pass < 3; ++pass) {
        const bool want_long = pass == 2;
        int np_max = want_long ? HX_MAX_NP64 : HX_MAX_NP32;
        // -m 2 on the Wu-Manber path: three bit-vectors per automaton, so only groups of <= 8 patterns stay in
        // registers (measured per 1.5 Gbases: config 5 108.7 -> 84.8 ms, built-in regions as 8 + 7 automata 14.3 -> 11.4 ms)
        // (packed, 16 automata are 8 words: the same 24 registers of state)
        if (pass == 1 && k == 2 && o.fast_small_k) np_max = 8;
        if (o.group_np_max >= 2) np_max = std::min<int>(np_max, (int)o.group_np_max);
        // connected components over the pairs of this class (union-find on pattern ids)
        std::vector<int> parent(pat_id.size());
        for (size_t j = 0; j < parent.size(); ++j) parent[j] = (int)j;
        auto root = [&](int x) {
            while (parent[x] != x) x = parent[x] = parent[parent[x]];
            return x;
        };
        for (uint32_t i = 0; i < n_pairs; ++i)
            if (pp[i].cls == pass) parent[root(fid[i])] = root(rid[i]);
        std::map<int, std::vector<uint32_t>> comps;  // root -> pairs, ascending
        for (uint32_t i = 0; i < n_pairs; ++i)
            if (pp[i].cls == pass) comps[root(fid[i])].push_back(i);
        // items = whole components, or the pair-order pieces of one that is too large for a group
        struct Item { std::vector<uint32_t> pairs; std::vector<int> pats; uint32_t first; };
        std::vector<Item> items;
        for (auto &c : comps) {
            Item it;
            it.first = c.second[0];
            for (uint32_t i : c.second) {
                const int need = (std::find(it.pats.begin(), it.pats.end(), fid[i]) == it.pats.end() ? 1 : 0) +
                                 ((rid[i] != fid[i] && std::find(it.pats.begin(), it.pats.end(), rid[i]) == it.pats.end()) ? 1 : 0);
                if (!it.pairs.empty() && ((int)it.pats.size() + need > np_max || it.pairs.size() >= HX_MAX_GROUP_PAIRS)) {
                    items.push_back(it);
                    it = Item();
                    it.first = i;
                }
                if (std::find(it.pats.begin(), it.pats.end(), fid[i]) == it.pats.end()) it.pats.push_back(fid[i]);
                if (std::find(it.pats.begin(), it.pats.end(), rid[i]) == it.pats.end()) it.pats.push_back(rid[i]);
                it.pairs.push_back(i);
            }
            items.push_back(it);
        }
        std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) {
            return x.pats.size() != y.pats.size() ? x.pats.size() > y.pats.size() : x.first < y.first;
        });
        struct Bin { std::vector<uint32_t> pairs; size_t npat; uint32_t first; };
        std::vector<Bin> bins;
        size_t first_open = 0;
        for (const Item &it : items) {
            size_t b = first_open;
            for (; b < bins.size(); ++b)
                if (bins[b].npat + it.pats.size() <= (size_t)np_max && bins[b].pairs.size() + it.pairs.size() <= HX_MAX_GROUP_PAIRS)
                    break;
            if (b == bins.size()) bins.push_back(Bin{{}, 0, it.first});
            bins[b].pairs.insert(bins[b].pairs.end(), it.pairs.begin(), it.pairs.end());
            bins[b].npat += it.pats.size();
            bins[b].first = std::min(bins[b].first, it.first);
            while (first_open < bins.size() && (bins[first_open].npat >= (size_t)np_max || bins[first_open].pairs.size() >= HX_MAX_GROUP_PAIRS ||
                                                bins.size() - first_open > 64))
                ++first_open;  // full, or too far back to be worth searching
        }
        std::sort(bins.begin(), bins.end(), [](const Bin &x, const Bin &y) { return x.first < y.first; });
        for (Bin &bin : bins) {
            std::sort(bin.pairs.begin(), bin.pairs.end());
            HxGroup g;
            memset(&g, 0, sizeof g);
            g.word64 = want_long ? 1 : 0;
            g.packed = pass == 0 ? 1 : 0;
            groups.push_back(g);
            gpats.emplace_back();
            const int cur = (int)groups.size() - 1;
            auto slot_of = [&](const PatternKey &key) -> int {
                for (size_t j = 0; j < gpats[cur].size(); ++j)
                    if (gpats[cur][j] == key) return (int)j;
                gpats[cur].push_back(key);
                return (int)gpats[cur].size() - 1;
            };
            for (uint32_t i : bin.pairs) {
                const int fi = slot_of(pp[i].f), ri = slot_of(pp[i].r);
                HxGroup &gg = groups[cur];
                gg.pair_f[gg.npairs] = (uint8_t)fi;
                gg.pair_r[gg.npairs] = (uint8_t)ri;
                gg.pair_global[gg.npairs] = (uint16_t)i;
                gg.rmask[ri] |= (uint16_t)(1u << gg.npairs);
                gg.npairs++;
                refs[i].len_f = (uint32_t)F[i].size();  // utils.rs:327 forward_len
                refs[i].pad = (uint32_t)cur;            // group index for now, slots resolved below
                refs[i].f_slot = (uint32_t)fi;
                refs[i].r_slot = (uint32_t)ri;
            }
            if ((int)gpats[cur].size() > np_max || groups[cur].npairs > HX_MAX_GROUP_PAIRS)
                return pfail(err, HX_ERR_STATE, "hx_set_primers: internal grouping error");
        }
    }
    // tables: [group][alphabet dna|rna][256 rows][row stride]; text bytes are case-folded
    // here (utils.rs:295) so the kernel indexes rows with the raw byte.
    std::vector<uint8_t> &blob = P.blob, &blob_wm = P.blob_wm;
    std::vector<uint64_t> &vblob = P.vblob;
    blob.clear();
    blob_wm.clear();
    vblob.clear();
    uint32_t slot_base = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        HxGroup &g = groups[gi];
        g.np = (int)gpats[gi].size();
        g.slot_base = slot_base;
        slot_base += (uint32_t)g.np;
        const int wb = g.word64 ? 8 : 4, bits = wb * 8;
        const int stride = hx_row_stride(g.np, wb);
        g.table_off = (uint32_t)blob.size();
        blob.resize(blob.size() + 2 * (size_t)hx_table_bytes(g.np, wb), 0);
        blob_wm.resize(blob.size(), 0);
        // what the group's automata scan: the primer, or its last `bits` symbols when it is longer than the word
        std::vector<PatternKey> scan(gpats[gi]);
        for (int p = 0; p < g.np; ++p) {
            const size_t full = gpats[gi][p].dna.size();
            g.mfull[p] = (uint8_t)full;
            if (full > (size_t)bits) {
                scan[p].dna = scan[p].dna.substr(full - bits);
                scan[p].rna = scan[p].rna.substr(full - bits);
                g.longmask |= 1u << p;
                g.vslot[p] = (uint32_t)(vblob.size() / 512);
                for (int alpha = 0; alpha < 2; ++alpha) {
                    uint64_t rows[256];
                    peq_rows(alpha ? gpats[gi][p].rna : gpats[gi][p].dna, 64, rows);
                    vblob.insert(vblob.end(), rows, rows + 256);
                }
            }
            g.m[p] = (uint8_t)scan[p].dna.size();
        }
        for (int alpha = 0; alpha < 2; ++alpha) {
            uint8_t *tab = blob.data() + g.table_off + (size_t)alpha * hx_table_bytes(g.np, wb);
            for (int p = 0; p < g.np; ++p) {
                uint64_t rows[256];
                peq_rows(alpha ? scan[p].rna : scan[p].dna, bits, rows);
                for (int b = 1; b < 256; ++b)  // row 0 stays all-zero: the "null" byte
                    memcpy(tab + (size_t)b * stride + (size_t)p * wb, &rows[b], wb);  // little endian: low bytes first
            }
            // fast-path table: complemented inside the pattern bits, 0 in the unused low bits; row 0 (the
            // null byte) matches nothing, i.e. all pattern bits set
            uint8_t *tabw = blob_wm.data() + g.table_off + (size_t)alpha * hx_table_bytes(g.np, wb);
            for (int b = 0; b < 256; ++b) {
                for (int p = 0; p < g.np; ++p) {
                    const int m = g.m[p];
                    const uint64_t pmask = m >= bits ? ~0ull : (((1ull << m) - 1) << (bits - m));
                    uint64_t w = 0;
                    memcpy(&w, tab + (size_t)b * stride + (size_t)p * wb, wb);
                    const uint64_t nw = ~w & pmask & (wb == 8 ? ~0ull : 0xffffffffull);
                    memcpy(tabw + (size_t)b * stride + (size_t)p * wb, &nw, wb);
                }
            }
        }
    }
    // packed tables of the groups of class 0: two 16-bit automata per word, pattern 2j in the low half — or, at -m 0,
    // four 8-bit automata per word (pattern 4j + h in byte h) when the last 8 symbols of every pattern of the group are
    // specific enough by themselves: >= 12 bits, i.e. a chance candidate every 4096 positions of random ACGT text at
    // most (a pure 8-mer has 16; the built-in primers have 12.8..16).  Half the words again: 15 automata are 4 words and
    // ONE 16-byte Peq row per text byte.  Candidates are verified against the rest of the primer in the drain.
    auto suffix8_ok = [&](const std::string &pat) {
        if (o.pack16 >= 3) return true;  // test-suite: force
        double bits = 0;
        for (size_t i = pat.size() > 8 ? pat.size() - 8 : 0; i < pat.size(); ++i) {
            const char c = pat[i];
            const int deg = strchr("RYSWKM", c) ? 2 : strchr("BDHV", c) ? 3 : c == 'N' ? 4 : 1;
            bits += std::log2(4.0 / deg);
        }
        return pat.size() <= 8 || bits >= 12.0;
    };
    P.blob_pk.clear();
    {
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            HxGroup &g = groups[gi];
            if (!g.packed) continue;
            g.pk_longmask = 0;
            bool quarters = k == 0 && o.pack8;
            for (int p = 0; p < g.np && quarters; ++p) quarters = suffix8_ok(gpats[gi][p].dna) && suffix8_ok(gpats[gi][p].rna);
            const int parts = quarters ? 4 : 2, pb = 32 / parts;
            g.pk_parts = parts;
            const int nw = (g.np + parts - 1) / parts, stride = hx_row_stride(nw, 4);
            g.pk_table_off = (uint32_t)P.blob_pk.size();
            P.blob_pk.resize(P.blob_pk.size() + 2 * (size_t)hx_table_bytes(nw, 4), 0);
            for (int p = 0; p < g.np; ++p) {
                const size_t full = gpats[gi][p].dna.size();
                g.pk_m[p] = (uint8_t)std::min<size_t>(full, (size_t)pb);
                if (full > (size_t)pb) {
                    g.pk_longmask |= 1u << p;
                    if (g.longmask >> p & 1u) {
                        g.pk_vslot[p] = g.vslot[p];  // the 64-bit table of the whole primer exists already
                    } else {
                        g.pk_vslot[p] = (uint32_t)(vblob.size() / 512);
                        for (int alpha = 0; alpha < 2; ++alpha) {
                            uint64_t rows[256];
                            peq_rows(alpha ? gpats[gi][p].rna : gpats[gi][p].dna, 64, rows);
                            vblob.insert(vblob.end(), rows, rows + 256);
                        }
                    }
                }
            }
            for (int alpha = 0; alpha < 2; ++alpha) {
                uint8_t *tab = P.blob_pk.data() + g.pk_table_off + (size_t)alpha * hx_table_bytes(nw, 4);
                const uint32_t part_all = parts == 4 ? 0xffu : 0xffffu;
                for (int j = 0; j < nw; ++j) {
                    uint64_t rows[4][256];
                    uint32_t pmask[4] = {part_all, part_all, part_all, part_all};  // a part without a pattern: all-ones in every row, never reports
                    for (int h = 0; h < parts; ++h) {
                        const int p = parts * j + h;
                        if (p >= g.np) {
                            for (int b = 0; b < 256; ++b) rows[h][b] = 0;
                            continue;
                        }
                        const std::string &full = alpha ? gpats[gi][p].rna : gpats[gi][p].dna;
                        const std::string suf = full.size() > (size_t)pb ? full.substr(full.size() - pb) : full;
                        peq_rows(suf, pb, rows[h]);  // top-aligned in pb bits
                        pmask[h] = (part_all << (pb - (int)suf.size())) & part_all;
                    }
                    for (int b = 0; b < 256; ++b) {  // complemented inside the pattern bits, 0 in the unused low bits
                        uint32_t w = 0;
                        for (int h = 0; h < parts; ++h) w |= (~(uint32_t)rows[h][b] & pmask[h]) << (pb * h);
                        memcpy(tab + (size_t)b * stride + (size_t)j * 4, &w, 4);
                    }
                }
            }
        }
    }
    for (uint32_t i = 0; i < n_pairs; ++i) {
        const HxGroup &g = groups[refs[i].pad];
        refs[i].f_slot += g.slot_base;
        refs[i].r_slot += g.slot_base;
        refs[i].pad = 0;
    }
    P.m_max = m_max;
    P.n_slots = slot_base;
    return HX_OK;
}

}  // namespace

extern "C" {

int hx_abi_version(void) { return HX_ABI_VERSION; }

const char *hx_last_error(const hx_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hx_create(hx_ctx **out, const int *dev_ids, int n_dev) {
    DeviceGuard device_guard;
    if (!out || n_dev < 1) return fail(nullptr, HX_ERR_ARG, "hx_create: bad arguments");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, HX_ERR_CUDA, "hx_create: no CUDA device (%s); there is no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    hx_ctx *ctx = new (std::nothrow) hx_ctx();
    if (!ctx) return fail(nullptr, HX_ERR_NOMEM, "hx_create: out of memory");
    ctx->devs.resize(n_dev);
    for (int i = 0; i < n_dev; ++i) {
        Device &D = ctx->devs[i];
        D.dev = dev_ids ? dev_ids[i] : i;
        if (D.dev < 0 || D.dev >= count) {
            fail(nullptr, HX_ERR_ARG, "hx_create: device %d not present (count %d)", D.dev, count);
            hx_destroy(ctx);
            return HX_ERR_ARG;
        }
    }
    // Per-device state on parallel host threads: the first call on a device creates its CUDA primary context, which
    // takes 0.4 s and more per device on a virtualised box and used to run device after device (8 GPUs: seconds).
    std::vector<cudaError_t> dev_err(n_dev, cudaSuccess);
    auto init_dev = [&](int i) {
        Device &D = ctx->devs[i];
        cudaError_t ce = cudaSetDevice(D.dev);
        auto init_slot = [&](Slot &S) {
            if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking);
            if (ce == cudaSuccess) ce = cudaHostAlloc((void **)&S.h_status, 16, cudaHostAllocDefault);
            if (ce == cudaSuccess) memset(S.h_status, 0, 16);
        };
        for (int s = 0; s < HX_NSLOT; ++s) init_slot(D.slots[s]);
        for (int l = 0; l < HX_NLANE; ++l) {
            init_slot(D.lanes[l].slot);
            if (ce == cudaSuccess) ce = cudaHostAlloc((void **)&D.lanes[l].h_totals, 16, cudaHostAllocDefault);
        }
        dev_err[i] = ce;
    };
    if (n_dev == 1) {
        init_dev(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_dev; ++i) th.emplace_back(init_dev, i);
        for (std::thread &t : th) t.join();
    }
    for (int i = 0; i < n_dev; ++i)
        if (dev_err[i] != cudaSuccess) {
            fail(nullptr, HX_ERR_CUDA, "hx_create: device %d: %s", ctx->devs[i].dev, cudaGetErrorString(dev_err[i]));
            hx_destroy(ctx);
            return HX_ERR_CUDA;
        }
    *out = ctx;
    return HX_OK;
}

void hx_destroy(hx_ctx *ctx) {
    DeviceGuard device_guard;
    if (!ctx) return;
    sync_timed(ctx);
    if (ctx->host_cache && ctx->host_cache_free) ctx->host_cache_free(ctx->host_cache);  // pinned memory: before the devices go
    ctx->host_cache = nullptr;
    if (!ctx->devs.empty()) cudaSetDevice(ctx->devs[0].dev);
    ctx->find_pool.release();
    ctx->pack_ring.release();
    for (Device &D : ctx->devs) {
        cudaSetDevice(D.dev);
        cudaDeviceSynchronize();
        for (Slot &S : D.slots) S.release();
        for (TextLane &L : D.lanes) L.release();
        for (cudaEvent_t e : D.ring_ev)
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : D.link_ev)
            if (e) cudaEventDestroy(e);
        if (D.d_tables) cudaFree(D.d_tables);
        if (D.d_tables_wm) cudaFree(D.d_tables_wm);
        if (D.d_tables_pk) cudaFree(D.d_tables_pk);
        if (D.d_vtables) cudaFree(D.d_vtables);
        if (D.d_groups) cudaFree(D.d_groups);
        if (D.d_pairs) cudaFree(D.d_pairs);
    }
    delete ctx;
}

int hx_set_option(hx_ctx *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return HX_ERR_ARG;
    const std::string n(name);
    if (n == "tile_len") ctx->opt_tile_len = value;
    else if (n == "single_max") ctx->opt_single_max = value;
    else if (n == "slab_bytes") ctx->opt_slab_bytes = std::max<int64_t>(value, 1);
    else if (n == "time_kernels") ctx->opt_time = value;
    else if (n == "text_tma") ctx->opt_text_tma = value ? 1 : 0;
    else if (n == "fast_small_k") ctx->opt_fast_small_k = value ? 1 : 0;
    else if (n == "long_filter") ctx->opt_long_filter = value < 0 ? 0 : (value > 2 ? 2 : value);  // read by hx_set_primers
    else if (n == "group_np_max") ctx->opt_group_np_max = value;         // read by hx_set_primers; 0 = automatic
    else if (n == "pack16") ctx->opt_pack16 = value < 0 ? 0 : (value > 3 ? 3 : value);  // hx_set_primers builds, every scan reads
    else if (n == "pack8") ctx->opt_pack8 = value ? 1 : 0;
    else if (n == "pack_h2d") {
        ctx->opt_pack_h2d = value < 0 ? -1 : (value >= 2 ? 2 : value ? 1 : 0);
        ctx->auto_pack = {};  // automatic: measure again
    } else if (n == "pack_threads") {
        ctx->opt_pack_threads = std::max<int64_t>(value, 0);
        ctx->auto_pack = {};
    }
    else if (n == "pack_piece_bytes") ctx->opt_pack_piece_bytes = std::min<int64_t>(std::max<int64_t>(value, 4096), 256ll << 20) & ~(int64_t)63;
    else return fail(ctx, HX_ERR_ARG, "unknown option %s", name);
    return HX_OK;
}

int hx_set_primers(hx_ctx *ctx, uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len,
                   const char *const *rev, const uint32_t *rev_len, uint8_t k) {
    DeviceGuard device_guard;
    if (!ctx) return HX_ERR_ARG;
    ctx->primers_set = false;
    PrimerPlan P;
    {
        const PlanOpts o = {ctx->opt_fast_small_k, ctx->opt_long_filter, ctx->opt_group_np_max, ctx->opt_pack16, ctx->opt_pack8};
        std::string err;
        const int rc = plan_primers(o, n_pairs, fwd, fwd_len, rev, rev_len, k, P, err);
        if (rc != HX_OK) return fail(ctx, rc, "%s", err.c_str());
    }
    std::vector<HxGroup> &groups = P.groups;
    std::vector<HxPairRef> &refs = P.refs;
    std::vector<uint8_t> &blob = P.blob, &blob_wm = P.blob_wm;
    std::vector<uint64_t> &vblob = P.vblob;
    const uint32_t m_max = P.m_max, slot_base = P.n_slots;
    // upload to every device
    for (Device &D : ctx->devs) {
        HX_CUDA(ctx, cudaSetDevice(D.dev));
        HX_CUDA(ctx, cudaDeviceSynchronize());
        if (D.d_tables) cudaFree(D.d_tables);
        if (D.d_tables_wm) cudaFree(D.d_tables_wm);
        if (D.d_tables_pk) cudaFree(D.d_tables_pk);
        if (D.d_vtables) cudaFree(D.d_vtables);
        if (D.d_groups) cudaFree(D.d_groups);
        if (D.d_pairs) cudaFree(D.d_pairs);
        D.d_tables = nullptr; D.d_tables_wm = nullptr; D.d_tables_pk = nullptr; D.d_vtables = nullptr; D.d_groups = nullptr; D.d_pairs = nullptr;
        if (!P.blob_pk.empty()) {
            HX_CUDA(ctx, cudaMalloc((void **)&D.d_tables_pk, P.blob_pk.size()));
            HX_CUDA(ctx, cudaMemcpy(D.d_tables_pk, P.blob_pk.data(), P.blob_pk.size(), cudaMemcpyHostToDevice));
        }
        if (!vblob.empty()) {
            HX_CUDA(ctx, cudaMalloc((void **)&D.d_vtables, vblob.size() * sizeof(uint64_t)));
            HX_CUDA(ctx, cudaMemcpy(D.d_vtables, vblob.data(), vblob.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        }
        HX_CUDA(ctx, cudaMalloc((void **)&D.d_tables, blob.size()));
        HX_CUDA(ctx, cudaMalloc((void **)&D.d_tables_wm, blob_wm.size()));
        HX_CUDA(ctx, cudaMemcpy(D.d_tables_wm, blob_wm.data(), blob_wm.size(), cudaMemcpyHostToDevice));
        HX_CUDA(ctx, cudaMalloc((void **)&D.d_groups, groups.size() * sizeof(HxGroup)));
        HX_CUDA(ctx, cudaMalloc((void **)&D.d_pairs, refs.size() * sizeof(HxPairRef)));
        HX_CUDA(ctx, cudaMemcpy(D.d_tables, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        HX_CUDA(ctx, cudaMemcpy(D.d_groups, groups.data(), groups.size() * sizeof(HxGroup), cudaMemcpyHostToDevice));
        HX_CUDA(ctx, cudaMemcpy(D.d_pairs, refs.data(), refs.size() * sizeof(HxPairRef), cudaMemcpyHostToDevice));
    }
    ctx->packable = true;
    for (uint32_t i = 0; i < n_pairs; ++i)
        for (const char *str : {fwd[i], rev[i]}) {
            const size_t len = (str == fwd[i] ? fwd_len : rev_len) ? (str == fwd[i] ? fwd_len[i] : rev_len[i]) : strlen(str);
            for (size_t j = 0; j < len; ++j)
                if (!strchr("ACGTURYSWKMBDHVN", str[j]) && !(str[j] >= 'a' && str[j] <= 'z')) ctx->packable = false;
        }
    ctx->n_pairs = n_pairs;
    ctx->k = k > 64 ? 64 : k;  // dist <= m <= 64: any larger bound accepts exactly the same hits
    ctx->m_max = m_max;
    ctx->n_slots = slot_base;
    ctx->groups.swap(groups);
    ctx->pair_refs.swap(refs);
    ctx->tables.swap(blob);
    ctx->primers_set = true;
    // what pack_h2d = automatic measured belongs to the primer set it was measured with: a set that makes the call
    // link-bound wants the mixed path, one that makes it kernel-bound does not (profiles/r02_mixed_slabs.md)
    ctx->auto_pack = {};
    return HX_OK;
}

int hx_run_batch_device(hx_ctx *ctx, int dev_slot, const uint8_t *d_arena, const uint64_t *d_offsets,
                        uint32_t n_records, uint64_t arena_bytes, hx_result *d_out, void *stream) {
    DeviceGuard device_guard;
    if (!ctx) return HX_ERR_ARG;
    if (!ctx->primers_set) return fail(ctx, HX_ERR_STATE, "hx_run_batch_device: call hx_set_primers first");
    if (dev_slot < 0 || dev_slot >= (int)ctx->devs.size() || (n_records && (!d_arena || !d_offsets || !d_out)))
        return fail(ctx, HX_ERR_ARG, "hx_run_batch_device: bad arguments");
    if (((uintptr_t)d_arena & 15u) != 0) return fail(ctx, HX_ERR_ARG, "hx_run_batch_device: d_arena must be 16-byte aligned");
    Device &D = ctx->devs[dev_slot];
    HX_CUDA(ctx, cudaSetDevice(D.dev));
    Slot &S = D.slots[0];
    cudaStream_t st = stream ? (cudaStream_t)stream : S.stream;
    const SlabPlan pl = plan_slab(ctx, n_records, arena_bytes);
    int rc = prepare_slab(ctx, S, pl, n_records, 0);
    if (rc == HX_OK) rc = enqueue_slab(ctx, D, S, pl, d_arena, d_offsets, 0, n_records, d_out, st);
    if (rc != HX_OK) return rc;
    if (!stream) {
        HX_CUDA(ctx, cudaStreamSynchronize(st));
        if (n_records) {
            rc = check_status(ctx, S);
            if (rc != HX_OK) return rc;
        }
        sync_timed(ctx);
    }
    return HX_OK;
}

// ---- the text path: scan + on-device FASTA / GFF3 assembly on one text lane ------------------------------------------
// Streams device text to the sink through the lane's two pinned staging buffers: the copy of chunk i + 1 is in
// flight while the sink consumes chunk i.
static const size_t HX_STAGE_BYTES = (size_t)16 << 20;
static int stream_text(hx_ctx *ctx, TextLane &L, const uint8_t *const d_src[2], const uint64_t len[2],
                       const HxhTextSink &sink) {
    cudaStream_t st = L.slot.stream;
    for (int i = 0; i < 2; ++i) {
        if (!L.ex_stage[i]) HX_CUDA(ctx, cudaHostAlloc((void **)&L.ex_stage[i], HX_STAGE_BYTES, cudaHostAllocDefault));
        if (!L.ex_stage_ev[i]) HX_CUDA(ctx, cudaEventCreateWithFlags(&L.ex_stage_ev[i], cudaEventDisableTiming));
    }
    struct Chunk { int kind; uint64_t off; size_t n; };
    auto chunk_at = [&](uint64_t i, Chunk &c) {
        const uint64_t n0 = (len[0] + HX_STAGE_BYTES - 1) / HX_STAGE_BYTES;
        c.kind = i < n0 ? 0 : 1;
        c.off = (c.kind ? i - n0 : i) * HX_STAGE_BYTES;
        c.n = (size_t)std::min<uint64_t>(HX_STAGE_BYTES, len[c.kind] - c.off);
    };
    const uint64_t n_chunks = (len[0] + HX_STAGE_BYTES - 1) / HX_STAGE_BYTES + (len[1] + HX_STAGE_BYTES - 1) / HX_STAGE_BYTES;
    auto issue = [&](uint64_t i) -> cudaError_t {
        Chunk c;
        chunk_at(i, c);
        cudaError_t e = cudaMemcpyAsync(L.ex_stage[i & 1], d_src[c.kind] + c.off, c.n, cudaMemcpyDeviceToHost, st);
        return e == cudaSuccess ? cudaEventRecord(L.ex_stage_ev[i & 1], st) : e;
    };
    if (n_chunks) HX_CUDA(ctx, issue(0));
    for (uint64_t i = 0; i < n_chunks; ++i) {
        if (i + 1 < n_chunks) HX_CUDA(ctx, issue(i + 1));
        HX_CUDA(ctx, cudaEventSynchronize(L.ex_stage_ev[i & 1]));
        Chunk c;
        chunk_at(i, c);
        if (sink.chunk(sink.user, c.kind, c.off, (const char *)L.ex_stage[i & 1], c.n) != 0) {
            cudaStreamSynchronize(st);
            return fail(ctx, HX_ERR_IO, "hx_run_batch_text_stream: the sink reported an error");
        }
    }
    return HX_OK;
}

// One batch through scan + extraction on text lane `lane_id` (= device index * HX_NLANE + lane).  With a sink the text is
// streamed (sink->begin receives both totals before the first chunk); without, it is returned whole in the lane's pinned
// buffers.  Lanes are independent: different host threads may drive different lanes of one context at the same time.
static int run_text_lane(hx_ctx *ctx, int lane_id, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                         hx_result *out, const char *ids, const uint64_t *id_off, const char *tails,
                         const uint32_t *tail_off, const char *gff_tails, const uint32_t *gff_tail_off,
                         const char **text, uint64_t *text_len, const char **gff_text, uint64_t *gff_len,
                         const HxhTextSink *sink) {
    DeviceGuard device_guard;
    if (!ctx) return HX_ERR_ARG;
    const bool want_gff = gff_tails && gff_tail_off;
    if (text) *text = nullptr;
    if (text_len) *text_len = 0;
    if (gff_text) *gff_text = nullptr;
    if (gff_len) *gff_len = 0;
    if (!ctx->primers_set) return fail(ctx, HX_ERR_STATE, "hx_run_batch_text: call hx_set_primers first");
    if (lane_id < 0 || lane_id >= (int)ctx->devs.size() * HX_NLANE) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text: bad lane");
    if (n_records == 0) return HX_OK;
    if (!offsets || !ids || !id_off || !tails || !tail_off) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text: NULL argument");
    if (offsets[n_records] < offsets[0]) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text: offsets must be non-decreasing");
    if (offsets[n_records] > offsets[0] && !arena) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text: NULL arena");
    Device &D = ctx->devs[lane_id / HX_NLANE];
    TextLane &L = D.lanes[lane_id % HX_NLANE];
    Slot &S = L.slot;
    cudaStream_t st = S.stream;
    HX_CUDA(ctx, cudaSetDevice(D.dev));
    const uint32_t np = ctx->n_pairs;
    const uint64_t n_entries = (uint64_t)n_records * np;
    const uint64_t base = offsets[0] & ~15ull;
    const uint64_t copy_bytes = offsets[n_records] - base, bytes = offsets[n_records] - offsets[0];
    const uint64_t id_bytes = id_off[n_records] - id_off[0];
    if (id_off[0] != 0 || tail_off[0] != 0) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text: id_off / tail_off must start at 0");
    if (want_gff && gff_tail_off[0] != 0) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text: gff_tail_off must start at 0");
    const uint32_t tail_bytes = tail_off[np];
    const uint32_t nblk = (uint32_t)((n_entries + 4095) / 4096);
    static const bool trace = getenv("HYPEREX_TIMING") && atoi(getenv("HYPEREX_TIMING")) >= 2;  // per-call phase times
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tt0 = now();
    HX_CUDA(ctx, cudaStreamSynchronize(st));
    const SlabPlan pl = plan_slab(ctx, n_records, bytes);
    int rc = prepare_slab(ctx, S, pl, n_records, copy_bytes + 64);
    if (rc != HX_OK) return rc;
    for (int pass = 0;; ++pass) {  // the extraction scratch: one block, carved per call
        DevPool &P = L.ex_pool;
        P.reset();
        L.ex_ids.carve(P, id_bytes + 16);
        L.ex_id_off.carve(P, (size_t)n_records + 1);
        L.ex_tails.carve(P, tail_bytes + 16);
        L.ex_tail_off.carve(P, np + 1);
        L.ex_lens.carve(P, n_entries + 1);
        L.ex_sums.carve(P, nblk + 2);
        if (want_gff) {
            L.ex_gtails.carve(P, gff_tail_off[np] + 16);
            L.ex_gtail_off.carve(P, np + 1);
            L.ex_glens.carve(P, n_entries + 1);
            L.ex_gsums.carve(P, nblk + 2);
        }
        if (P.fits()) break;
        if (pass) return fail(ctx, HX_ERR_STATE, "hx_run_batch_text: pool did not fit after growing");
        HX_CUDA(ctx, P.grow());
    }
    if (want_gff) {
        if (gff_tail_off[np])
            HX_CUDA(ctx, cudaMemcpyAsync(L.ex_gtails.p, gff_tails, gff_tail_off[np], cudaMemcpyHostToDevice, st));
        HX_CUDA(ctx, cudaMemcpyAsync(L.ex_gtail_off.p, gff_tail_off, ((size_t)np + 1) * 4, cudaMemcpyHostToDevice, st));
    }
    if (copy_bytes) HX_CUDA(ctx, cudaMemcpyAsync(S.arena.p, arena + base, copy_bytes, cudaMemcpyHostToDevice, st));
    HX_CUDA(ctx, cudaMemcpyAsync(S.offsets.p, offsets, ((size_t)n_records + 1) * 8, cudaMemcpyHostToDevice, st));
    if (id_bytes) HX_CUDA(ctx, cudaMemcpyAsync(L.ex_ids.p, ids, id_bytes, cudaMemcpyHostToDevice, st));
    HX_CUDA(ctx, cudaMemcpyAsync(L.ex_id_off.p, id_off, ((size_t)n_records + 1) * 8, cudaMemcpyHostToDevice, st));
    if (tail_bytes) HX_CUDA(ctx, cudaMemcpyAsync(L.ex_tails.p, tails, tail_bytes, cudaMemcpyHostToDevice, st));
    HX_CUDA(ctx, cudaMemcpyAsync(L.ex_tail_off.p, tail_off, ((size_t)np + 1) * 4, cudaMemcpyHostToDevice, st));
    rc = enqueue_slab(ctx, D, S, pl, S.arena.p, S.offsets.p, base, n_records, S.out.p, st);
    if (rc != HX_OK) return rc;
    if (out) HX_CUDA(ctx, cudaMemcpyAsync(out, S.out.p, (size_t)n_entries * sizeof(hx_result), cudaMemcpyDeviceToHost, st));
    ctx->launches += hx_launch_fasta_offsets(S.out.p, L.ex_id_off.p, L.ex_tail_off.p, np, n_entries, L.ex_lens.p,
                                             L.ex_sums.p, st);
    // total = offset of the last entry + its length = block offset + in-block prefix; read both pieces
    L.h_totals[0] = L.h_totals[1] = 0;
    HX_CUDA(ctx, cudaMemcpyAsync(&L.h_totals[0], L.ex_sums.p + nblk, 8, cudaMemcpyDeviceToHost, st));
    if (want_gff) {
        ctx->launches += hx_launch_gff_offsets(S.out.p, L.ex_id_off.p, L.ex_gtail_off.p, np, n_entries, L.ex_glens.p,
                                               L.ex_gsums.p, st);
        HX_CUDA(ctx, cudaMemcpyAsync(&L.h_totals[1], L.ex_gsums.p + nblk, 8, cudaMemcpyDeviceToHost, st));
    }
    const double tt1 = now();
    HX_CUDA(ctx, cudaStreamSynchronize(st));
    const double tt2 = now();
    rc = check_status(ctx, S);
    if (rc != HX_OK) return rc;
    const uint64_t total = L.h_totals[0], gtotal = L.h_totals[1];
    if (sink && sink->begin && sink->begin(sink->user, total, gtotal) != 0)
        return fail(ctx, HX_ERR_IO, "hx_run_batch_text_stream: the sink reported an error");
    for (int pass = 0; total || gtotal; ++pass) {  // both output texts share one device block
        DevPool &P = L.ex_text_pool;
        P.reset();
        L.ex_text.carve(P, total + 16);
        L.ex_gtext.carve(P, gtotal + 16);
        if (P.fits()) break;
        if (pass) return fail(ctx, HX_ERR_STATE, "hx_run_batch_text: text pool did not fit after growing");
        HX_CUDA(ctx, P.grow());
    }
    if (total) {
        if (!sink && total > L.ex_host_cap) {
            if (L.ex_host) cudaFreeHost(L.ex_host);
            L.ex_host = nullptr;
            L.ex_host_cap = 0;
            const size_t want = (size_t)(total + total / 4 + 4096);
            HX_CUDA(ctx, cudaHostAlloc((void **)&L.ex_host, want, cudaHostAllocDefault));
            L.ex_host_cap = want;
        }
        ctx->launches += hx_launch_fasta_write(S.arena.p, S.offsets.p, base, S.out.p, L.ex_ids.p, L.ex_id_off.p,
                                               L.ex_tails.p, L.ex_tail_off.p, np, n_entries, L.ex_lens.p, L.ex_text.p, st);
        if (!sink) HX_CUDA(ctx, cudaMemcpyAsync(L.ex_host, L.ex_text.p, total, cudaMemcpyDeviceToHost, st));
    }
    if (gtotal) {
        if (!sink && gtotal > L.ex_ghost_cap) {
            if (L.ex_ghost) cudaFreeHost(L.ex_ghost);
            L.ex_ghost = nullptr;
            L.ex_ghost_cap = 0;
            const size_t want = (size_t)(gtotal + gtotal / 4 + 4096);
            HX_CUDA(ctx, cudaHostAlloc((void **)&L.ex_ghost, want, cudaHostAllocDefault));
            L.ex_ghost_cap = want;
        }
        ctx->launches += hx_launch_gff_write(S.out.p, L.ex_ids.p, L.ex_id_off.p, L.ex_gtails.p, L.ex_gtail_off.p, np,
                                             n_entries, L.ex_glens.p, L.ex_gtext.p, st);
        if (!sink) HX_CUDA(ctx, cudaMemcpyAsync(L.ex_ghost, L.ex_gtext.p, gtotal, cudaMemcpyDeviceToHost, st));
    }
    const double tt3 = now();
    if (sink && (total || gtotal)) {
        const uint8_t *const src[2] = {L.ex_text.p, L.ex_gtext.p};
        const uint64_t lens[2] = {total, gtotal};
        rc = stream_text(ctx, L, src, lens, *sink);
        if (rc != HX_OK) return rc;
    }
    HX_CUDA(ctx, cudaStreamSynchronize(st));  // also covers the result-table copy into `out`
    HX_CUDA(ctx, cudaGetLastError());
    if (trace)
        fprintf(stderr, "[hx_run_batch_text lane %d] records=%u bases=%llu text=%llu: carve+enqueue %.1f ms, scan sync %.1f ms, "
                        "enqueue text %.1f ms, text copy/sink %.1f ms\n",
                lane_id, n_records, (unsigned long long)bytes, (unsigned long long)(total + gtotal), (tt1 - tt0) * 1e3,
                (tt2 - tt1) * 1e3, (tt3 - tt2) * 1e3, (now() - tt3) * 1e3);
    if (ctx->opt_time) sync_timed(ctx);
    if (text) *text = sink ? nullptr : (const char *)L.ex_host;
    if (text_len) *text_len = total;
    if (want_gff) {
        if (gff_text) *gff_text = sink ? nullptr : (const char *)L.ex_ghost;
        if (gff_len) *gff_len = gtotal;
    }
    return HX_OK;
}

int hx_run_batch_fasta(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out,
                       const char *ids, const uint64_t *id_off, const char *tails, const uint32_t *tail_off,
                       const char **text, uint64_t *text_len) {
    if (ctx && (!text || !text_len || (n_records && !out))) return fail(ctx, HX_ERR_ARG, "hx_run_batch_fasta: NULL out/text/text_len");
    return run_text_lane(ctx, 0, arena, offsets, n_records, out, ids, id_off, tails, tail_off, nullptr, nullptr, text,
                         text_len, nullptr, nullptr, nullptr);
}

int hx_run_batch_text(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out,
                      const char *ids, const uint64_t *id_off, const char *tails, const uint32_t *tail_off,
                      const char *gff_tails, const uint32_t *gff_tail_off, const char **text, uint64_t *text_len,
                      const char **gff_text, uint64_t *gff_len) {
    if (ctx && (!text || !text_len || (n_records && !out))) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text: NULL out/text/text_len");
    const bool want_gff = gff_tails && gff_tail_off && gff_text && gff_len;
    return run_text_lane(ctx, 0, arena, offsets, n_records, out, ids, id_off, tails, tail_off, want_gff ? gff_tails : nullptr,
                         want_gff ? gff_tail_off : nullptr, text, text_len, gff_text, gff_len, nullptr);
}

namespace {
struct PublicSink {  // hx_text_sink (kind, bytes, n) behind the internal sink (which also carries totals and offsets)
    hx_text_sink fn;
    void *user;
    static int chunk(void *u, int kind, uint64_t, const char *bytes, size_t n) {
        PublicSink *p = (PublicSink *)u;
        return p->fn(p->user, kind, bytes, n);
    }
};
}  // namespace

int hx_run_batch_text_stream(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                             hx_result *out, const char *ids, const uint64_t *id_off, const char *tails,
                             const uint32_t *tail_off, const char *gff_tails, const uint32_t *gff_tail_off,
                             hx_text_sink sink, void *user, uint64_t *text_len, uint64_t *gff_len) {
    if (!ctx) return HX_ERR_ARG;
    if (!sink || (n_records && !out)) return fail(ctx, HX_ERR_ARG, "hx_run_batch_text_stream: NULL sink/out");
    PublicSink ps{sink, user};
    const HxhTextSink s{nullptr, &PublicSink::chunk, &ps};
    const bool want_gff = gff_tails && gff_tail_off;
    return run_text_lane(ctx, 0, arena, offsets, n_records, out, ids, id_off, tails, tail_off, want_gff ? gff_tails : nullptr,
                         want_gff ? gff_tail_off : nullptr, nullptr, text_len, nullptr, gff_len, &s);
}

}  // extern "C"

// hx_run_batch / hx_run_batch_packed: rec_flags != NULL means `arena` is a packed arena (4-bit codes, nibble i = base i).
// An ASCII arena may still cross PCIe packed (option pack_h2d): a producer thread packs slab after slab into a ring of
// pinned staging buffers with all host threads while this thread enqueues the copies and kernels of the slabs that are
// ready, and the device runs its nibble-text kernels.  The bytes per base over PCIe halve; the caller's buffers and the
// result table are what they were.
namespace {
struct SlabCut {
    uint32_t r0, r1;
    uint64_t base, copy_bytes;
};

// A piece of a slab on the pack-on-the-way path: bases [p, q) of the batch, packed range by range into ring buffer
// `slot` and copied from there to byte dst_off of the slab's arena
struct PackPiece {
    int64_t slab;
    uint64_t p, q, al;  // al: the base the first byte of the ring buffer stands for (p rounded down to the slab's base)
    uint64_t dst_off, bytes;
    int slot;
    int64_t first_range, n_ranges;
    bool last_of_slab;
};
struct PackRange {
    uint64_t p, q;
    int64_t piece;
    bool slab_first;
};
}  // namespace

// ---- hx_run_batch, pack_h2d = 2: BOTH the link and the packer threads work all the time ----------------------------------
// ASCII slabs keep the link busy and cost no host work; packed slabs cost host work and half the link time.  With a link
// that moves L bases per second as ASCII and packers that make P, sending the fraction f = 2P / (2L + P) of the batch packed
// finishes both at the same moment: L + P / 2 bases per second, against max(L, P) for either pure mode.  f is not computed:
// the issuing thread takes ASCII slabs from the FRONT of the batch whenever no packed piece is waiting, the packer threads
// claim slabs from the BACK, and the two meet wherever the host they run on makes them meet (all packed on a host with many
// idle cores, nearly all ASCII on a host whose cores are busy).  Slabs get (device, slot) pairs as those become free.
namespace {
// The issuing thread is the link's scheduler: an ASCII slab crosses in chunks, and at most HYB_LINKQ copies (chunks and
// packed pieces together) are queued at any time, so a packed piece that becomes ready waits for a fraction of a
// millisecond, not for a whole slab — the packers' ring holds less than a millisecond of their work.
constexpr int HYB_LINKQ_MAX = 8;
int hyb_linkq() {  // HYPEREX_MIXED_LINKQ / HYPEREX_MIXED_CHUNK_MB: A/B knobs (tools/mixed_ab.py)
    static const int v = [] {
        const char *e = getenv("HYPEREX_MIXED_LINKQ");
        return e ? std::max(1, std::min(HYB_LINKQ_MAX, atoi(e))) : 3;
    }();
    return v;
}
uint64_t hyb_chunk() {
    static const uint64_t v = [] {
        const char *e = getenv("HYPEREX_MIXED_CHUNK_MB");
        return (uint64_t)(e ? std::max(1, atoi(e)) : 8) << 20;
    }();
    return v;
}
struct HybSlab {
    uint32_t r0, r1;
    uint64_t base_a, copy_a;  // ASCII: 16-byte phase of the arena kept
    uint64_t base_p, copy_p;  // packed on the way: base rounded down to 128 symbols
    int64_t first_piece = 0;
    int di = -1, si = -1;     // where it runs (known once issued / first piece issued)
};
}  // namespace

static int run_batch_hybrid(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out) {
    const size_t np = ctx->n_pairs;
    const int ndev = (int)ctx->devs.size();
    const int pack_threads = ctx->opt_pack_threads > 0 ? (int)ctx->opt_pack_threads : (int)std::thread::hardware_concurrency();
    const uint64_t total = offsets[n_records] - offsets[0];
    // slabs a quarter of the usual size: the two sides meet inside one slab's worth of time, and that time is the tail
    uint64_t slab_bytes = std::max<uint64_t>((uint64_t)ctx->opt_slab_bytes / 4, 1);
    slab_bytes = std::min(slab_bytes, std::max<uint64_t>(total / (8ull * ndev) + 1, (uint64_t)8 << 20));
    std::vector<HybSlab> slabs;
    for (uint32_t r0 = 0; r0 < n_records;) {
        const uint64_t limit = offsets[r0] + slab_bytes;
        uint32_t r1 = (uint32_t)(std::upper_bound(offsets + r0 + 1, offsets + n_records + 1, limit) - offsets) - 1;
        if (r1 <= r0) r1 = r0 + 1;
        for (uint32_t i = r0; i < r1; ++i)
            if (offsets[i + 1] < offsets[i]) return fail(ctx, HX_ERR_ARG, "hx_run_batch: offsets must be non-decreasing");
        HybSlab c;
        c.r0 = r0;
        c.r1 = r1;
        c.base_a = offsets[r0] & ~15ull;
        c.copy_a = offsets[r1] - c.base_a;
        c.base_p = offsets[r0] & ~127ull;
        c.copy_p = (offsets[r1] + 1) / 2 - c.base_p / 2;
        slabs.push_back(c);
        r0 = r1;
    }
    const int64_t n_slabs = (int64_t)slabs.size();

    // pieces and ranges of every slab in the packers' order: last slab first (most of them are never packed)
    std::vector<PackPiece> pieces;
    std::vector<PackRange> ranges;
    const uint64_t piece_bases = 2ull * (uint64_t)ctx->opt_pack_piece_bytes;  // a multiple of 128
    const uint64_t range_bases = 128u << 10;
    const int n_buf = HX_NRING * ndev;
    for (int64_t j = n_slabs - 1; j >= 0; --j) {
        HybSlab &c = slabs[j];
        const uint64_t begin = offsets[c.r0], end = offsets[c.r1];
        c.first_piece = (int64_t)pieces.size();
        for (uint64_t al = c.base_p;; al += piece_bases) {
            PackPiece P;
            P.slab = j;
            P.al = al;
            P.p = std::max(al, begin);
            P.q = std::max(P.p, std::min(end, al + piece_bases));
            P.dst_off = (al - c.base_p) / 2;
            P.bytes = P.q > P.al ? (P.q + 1) / 2 - P.al / 2 : 0;
            P.slot = (int)(pieces.size() % (size_t)n_buf);
            P.first_range = (int64_t)ranges.size();
            for (uint64_t x = P.p; x < P.q;) {
                const uint64_t y = std::min(P.q, (x + range_bases) & ~(uint64_t)127);
                ranges.push_back({x, y, (int64_t)pieces.size(), x == begin});
                x = y;
            }
            P.n_ranges = (int64_t)ranges.size() - P.first_range;
            P.last_of_slab = al + piece_bases >= end;
            pieces.push_back(P);
            if (P.last_of_slab) break;
        }
    }
    const int64_t n_pieces = (int64_t)pieces.size();
    std::vector<HxhPackPart> parts(ranges.size());
    std::unique_ptr<std::atomic<int>[]> piece_done(new std::atomic<int>[pieces.size()]);
    for (size_t g = 0; g < pieces.size(); ++g) piece_done[g].store(0, std::memory_order_relaxed);

    // ring, flags, events (as on the all-packed path; nothing is enqueued yet, so an error can simply return)
    PackRing &R = ctx->pack_ring;
    {
        const size_t buf_bytes = (size_t)ctx->opt_pack_piece_bytes + 64;
        if (R.n_buf != n_buf || R.buf_bytes != buf_bytes) {
            if (R.p) cudaFreeHost(R.p);
            R.p = nullptr;
            R.n_buf = 0;
            if (cudaHostAlloc((void **)&R.p, buf_bytes * (size_t)n_buf, cudaHostAllocPortable) != cudaSuccess)
                return fail(ctx, HX_ERR_NOMEM, "hx_run_batch: cannot allocate %zu bytes of pinned staging", buf_bytes * (size_t)n_buf);
            R.n_buf = n_buf;
            R.buf_bytes = buf_bytes;
        }
        if (R.flags_cap < (size_t)n_records + 64) {
            if (R.flags) cudaFreeHost(R.flags);
            R.flags = nullptr;
            R.flags_cap = 0;
            const size_t want = (size_t)n_records + (size_t)n_records / 4 + 64;
            if (cudaHostAlloc((void **)&R.flags, want, cudaHostAllocPortable) != cudaSuccess)
                return fail(ctx, HX_ERR_NOMEM, "hx_run_batch: cannot allocate %zu bytes of pinned staging", want);
            R.flags_cap = want;
        }
        for (Device &D : ctx->devs) {
            if (cudaSetDevice(D.dev) != cudaSuccess) return fail(ctx, HX_ERR_CUDA, "hx_run_batch: cudaSetDevice failed");
            while ((int)D.ring_ev.size() < n_buf) {
                cudaEvent_t e = nullptr;
                if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(ctx, HX_ERR_CUDA, "hx_run_batch: cudaEventCreate failed");
                D.ring_ev.push_back(e);
            }
            while ((int)D.link_ev.size() < HYB_LINKQ_MAX) {
                cudaEvent_t e = nullptr;
                if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(ctx, HX_ERR_CUDA, "hx_run_batch: cudaEventCreate failed");
                D.link_ev.push_back(e);
            }
            for (int si = 0; si < HX_NSLOT; ++si)
                if (!D.slots[si].done_ev && cudaEventCreateWithFlags(&D.slots[si].done_ev, cudaEventDisableTiming) != cudaSuccess)
                    return fail(ctx, HX_ERR_CUDA, "hx_run_batch: cudaEventCreate failed");
        }
    }

    // who owns which slabs: [0, front) were taken as ASCII by the issuing thread, [back, n_slabs) by the packers
    std::mutex claim_mu;
    int64_t front = 0;
    std::atomic<int64_t> back{n_slabs};
    std::atomic<int64_t> next_range{0}, allowed_piece{n_buf};
    std::atomic<bool> pack_abort{false};
    bool pack_stop = false;                            // under claim_mu: a claim was refused, the packers are leaving
    std::atomic<int64_t> ascii_ns{0}, pack_ns{0};      // issue-to-issue time of the last ASCII / packed slab
    const int workers = std::max(1, pack_threads - 1);  // the issuing thread polls, so it is not idle
    std::thread packer([&, workers] {
        hxh_parallel_run(workers, workers, [&](int) {
            for (;;) {
                const int64_t i = next_range.fetch_add(1, std::memory_order_relaxed);
                if (i >= (int64_t)ranges.size() || pack_abort.load(std::memory_order_relaxed)) return;
                const PackRange &rg = ranges[i];
                const PackPiece &P = pieces[rg.piece];
                if (P.slab < back.load(std::memory_order_acquire)) {  // ranges come in order: this is slab back - 1 or the claim fails
                    std::lock_guard<std::mutex> lk(claim_mu);
                    if (P.slab < back.load(std::memory_order_relaxed)) {
                        // The slab is taken only if the packers can be through with it before the ASCII side has eaten its way
                        // up to it (a claim cannot be given back, and whatever the packers still hold when the sides meet is
                        // the tail of the call): (slabs still between the sides) x (time per ASCII slab) >= time per packed slab,
                        // both measured in this call; one slab of margin until they are.  A refusal is final.
                        const int64_t between = P.slab - front;  // unclaimed slabs below this one
                        const int64_t ta = ascii_ns.load(std::memory_order_relaxed), tp = pack_ns.load(std::memory_order_relaxed);
                        const bool ok = !pack_stop && front <= P.slab && (ta > 0 && tp > 0 ? between * ta >= tp : between >= 1 || n_slabs == 1);
                        if (!ok) {
                            pack_stop = true;
                            return;  // every later range belongs to the ASCII side too
                        }
                        back.store(P.slab, std::memory_order_release);
                    }
                }
                while (rg.piece >= allowed_piece.load(std::memory_order_acquire)) {  // the ring buffer is still being read
                    if (pack_abort.load(std::memory_order_relaxed)) return;
                    hx_cpu_relax();
                }
                uint8_t *buf = R.p + (size_t)P.slot * R.buf_bytes;
                hxh_pack_range(arena, offsets, n_records, rg.p, rg.q, (uint8_t *)((uintptr_t)buf - (uintptr_t)(P.al / 2)), R.flags, false,
                               rg.slab_first, &parts[i]);
                piece_done[rg.piece].fetch_add(1, std::memory_order_release);
            }
        });
    });

    // slots: 0 free, 1 in flight (done_ev recorded), 2 being filled by a packed slab
    std::vector<int> state(ndev * HX_NSLOT, 0), used(ndev * HX_NSLOT, 0);
    int rc = HX_OK;
    int rr = 0;  // round-robin start of the device search
    auto retire = [&] {  // slabs whose result copy has completed give their slots back
        for (int di = 0; di < ndev && rc == HX_OK; ++di)
            for (int si = 0; si < HX_NSLOT && rc == HX_OK; ++si) {
                int &st = state[di * HX_NSLOT + si];
                if (st != 1) continue;
                const cudaError_t q = cudaEventQuery(ctx->devs[di].slots[si].done_ev);
                if (q == cudaErrorNotReady) continue;
                if (q != cudaSuccess) rc = fail(ctx, HX_ERR_CUDA, "hx_run_batch: %s", cudaGetErrorString(q));
                else rc = check_status(ctx, ctx->devs[di].slots[si]);
                st = 0;
            }
    };
    auto find_slot = [&](int &di_out, int &si_out) -> bool {  // a free (device, slot), devices in turn
        for (int t = 0; t < ndev; ++t) {
            const int di = (rr + t) % ndev;
            for (int si = 0; si < HX_NSLOT; ++si)
                if (state[di * HX_NSLOT + si] == 0) {
                    di_out = di;
                    si_out = si;
                    rr = (di + 1) % ndev;
                    return true;
                }
        }
        return false;
    };
    int64_t oldest = 0, issued = 0;  // ring: first piece whose copy is not known to be complete / pieces whose copies are enqueued
    auto piece_ev = [&](int64_t g) -> cudaEvent_t { return ctx->devs[slabs[pieces[g].slab].di].ring_ev[pieces[g].slot]; };
    auto poll_ring = [&] {
        while (oldest < issued && cudaEventQuery(piece_ev(oldest)) == cudaSuccess) ++oldest;
        allowed_piece.store(oldest + n_buf, std::memory_order_release);
    };
    uint64_t h2d = 0, n_packed_slabs = 0;
    int64_t g = 0;  // next piece to issue, in the packers' order
    int64_t af_j = -1;  // the ASCII slab that is crossing chunk by chunk, and how far it is
    uint64_t af_off = 0;
    int64_t a_issued = 0, a_done = 0;  // ASCII chunk copies enqueued / known to be complete
    int a_dev[HYB_LINKQ_MAX] = {0};
    const int HYB_LINKQ = hyb_linkq();
    const uint64_t HYB_CHUNK = hyb_chunk();
    const bool dbg = getenv("HYPEREX_HYB_DEBUG") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    uint64_t n_iter = 0, n_no_slot = 0, n_queue_full = 0, n_chunks = 0, n_idle = 0, n_pk_no_slot = 0;
    double t_first_piece = -1, t_met = -1, t_last_ascii = -1;
    int64_t t_pack_prev = 0, t_ascii_prev = 0;
    uint64_t n_ascii_slabs = 0;
    while (rc == HX_OK) {
        ++n_iter;
        poll_ring();
        retire();
        if (rc != HX_OK) break;
        bool did = false;
        // (a) a packed piece that is ready goes first: it is what the packer threads are waiting to get rid of
        if (g < n_pieces && pieces[g].slab >= back.load(std::memory_order_acquire) &&
            piece_done[g].load(std::memory_order_acquire) >= (int)pieces[g].n_ranges) {
            const PackPiece &P = pieces[g];
            HybSlab &c = slabs[P.slab];
            const uint32_t n = c.r1 - c.r0;
            if (t_first_piece < 0) t_first_piece = since();
            if (c.di < 0) {  // first piece of its slab: the slab needs a slot
                int di, si;
                if (!find_slot(di, si)) ++n_pk_no_slot;
                else {
                    Device &D = ctx->devs[di];
                    HX_CUDA_BREAK(ctx, rc, cudaSetDevice(D.dev));
                    const SlabPlan pl = plan_slab(ctx, n, offsets[c.r1] - offsets[c.r0]);
                    rc = prepare_slab(ctx, D.slots[si], pl, n, c.copy_p + 64, true);
                    if (rc != HX_OK) break;
                    c.di = di;
                    c.si = si;
                    state[di * HX_NSLOT + si] = 2;
                    used[di * HX_NSLOT + si] = 1;
                }
            }
            if (rc != HX_OK) break;
            if (c.di >= 0) {
                Device &D = ctx->devs[c.di];
                Slot &S = D.slots[c.si];
                HX_CUDA_BREAK(ctx, rc, cudaSetDevice(D.dev));
                if (P.bytes) HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.arena.p + P.dst_off, R.p + (size_t)P.slot * R.buf_bytes, P.bytes, cudaMemcpyHostToDevice, S.stream));
                HX_CUDA_BREAK(ctx, rc, cudaEventRecord(piece_ev(g), S.stream));
                issued = g + 1;
                ++g;
                did = true;
                if (P.last_of_slab) {  // evidence merge, rec_flags, offsets, kernels, result copy
                    const PackPiece &P0 = pieces[c.first_piece];
                    uint32_t open_rec = UINT32_MAX, open_acc = 0;
                    hxh_pack_merge(R.flags, parts.data() + P0.first_range, (size_t)(P.first_range + P.n_ranges - P0.first_range), open_rec, open_acc);
                    // empty records at the head of the slab end before its first base: no range of THIS slab finalises them (the
                    // last range of the slab before it does, with the same 0, if that slab is packed at all)
                    for (uint32_t r = c.r0; r < c.r1 && offsets[r + 1] == offsets[c.r0]; ++r) R.flags[r] = 0;
                    const SlabPlan pl = plan_slab(ctx, n, offsets[c.r1] - offsets[c.r0]);
                    HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.flags.p, R.flags + c.r0, n, cudaMemcpyHostToDevice, S.stream));
                    HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.offsets.p, offsets + c.r0, ((size_t)n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, S.stream));
                    rc = enqueue_slab(ctx, D, S, pl, S.arena.p, S.offsets.p, c.base_p, n, S.out.p, S.stream, S.flags.p);
                    if (rc != HX_OK) break;
                    HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(out + (size_t)c.r0 * np, S.out.p, (size_t)n * np * sizeof(hx_result), cudaMemcpyDeviceToHost, S.stream));
                    HX_CUDA_BREAK(ctx, rc, cudaEventRecord(S.done_ev, S.stream));
                    state[c.di * HX_NSLOT + c.si] = 1;
                    h2d += c.copy_p + ((uint64_t)n + 1) * sizeof(uint64_t) + n;
                    ++n_packed_slabs;
                    {
                        const int64_t now = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
                        if (n_packed_slabs > 1) pack_ns.store(std::max<int64_t>(now - t_pack_prev, 1), std::memory_order_relaxed);
                        t_pack_prev = now;
                    }
                }
            }
        }
        // (b) otherwise the link gets the next chunk of an ASCII slab from the front.  One free slot per device is left to the
        //     packed side while it has work, so that a finished piece never waits for its slab's slot.
        if (!did && rc == HX_OK) {
            while (a_done < a_issued && cudaEventQuery(ctx->devs[a_dev[a_done % HYB_LINKQ]].link_ev[a_done % HYB_LINKQ]) == cudaSuccess) ++a_done;
            if (af_j < 0) {
                int di = -1, si = -1, n_free = 0;
                for (int x : state) n_free += x == 0;
                bool packed_pending;
                {
                    std::lock_guard<std::mutex> lk(claim_mu);
                    const int64_t b = back.load(std::memory_order_relaxed);
                    packed_pending = g < n_pieces && (pieces[g].slab >= b || front < b);
                }
                int64_t j = -1;
                if (n_free >= (packed_pending ? 2 : 1) && find_slot(di, si)) {
                    std::lock_guard<std::mutex> lk(claim_mu);
                    if (front < back.load(std::memory_order_relaxed)) j = front++;
                    else if (t_met < 0) t_met = since();
                } else {
                    ++n_no_slot;
                }
                if (j >= 0) {
                    HybSlab &c = slabs[j];
                    HX_CUDA_BREAK(ctx, rc, cudaSetDevice(ctx->devs[di].dev));
                    const SlabPlan pl = plan_slab(ctx, c.r1 - c.r0, offsets[c.r1] - offsets[c.r0]);
                    rc = prepare_slab(ctx, ctx->devs[di].slots[si], pl, c.r1 - c.r0, c.copy_a + 64, false);
                    if (rc != HX_OK) break;
                    c.di = di;
                    c.si = si;
                    used[di * HX_NSLOT + si] = 1;
                    state[di * HX_NSLOT + si] = 2;
                    af_j = j;
                    af_off = 0;
                    did = true;
                }
            }
            if (af_j >= 0 && (a_issued - a_done) + (issued - oldest) >= HYB_LINKQ) ++n_queue_full;
            if (af_j >= 0 && (a_issued - a_done) + (issued - oldest) < HYB_LINKQ) {
                HybSlab &c = slabs[af_j];
                const uint32_t n = c.r1 - c.r0;
                Device &D = ctx->devs[c.di];
                Slot &S = D.slots[c.si];
                HX_CUDA_BREAK(ctx, rc, cudaSetDevice(D.dev));
                if (af_off < c.copy_a) {
                    const uint64_t chunk = std::min<uint64_t>(HYB_CHUNK, c.copy_a - af_off);
                    HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.arena.p + af_off, arena + c.base_a + af_off, chunk, cudaMemcpyHostToDevice, S.stream));
                    HX_CUDA_BREAK(ctx, rc, cudaEventRecord(D.link_ev[a_issued % HYB_LINKQ], S.stream));
                    a_dev[a_issued % HYB_LINKQ] = c.di;
                    ++a_issued;
                    ++n_chunks;
                    t_last_ascii = since();
                    af_off += chunk;
                }
                if (af_off >= c.copy_a) {  // the whole slab is on its way: offsets, kernels, result copy
                    const SlabPlan pl = plan_slab(ctx, n, offsets[c.r1] - offsets[c.r0]);
                    HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.offsets.p, offsets + c.r0, ((size_t)n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, S.stream));
                    rc = enqueue_slab(ctx, D, S, pl, S.arena.p, S.offsets.p, c.base_a, n, S.out.p, S.stream, nullptr);
                    if (rc != HX_OK) break;
                    HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(out + (size_t)c.r0 * np, S.out.p, (size_t)n * np * sizeof(hx_result), cudaMemcpyDeviceToHost, S.stream));
                    HX_CUDA_BREAK(ctx, rc, cudaEventRecord(S.done_ev, S.stream));
                    state[c.di * HX_NSLOT + c.si] = 1;
                    h2d += c.copy_a + ((uint64_t)n + 1) * sizeof(uint64_t);
                    af_j = -1;
                    {
                        const int64_t now = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
                        if (++n_ascii_slabs > 1) ascii_ns.store(std::max<int64_t>(now - t_ascii_prev, 1), std::memory_order_relaxed);
                        t_ascii_prev = now;
                    }
                }
                did = true;
            }
        }
        if (rc != HX_OK) break;
        // (c) done when the two sides have met and every piece of the packed side is on its way
        if (!did) {
            bool finished;
            {
                std::lock_guard<std::mutex> lk(claim_mu);
                const int64_t b = back.load(std::memory_order_relaxed);
                finished = af_j < 0 && front >= b && (g >= n_pieces || pieces[g].slab < b);
            }
            if (finished) break;
            ++n_idle;
            hx_cpu_relax();
        }
    }
    const double t_loop = since();
    if (rc != HX_OK) pack_abort.store(true);
    packer.join();
    for (int di = 0; di < ndev; ++di) {  // epilogue on every path, as in run_batch_impl
        cudaSetDevice(ctx->devs[di].dev);
        for (int si = 0; si < HX_NSLOT; ++si) {
            if (!used[di * HX_NSLOT + si]) continue;
            cudaError_t e = cudaStreamSynchronize(ctx->devs[di].slots[si].stream);
            if (e != cudaSuccess && rc == HX_OK) rc = fail(ctx, HX_ERR_CUDA, "hx_run_batch: %s", cudaGetErrorString(e));
            if (rc == HX_OK && e == cudaSuccess && state[di * HX_NSLOT + si] != 2) rc = check_status(ctx, ctx->devs[di].slots[si]);
        }
    }
    sync_timed(ctx);
    if (dbg)
        fprintf(stderr, "[hyb] slabs %lld packed %llu pieces %lld | iter %llu idle %llu ascii: chunks %llu no_slot %llu queue_full %llu | packed first-slot misses %llu | ms: first piece %.2f last ascii chunk %.2f met %.2f loop %.2f total %.2f\n",
                (long long)n_slabs, (unsigned long long)n_packed_slabs, (long long)g, (unsigned long long)n_iter, (unsigned long long)n_idle, (unsigned long long)n_chunks,
                (unsigned long long)n_no_slot, (unsigned long long)n_queue_full, (unsigned long long)n_pk_no_slot, t_first_piece, t_last_ascii, t_met, t_loop, since());
    ctx->batch_stats[0] = h2d;
    ctx->batch_stats[1] = (uint64_t)n_records * np * sizeof(hx_result);
    ctx->batch_stats[2] = n_packed_slabs ? 1 : 0;
    ctx->batch_stats[3] = (uint64_t)n_slabs;
    ctx->batch_stats[4] = (uint64_t)pack_threads;
    ctx->batch_stats[5] = n_packed_slabs;
    ctx->batch_stats[6] = 2;
    return rc;
}

static int run_batch_impl(hx_ctx *ctx, const uint8_t *arena, const uint8_t *rec_flags, const uint64_t *offsets, uint32_t n_records,
                          hx_result *out) {
    DeviceGuard device_guard;
    const bool packed_in = rec_flags != nullptr;
    if (!ctx->primers_set) return fail(ctx, HX_ERR_STATE, "hx_run_batch: call hx_set_primers first");
    if (packed_in && !ctx->packable)
        return fail(ctx, HX_ERR_STATE, "hx_run_batch_packed: the primer set has symbols outside ACGTURYSWKMBDHVN, which packed text "
                                       "cannot represent; use hx_run_batch");
    if (n_records == 0) return HX_OK;
    if (!offsets || !out) return fail(ctx, HX_ERR_ARG, "hx_run_batch: NULL offsets/out");
    if (offsets[n_records] < offsets[0]) return fail(ctx, HX_ERR_ARG, "hx_run_batch: offsets must be non-decreasing");
    if (offsets[n_records] > offsets[0] && !arena) return fail(ctx, HX_ERR_ARG, "hx_run_batch: NULL arena");
    // pack on the way: every slab (1) / never (0) / mixed (2) / automatic (-1).  Whether packing pays is a property of the host,
    // not of the library: on a 16-thread B200 host one process gets 66-79 Gbases/s mixed against 50-53 ASCII, on a 24-thread
    // host shared by two processes (one GPU each) mixed equals ASCII and packing every slab halves it
    // (profiles/r02_mixed_slabs.md, r02_pack_h2d_auto.md) — a process cannot see what its siblings do to the memory system.
    // So automatic = measure: batches large enough for the pipeline to fill (>= 64 MB), with at least 2 packer threads per
    // device, alternate ASCII and mixed six times; calls three to six are timed (buffers exist, pages are mapped), the
    // better of its two runs counts for each mode, and mixed serves the later calls iff it was at least 5 % faster.
    const int pack_threads = ctx->opt_pack_threads > 0 ? (int)ctx->opt_pack_threads : (int)std::thread::hardware_concurrency();
    const bool can_pack = !packed_in && ctx->packable && arena && offsets[n_records] > offsets[0];
    const bool pack_probe = can_pack && ctx->opt_pack_h2d < 0 && ctx->auto_pack.decided < 0 &&
                            offsets[n_records] - offsets[0] >= ((uint64_t)64 << 20) && pack_threads >= 2 * (int)ctx->devs.size();
    // what automatic measures against ASCII slabs is the mixed path (pack_h2d = 2): it is at least as fast as packing every
    // slab at every thread count measured (profiles/r02_mixed_slabs.md), so pack_h2d = 1 only runs when asked for
    const bool mixed = can_pack && (ctx->opt_pack_h2d == 2 ||
                                    (ctx->opt_pack_h2d < 0 && (pack_probe ? (ctx->auto_pack.calls & 1) != 0 : ctx->auto_pack.decided == 1)));
    const bool pack_otf = can_pack && ctx->opt_pack_h2d == 1;
    const auto t_call = std::chrono::steady_clock::now();
    auto probe_done = [&](int rc_) {
        if (!pack_probe || rc_ != HX_OK) return;
        auto &ap = ctx->auto_pack;
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_call).count();
        if (ap.calls >= 2)  // (the first call of either mode allocates and maps its buffers: not timed)
            ap.rate[mixed ? 1 : 0] = std::max(ap.rate[mixed ? 1 : 0], (double)(offsets[n_records] - offsets[0]) / std::max(sec, 1e-9));
        if (++ap.calls == 6) ap.decided = ap.rate[1] > 1.05 * ap.rate[0] ? 1 : 0;
    };
    if (mixed) {
        const int rc_ = run_batch_hybrid(ctx, arena, offsets, n_records, out);
        probe_done(rc_);
        return rc_;
    }
    const bool nib = packed_in || pack_otf;
    const size_t np = ctx->n_pairs;
    const int ndev = (int)ctx->devs.size();
    // Slab size: opt_slab_bytes for large batches (half of it when the slabs are packed on the way: the packed slab is
    // what crosses the link, and a finer pipeline fills sooner); a batch of only a few slabs is cut finer (>= 8 slabs per
    // device, >= 8 MB each) so that the first copy and the last kernel, which nothing overlaps, stay short and every
    // device of the context gets work.
    uint64_t slab_bytes = std::max<uint64_t>((uint64_t)ctx->opt_slab_bytes / (pack_otf ? 2 : 1), 1);
    {
        const uint64_t share = (offsets[n_records] - offsets[0]) / (8ull * ndev) + 1;
        slab_bytes = std::min(slab_bytes, std::max<uint64_t>(share, (uint64_t)8 << 20));
    }
    // ---- cut the batch into slabs (records only; nothing is enqueued yet, so an error can simply return) ----
    std::vector<SlabCut> cuts;
    for (uint32_t r0 = 0; r0 < n_records;) {
        // largest r1 with offsets[r1] - offsets[r0] <= this slab's size (at least one record).  The last slabs of
        // a batch taper (half of what is left per device, >= 8 MB): only the kernels and the result copy of the very
        // last slab are not hidden under a host->device copy, so that slab should be small.
        const uint64_t left = (offsets[n_records] - offsets[r0]) / ndev;
        const uint64_t limit = offsets[r0] + std::min(slab_bytes, std::max<uint64_t>(left / 2, (uint64_t)8 << 20));
        uint32_t r1 = (uint32_t)(std::upper_bound(offsets + r0 + 1, offsets + n_records + 1, limit) - offsets) - 1;
        if (r1 <= r0) r1 = r0 + 1;
        for (uint32_t i = r0; i < r1; ++i)  // cheap sanity: monotone offsets inside the slab edges
            if (offsets[i + 1] < offsets[i]) return fail(ctx, HX_ERR_ARG, "hx_run_batch: offsets must be non-decreasing");
        // the host arena slice is copied with its 16-byte phase preserved so that device addresses keep the alignment of
        // the arena offsets (packed: 32 symbols = 16 bytes; packed on the way: 128 symbols, the staging buffer then takes
        // the packer's 64-byte stores aligned)
        const uint64_t base = offsets[r0] & (pack_otf ? ~127ull : nib ? ~31ull : ~15ull);
        const uint64_t copy_bytes = nib ? (offsets[r1] + 1) / 2 - base / 2 : offsets[r1] - base;
        cuts.push_back({r0, r1, base, copy_bytes});
        r0 = r1;
    }
    const int64_t n_slabs = (int64_t)cuts.size();
    {
        uint64_t h2d = 0;
        for (const SlabCut &c : cuts) h2d += c.copy_bytes + ((uint64_t)(c.r1 - c.r0) + 1) * sizeof(uint64_t) + (nib ? c.r1 - c.r0 : 0);
        ctx->batch_stats[0] = h2d;
        ctx->batch_stats[1] = (uint64_t)n_records * np * sizeof(hx_result);
        ctx->batch_stats[2] = nib ? 1 : 0;
        ctx->batch_stats[3] = (uint64_t)n_slabs;
        ctx->batch_stats[4] = pack_otf ? (uint64_t)pack_threads : 0;
        ctx->batch_stats[5] = nib ? (uint64_t)n_slabs : 0;
        ctx->batch_stats[6] = packed_in ? 3 : pack_otf ? 1 : 0;
    }
    // per (device, slot): the slab it last ran, so its status word can be checked before reuse
    std::vector<int> used(ndev * HX_NSLOT, 0);
    int rc = HX_OK;

    // ---- pack on the way ----
    // Every slab is cut into pieces (one ring buffer each) and every piece into ranges of 128 Ki bases.  The packer threads
    // take the ranges of the WHOLE batch in order from one counter (one fork / join per call) and only wait when the ring
    // buffer of their range is still being read by an earlier copy; this thread issues the copy of a piece as soon as its
    // last range is done, frees ring buffers as their copies complete, and finishes a slab (evidence merge, rec_flags,
    // offsets, kernels, result copy) after its last piece.
    std::vector<PackPiece> pieces;
    std::vector<PackRange> ranges;
    std::vector<int64_t> first_piece;  // per slab
    std::vector<HxhPackPart> parts;
    std::unique_ptr<std::atomic<int>[]> piece_done;  // ranges of the piece that are packed
    std::atomic<int64_t> next_range{0}, allowed_piece{0};
    std::atomic<bool> pack_abort{false};
    std::thread packer;
    PackRing &R = ctx->pack_ring;
    int64_t oldest = 0;  // first piece whose copy is not known to be complete
    auto piece_ev = [&](int64_t g) -> cudaEvent_t { return ctx->devs[pieces[g].slab % ndev].ring_ev[pieces[g].slot]; };
    int64_t issued = 0;  // pieces whose copies are enqueued
    auto poll_ring = [&] {
        while (oldest < issued && cudaEventQuery(piece_ev(oldest)) == cudaSuccess) ++oldest;
        allowed_piece.store(oldest + R.n_buf, std::memory_order_release);
    };
    if (pack_otf) {
        const uint64_t piece_bases = 2ull * (uint64_t)ctx->opt_pack_piece_bytes;  // a multiple of 128
        const uint64_t range_bases = 128u << 10;
        const int n_buf = HX_NRING * ndev;
        for (int64_t j = 0; j < n_slabs; ++j) {
            const SlabCut &c = cuts[j];
            const uint64_t begin = offsets[c.r0], end = offsets[c.r1];
            first_piece.push_back((int64_t)pieces.size());
            for (uint64_t al = c.base;; al += piece_bases) {
                PackPiece P;
                P.slab = j;
                P.al = al;
                P.p = std::max(al, begin);
                P.q = std::max(P.p, std::min(end, al + piece_bases));
                P.dst_off = (al - c.base) / 2;
                P.bytes = P.q > P.al ? (P.q + 1) / 2 - P.al / 2 : 0;
                P.slot = (int)(pieces.size() % (size_t)n_buf);
                P.first_range = (int64_t)ranges.size();
                for (uint64_t x = P.p; x < P.q;) {  // range edges at multiples of 128 (whole vector blocks)
                    const uint64_t y = std::min(P.q, (x + range_bases) & ~(uint64_t)127);
                    ranges.push_back({x, y, (int64_t)pieces.size(), x == begin});
                    x = y;
                }
                P.n_ranges = (int64_t)ranges.size() - P.first_range;
                P.last_of_slab = al + piece_bases >= end;
                pieces.push_back(P);
                if (P.last_of_slab) break;
            }
        }
        parts.resize(ranges.size());
        piece_done.reset(new std::atomic<int>[pieces.size()]);
        for (size_t g = 0; g < pieces.size(); ++g) piece_done[g].store(0, std::memory_order_relaxed);
        // ring + flags + events (first call, or a larger piece size / batch than before)
        const size_t buf_bytes = (size_t)ctx->opt_pack_piece_bytes + 64;
        if (R.n_buf != n_buf || R.buf_bytes != buf_bytes) {
            if (R.p) cudaFreeHost(R.p);
            R.p = nullptr;
            R.n_buf = 0;
            if (cudaHostAlloc((void **)&R.p, buf_bytes * (size_t)n_buf, cudaHostAllocPortable) != cudaSuccess)
                return fail(ctx, HX_ERR_NOMEM, "hx_run_batch: cannot allocate %zu bytes of pinned staging", buf_bytes * (size_t)n_buf);
            R.n_buf = n_buf;
            R.buf_bytes = buf_bytes;
        }
        if (R.flags_cap < (size_t)n_records + 64) {
            if (R.flags) cudaFreeHost(R.flags);
            R.flags = nullptr;
            R.flags_cap = 0;
            const size_t want = (size_t)n_records + (size_t)n_records / 4 + 64;
            if (cudaHostAlloc((void **)&R.flags, want, cudaHostAllocPortable) != cudaSuccess)
                return fail(ctx, HX_ERR_NOMEM, "hx_run_batch: cannot allocate %zu bytes of pinned staging", want);
            R.flags_cap = want;
        }
        for (Device &D : ctx->devs) {
            if ((int)D.ring_ev.size() >= n_buf) continue;
            if (cudaSetDevice(D.dev) != cudaSuccess) return fail(ctx, HX_ERR_CUDA, "hx_run_batch: cudaSetDevice failed");
            while ((int)D.ring_ev.size() < n_buf) {
                cudaEvent_t e = nullptr;
                if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(ctx, HX_ERR_CUDA, "hx_run_batch: cudaEventCreate failed");
                D.ring_ev.push_back(e);
            }
        }
        // records that end before the first base of the batch: no range finalises them
        for (uint32_t r = 0; r < n_records && offsets[r + 1] == offsets[0]; ++r) R.flags[r] = 0;
        allowed_piece.store(n_buf);
        const int workers = std::max(1, pack_threads - 1);  // this thread issues; it polls, so it is not idle
        packer = std::thread([&, workers] {
            hxh_parallel_run(workers, workers, [&](int) {
                for (;;) {
                    const int64_t i = next_range.fetch_add(1, std::memory_order_relaxed);
                    if (i >= (int64_t)ranges.size() || pack_abort.load(std::memory_order_relaxed)) return;
                    const PackRange &rg = ranges[i];
                    while (rg.piece >= allowed_piece.load(std::memory_order_acquire)) {  // the ring buffer is still being read
                        if (pack_abort.load(std::memory_order_relaxed)) return;
                        hx_cpu_relax();
                    }
                    const PackPiece &P = pieces[rg.piece];
                    uint8_t *buf = R.p + (size_t)P.slot * R.buf_bytes;
                    // ordinary stores: the buffer is meant to stay in the cache until the DMA engine has read it
                    hxh_pack_range(arena, offsets, n_records, rg.p, rg.q, (uint8_t *)((uintptr_t)buf - (uintptr_t)(P.al / 2)), R.flags, false,
                                   rg.slab_first, &parts[i]);
                    piece_done[rg.piece].fetch_add(1, std::memory_order_release);
                }
            });
        });
    }
    uint32_t merge_open_rec = UINT32_MAX, merge_open_acc = 0;

    for (int64_t j = 0; j < n_slabs && rc == HX_OK; ++j) {
        const SlabCut &c = cuts[j];
        const uint32_t r0 = c.r0, n = c.r1 - c.r0;
        const uint64_t bytes = offsets[c.r1] - offsets[c.r0];
        const int di = (int)(j % ndev);
        const int si = (int)((j / ndev) % HX_NSLOT);
        Device &D = ctx->devs[di];
        Slot &S = D.slots[si];
        HX_CUDA_BREAK(ctx, rc, cudaSetDevice(D.dev));
        if (used[di * HX_NSLOT + si]) {
            // buffers of this slot may be regrown below: wait for its previous slab and check it
            if (pack_otf) {  // keep freeing ring buffers while waiting
                cudaError_t q;
                while ((q = cudaEventQuery(S.done_ev)) == cudaErrorNotReady) {
                    poll_ring();
                    hx_cpu_relax();
                }
                HX_CUDA_BREAK(ctx, rc, q);
            } else {
                HX_CUDA_BREAK(ctx, rc, cudaStreamSynchronize(S.stream));
            }
            rc = check_status(ctx, S);
            if (rc != HX_OK) break;
        }
        if (pack_otf && !S.done_ev) HX_CUDA_BREAK(ctx, rc, cudaEventCreateWithFlags(&S.done_ev, cudaEventDisableTiming));
        const SlabPlan pl = plan_slab(ctx, n, bytes);
        rc = prepare_slab(ctx, S, pl, n, c.copy_bytes + 64, nib);
        if (rc != HX_OK) break;
        used[di * HX_NSLOT + si] = 1;  // from here on the slot's stream may hold work on caller memory
        if (pack_otf) {
            // piece by piece as the packer threads finish them, then the rec_flags
            cudaError_t e = cudaSuccess;
            for (int64_t g = first_piece[j]; e == cudaSuccess; ++g) {
                const PackPiece &P = pieces[g];
                while (piece_done[g].load(std::memory_order_acquire) < (int)P.n_ranges) {
                    poll_ring();
                    hx_cpu_relax();
                }
                if (P.bytes) e = cudaMemcpyAsync(S.arena.p + P.dst_off, R.p + (size_t)P.slot * R.buf_bytes, P.bytes, cudaMemcpyHostToDevice, S.stream);
                if (e == cudaSuccess) e = cudaEventRecord(piece_ev(g), S.stream);
                issued = g + 1;
                if (P.last_of_slab) break;
            }
            if (e != cudaSuccess) {
                rc = fail(ctx, HX_ERR_CUDA, "hx_run_batch: %s", cudaGetErrorString(e));
                break;
            }
            const PackPiece &P0 = pieces[first_piece[j]], &P1 = pieces[issued - 1];
            hxh_pack_merge(R.flags, parts.data() + P0.first_range, (size_t)(P1.first_range + P1.n_ranges - P0.first_range), merge_open_rec,
                           merge_open_acc);
            HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.flags.p, R.flags + r0, n, cudaMemcpyHostToDevice, S.stream));
        } else {
            const uint8_t *h_src = arena + (nib ? c.base / 2 : c.base);
            if (c.copy_bytes) HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.arena.p, h_src, c.copy_bytes, cudaMemcpyHostToDevice, S.stream));
            if (nib) HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.flags.p, rec_flags + r0, n, cudaMemcpyHostToDevice, S.stream));
        }
        HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(S.offsets.p, offsets + r0, ((size_t)n + 1) * sizeof(uint64_t),
                                               cudaMemcpyHostToDevice, S.stream));
        rc = enqueue_slab(ctx, D, S, pl, S.arena.p, S.offsets.p, c.base, n, S.out.p, S.stream, nib ? S.flags.p : nullptr);
        if (rc != HX_OK) break;
        HX_CUDA_BREAK(ctx, rc, cudaMemcpyAsync(out + (size_t)r0 * np, S.out.p, (size_t)n * np * sizeof(hx_result),
                                               cudaMemcpyDeviceToHost, S.stream));
        if (pack_otf) HX_CUDA_BREAK(ctx, rc, cudaEventRecord(S.done_ev, S.stream));
    }
    if (packer.joinable()) {
        if (rc != HX_OK) pack_abort.store(true);  // threads waiting for a ring buffer leave; otherwise every range is packed
        packer.join();
    }
    // epilogue on every path: drain each (device, slot) stream that was used, so that no copy from the caller's arena
    // or into the caller's table is in flight when this function returns, and the slots are idle for the next call
    for (int di = 0; di < ndev; ++di) {
        cudaSetDevice(ctx->devs[di].dev);
        for (int si = 0; si < HX_NSLOT; ++si) {
            if (!used[di * HX_NSLOT + si]) continue;
            cudaError_t e = cudaStreamSynchronize(ctx->devs[di].slots[si].stream);
            if (e != cudaSuccess && rc == HX_OK) rc = fail(ctx, HX_ERR_CUDA, "hx_run_batch: %s", cudaGetErrorString(e));
            if (rc == HX_OK && e == cudaSuccess) rc = check_status(ctx, ctx->devs[di].slots[si]);
        }
    }
    sync_timed(ctx);
    probe_done(rc);
    return rc;
}

extern "C" {

int hx_run_batch(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out) {
    if (!ctx) return HX_ERR_ARG;
    return run_batch_impl(ctx, arena, nullptr, offsets, n_records, out);
}

int hx_run_batch_packed(hx_ctx *ctx, const uint8_t *packed, const uint8_t *rec_flags, const uint64_t *offsets,
                        uint32_t n_records, hx_result *out) {
    if (!ctx) return HX_ERR_ARG;
    if (n_records && !rec_flags) return fail(ctx, HX_ERR_ARG, "hx_run_batch_packed: NULL rec_flags");
    static const uint8_t none = 0;
    return run_batch_impl(ctx, packed, n_records ? rec_flags : &none, offsets, n_records, out);
}

int64_t hx_find_all_end(hx_ctx *ctx, const char *pattern, uint32_t pattern_len, const uint8_t *text, uint64_t text_len,
                        uint8_t k, int fold_case, uint64_t *out_end, uint8_t *out_dist, uint64_t cap) {
    DeviceGuard device_guard;
    if (!ctx) return HX_ERR_ARG;
    if (!pattern || pattern_len == 0 || pattern_len > 64)
        return fail(ctx, HX_ERR_ARG, "hx_find_all_end: pattern length must be 1..64");
    if (text_len && !text) return fail(ctx, HX_ERR_ARG, "hx_find_all_end: NULL text");
    if (cap && (!out_end || !out_dist)) return fail(ctx, HX_ERR_ARG, "hx_find_all_end: NULL output");
    if (text_len == 0) return 0;
    Device &D = ctx->devs[0];
    HX_CUDA(ctx, cudaSetDevice(D.dev));
    cudaStream_t st = D.slots[0].stream;
    const std::string pat(pattern, pattern_len);
    std::vector<uint64_t> table(256, 0);
    for (int b = 0; b < 256; ++b) {
        const uint8_t u = (fold_case && b >= 'a' && b <= 'z') ? (uint8_t)(b - 32) : (uint8_t)b;
        table[b] = peq_word(pat, u, 64);
    }
    const uint32_t tile_len = ctx->opt_tile_len > 0 ? (uint32_t)ctx->opt_tile_len : 1024u;
    const uint64_t ntiles64 = (text_len + tile_len - 1) / tile_len;
    if (ntiles64 > 0x7fffffffull) return fail(ctx, HX_ERR_ARG, "hx_find_all_end: text too long");
    const uint32_t ntiles = (uint32_t)ntiles64;
    int64_t result = HX_ERR_CUDA;
    uint64_t total = 0;
    const uint64_t ocap = std::min<uint64_t>(cap, text_len);
    DevBuf<uint8_t> b_text, b_dist;
    DevBuf<uint64_t> b_table, b_counts, b_end;
    for (int pass = 0;; ++pass) {  // one context-owned block, carved per call
        DevPool &P = ctx->find_pool;
        P.reset();
        b_text.carve(P, text_len);
        b_table.carve(P, 256);
        b_counts.carve(P, (size_t)ntiles + 1);
        b_end.carve(P, std::max<uint64_t>(ocap, 1));
        b_dist.carve(P, std::max<uint64_t>(ocap, 1));
        if (P.fits()) break;
        if (pass) return fail(ctx, HX_ERR_STATE, "hx_find_all_end: pool did not fit after growing");
        HX_CUDA(ctx, P.grow());
    }
    uint8_t *d_text = b_text.p, *d_dist = b_dist.p;
    uint64_t *d_table = b_table.p, *d_counts = b_counts.p, *d_end = b_end.p;
    cudaError_t e = cudaMemcpyAsync(d_text, text, text_len, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_table, table.data(), 256 * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        ctx->launches += hx_launch_find(d_text, text_len, (const uint8_t *)d_table, pattern_len, k, tile_len, ntiles,
                                        d_counts, d_end, d_dist, ocap, 0, st);
        ctx->launches += hx_launch_find(d_text, text_len, (const uint8_t *)d_table, pattern_len, k, tile_len, ntiles,
                                        d_counts, d_end, d_dist, ocap, 1, st);
        e = cudaMemcpyAsync(&total, d_counts + ntiles, 8, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
        const uint64_t ncopy = std::min<uint64_t>(total, ocap);
        if (ncopy) {
            e = cudaMemcpy(out_end, d_end, ncopy * 8, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(out_dist, d_dist, ncopy, cudaMemcpyDeviceToHost);
        }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) result = (int64_t)total;
    else fail(ctx, HX_ERR_CUDA, "hx_find_all_end: %s", cudaGetErrorString(e));
    return result;
}

void *hx_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void hx_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}

int hx_batch_stats(const hx_ctx *ctx, uint64_t *out, int n) {
    if (!ctx || !out || n < 0) return HX_ERR_ARG;
    for (int i = 0; i < n && i < 7; ++i) out[i] = ctx->batch_stats[i];
    return HX_OK;
}

uint64_t hx_launch_count(const hx_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int hx_group_info(const hx_ctx *ctx, uint32_t *n_groups, uint32_t *steps32_per_base, uint32_t *steps64_per_base) {
    if (!ctx || !ctx->primers_set) return HX_ERR_STATE;
    uint32_t s32 = 0, s64 = 0;
    for (const HxGroup &g : ctx->groups) (g.word64 ? s64 : s32) += (uint32_t)g.np;
    if (n_groups) *n_groups = (uint32_t)ctx->groups.size();
    if (steps32_per_base) *steps32_per_base = s32;
    if (steps64_per_base) *steps64_per_base = s64;
    return HX_OK;
}

int hx_plan_groups(uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len, const char *const *rev,
                   const uint32_t *rev_len, uint8_t k, int fast_small_k, int long_filter, int group_np_max,
                   uint32_t *n_groups, uint32_t *steps32_per_base, uint32_t *steps64_per_base, uint32_t *pair_group) {
    PrimerPlan P;
    const PlanOpts o = {fast_small_k, long_filter, group_np_max, 1, 1};
    std::string err;
    const int rc = plan_primers(o, n_pairs, fwd, fwd_len, rev, rev_len, k, P, err);
    if (rc != HX_OK) return fail(nullptr, rc, "%s", err.c_str());
    uint32_t s32 = 0, s64 = 0;
    for (size_t gi = 0; gi < P.groups.size(); ++gi) {
        const HxGroup &g = P.groups[gi];
        (g.word64 ? s64 : s32) += (uint32_t)g.np;
        if (pair_group)
            for (int q = 0; q < g.npairs; ++q) pair_group[g.pair_global[q]] = (uint32_t)gi;
    }
    if (n_groups) *n_groups = (uint32_t)P.groups.size();
    if (steps32_per_base) *steps32_per_base = s32;
    if (steps64_per_base) *steps64_per_base = s64;
    return HX_OK;
}

int hx_plan_packing(uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len, const char *const *rev,
                    const uint32_t *rev_len, uint8_t k, int pack16, int pack8, uint32_t pair, int role, int alphabet, uint8_t byte,
                    uint32_t *parts_per_word, uint32_t *part_index, uint32_t *m_scan, uint32_t *packed_word) {
    PrimerPlan P;
    const PlanOpts o = {1, 1, 0, pack16, pack8};
    std::string err;
    const int rc = plan_primers(o, n_pairs, fwd, fwd_len, rev, rev_len, k, P, err);
    if (rc != HX_OK) return fail(nullptr, rc, "%s", err.c_str());
    if (pair >= n_pairs || alphabet < 0 || alphabet > 1) return fail(nullptr, HX_ERR_ARG, "hx_plan_packing: bad arguments");
    for (const HxGroup &g : P.groups)
        for (int q = 0; q < g.npairs; ++q) {
            if (g.pair_global[q] != pair) continue;
            const int p = role ? g.pair_r[q] : g.pair_f[q];
            if (!g.packed) {  // the pair's group scans one automaton per word
                if (parts_per_word) *parts_per_word = 1;
                if (part_index) *part_index = (uint32_t)p;
                if (m_scan) *m_scan = g.m[p];
                if (packed_word) *packed_word = 0;
                return HX_OK;
            }
            const int parts = g.pk_parts, nw = (g.np + parts - 1) / parts, stride = hx_row_stride(nw, 4);
            uint32_t w = 0;
            memcpy(&w, P.blob_pk.data() + g.pk_table_off + (size_t)alphabet * hx_table_bytes(nw, 4) + (size_t)byte * stride + (size_t)(p / parts) * 4, 4);
            if (parts_per_word) *parts_per_word = (uint32_t)parts;
            if (part_index) *part_index = (uint32_t)p;
            if (m_scan) *m_scan = g.pk_m[p];
            if (packed_word) *packed_word = w;
            return HX_OK;
        }
    return fail(nullptr, HX_ERR_STATE, "hx_plan_packing: pair not found in any group");
}

int hx_plan_peq_word(uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len, const char *const *rev,
                     const uint32_t *rev_len, uint8_t k, int fast_small_k, int long_filter, int group_np_max, uint32_t pair,
                     int role, int alphabet, uint8_t byte, uint64_t *scan_word, uint64_t *wm_word, uint64_t *verify_word,
                     uint32_t *word_bits, uint32_t *m_scan, uint32_t *m_full) {
    PrimerPlan P;
    const PlanOpts o = {fast_small_k, long_filter, group_np_max, 1, 1};
    std::string err;
    const int rc = plan_primers(o, n_pairs, fwd, fwd_len, rev, rev_len, k, P, err);
    if (rc != HX_OK) return fail(nullptr, rc, "%s", err.c_str());
    if (pair >= n_pairs || alphabet < 0 || alphabet > 1) return fail(nullptr, HX_ERR_ARG, "hx_plan_peq_word: bad arguments");
    for (const HxGroup &g : P.groups)
        for (int q = 0; q < g.npairs; ++q) {
            if (g.pair_global[q] != pair) continue;
            const int p = role ? g.pair_r[q] : g.pair_f[q];
            const int wb = g.word64 ? 8 : 4, stride = hx_row_stride(g.np, wb);
            const size_t off = g.table_off + (size_t)alphabet * hx_table_bytes(g.np, wb) + (size_t)byte * stride + (size_t)p * wb;
            uint64_t w = 0, ww = 0;
            memcpy(&w, P.blob.data() + off, wb);
            memcpy(&ww, P.blob_wm.data() + off, wb);
            if (scan_word) *scan_word = w;
            if (wm_word) *wm_word = ww;
            if (verify_word) *verify_word = (g.longmask >> p & 1u) ? P.vblob[((size_t)g.vslot[p] * 2 + alphabet) * 256 + byte] : 0;
            if (word_bits) *word_bits = (uint32_t)wb * 8;
            if (m_scan) *m_scan = g.m[p];
            if (m_full) *m_full = g.mfull[p];
            return HX_OK;
        }
    return fail(nullptr, HX_ERR_STATE, "hx_plan_peq_word: pair not placed");
}

int hx_kernel_time(hx_ctx *ctx, double *myers_ms, uint64_t *myers_launches, int reset) {
    DeviceGuard device_guard;
    if (!ctx) return HX_ERR_ARG;
    sync_timed(ctx);
    if (myers_ms) *myers_ms = ctx->myers_ms;
    if (myers_launches) *myers_launches = ctx->myers_launches;
    if (reset) {
        ctx->myers_ms = 0;
        ctx->myers_launches = 0;
    }
    return HX_OK;
}

}  // extern "C"

// ---- internal (C++ linkage, hx_host.h): used by the file-level driver in hx_host.cpp ----------------------------------
int hxh_device_count(const hx_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }
int hxh_lane_count(const hx_ctx *ctx) { return ctx ? (int)ctx->devs.size() * HX_NLANE : 0; }

int hxh_run_text_lane(hx_ctx *ctx, int lane, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                      hx_result *out, const char *ids, const uint64_t *id_off, const char *fa_tails,
                      const uint32_t *fa_tail_off, const char *gff_tails, const uint32_t *gff_tail_off,
                      const HxhTextSink *sink) {
    return run_text_lane(ctx, lane, arena, offsets, n_records, out, ids, id_off, fa_tails, fa_tail_off, gff_tails,
                         gff_tail_off, nullptr, nullptr, nullptr, nullptr, sink);
}

// result table only, on the lane's own device and stream (the host-formatter path of the file-level driver)
int hxh_run_batch_lane(hx_ctx *ctx, int lane, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                       hx_result *out) {
    DeviceGuard device_guard;
    if (!ctx) return HX_ERR_ARG;
    if (!ctx->primers_set) return fail(ctx, HX_ERR_STATE, "hx_run_batch: call hx_set_primers first");
    if (lane < 0 || lane >= (int)ctx->devs.size() * HX_NLANE) return fail(ctx, HX_ERR_ARG, "hx_run_batch: bad lane");
    if (n_records == 0) return HX_OK;
    if (!offsets || !out || offsets[n_records] < offsets[0] || (offsets[n_records] > offsets[0] && !arena))
        return fail(ctx, HX_ERR_ARG, "hx_run_batch: bad arguments");
    Device &D = ctx->devs[lane / HX_NLANE];
    Slot &S = D.lanes[lane % HX_NLANE].slot;
    HX_CUDA(ctx, cudaSetDevice(D.dev));
    HX_CUDA(ctx, cudaStreamSynchronize(S.stream));
    const uint64_t base = offsets[0] & ~15ull, copy_bytes = offsets[n_records] - base;
    const SlabPlan pl = plan_slab(ctx, n_records, offsets[n_records] - offsets[0]);
    int rc = prepare_slab(ctx, S, pl, n_records, copy_bytes + 64);
    if (rc != HX_OK) return rc;
    if (copy_bytes) HX_CUDA(ctx, cudaMemcpyAsync(S.arena.p, arena + base, copy_bytes, cudaMemcpyHostToDevice, S.stream));
    HX_CUDA(ctx, cudaMemcpyAsync(S.offsets.p, offsets, ((size_t)n_records + 1) * 8, cudaMemcpyHostToDevice, S.stream));
    rc = enqueue_slab(ctx, D, S, pl, S.arena.p, S.offsets.p, base, n_records, S.out.p, S.stream);
    if (rc != HX_OK) return rc;
    HX_CUDA(ctx, cudaMemcpyAsync(out, S.out.p, (size_t)n_records * ctx->n_pairs * sizeof(hx_result), cudaMemcpyDeviceToHost,
                                 S.stream));
    HX_CUDA(ctx, cudaStreamSynchronize(S.stream));
    return check_status(ctx, S);
}

void **hxh_host_cache(hx_ctx *ctx, void (*free_fn)(void *)) {
    if (!ctx) return nullptr;
    ctx->host_cache_free = free_fn;
    return &ctx->host_cache;
}

void hxh_set_error(hx_ctx *ctx, const char *msg) {
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err = msg ? msg : "";
}
