// hx_host.h — host-side pieces of the product that are not device code: the reference's
// primer/region tables, IUPAC match predicate, FASTA ingest and the FASTA/GFF3 writers.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

struct hx_ctx;
struct hx_result;

// utils.rs:251-263 + rust-bio build_64: does pattern symbol p match text byte b?
bool hxh_symbol_matches(uint8_t p, uint8_t b);

// ---- internal interface between the file-level driver (hx_host.cpp) and the device runtime (hx_api.cu) ----------
// Sink of the streamed text path.  begin (may be NULL) receives the FASTA / GFF3 byte totals of the batch before the
// first chunk; chunk receives `n` bytes of text `kind` (0 FASTA, 1 GFF3) that belong at byte `off` of that kind's text
// for this batch, valid only during the call.  Non-zero return aborts the batch with HX_ERR_IO.
struct HxhTextSink {
    int (*begin)(void *user, uint64_t fa_total, uint64_t gff_total);
    int (*chunk)(void *user, int kind, uint64_t off, const char *bytes, size_t n);
    void *user;
};
int hxh_device_count(const hx_ctx *ctx);
int hxh_lane_count(const hx_ctx *ctx);  // text lanes of the context: devices x lanes per device
// hx_run_batch_text_stream on text lane `lane` (its own device, stream, buffers); lanes may run concurrently.  `out`
// may be NULL when the caller does not need the result table on the host.
int hxh_run_text_lane(hx_ctx *ctx, int lane, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                      hx_result *out, const char *ids, const uint64_t *id_off, const char *fa_tails,
                      const uint32_t *fa_tail_off, const char *gff_tails, const uint32_t *gff_tail_off,
                      const HxhTextSink *sink);
// hx_run_batch (result table only) on the lane's device
int hxh_run_batch_lane(hx_ctx *ctx, int lane, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                       hx_result *out);
// The packer's building blocks (hx_host.cpp): one range of bases [p, q) — edges at multiples of 128, or where a slab or the
// batch begins / ends — packed into packed_virtual[i / 2] for base i; ranges are independent of each other (any thread,
// any order), the per-record evidence is completed by merging the parts in range order.
struct HxhPackPart {
    int64_t first_final = -1;  // first record this range finalised (-1: none, -2: empty range)
    uint32_t open_rec = 0, open_acc = 0;
};
void hxh_pack_range(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint64_t p, uint64_t q, uint8_t *packed_virtual,
                    uint8_t *rec_flags, bool stream_stores, bool clear_low_nibble, HxhPackPart *part);
void hxh_pack_merge(uint8_t *rec_flags, const HxhPackPart *parts, size_t n, uint32_t &open_rec, uint32_t &open_acc);
// fn(i) for i in [0, n) on up to `threads` threads of the process-wide pool (the caller is one of them)
void hxh_parallel_run(int n, int threads, const std::function<void(int)> &fn);
// hx_pack_records (include/hyperex_b200.h) in pieces, for a caller that ships the packed bytes while they are produced
// (hx_run_batch with pack_h2d): see hx_host.cpp
struct HxhPackJob {
    const uint8_t *arena = nullptr;
    const uint64_t *offsets = nullptr;
    uint32_t n_records = 0;
    uint8_t *rec_flags = nullptr;
    uint64_t begin = 0, end = 0;      // base range of the records
    uint32_t open_rec = 0, open_acc = 0;  // the record the last piece left open and its evidence so far
};
int hxh_pack_begin(HxhPackJob &J, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint8_t *rec_flags);
void hxh_pack_piece(HxhPackJob &J, uint64_t p, uint64_t q, uint8_t *packed_virtual, int n_threads, bool stream_stores);
// ... and in one go, into a buffer that starts at packed byte packed_base / 2 of the batch
int hxh_pack_slab(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint8_t *packed, uint64_t packed_base,
                  uint8_t *rec_flags, int n_threads);
// slot for host-side state the driver keeps in the context across calls (freed with free_fn by hx_destroy)
void **hxh_host_cache(hx_ctx *ctx, void (*free_fn)(void *));
void hxh_set_error(hx_ctx *ctx, const char *msg);

// ---- FASTA ingest: rust-bio fasta::Reader rules (call sites utils.rs:238,270) over a decompressed stream ----------
struct HxhRecordBatch {
    uint8_t *arena = nullptr;      // sequence bytes back to back; pinned when possible (hx_alloc_pinned), else malloc
    size_t arena_cap = 0;
    bool pinned = false;
    std::vector<uint64_t> offsets; // n + 1
    std::vector<char> ids;         // record ids back to back, no separators
    std::vector<uint64_t> id_off;  // n + 1
    std::vector<std::string> descs; // only filled when the parser is asked to keep descriptions
    std::vector<uint8_t> has_desc;  // ... 0: the header had no description (rust-bio: desc == None), 1: it had (maybe "")
    bool stop = false;             // the reader's iteration ends after this batch (malformed / empty record)
    size_t n_records() const { return offsets.size() - 1; }
    void clear();
    bool reserve(size_t bytes);    // arena capacity; false when the allocation fails
    ~HxhRecordBatch();
};

// Parses FASTA text [p, p + n) that starts at a record start (or at the start of the input) into `b`.
// drop_last: the byte stream ended with a read error right after these bytes — the record in progress is lost, like in
// rust-bio where the failing read_line fails the whole record.  Returns 0, or HX_ERR_NOMEM; b.stop is set when the
// reader's iteration ends inside this text ("Expected > at record start." / an empty record).
int hxh_parse_fasta_text(const uint8_t *p, size_t n, bool drop_last, bool keep_desc, HxhRecordBatch &b);

class HxhByteSource;  // plain / gzip / bzip2 / xz / zstd byte stream (utils.rs:148-154)

// Cuts the (decompressed) input into chunks that end at record boundaries — a line that starts with '>' always starts
// a new record in rust-bio's reader — so that chunks can be parsed independently, on parallel threads, and their
// batches simply follow each other in input order.  A plain file is memory-mapped and chunks are views into the
// mapping; compressed input is decoded into the caller's buffer.
class HxhChunker {
  public:
    HxhChunker();
    ~HxhChunker();
    int open(const char *path, std::string &err);  // 0 or a negative hx_status (niffler: file too short to sniff, ...)
    // Next chunk of about `target` bytes (at least one whole record).  Returns false at the end of the input.  `buf` is
    // scratch owned by the caller (used for compressed input); [p, p + n) stays valid until the next call with the
    // same `buf` (forever for a mapped file).  ended_by_error: see hxh_parse_fasta_text(drop_last).
    bool next(size_t target, std::vector<uint8_t> &buf, const uint8_t *&p, size_t &n, bool &ended_by_error);
    bool mapped() const { return map_ != nullptr; }
    size_t mapped_bytes() const { return map_len_; }

  private:
    HxhByteSource *src_ = nullptr;
    const uint8_t *map_ = nullptr;
    size_t map_len_ = 0, map_pos_ = 0;
    std::vector<uint8_t> carry_;  // bytes after the last record boundary of the previous chunk
    bool eof_ = false, error_ = false;
};
