// hx_host.h — host-side pieces of the product that are not device code: the reference's
// primer/region tables, IUPAC match predicate, FASTA ingest and the FASTA/GFF3 writers.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

struct hx_ctx;
int hxh_device_count(const hx_ctx *ctx);  // hx_api.cu

// utils.rs:251-263 + rust-bio build_64: does pattern symbol p match text byte b?
bool hxh_symbol_matches(uint8_t p, uint8_t b);

// rust-bio fasta::Reader rules (call sites utils.rs:238,270) over a decompressed stream.
struct HxhRecordBatch {
    uint8_t *arena = nullptr;      // pinned when possible (hx_alloc_pinned), else malloc
    size_t arena_cap = 0;
    bool pinned = false;
    std::vector<uint64_t> offsets; // n + 1
    std::vector<std::string> ids;
    void clear();
    ~HxhRecordBatch();
};

class HxhByteSource;  // plain / gzip / bzip2 / xz / zstd byte stream (utils.rs:148-154)

class HxhFastaReader {
  public:
    HxhFastaReader();
    ~HxhFastaReader();
    // returns 0 or a negative hx_status; err receives the message
    int open(const char *path, std::string &err);
    // appends records until the arena holds >= max_bytes or max_records; returns the number of
    // records appended, 0 at the end of the input.
    size_t next_batch(HxhRecordBatch &batch, size_t max_bytes, size_t max_records);

  private:
    bool next_line(const uint8_t *&line, size_t &len);  // zero-copy view into buf_, '\n' excluded
    HxhByteSource *src_ = nullptr;
    std::vector<uint8_t> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
    std::string header_;  // look-ahead header line (rust-bio keeps the next header in self.line)
    bool have_line_ = false;
    size_t size_hint_ = 0;
    bool finished_ = false;
};
