/*
 * hyperex_b200.h — C ABI of the B200-native HyperEx primer-search-and-extract hot path.
 *
 * Drop-in boundary.  The reference (Ebedthan/hyperex v0.2.0) has no FFI;
 * its seam is the Rust call
 *     utils::get_hypervar_regions(file, primers, prefix, mismatch)     src/utils.rs:231-236
 * (called once from src/main.rs:38) and, inside it, the rust-bio 1.6.0 operator API
 *     MyersBuilder::new/.ambig/.build_64        src/utils.rs:265-268, 301-302
 *     Myers<u64>::find_all_lazy(text, k)        src/utils.rs:305-309
 * Every entry point below names the reference lines it replaces.  All functions are
 * `extern "C"`, take plain pointers and sizes, never throw or unwind, and return 0 (HX_OK)
 * or a negative hx_status; hx_last_error() gives the message.  There is NO CPU fallback:
 * every compute entry point fails with HX_ERR_CUDA when no CUDA device is usable.
 */
#ifndef HYPEREX_B200_H
#define HYPEREX_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HX_ABI_VERSION 1

typedef enum {
    HX_OK = 0,
    HX_ERR_ARG = -1,      /* bad argument (NULL, zero-length / >64-nt primer, unsorted offsets ...) */
    HX_ERR_CUDA = -2,     /* CUDA runtime error or no device */
    HX_ERR_NOMEM = -3,    /* host or device allocation failed */
    HX_ERR_STATE = -4,    /* call order (no primers set, ...) */
    HX_ERR_IO = -5,       /* file errors in the host driver */
    HX_ERR_CAPACITY = -6  /* an output buffer was too small; never silently truncates */
} hx_status;

/* flags of hx_result */
#define HX_FWD_ANY 1u        /* forward_matches non-empty           (utils.rs:312) */
#define HX_REV_ANY 2u        /* reverse_matches non-empty           (utils.rs:313) */
#define HX_PAIR_VALID 4u     /* best_pair is Some                   (utils.rs:328-351) */
#define HX_ALPHABET_NONE 8u  /* sequence_type == None, record skipped (utils.rs:273-280) */
#define HX_RNA 16u           /* record classified as RNA            (utils.rs:214-215) */

/* One (record, primer pair) outcome: everything utils.rs:353-409 needs to write the FASTA
 * record, the GFF line or one of the four warnings. */
typedef struct hx_result {
    uint32_t f_start;  /* 0-based region start = f_end - (len(forward)-1)   (utils.rs:334) */
    uint32_t r_end;    /* 0-based inclusive region end                      (utils.rs:348) */
    uint8_t flags;     /* HX_* bits above */
    uint8_t combined;  /* f_dist + r_dist of the chosen pair                (utils.rs:342) */
    uint16_t reserved; /* 0 */
} hx_result;

typedef struct hx_ctx hx_ctx;

/* ---- context ------------------------------------------------------------------------ */
/* Creates a context driving n_dev CUDA devices (dev_ids may be NULL => devices 0..n_dev-1).
 * Records of a batch are sharded over the devices with no collective. */
int hx_create(hx_ctx **out, const int *dev_ids, int n_dev);
void hx_destroy(hx_ctx *ctx);
const char *hx_last_error(const hx_ctx *ctx); /* ctx may be NULL: last creation error */
int hx_abi_version(void);

/* Replaces MyersBuilder::new + .ambig x11 (utils.rs:251-268), to_reverse_complement
 * (utils.rs:299) and build_64 x2 (utils.rs:301-302) for every primer pair, once per run
 * instead of once per record x pair.  Builds IUPAC Peq tables for fwd, rc_dna(rev) and
 * rc_rna(rev).  A primer of length 0 or > 64 is HX_ERR_ARG (the reference aborts inside
 * rust-bio).  k is `mismatch` of utils.rs:235 (unit-cost edit distance bound). */
int hx_set_primers(hx_ctx *ctx, uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len,
                   const char *const *rev, const uint32_t *rev_len, uint8_t k);

/* Tuning knobs (tests use them to force window seams). name: "tile_len" (owned bytes per
 * window of a long record), "single_max" (records up to this length are one tile),
 * "slab_bytes" (host->device pipeline granule), "time_kernels" (CUDA-event timing of the scan),
 * "text_tma" (1: stage the text through shared memory with TMA bulk copies; 0, default: per-thread
 * 16-byte vector loads, measured 3 % faster), "fast_small_k" (1, default: for -m 0..2 and primers <= 32 nt scan
 * with Wu-Manber / Shift-Or automata, which accept exactly the same hits with the same distances in fewer ALU
 * operations; 0: always the Myers automaton), "long_filter" (read by hx_set_primers; 1, default: for -m <= 8 a
 * primer of 33..64 nt is scanned as the 32-bit automaton of its last 32 symbols and every candidate end is verified
 * against the whole primer on the device, which accepts exactly the same hits with the same distances — unless those
 * 32 symbols are so degenerate (mostly N) that too many positions would be candidates; 2: filter those too; 0: such
 * primers are stepped as 64-bit automata over the whole text), "group_np_max" (read by hx_set_primers; cap on the
 * automata one scan thread steps per text byte, 0 = automatic: 16 with 32-bit words, 8 with 64-bit words and on the
 * Wu-Manber path at -m 2), "pack16" / "pack8" (read by hx_set_primers; pattern packing of the -m 0..2 path: automata of
 * the primers' last symbols share a 32-bit word and every candidate end is verified against the whole primer on the
 * device, identical hits; pack16: 0 = off, 1, default = two <= 16-symbol automata per word when every suffix is specific
 * enough, 2 / 3 = force two / four per word; pack8: 1, default = at -m 0 four <= 8-symbol automata per word when every
 * 8-symbol suffix is specific enough (>= 12 bits), 0 = never), "pack_h2d" (hx_run_batch only: 1 = the host threads pack every slab of the caller's ASCII
 * arena to 4 bits per base while it is on its way to the device — see hx_pack_records for the code — so half the bytes
 * cross PCIe and the device runs its packed-text kernels; 2 = mixed: ASCII slabs from the front of the batch keep the link
 * busy while the host threads pack slabs from its back, the two sides meet where this host makes them meet (link rate +
 * half the packers' rate instead of the larger of the two); identical results in every mode; 0 = ASCII slabs; -1,
 * default = automatic: measured — the first six batches >= 64 MB of a context with >= 2 packer threads per device
 * alternate ASCII and mixed, the last four are timed, and the faster mode (mixed only when >= 5 % faster) serves every
 * later call; setting the option again, or another primer set, measures again; never for a primer set that packed text cannot represent), "pack_threads" (host threads of that path, 0 = all hardware threads; a process that shares the host
 * with other ranks passes its share), "pack_piece_bytes" (size of the pinned ring buffers the packed text goes through,
 * default 4 MiB: small enough to stay in the last-level cache between the packer's stores and the DMA reads). */
int hx_set_option(hx_ctx *ctx, const char *name, int64_t value);

/* Replaces the record loop body utils.rs:270-351 (sequence_type, to_ascii_uppercase, both
 * find_all_lazy scans and the pairing loop) for a batch of records held in HOST memory:
 * record i = arena[offsets[i] .. offsets[i+1]).  out has n_records * n_pairs entries,
 * record-major, pairs in hx_set_primers order (== the reference's output order).
 * Internally: slabs are copied host->device on alternating streams while the previous slab
 * is being scanned; results are copied back in input order.  Synchronous for the caller. */
int hx_run_batch(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                 hx_result *out);

/* hx_run_batch plus on-device extraction: besides the result table, the FASTA text that
 * utils.rs:354-369 writes for the batch is assembled on the GPU and returned in pinned host memory owned by the
 * context (valid until the next call on ctx), so the host only has to write() it.
 *   ids / id_off[n_records + 1]   record ids, concatenated (no separators)
 *   tails / tail_off[n_pairs + 1] per pair the text that follows the id, up to and including the '\n'
 *                                 (" region=v3v4 forward=... reverse=...\n", utils.rs:356-363)
 * Entries appear in output order (record-major, pair order); a (record, pair) without a region contributes
 * nothing.  Runs on the context's first device. */
int hx_run_batch_fasta(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out,
                       const char *ids, const uint64_t *id_off, const char *tails, const uint32_t *tail_off,
                       const char **text, uint64_t *text_len);

/* hx_run_batch_fasta plus the GFF3 lines of utils.rs:378-385 ("{id}\thyperex\tregion\t{f_start+1}\t{r_end+1}" followed
 * by the per-pair gff_tails entry "\t.\t.\t.\tNote=...\n"), also assembled on the GPU.  The "##gff-version 3" header
 * (utils.rs:247) is not part of the text. */
int hx_run_batch_text(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, hx_result *out,
                      const char *ids, const uint64_t *id_off, const char *fa_tails, const uint32_t *fa_tail_off,
                      const char *gff_tails, const uint32_t *gff_tail_off, const char **fa_text, uint64_t *fa_len,
                      const char **gff_text, uint64_t *gff_len);

/* hx_run_batch_text with the text delivered in order through a callback instead of one pinned buffer: first the
 * FASTA bytes (kind 0), then the GFF3 bytes (kind 1), in chunks of at most 16 MiB that are valid only during the
 * call.  The device->host copy of the next chunk runs while the sink consumes the current one, so a sink that
 * write()s (what fasta::Writer / the GFF File do at utils.rs:365-369, 378-385) overlaps with the transfer.  A
 * non-zero return from the sink aborts with HX_ERR_IO.  text_len / gff_len receive the totals (may be NULL). */
typedef int (*hx_text_sink)(void *user, int kind, const char *bytes, size_t n);
int hx_run_batch_text_stream(hx_ctx *ctx, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records,
                             hx_result *out, const char *ids, const uint64_t *id_off, const char *fa_tails,
                             const uint32_t *fa_tail_off, const char *gff_tails, const uint32_t *gff_tail_off,
                             hx_text_sink sink, void *user, uint64_t *text_len, uint64_t *gff_len);

/* ---- packed arena (BASELINE.json north_star: "2-bit/ASCII arena in pinned memory") ------------------------------
 * Four bits per base, two bases per byte: nibble i of `packed` is base i of the arena (low nibble = even index), so
 * the same offsets array describes both.  Codes: 0 = any byte outside the 16 letters, 1..15 = A C G T|U R Y S W K M B
 * D H V N after to_ascii_uppercase (utils.rs:295).  'T' and 'U' share a code: a record that holds both is alphabet
 * None and is skipped (utils.rs:225-226), so within a scanned record the code is unambiguous.  rec_flags[r] receives
 * the evidence utils.rs:207-228 (sequence_type) needs, HX_REC_* below, so the device does not have to look at the
 * text again to classify it.  Pure host code (AVX2 when the CPU has it), n_threads <= 0: all hardware threads.
 * packed must hold offsets[n_records] / 2 + 1 bytes (it is indexed like the arena, from offset 0). */
#define HX_REC_HAS_U 1u
#define HX_REC_HAS_T 2u
#define HX_REC_NON_IUPAC 4u
int hx_pack_records(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint8_t *packed,
                    uint8_t *rec_flags, int n_threads);

/* hx_run_batch on a packed arena: half the bytes cross PCIe, the device scans nibbles and skips its own
 * classification pass.  Same result table, bit for bit.  A byte outside the 16 letters can only ever match a primer
 * symbol that is that very byte; a primer set with such symbols (anything but ACGTURYSWKMBDHVN) cannot be matched
 * on packed text and the call fails with HX_ERR_STATE — use hx_run_batch. */
int hx_run_batch_packed(hx_ctx *ctx, const uint8_t *packed, const uint8_t *rec_flags, const uint64_t *offsets,
                        uint32_t n_records, hx_result *out);

/* Same, for an arena already resident in device memory of the context's device `dev_slot`
 * (index into dev_ids).  d_arena must be 16-byte aligned and readable up to the next
 * multiple of 16 bytes past offsets[n_records]; d_out is device memory.  Work is enqueued on
 * `stream` (a cudaStream_t, may be NULL for the context's own stream) and is asynchronous
 * unless stream is NULL. */
int hx_run_batch_device(hx_ctx *ctx, int dev_slot, const uint8_t *d_arena, const uint64_t *d_offsets,
                        uint32_t n_records, uint64_t arena_bytes, hx_result *d_out, void *stream);

/* Known-answer entry point mirroring Myers<u64>::find_all_lazy(text, k).collect()
 * (utils.rs:305-309) 1:1 on the GPU: every (end 0-based inclusive, dist) with dist <= k,
 * ascending.  The pattern is used as given (no reverse complement); IUPAC table of
 * utils.rs:251-263; text bytes are matched as given when fold_case == 0 and upper-cased
 * (utils.rs:295) otherwise.  Returns the total number of hits (>= 0) or a negative
 * hx_status; if the total exceeds cap only the first cap hits are stored. */
int64_t hx_find_all_end(hx_ctx *ctx, const char *pattern, uint32_t pattern_len, const uint8_t *text,
                        uint64_t text_len, uint8_t k, int fold_case, uint64_t *out_end, uint8_t *out_dist,
                        uint64_t cap);

/* pinned host memory for arenas / result tables (cudaHostAlloc) */
void *hx_alloc_pinned(size_t bytes);
void hx_free_pinned(void *p);

/* how hx_set_primers packed the primer set: pattern groups (= scan launches per slab) and the automata stepped
 * per text byte with 32-bit / 64-bit words (de-duplicated patterns) */
int hx_group_info(const hx_ctx *ctx, uint32_t *n_groups, uint32_t *steps32_per_base, uint32_t *steps64_per_base);

/* The same packing computed on the host alone (no CUDA, no context): what hx_set_primers would do with this primer set,
 * -m value and these options ("fast_small_k", "long_filter", "group_np_max" of hx_set_option).  pair_group (n_pairs
 * entries, may be NULL) receives the group of every pair.  Used by the CPU test-suite and for capacity planning. */
int hx_plan_groups(uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len, const char *const *rev,
                   const uint32_t *rev_len, uint8_t k, int fast_small_k, int long_filter, int group_np_max,
                   uint32_t *n_groups, uint32_t *steps32_per_base, uint32_t *steps64_per_base, uint32_t *pair_group);

/* One entry of the tables that plan would upload (host only; test-suite): for primer `role` (0 forward, 1 reverse) of
 * `pair`, alphabet (0 dna, 1 rna) and raw text byte — the Peq word of the scanned automaton (top-aligned in word_bits
 * bits; a primer longer than the word is scanned as its last word_bits symbols: m_scan < m_full), its complemented
 * Wu-Manber form, and the 64-bit word of the whole primer used by the verification (0 when the primer is not filtered). */
int hx_plan_peq_word(uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len, const char *const *rev,
                     const uint32_t *rev_len, uint8_t k, int fast_small_k, int long_filter, int group_np_max, uint32_t pair,
                     int role, int alphabet, uint8_t byte, uint64_t *scan_word, uint64_t *wm_word, uint64_t *verify_word,
                     uint32_t *word_bits, uint32_t *m_scan, uint32_t *m_full);

/* How that plan packs the small-k automata (host only; test-suite): for primer `role` of `pair` under options "pack16" /
 * "pack8" — the automata per 32-bit word of its group's packed table (1: the group is not packed, 2, or 4 at -m 0), the
 * primer's part index in the group (part p lives in bits [w * (p % parts), w * (p % parts + 1)) of word p / parts, w = 32 /
 * parts), the symbols scanned (the primer's last w symbols when it is longer), and the packed table word of text byte
 * `byte`: complemented Peq bits, top-aligned inside every part, 0 in a part's unused low bits. */
int hx_plan_packing(uint32_t n_pairs, const char *const *fwd, const uint32_t *fwd_len, const char *const *rev,
                    const uint32_t *rev_len, uint8_t k, int pack16, int pack8, uint32_t pair, int role, int alphabet, uint8_t byte,
                    uint32_t *parts_per_word, uint32_t *part_index, uint32_t *m_scan, uint32_t *packed_word);

/* number of kernels the context has launched so far (bench.py's gpu_launches) */
uint64_t hx_launch_count(const hx_ctx *ctx);
/* device time (ms) of the dominant kernel (hx_myers_pair_kernel) accumulated by CUDA events
 * on the launching stream since the last reset, and the number of launches timed; only
 * collected after hx_set_option(ctx, "time_kernels", 1). */
int hx_kernel_time(hx_ctx *ctx, double *myers_ms, uint64_t *myers_launches, int reset);

/* ---- host helpers that shape the output bytes (pure host code) ---------------------- */
/* utils.rs:207-228: 0 dna, 1 rna, 2 none */
int hx_sequence_type(const uint8_t *seq, size_t n);
/* utils.rs:169-199: alphabet 0 dna / 1 rna; out holds n bytes */
void hx_to_complement(const char *primer, size_t n, int alphabet, char *out);
void hx_to_reverse_complement(const char *primer, size_t n, int alphabet, char *out);
/* utils.rs:157-166 (+29-45): label of a primer pair, "" when unknown; out holds >= 8 bytes */
void hx_primers_to_region(const char *fwd, const char *rev, char *out);
/* utils.rs:112-131: 0 and the two primers, or HX_ERR_ARG for an unsupported name */
int hx_region_to_primer(const char *region, const char **fwd, const char **rev);
/* utils.rs:23-25 */
const char *const *hx_supported_regions(uint32_t *n);
/* utils.rs:134-145: parses a primer CSV; returns n_pairs >= 0 or HX_ERR_IO / HX_ERR_ARG
 * ("Primer file must be comma-separated").  Strings are owned by the library until
 * hx_free_primer_file. */
int hx_file_to_vec(const char *path, char ***fwd, char ***rev);
void hx_free_primer_file(char **fwd, char **rev, int n);

/* ---- FASTA ingest -> arena packer (replaces fasta::Reader::new(reader).records(), utils.rs:237-238,270)
 * rust-bio reader rules: id = header up to the first whitespace, sequence lines joined with
 * trim_end(), iteration stops at the first malformed or empty record; input may be plain, gzip,
 * bzip2, xz or zstd (niffler sniffing, utils.rs:148-154).  hx_fasta_next packs the records of the next
 * ~max_bytes of FASTA text (always whole records, at least one; 0 = 256 MiB) into a pinned arena owned by the
 * reader (valid until the next call / close) and returns the number of records (0 at the end of the
 * input, negative hx_status on error, e.g. HX_ERR_NOMEM). */
typedef struct hx_fasta hx_fasta;
int hx_fasta_open(const char *path, hx_fasta **out);
/* Before the first hx_fasta_next*: parse (and, with pack_ahead != 0, pack to 4 bits per base) the chunks ahead of the
 * caller on n_threads background threads (<= 0: hardware threads - 1, at most 8; 1, the default: everything happens on the
 * calling thread inside hx_fasta_next).  Batches still arrive in input order and are byte for byte the ones the single
 * thread returns for the same max_bytes; the reader then owns n_threads + 2 arenas of about max_bytes each (kept for
 * the next reader of the process after hx_fasta_close, up to HYPEREX_READER_CACHE_MB = 1024 MiB: pinning costs more
 * than parsing).
 * HX_ERR_STATE once reading has begun. */
int hx_fasta_configure(hx_fasta *r, int n_threads, int pack_ahead);
int64_t hx_fasta_next(hx_fasta *r, size_t max_bytes, const uint8_t **arena, const uint64_t **offsets,
                      const char *const **ids);
/* hx_fasta_next that hands the batch over as the packed arena of hx_run_batch_packed (north_star: the producer emits
 * the packed arena): *packed = 4-bit codes (offsets[n] / 2 + 1 bytes, pinned), *rec_flags = the HX_REC_* evidence of
 * every record — exactly what hx_pack_records makes of the batch's ASCII arena, which stays available through *arena
 * (may be NULL) for a caller that extracts region text.  Packed by the background threads when configured with
 * pack_ahead, otherwise inside the call. */
int64_t hx_fasta_next_packed(hx_fasta *r, size_t max_bytes, const uint8_t **packed, const uint8_t **rec_flags,
                             const uint64_t **offsets, const char *const **ids, const uint8_t **arena);
/* descriptions of the records of the last hx_fasta_next batch: the header text after the first whitespace, NULL for
 * a header without one (rust-bio Record::desc; fasta::Writer::write_record prints ">id desc", utils.rs:447-449) */
const char *const *hx_fasta_descs(hx_fasta *r);
void hx_fasta_close(hx_fasta *r);

/* ---- the reference's seam itself ----------------------------------------------------- */
/* utils.rs:231-414 get_hypervar_regions(file, primers, prefix, mismatch): reads FASTA
 * (plain / gzip / bzip2 / xz, utils.rs:148-154), runs the batch path on the GPU(s) of ctx,
 * writes <prefix>.fa and <prefix>.gff with the bytes the CPU restatement of the reference (oracle/) writes — the
 * reference binary itself cannot be built here (no Rust toolchain; oracle/ref_recipe makes its goldens elsewhere).  log_fn (may be
 * NULL) receives the reference's warn!/error! messages (level 1 = WARN, 2 = ERROR). */
typedef void (*hx_log_fn)(int level, const char *msg, void *user);
int hx_get_hypervar_regions(hx_ctx *ctx, const char *file, uint32_t n_pairs, const char *const *fwd,
                            const char *const *rev, const char *prefix, uint8_t mismatch, hx_log_fn log_fn,
                            void *user);

/* Where the time of the last hx_get_hypervar_regions call on ctx went (bench.py's file_e2e block; HYPEREX_TIMING=1
 * prints the same).  The call is a pipeline of host threads — chunker, parsers, two text lanes per device, writers —
 * so the per-stage seconds are SUMS over the threads of a stage and may exceed the wall time.  Fills
 * out[0 .. min(n, HX_FILE_STATS_N)) in the order of the enum. */
enum {
    HX_FS_WALL_S = 0,    /* wall time of the call */
    HX_FS_BASES,         /* sequence bytes read */
    HX_FS_RECORDS,
    HX_FS_BATCHES,
    HX_FS_CHUNK_S,       /* cutting the input at record boundaries (compressed input: the decoder) */
    HX_FS_PARSE_S,       /* FASTA text -> pinned record arenas, summed over parser threads */
    HX_FS_DEVICE_S,      /* H2D + kernels + D2H, summed over lanes, waiting for writes excluded */
    HX_FS_WRITE_S,       /* lanes waiting for pwrite() of the output text */
    HX_FS_LOG_S,         /* log_fn calls */
    HX_FS_FASTA_BYTES,   /* size of <prefix>.fa */
    HX_FS_GFF_BYTES,     /* size of <prefix>.gff */
    HX_FS_PARSERS,       /* threads per stage */
    HX_FS_LANES,
    HX_FS_WRITERS,
    HX_FS_SETUP_S,       /* hx_set_primers, opening the files, thread start */
    HX_FS_CLOSE_S,       /* unmapping and closing the output files */
    HX_FILE_STATS_N
};
int hx_file_stats(hx_ctx *ctx, double *out, int n);

/* What the last hx_run_batch / hx_run_batch_packed call of this context moved: out[0] bytes copied host -> device
 * (text, offsets, rec_flags), out[1] bytes copied device -> host (the result table), out[2] 1 when the text crossed
 * PCIe as 4-bit codes (a packed arena, or an ASCII arena packed on the way: option pack_h2d), out[3] slabs,
 * out[4] host threads that packed (0: none), out[5] slabs that crossed packed (pack_h2d = 2 mixes both kinds in one
 * call), out[6] how the call ran: 0 ASCII slabs, 1 every slab packed on the way, 2 mixed, 3 packed arena.  n = entries
 * of out (fewer are fine). */
int hx_batch_stats(const hx_ctx *ctx, uint64_t *out, int n);

#ifdef __cplusplus
}
#endif
#endif /* HYPEREX_B200_H */
