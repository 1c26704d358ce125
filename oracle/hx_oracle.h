/*
 * hx_oracle.h — CPU ORACLE for the HyperEx primer-search-and-extract hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hyperex_b200/ (the product) may include,
 * link, import or execute this code.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs use it, and only as the checker
 * or the reported CPU baseline.
 *
 * It is a plain-C restatement of the reference's algorithm (Ebedthan/hyperex v0.2.0,
 * src/utils.rs) plus the published algorithm of its un-vendored dependency rust-bio
 * `bio` 1.6.0 (Cargo.toml:17, Cargo.lock:133-136): bio::pattern_matching::myers
 * {MyersBuilder::ambig, build_64, Myers<u64>::find_all_lazy}.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests — no reference test
 * asserts a hit coordinate.  The Myers restatement is pinned by the
 * rust-bio documentation known-answer vectors and by an independent O(mn) DP
 * (hxo_dp_find_all_end); the host helpers are pinned by the reference unit-test
 * strings (src/utils.rs:593-683).
 */
#ifndef HX_ORACLE_H
#define HX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { HXO_DNA = 0, HXO_RNA = 1, HXO_NONE = 2 };

/* result of one (record, primer pair); same meaning as the product's hx_result */
typedef struct {
    uint32_t f_start;  /* 0-based start of the region  (utils.rs:334)            */
    uint32_t r_end;    /* 0-based inclusive end        (utils.rs:348)            */
    uint8_t flags;     /* b0 fwd_any b1 rev_any b2 pair_valid b3 alphabet_none b4 rna */
    uint8_t combined;  /* f_dist + r_dist of the chosen pair                     */
    uint16_t reserved;
} hxo_result;

/* rust-bio MyersBuilder (with the ambiguity table of utils.rs:251-263) + build_64 */
typedef struct {
    uint64_t peq[256];
    uint64_t bound;
    uint32_t m;
} hxo_myers;

int hxo_myers_build(hxo_myers *my, const uint8_t *pattern, uint32_t m, int use_hyperex_ambigs);
/* generic builder used by the rust-bio known-answer tests: ambig_codes[i] matches the
 * bytes of ambig_equiv[i] (NUL-terminated) in addition to itself */
int hxo_myers_build_custom(hxo_myers *my, const uint8_t *pattern, uint32_t m, uint32_t n_ambig,
                           const uint8_t *ambig_codes, const char *const *ambig_equiv);

/* Myers<u64>::find_all_lazy(text, k).collect(): (end 0-based inclusive, dist) for every
 * position with dist <= k, ascending.  Returns the total count (may exceed cap; only
 * the first cap hits are stored). */
size_t hxo_find_all_end(const hxo_myers *my, const uint8_t *text, size_t n, uint8_t k,
                        uint64_t *ends, uint8_t *dists, size_t cap);

/* independent cross-check: O(mn) semi-global unit-cost DP with the same match predicate */
size_t hxo_dp_find_all_end(const uint8_t *pattern, uint32_t m, const uint8_t *text, size_t n,
                           uint8_t k, uint64_t *ends, uint8_t *dists, size_t cap);

/* utils.rs:207-228 */
int hxo_sequence_type(const uint8_t *seq, size_t n);
/* utils.rs:169-199; out must hold n bytes (no terminator written) */
void hxo_to_complement(const uint8_t *primer, size_t n, int alphabet, uint8_t *out);
void hxo_to_reverse_complement(const uint8_t *primer, size_t n, int alphabet, uint8_t *out);
/* utils.rs:29-45,157-166; writes a NUL-terminated label (<= 7 chars) */
void hxo_primers_to_region(const char *fwd, const char *rev, char *out8);
/* utils.rs:112-131; returns 0 and sets the two primer strings, or -1 if unknown */
int hxo_region_to_primer(const char *region, const char **fwd, const char **rev);
/* utils.rs:23-25 */
const char *const *hxo_supported_regions(uint32_t *n);

/* the literal pairing loop of utils.rs:327-351.  returns 1 if a pair was chosen. */
int hxo_pair_literal(const uint64_t *f_end, const uint8_t *f_dist, size_t nf, const uint64_t *r_end,
                     const uint8_t *r_dist, size_t nr, size_t forward_len, uint64_t *f_start,
                     uint64_t *r_end_out, uint8_t *combined);

/* utils.rs:270-351 for a batch of records held in an arena: record i = arena[offsets[i]..offsets[i+1]).
 * out is n_records x n_pairs, record-major.  nthreads <= 1 runs single-threaded like the
 * reference; otherwise records are split over pthreads (CPU baseline only).
 * rebuild_per_record != 0 also rebuilds both Peq tables for every record x pair as the
 * reference does (utils.rs:301-302); 0 builds them once per (pair, alphabet). */
int hxo_run_batch(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint32_t n_pairs,
                  const char *const *fwd, const char *const *rev, uint8_t k, hxo_result *out,
                  int nthreads, int rebuild_per_record);

/* utils.rs:353-385: append the FASTA record / GFF line of every valid result to two
 * growing byte buffers.  ids: NUL-terminated record ids.  Buffers are malloc'd/realloc'd;
 * the caller frees them with hxo_free. */
int hxo_format_outputs(const uint8_t *arena, const uint64_t *offsets, const char *const *ids,
                       uint32_t n_records, uint32_t n_pairs, const char *const *fwd,
                       const char *const *rev, const hxo_result *res, uint8_t **fa, size_t *fa_len,
                       uint8_t **gff, size_t *gff_len);
void hxo_free(void *p);

/* rust-bio fasta::Reader rules: parse an uncompressed FASTA buffer.
 * Returns the record count; arena/offsets/ids are malloc'd (ids: concatenated
 * NUL-terminated strings, id_off[n] offsets into it). */
int64_t hxo_parse_fasta(const uint8_t *buf, size_t len, uint8_t **arena, uint64_t **offsets,
                        char **ids, uint64_t **id_off);

#ifdef __cplusplus
}
#endif
#endif
