// hx_device.h — structures shared between the host runtime (hx_api.cu) and the sm_100a
// kernels (hx_kernels.cu).  Vocabulary follows the reference's domain: records, primer
// pairs, patterns (a forward primer or a reverse-complemented reverse primer), hits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HX_MAX_NP32 16     // patterns per thread in the 32-bit kernel (m <= 32)
#define HX_MAX_NP64 8      // patterns per thread in the 64-bit kernel (33 <= m <= 64)
#define HX_MAX_GROUP_PAIRS 16
#ifndef HX_THREADS
#define HX_THREADS 128     // threads per CTA of the scan kernel
#endif
#define HX_NBUCKETS 1024   // length buckets of the tile ordering
#define HX_NONE64 0xFFFFFFFFFFFFFFFFull
#define HX_LONG_TILES 64u  // records with more windows than this are joined by a whole warp

// raw per-record alphabet evidence gathered by hx_classify_kernel (utils.rs:207-228)
#define HX_RAW_U 1u
#define HX_RAW_T 2u
#define HX_RAW_NONIUPAC 4u

// A tile is the unit of work of one thread: a whole short record, or one window of a long
// record.  Hits are only accepted at record positions [own_start, own_start + own_len);
// the `warm` bytes before own_start only bring the automaton into the exact state
// (SURVEY.md §5 "overlapped windows": warm = m_max + k - 1).
struct HxTile {
    uint32_t rec;        // record index inside the batch
    uint32_t own_start;  // record-relative first owned position
    uint32_t own_len;    // number of owned positions
    uint32_t warm;       // warm-up bytes scanned before own_start (0 at the record start)
};

// One pattern group = the patterns one thread steps per text byte and the primer pairs
// formed among them.
struct HxGroup {
    int32_t np;                               // patterns in the group
    int32_t npairs;                           // pairs in the group
    int32_t word64;                           // 1 => 64-bit automaton words
    uint32_t table_off;                       // byte offset of the group's Peq table in the blob
    uint32_t slot_base;                       // first per-pattern summary slot of the group
    uint8_t m[HX_MAX_NP32];                   // SCANNED pattern lengths (<= 32 in a 32-bit group: a primer longer
                                              // than the word is scanned as its 32-symbol suffix, see longmask)
    uint8_t mfull[HX_MAX_NP32];               // primer lengths as given (utils.rs:327, 334-337 use these)
    uint8_t pair_f[HX_MAX_GROUP_PAIRS];       // group-local forward pattern of each pair
    uint8_t pair_r[HX_MAX_GROUP_PAIRS];       // group-local reverse pattern of each pair
    uint16_t pair_global[HX_MAX_GROUP_PAIRS]; // index of the pair in hx_set_primers order
    uint16_t rmask[HX_MAX_NP32];              // bit q set: pair q of the group has this pattern as reverse
    uint32_t vslot[HX_MAX_NP32];              // suffix-filtered pattern: index of its 64-bit verification table
    uint32_t longmask;                        // bit p set: pattern p is the 32-symbol suffix of a longer primer; its
                                              // events are candidates, verified against the whole primer in the drain
    // -m 0 pattern packing: two Shift-Or automata of <= 16 symbols per 32-bit word, pattern 2j in the low
    // half of word j and pattern 2j + 1 in the high half.  A primer longer than 16 symbols is scanned as its last 16
    // symbols and every candidate end is verified against the rest of the primer in the drain.
    int32_t packed;                           // 1: the group has a packed table (hx_set_primers decided it pays)
    int32_t pk_parts;                         // automata per word of that table: 2 (<= 16 symbols each), or at -m 0 4
                                              // (<= 8 symbols each) when every 8-symbol suffix is specific enough
    uint32_t pk_table_off;                    // byte offset of the packed table in the packed blob
    uint32_t pk_longmask;                     // bit p set: pattern p is the 32 / pk_parts-symbol suffix of a longer primer
    uint8_t pk_m[HX_MAX_NP32];                // scanned lengths (<= 32 / pk_parts) in the packed table
    uint32_t pk_vslot[HX_MAX_NP32];           // verification table of a packed pattern with pk_longmask set
};

// global pair -> where its summaries live (used by the pairing/combine kernel)
struct HxPairRef {
    uint32_t f_slot;  // per-pattern summary slot of the forward pattern
    uint32_t r_slot;  // per-pattern summary slot of the reverse pattern
    uint32_t len_f;   // byte length of the forward primer (utils.rs:327)
    uint32_t pad;
};

// Peq row geometry: a row holds the eq words of all patterns of the group for one text byte.
// Row stride is an odd number of 16-byte units so that the rows of A/C/G/T/N fall into
// distinct shared-memory bank groups (no conflicts among the common text symbols).
__host__ __device__ constexpr int hx_row_bytes(int np, int wordbytes) { return ((np * wordbytes + 15) / 16) * 16; }
__host__ __device__ constexpr int hx_row_stride(int np, int wordbytes) { return (hx_row_bytes(np, wordbytes) / 16 | 1) * 16; }
__host__ __device__ constexpr int hx_table_bytes(int np, int wordbytes) { return 256 * hx_row_stride(np, wordbytes); }

struct HxScanArgs {
    const uint8_t *arena;
    const uint64_t *offsets;   // n_records + 1 (arena offsets as given by the caller)
    uint64_t off_base;         // arena offset that device address `arena` corresponds to
    const HxTile *tiles;
    const uint32_t *order;     // processing order (longest tile first)
    const uint32_t *ntiles;    // device-side tile count
    const uint32_t *rec_raw;   // HX_RAW_* per record
    const uint32_t *tile_cnt;  // tiles per record (1: the scan kernel writes the final result)
    void *out;                 // hx_result[n_records][n_pairs_total]
    uint32_t n_pairs_total;
    const uint8_t *tables;     // Peq blob
    const uint8_t *tables_wm;  // complemented Peq blob of the small-k fast path (Shift-Or convention)
    const uint8_t *tables_pk;  // -m 0 packed blob: two 16-bit complemented Peq half-words per 32-bit word
    int32_t packed;            // 1: run the packed Shift-Or scan (wm_k == 0 and the group has a packed table)
    int32_t nib;               // 1: `arena` holds 4-bit codes (hx_run_batch_packed); offsets / off_base count symbols
    const uint64_t *vtables;   // [vslot][alphabet dna|rna][256] top-aligned 64-bit Peq words of the suffix-filtered primers
    int32_t wm_k;              // >= 0: run the Wu-Manber fast path with this many errors (32-bit groups only)
    const HxGroup *group;      // this launch's group (device memory)
    uint64_t *sum_f;           // [slot][tile] best valid forward hit  key = dist<<32 | end
    uint64_t *sum_r;           // [slot][tile] best hit of any position
    uint64_t *pair_hi;         // [pair][tile] combined<<32 | f_end
    uint32_t *pair_lo;         // [pair][tile] r_end
    uint32_t tile_stride;      // capacity used as the [slot]/[pair] stride
    uint32_t k;
    int32_t variant;           // text path: 0 per-thread LDG.128 + prefetch, 1 TMA bulk copies into shared memory
    uint32_t two;              // == 2, passed at run time (a constant ptxas cannot fold)
    uint32_t *sched;           // {next batch, CTAs that left}: both 0 at launch, zeroed again by the last CTA to leave
};

struct HxBuildArgs {
    const uint64_t *offsets;
    uint32_t n_records;
    uint32_t single_max;
    uint32_t tile_len;
    uint32_t warm;          // m_max + k - 1
    uint32_t bucket_gran;   // bytes per length bucket
    uint32_t tile_cap;
    HxTile *tiles;
    uint32_t *tile_first;   // per record
    uint32_t *tile_cnt;     // per record
    uint32_t *ntiles;       // counter (zeroed)
    uint32_t *hist;         // HX_NBUCKETS (zeroed)
    uint32_t *overflow;     // set to 1 when tile_cap would be exceeded
    uint32_t *long_list;    // records with more than HX_LONG_TILES windows
    uint32_t *n_long;       // counter (zeroed)
};

// launchers (hx_kernels.cu); each returns the number of kernels it launched
int hx_launch_tile_build(const HxBuildArgs &a, uint32_t *cursor, uint32_t *order, cudaStream_t st);
int hx_launch_classify(const uint8_t *arena, const uint64_t *offsets, uint64_t off_base, const HxTile *tiles,
                       const uint32_t *ntiles, uint32_t tile_cap, uint32_t *rec_raw, const uint32_t *tile_cnt,
                       uint32_t n_records, uint32_t *n_plain, cudaStream_t st);
int hx_launch_flags(const uint8_t *flags, uint32_t *rec_raw, uint32_t n_records, cudaStream_t st);
int hx_launch_scan(const HxScanArgs &a, const HxGroup &hostgroup, uint32_t tile_cap, cudaStream_t st);
int hx_launch_combine(const uint32_t *tile_first, const uint32_t *tile_cnt, const uint32_t *rec_raw,
                      const HxPairRef *pairs, uint32_t n_pairs, uint32_t n_records, const uint64_t *sum_f,
                      const uint64_t *sum_r, const uint64_t *pair_hi, const uint32_t *pair_lo, uint32_t tile_stride,
                      const uint32_t *long_list, const uint32_t *n_long, void *out, cudaStream_t st);
int hx_launch_find(const uint8_t *text, uint64_t n, const uint8_t *table /*256 x u64*/, uint32_t m, uint32_t k,
                   uint32_t tile_len, uint32_t ntiles, uint64_t *counts /*ntiles+1*/, uint64_t *out_end,
                   uint8_t *out_dist, uint64_t cap, int pass, cudaStream_t st);

// on-device FASTA extraction
int hx_launch_fasta_offsets(const void *res, const uint64_t *id_off, const uint32_t *tail_off, uint32_t n_pairs,
                            uint64_t n_entries, uint64_t *lens, uint64_t *sums, cudaStream_t st);
int hx_launch_fasta_write(const uint8_t *arena, const uint64_t *offsets, uint64_t off_base, const void *res,
                          const uint8_t *ids, const uint64_t *id_off, const uint8_t *tails, const uint32_t *tail_off,
                          uint32_t n_pairs, uint64_t n_entries, const uint64_t *where, uint8_t *out, cudaStream_t st);
int hx_launch_gff_offsets(const void *res, const uint64_t *id_off, const uint32_t *tail_off, uint32_t n_pairs,
                          uint64_t n_entries, uint64_t *lens, uint64_t *sums, cudaStream_t st);
int hx_launch_gff_write(const void *res, const uint8_t *ids, const uint64_t *id_off, const uint8_t *tails,
                        const uint32_t *tail_off, uint32_t n_pairs, uint64_t n_entries, const uint64_t *where, uint8_t *out,
                        cudaStream_t st);
