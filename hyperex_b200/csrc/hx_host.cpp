// hx_host.cpp — host services of the product around the device path: the reference's
// primer/region tables and string helpers (restated 1:1 because they shape the output
// bytes), FASTA ingest with compression sniffing, the FASTA/GFF3 writers and the file-level
// driver that replaces utils::get_hypervar_regions (src/utils.rs:231-414).
#include "hx_host.h"

#include "../../include/hyperex_b200.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <thread>

// ------------------------------------------------------------------------------------
// IUPAC ambiguity table given to MyersBuilder::ambig (utils.rs:251-263)
// ------------------------------------------------------------------------------------
namespace {
struct Ambig {
    char code;
    const char *equiv;
};
const Ambig kAmbig[] = {{'M', "AC"},     {'R', "AG"},     {'W', "AT"},     {'S', "CG"},     {'Y', "CT"},
                        {'K', "GT"},     {'V', "ACGMRS"}, {'H', "ACTMWY"}, {'D', "AGTRWK"}, {'B', "CGTSYK"},
                        {'N', "ACGTMRWSYKVHDB"}};
}  // namespace

bool hxh_symbol_matches(uint8_t p, uint8_t b) {
    if (p == b) return true;
    if (b == 0) return false;
    for (const Ambig &a : kAmbig)
        if ((uint8_t)a.code == p) return strchr(a.equiv, (int)b) != nullptr;
    return false;
}

// ------------------------------------------------------------------------------------
// primer / region tables (utils.rs:23-66, 112-131)
// ------------------------------------------------------------------------------------
namespace {
const char *const kRegions[10] = {"v1v2", "v1v3", "v1v9", "v3v4", "v3v5", "v4", "v4v5", "v5v7", "v6v9", "v7v9"};

struct NamedPrimer {
    const char *name, *seq, *label;
};
// FORWARD_PRIMERS (utils.rs:47-55) with their PRIMER_TO_REGION label (utils.rs:29-45)
const NamedPrimer kForward[7] = {{"27F", "AGAGTTTGATCMTGGCTCAG", "v1"},    {"341F", "CCTACGGGNGGCWGCAG", "v3"},
                                 {"515F", "GTGCCAGCMGCCGCGGTAA", "v4"},    {"515F-Y", "GTGYCAGCMGCCGCGGTAA", "v4"},
                                 {"799F", "AACMGGATTAGATACCCKG", "v5"},    {"928F", "TAAAACTYAAAKGAATTGACGGGG", "v6"},
                                 {"1100F", "YAACGAGCGCAACCC", "v7"}};
// REVERSE_PRIMERS (utils.rs:57-66)
const NamedPrimer kReverse[8] = {{"336R", "ACTGCTGCSYCCCGTAGGAGTCT", "v2"}, {"534R", "ATTACCGCGGCTGCTGG", "v3"},
                                 {"805R", "GACTACHVGGGTATCTAATCC", "v4"},   {"926Rb", "CCGTCAATTYMTTTRAGT", "v5"},
                                 {"806R", "GGACTACHVGGGTWTCTAAT", "v4"},    {"909-928R", "CCCCGYCAATTCMTTTRAGT", "v5"},
                                 {"1193R", "ACGTCATCCCCACCTTCC", "v7"},     {"1492Rmod", "TACGGYTACCTTGTTAYGACTT", "v9"}};
struct RegionDef {
    const char *region, *f, *r;
};
const RegionDef kRegionDef[10] = {{"v1v2", "27F", "336R"},      {"v1v3", "27F", "534R"},   {"v1v9", "27F", "1492Rmod"},
                                  {"v3v4", "341F", "805R"},     {"v3v5", "341F", "926Rb"}, {"v4", "515F", "806R"},
                                  {"v4v5", "515F-Y", "909-928R"}, {"v5v7", "799F", "1193R"}, {"v6v9", "928F", "1492Rmod"},
                                  {"v7v9", "1100F", "1492Rmod"}};

const char *label_of(const char *primer) {
    for (const NamedPrimer &p : kForward)
        if (strcmp(p.seq, primer) == 0) return p.label;
    for (const NamedPrimer &p : kReverse)
        if (strcmp(p.seq, primer) == 0) return p.label;
    return "";
}
}  // namespace

extern "C" {

const char *const *hx_supported_regions(uint32_t *n) {
    if (n) *n = 10;
    return kRegions;
}

int hx_region_to_primer(const char *region, const char **fwd, const char **rev) {
    if (!region || !fwd || !rev) return HX_ERR_ARG;
    for (const RegionDef &d : kRegionDef) {
        if (strcmp(d.region, region) != 0) continue;
        for (const NamedPrimer &p : kForward)
            if (strcmp(p.name, d.f) == 0) *fwd = p.seq;
        for (const NamedPrimer &p : kReverse)
            if (strcmp(p.name, d.r) == 0) *rev = p.seq;
        return HX_OK;
    }
    return HX_ERR_ARG;  // the reference returns an empty Vec (utils.rs:124)
}

void hx_primers_to_region(const char *fwd, const char *rev, char *out) {
    const char *a = label_of(fwd), *b = label_of(rev);
    if (strcmp(a, "v4") == 0 && strcmp(b, "v4") == 0) snprintf(out, 8, "%s", a);  // utils.rs:161-162
    else snprintf(out, 8, "%s%s", a, b);
}

int hx_sequence_type(const uint8_t *seq, size_t n) {
    bool has_u = false, has_t = false, all_iupac = true;
    static const char kIupac[] = "ACGTRYSWKMBDHVN";
    for (size_t i = 0; i < n; ++i) {
        uint8_t c = seq[i];
        if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
        has_u |= c == 'U';
        has_t |= c == 'T';
        if (c == 0 || !memchr(kIupac, c, sizeof(kIupac) - 1)) all_iupac = false;
    }
    if (has_u && !has_t) return 1;
    if (has_t && !has_u) return 0;
    if (!has_u && !has_t) return all_iupac ? 0 : 2;
    return 2;
}

void hx_to_complement(const char *primer, size_t n, int alphabet, char *out) {
    for (size_t i = 0; i < n; ++i) {
        const char c = primer[i];
        char o = c;
        switch (c) {
        case 'A': o = alphabet == 0 ? 'T' : alphabet == 1 ? 'U' : 'A'; break;
        case 'T': o = alphabet == 0 ? 'A' : 'T'; break;  // 'T' is left alone under "rna"
        case 'U': o = alphabet == 1 ? 'A' : 'U'; break;  // 'U' is left alone under "dna"
        case 'C': o = (alphabet == 0 || alphabet == 1) ? 'G' : 'C'; break;
        case 'G': o = (alphabet == 0 || alphabet == 1) ? 'C' : 'G'; break;
        case 'R': o = 'Y'; break;
        case 'Y': o = 'R'; break;
        case 'K': o = 'M'; break;
        case 'M': o = 'K'; break;
        case 'B': o = 'V'; break;
        case 'V': o = 'B'; break;
        case 'D': o = 'H'; break;
        case 'H': o = 'D'; break;
        default: break;
        }
        out[i] = o;
    }
}

void hx_to_reverse_complement(const char *primer, size_t n, int alphabet, char *out) {
    std::string tmp(n, '\0');
    hx_to_complement(primer, n, alphabet, &tmp[0]);
    for (size_t i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
}

int hx_file_to_vec(const char *path, char ***fwd, char ***rev) {
    if (!path || !fwd || !rev) return HX_ERR_ARG;
    FILE *f = fopen(path, "rb");
    if (!f) return HX_ERR_IO;
    std::string all;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, n);
    fclose(f);
    // str::lines(): split on '\n', a trailing "\r" is stripped, no empty last line
    std::vector<std::string> lines;
    size_t p = 0;
    while (p < all.size()) {
        size_t e = all.find('\n', p);
        std::string l = all.substr(p, e == std::string::npos ? std::string::npos : e - p);
        if (!l.empty() && l.back() == '\r') l.pop_back();
        lines.push_back(l);
        if (e == std::string::npos) break;
        p = e + 1;
    }
    std::vector<std::string> F, R;
    for (const std::string &l : lines) {
        size_t c = l.find(',');
        if (c == std::string::npos) return HX_ERR_ARG;  // "Primer file must be comma-separated"
        size_t c2 = l.find(',', c + 1);
        F.push_back(l.substr(0, c));
        R.push_back(l.substr(c + 1, c2 == std::string::npos ? std::string::npos : c2 - c - 1));
    }
    const int cnt = (int)F.size();
    *fwd = (char **)calloc((size_t)cnt + 1, sizeof(char *));
    *rev = (char **)calloc((size_t)cnt + 1, sizeof(char *));
    for (int i = 0; i < cnt; ++i) {
        (*fwd)[i] = strdup(F[i].c_str());
        (*rev)[i] = strdup(R[i].c_str());
    }
    return cnt;
}

void hx_free_primer_file(char **fwd, char **rev, int n) {
    for (int i = 0; i < n; ++i) {
        if (fwd) free(fwd[i]);
        if (rev) free(rev[i]);
    }
    free(fwd);
    free(rev);
}

}  // extern "C"

// ------------------------------------------------------------------------------------
// Packed arena (BASELINE.json north_star: "2-bit/ASCII arena in pinned memory"): 4-bit codes, two bases per byte.
//   code 0        any byte that is not one of the 16 letters below (after to_ascii_uppercase, utils.rs:295)
//   codes 1..15   A C G T|U R Y S W K M B D H V N
// 'T' and 'U' share code 4: a record that holds both is alphabet None and is skipped (utils.rs:225-226), so inside a
// record that is scanned the code is unambiguous and the device resolves it with the record's alphabet.  A byte
// outside the 16 letters matches a primer symbol only if the primer holds that very byte (rust-bio build_64 + the
// ambiguity table of utils.rs:251-263 list nothing else); hx_run_batch_packed refuses primer sets that do, so code
// 0 — a text symbol that matches nothing — is exact.  While packing, the evidence sequence_type needs is gathered per
// record (utils.rs:207-228): has 'U', has 'T', has a byte outside "ACGTRYSWKMBDHVN" + 'U'.
// Nibble i of the packed buffer is base i of the arena (low nibble = even index), so the offsets array serves both.
// ------------------------------------------------------------------------------------
namespace {

// byte -> 4-bit code in the low nibble, HX_REC_* evidence bits in bits 4..6
struct PackTables {
    uint8_t lut[256];
    PackTables() {
        static const char letters[] = "ACGTRYSWKMBDHVN";
        for (int b = 0; b < 256; ++b) {
            const int u = (b >= 'a' && b <= 'z') ? b - 32 : b;
            const char *q = (u && u < 128) ? strchr(letters, u) : nullptr;
            const int code = q ? (int)(q - letters + 1) : (u == 'U' ? 4 : 0);
            const int flag = (u == 'U' ? HX_REC_HAS_U : 0) | (u == 'T' ? HX_REC_HAS_T : 0) | ((!q && u != 'U') ? HX_REC_NON_IUPAC : 0);
            lut[b] = (uint8_t)(code | (flag << 4));
        }
    }
};
const PackTables &pack_tables() {
    static const PackTables t;
    return t;
}

// The packer is a STREAM over bases, not over records: nibble i is base i whatever record it belongs to, so a thread
// packs its base range [start, stop) in whole vector blocks (64 bases with AVX2, 128 with AVX-512 VBMI; block starts are
// absolute multiples of the block size, so the packed stores are aligned whenever `packed` is) and only the evidence
// bits are per record.  A cursor walks the records alongside: blocks that lie inside one record — 22 of 23 for 1.5 kb
// reads — OR their table bytes into a vector accumulator; a block that holds record ends turns the three evidence bits
// into bit masks over its bases and splits them at the ends.  A record is finalised (rec_flags[r] stored) by the thread
// whose range holds its last base; what a thread saw of a record that ends later is handed to the merge after the join.
struct PackCursor {
    const uint64_t *off;
    uint32_t n;
    uint8_t *rec_flags;
    uint32_t cur;           // record that holds the next base
    uint64_t e;             // its end (UINT64_MAX past the last record)
    uint32_t acc = 0;       // evidence of `cur` so far (HX_REC_* bits)
    int64_t first_final = -1;
    void seek(uint64_t start) {
        // last r with off[r] <= start: empty records that end at `start` belong to whoever holds base start - 1
        cur = (uint32_t)(std::upper_bound(off, off + n + 1, start) - off) - 1;
        e = cur < n ? off[cur + 1] : UINT64_MAX;
    }
    void finalize() {  // `cur` ends here
        rec_flags[cur] = (uint8_t)acc;
        if (first_final < 0) first_final = cur;
        acc = 0;
        ++cur;
        e = cur < n ? off[cur + 1] : UINT64_MAX;
    }
};

// bases [p, q) one by one (heads and tails of a thread's range); q is even or the end of the batch
void pack_scalar(const uint8_t *arena, uint64_t p, uint64_t q, uint8_t *packed, PackCursor &c) {
    const uint8_t *lut = pack_tables().lut;
    for (uint64_t i = p; i < q; ++i) {
        while (c.e <= i) c.finalize();
        const uint8_t v = lut[arena[i]];
        c.acc |= v >> 4;
        if (i & 1) packed[i >> 1] = (uint8_t)((packed[i >> 1] & 0x0f) | (v << 4));
        else packed[i >> 1] = (uint8_t)(v & 0x0f);  // the high nibble follows, or stays 0 at the end of the batch
    }
    while (c.e <= q) c.finalize();
}

inline uint32_t seg_any(const uint64_t *m, int words, uint32_t a, uint32_t b) {  // any bit of m in [a, b)?
    uint64_t r = 0;
    for (int w = 0; w < words; ++w) {
        const uint32_t lo = (uint32_t)w * 64u, hi = lo + 64u;
        if (b <= lo || a >= hi) continue;
        uint64_t mask = ~0ull;
        if (a > lo) mask &= ~0ull << (a - lo);
        if (b < hi) mask &= ~0ull >> (hi - b);
        r |= m[w] & mask;
    }
    return r != 0;
}

// record ends inside the block [p, p + bases): mU / mT / mB hold one bit per base of the block
inline void pack_split(PackCursor &c, uint64_t p, uint32_t bases, const uint64_t *mU, const uint64_t *mT, const uint64_t *mB,
                       uint32_t vec_acc) {
    const int words = (int)(bases / 64u);
    c.acc |= vec_acc;
    uint32_t pos = 0;
    while (c.e <= p + bases) {
        const uint32_t end = (uint32_t)(c.e - p);
        if (end > pos)
            c.acc |= (seg_any(mU, words, pos, end) ? HX_REC_HAS_U : 0u) | (seg_any(mT, words, pos, end) ? HX_REC_HAS_T : 0u) |
                     (seg_any(mB, words, pos, end) ? HX_REC_NON_IUPAC : 0u);
        pos = end;
        c.finalize();
    }
    if (pos < bases)
        c.acc |= (seg_any(mU, words, pos, bases) ? HX_REC_HAS_U : 0u) | (seg_any(mT, words, pos, bases) ? HX_REC_HAS_T : 0u) |
                 (seg_any(mB, words, pos, bases) ? HX_REC_NON_IUPAC : 0u);
}

// Software prefetch distance of the vector loops in bytes (HYPEREX_PACK_PF; < 0: into all cache levels, > 0: with the
// non-temporal hint, 0: none).  A core streams 9-10 GB/s of text with the hardware prefetcher alone and 14 GB/s when it
// also asks 2 KB ahead (16 threads of the B200 host: 76-89 -> 103-111 Gbases/s); the non-temporal hint measured slower
// (profiles/r02_pack_h2d_auto.md, r02_pack_h2d_ab_ring.json).
int pack_prefetch() {
    static const int v = [] {
        const char *e = getenv("HYPEREX_PACK_PF");
        return e ? atoi(e) : -2048;
    }();
    return v;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline uint32_t pack_reduce_avx2(__m256i vacc) {  // OR of the evidence bits of all bytes
    alignas(32) uint64_t w[4];
    _mm256_store_si256((__m256i *)w, vacc);
    uint64_t r = w[0] | w[1] | w[2] | w[3];
    r |= r >> 32;
    r |= r >> 16;
    r |= r >> 8;
    return (uint32_t)(r >> 4) & 7u;
}
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) inline uint32_t pack_reduce_avx512(__m512i vacc) {
    return (_mm512_test_epi8_mask(vacc, _mm512_set1_epi8(HX_REC_HAS_U << 4)) ? HX_REC_HAS_U : 0u) |
           (_mm512_test_epi8_mask(vacc, _mm512_set1_epi8(HX_REC_HAS_T << 4)) ? HX_REC_HAS_T : 0u) |
           (_mm512_test_epi8_mask(vacc, _mm512_set1_epi8(HX_REC_NON_IUPAC << 4)) ? HX_REC_NON_IUPAC : 0u);
}
// 64 bases per trip.  Table byte of a base: letters (0x40..0x7f) through two 16-entry shuffles on the low five bits, which
// also folds the case; everything else is "not one of the 16" (code 0, HX_REC_NON_IUPAC).
__attribute__((target("avx2"))) void pack_blocks_avx2(const uint8_t *arena, uint64_t p, uint64_t q, uint8_t *packed, PackCursor &c, bool nt) {
    const uint8_t *lut = pack_tables().lut;
    alignas(32) uint8_t t_lo[32], t_hi[32];
    for (int i = 0; i < 16; ++i) {
        t_lo[i] = t_lo[16 + i] = lut[0x40 + i];  // '@' A..O
        t_hi[i] = t_hi[16 + i] = lut[0x50 + i];  // P..Z [ \ ] ^ _
    }
    const __m256i lut_lo = _mm256_load_si256((const __m256i *)t_lo), lut_hi = _mm256_load_si256((const __m256i *)t_hi);
    const __m256i m_0f = _mm256_set1_epi8(0x0f), m_c0 = _mm256_set1_epi8((char)0xc0), c_40 = _mm256_set1_epi8(0x40);
    const __m256i mult = _mm256_set1_epi16(0x1001);  // bytes (1, 16): low code + 16 * high code
    __m256i vacc = _mm256_setzero_si256();
    const int pf = pack_prefetch();
    for (; p + 64 <= q; p += 64) {
        if (pf > 0) _mm_prefetch((const char *)(arena + p + pf), _MM_HINT_NTA);
        else if (pf < 0) _mm_prefetch((const char *)(arena + p - pf), _MM_HINT_T0);
        __m256i o[2];
        for (int h = 0; h < 2; ++h) {
            const __m256i x = _mm256_loadu_si256((const __m256i *)(arena + p + 32 * h));
            const __m256i idx = _mm256_and_si256(x, m_0f);
            const __m256i t = _mm256_blendv_epi8(_mm256_shuffle_epi8(lut_lo, idx), _mm256_shuffle_epi8(lut_hi, idx),
                                                 _mm256_slli_epi16(x, 3));  // bit 4 of the byte selects P..Z
            const __m256i letter = _mm256_cmpeq_epi8(_mm256_and_si256(x, m_c0), c_40);
            o[h] = _mm256_blendv_epi8(c_40, t, letter);  // 0x40 = code 0 + HX_REC_NON_IUPAC << 4
        }
        if (c.e > p + 64) {
            vacc = _mm256_or_si256(vacc, _mm256_or_si256(o[0], o[1]));
        } else {  // record ends inside the block: evidence of the record's finished blocks, then the split of this one
            const uint32_t va = pack_reduce_avx2(vacc);
            vacc = _mm256_setzero_si256();
            const uint64_t mU = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(o[0], 3)) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(o[1], 3)) << 32);
            const uint64_t mT = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(o[0], 2)) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(o[1], 2)) << 32);
            const uint64_t mB = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(o[0], 1)) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(o[1], 1)) << 32);
            pack_split(c, p, 64, &mU, &mT, &mB, va);
        }
        const __m256i a = _mm256_maddubs_epi16(_mm256_and_si256(o[0], m_0f), mult), b = _mm256_maddubs_epi16(_mm256_and_si256(o[1], m_0f), mult);
        const __m256i v = _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xd8);
        if (nt) _mm256_stream_si256((__m256i *)(packed + p / 2), v);
        else _mm256_storeu_si256((__m256i *)(packed + p / 2), v);
    }
    if (nt) _mm_sfence();
    // hand the vector accumulator to the cursor (the open record goes on in the scalar tail or in the next thread)
    c.acc |= pack_reduce_avx2(vacc);
}

// 128 bases per trip: one VPERMI2B looks a 7-bit byte up in the 128-entry table, bytes >= 0x80 are masked to 0x40
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) void pack_blocks_avx512(const uint8_t *arena, uint64_t p, uint64_t q, uint8_t *packed,
                                                                                PackCursor &c, bool nt) {
    const uint8_t *lut = pack_tables().lut;
    const __m512i tab0 = _mm512_loadu_si512(lut), tab1 = _mm512_loadu_si512(lut + 64);
    const __m512i c_40 = _mm512_set1_epi8(0x40), m_0f = _mm512_set1_epi8(0x0f), mult = _mm512_set1_epi16(0x1001);
    const __m512i bU = _mm512_set1_epi8(HX_REC_HAS_U << 4), bT = _mm512_set1_epi8(HX_REC_HAS_T << 4), bB = _mm512_set1_epi8(HX_REC_NON_IUPAC << 4);
    __m512i vacc = _mm512_setzero_si512();
    const int pf = pack_prefetch();
    for (; p + 128 <= q; p += 128) {
        if (pf > 0) {  // the text is read once: ask for it ahead of the hardware prefetcher, without keeping it in the caches
            _mm_prefetch((const char *)(arena + p + pf), _MM_HINT_NTA);
            _mm_prefetch((const char *)(arena + p + pf + 64), _MM_HINT_NTA);
        } else if (pf < 0) {
            _mm_prefetch((const char *)(arena + p - pf), _MM_HINT_T0);
            _mm_prefetch((const char *)(arena + p - pf + 64), _MM_HINT_T0);
        }
        const __m512i x0 = _mm512_loadu_si512(arena + p), x1 = _mm512_loadu_si512(arena + p + 64);
        const __m512i o0 = _mm512_mask_mov_epi8(_mm512_permutex2var_epi8(tab0, x0, tab1), _mm512_movepi8_mask(x0), c_40);
        const __m512i o1 = _mm512_mask_mov_epi8(_mm512_permutex2var_epi8(tab0, x1, tab1), _mm512_movepi8_mask(x1), c_40);
        if (c.e > p + 128) {
            vacc = _mm512_ternarylogic_epi64(vacc, o0, o1, 0xfe);
        } else {
            const uint64_t mU[2] = {_mm512_test_epi8_mask(o0, bU), _mm512_test_epi8_mask(o1, bU)};
            const uint64_t mT[2] = {_mm512_test_epi8_mask(o0, bT), _mm512_test_epi8_mask(o1, bT)};
            const uint64_t mB[2] = {_mm512_test_epi8_mask(o0, bB), _mm512_test_epi8_mask(o1, bB)};
            pack_split(c, p, 128, mU, mT, mB, pack_reduce_avx512(vacc));
            vacc = _mm512_setzero_si512();
        }
        const __m512i a = _mm512_maddubs_epi16(_mm512_and_si512(o0, m_0f), mult), b = _mm512_maddubs_epi16(_mm512_and_si512(o1, m_0f), mult);
        const __m512i v = _mm512_inserti64x4(_mm512_castsi256_si512(_mm512_cvtepi16_epi8(a)), _mm512_cvtepi16_epi8(b), 1);
        if (nt) _mm512_stream_si512((__m512i *)(packed + p / 2), v);
        else _mm512_storeu_si512(packed + p / 2, v);
    }
    if (nt) _mm_sfence();
    c.acc |= pack_reduce_avx512(vacc);
}
#endif

int pack_isa() {  // 0 scalar, 1 AVX2, 2 AVX-512 VBMI; HYPEREX_PACK_ISA caps it (tests run every level the CPU has)
#if defined(__x86_64__)
    static const int v = [] {
        int isa = 0;
        if (__builtin_cpu_supports("avx2")) isa = 1;
        if (isa == 1 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi")) isa = 2;
        return isa;
    }();
    const char *cap = getenv("HYPEREX_PACK_ISA");
    return cap ? std::min(v, atoi(cap)) : v;
#else
    return 0;
#endif
}

typedef HxhPackPart PackPart;  // what one range reports to the merge (hx_host.h)

// bases [start, stop) of the batch; interior range edges are multiples of 128 bases
void pack_range(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint64_t start, uint64_t stop, uint8_t *packed,
                uint8_t *rec_flags, int isa, bool stream_stores, PackPart *part) {
    PackCursor c{offsets, n_records, rec_flags, 0, 0};
    c.seek(start);
    const uint64_t blk = isa == 2 ? 128 : 64;
    uint64_t p = start;
    const uint64_t head_end = std::min(stop, (start + blk - 1) / blk * blk);
    if (isa == 0 || head_end > p) {
        const uint64_t q = isa == 0 ? stop : head_end;
        pack_scalar(arena, p, q, packed, c);
        p = q;
    }
    const uint64_t body_end = p + (stop - p) / blk * blk;
#if defined(__x86_64__)
    if (body_end > p) {
        // large outputs bypass the cache (no read-for-ownership of the packed lines; the next reader is a DMA engine or
        // another pass long after); the block stores are 64- / 32-byte aligned whenever the buffer is
        const bool nt = stream_stores && (((uintptr_t)(packed + p / 2)) & (blk / 2 - 1)) == 0;
        if (isa == 2) pack_blocks_avx512(arena, p, body_end, packed, c, nt);
        else pack_blocks_avx2(arena, p, body_end, packed, c, nt);
        p = body_end;
    }
#endif
    if (stop > p) pack_scalar(arena, p, stop, packed, c);
    while (c.e <= stop) c.finalize();  // records (also empty ones) that end exactly at the end of the range
    part->first_final = c.first_final;
    part->open_rec = c.cur;
    part->open_acc = c.acc;
}

}  // namespace

// ------------------------------------------------------------------------------------
// A small persistent pool for data-parallel host loops that run once per slab (packing a slab takes about a
// millisecond: starting threads for it would cost as much as the work).  One pool per process, created on first use;
// callers are serialised.
// ------------------------------------------------------------------------------------
namespace {
class HxhPool {
  public:
    static HxhPool &get() {
        static HxhPool *p = new HxhPool();  // never destroyed: worker threads must not outlive their state at exit
        return *p;
    }
    int size() const { return (int)workers_.size() + 1; }
    // runs fn(i) for i in [0, n) on up to `threads` threads (the caller is one of them)
    void run(int n, int threads, const std::function<void(int)> &fn) {
        if (n <= 0) return;
        if (n == 1 || threads <= 1 || workers_.empty()) {
            for (int i = 0; i < n; ++i) fn(i);
            return;
        }
        std::lock_guard<std::mutex> serial(call_);
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            done_ = 0;
            want_ = std::min((int)workers_.size(), threads - 1);
            woken_ = 0;
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        work(true);
        std::unique_lock<std::mutex> g(m_);
        cv_done_.wait(g, [&] { return done_ == n_ && active_ == 0; });
        fn_ = nullptr;
    }

  private:
    HxhPool() {
        const int hw = (int)std::thread::hardware_concurrency();
        for (int i = 0; i + 1 < std::max(1, hw); ++i) workers_.emplace_back([this] { loop(); });
        for (std::thread &t : workers_) t.detach();
    }
    void work(bool caller) {
        int mine = 0;
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n_) break;
            (*fn_)(i);
            ++mine;
        }
        std::lock_guard<std::mutex> g(m_);
        done_ += mine;
        if (!caller) --active_;
        if (done_ == n_ && active_ == 0) cv_done_.notify_all();
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // a caller that packs slab after slab comes back within microseconds: look for the next call for a moment
            // before going to sleep (a futex wake-up of 15 sleepers costs more than packing a small slab)
            for (int spin = 0; spin < 4000 && epoch_.load(std::memory_order_acquire) == seen; ++spin) {
#if defined(__x86_64__)
                _mm_pause();
#endif
            }
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
                seen = epoch_.load(std::memory_order_acquire);
                if (woken_ >= want_ || fn_ == nullptr) continue;  // enough helpers for this call
                ++woken_;
                ++active_;
            }
            work(false);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex call_, m_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, done_ = 0, want_ = 0, woken_ = 0, active_ = 0;
    std::atomic<int> next_{0};
    std::atomic<uint64_t> epoch_{0};
};
}  // namespace

extern "C" int hx_pack_records(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint8_t *packed,
                               uint8_t *rec_flags, int n_threads) {
    return hxh_pack_slab(arena, offsets, n_records, packed, 0, rec_flags, n_threads);
}

// ---- the packer's building blocks for a caller that schedules the ranges itself (hx_run_batch with pack_h2d) ----------
void hxh_pack_range(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint64_t p, uint64_t q, uint8_t *packed_virtual,
                    uint8_t *rec_flags, bool stream_stores, bool clear_low_nibble, HxhPackPart *part) {
    if (q <= p) {
        part->first_final = -2;  // empty range
        return;
    }
    if (clear_low_nibble && (p & 1)) packed_virtual[p / 2] = 0;
    pack_range(arena, offsets, n_records, p, q, packed_virtual, rec_flags, pack_isa(), stream_stores, part);
}

void hxh_pack_merge(uint8_t *rec_flags, const HxhPackPart *parts, size_t n, uint32_t &open_rec, uint32_t &open_acc) {
    // a record that spans range edges is finalised by the range that holds its last base; OR in what the ranges before
    // it saw of it
    for (size_t t = 0; t < n; ++t) {
        if (parts[t].first_final == -2) continue;
        if (parts[t].first_final >= 0) {
            if ((uint32_t)parts[t].first_final == open_rec) rec_flags[open_rec] |= (uint8_t)open_acc;
            open_rec = parts[t].open_rec;
            open_acc = parts[t].open_acc;
        } else if (parts[t].open_rec == open_rec) {
            open_acc |= parts[t].open_acc;
        } else {
            open_rec = parts[t].open_rec;
            open_acc = parts[t].open_acc;
        }
    }
}

void hxh_parallel_run(int n, int threads, const std::function<void(int)> &fn) { HxhPool::get().run(n, threads, fn); }

// hx_pack_records in pieces (hx_host.h): begin validates and settles the records that end before the first base; a
// piece packs the base range [p, q) — p is `begin` or a multiple of 128, q is `end` or a multiple of 128, pieces follow
// each other without gaps — into packed_virtual[i / 2] for base i and folds the evidence of the records into rec_flags;
// after the last piece every rec_flags entry is final.
int hxh_pack_begin(HxhPackJob &J, const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint8_t *rec_flags) {
    if (!offsets || (n_records && !rec_flags)) return HX_ERR_ARG;
    J.arena = arena;
    J.offsets = offsets;
    J.n_records = n_records;
    J.rec_flags = rec_flags;
    J.begin = J.end = n_records ? offsets[0] : 0;
    J.open_rec = UINT32_MAX;
    J.open_acc = 0;
    if (n_records == 0) return HX_OK;
    if (offsets[n_records] < offsets[0] || (offsets[n_records] > offsets[0] && !arena)) return HX_ERR_ARG;
    for (uint32_t r = 0; r < n_records; ++r)
        if (offsets[r + 1] < offsets[r]) return HX_ERR_ARG;
    J.end = offsets[n_records];
    // leading empty records end before the first base: no range finalises them
    for (uint32_t r = 0; r < n_records && offsets[r + 1] == J.begin; ++r) rec_flags[r] = 0;
    return HX_OK;
}

void hxh_pack_piece(HxhPackJob &J, uint64_t p, uint64_t q, uint8_t *packed_virtual, int n_threads, bool stream_stores) {
    if (q <= p) return;
    uint8_t *packed = packed_virtual;
    if (p == J.begin && (p & 1)) packed[p / 2] = 0;  // the batch starts on an odd base: nobody owns the low nibble of that byte
    const int isa = pack_isa();
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    // The piece is cut into ranges of about `grain` bases at multiples of 128 (whole vector blocks, aligned packed
    // stores) that the threads take one after the other.  Small ranges taken in order rather than one large range per
    // thread: neighbours balance the load, and threads that stream through addresses a fixed large distance apart were
    // seen to slow each other down.
    const char *grain_env = getenv("HYPEREX_PACK_GRAIN");  // tests cut finer
    uint64_t grain = std::max<uint64_t>(grain_env ? strtoull(grain_env, nullptr, 10) : (1u << 20), 128);
    if (!grain_env) grain = std::min<uint64_t>(grain, std::max<uint64_t>((q - p) / (4ull * (uint64_t)nt), 64u << 10));  // small pieces: 4 ranges per thread
    const int np = (int)std::min<uint64_t>((q - p + grain - 1) / grain, 1u << 20);
    nt = std::max(1, std::min(nt, np));
    std::vector<uint64_t> cut(np + 1);
    for (int t = 0; t <= np; ++t) {
        const uint64_t x = p + (uint64_t)((__uint128_t)(q - p) * (uint64_t)t / (uint64_t)np);
        cut[t] = t == 0 ? p : t == np ? q : std::min<uint64_t>(q, std::max<uint64_t>(p, x & ~(uint64_t)127));
    }
    std::vector<PackPart> parts(np);
    HxhPool::get().run(np, nt, [&](int t) {
        if (cut[t + 1] > cut[t])
            pack_range(J.arena, J.offsets, J.n_records, cut[t], cut[t + 1], packed, J.rec_flags, isa, stream_stores, &parts[t]);
        else parts[t].first_final = -2;  // empty range
    });
    hxh_pack_merge(J.rec_flags, parts.data(), parts.size(), J.open_rec, J.open_acc);
}

// hx_pack_records into a buffer whose first byte is packed byte `packed_base / 2` of the batch (packed_base even and
// <= offsets[0])
int hxh_pack_slab(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint8_t *packed, uint64_t packed_base,
                  uint8_t *rec_flags, int n_threads) {
    if (n_records && !packed) return HX_ERR_ARG;
    if ((packed_base & 1) || (n_records && offsets && packed_base > offsets[0])) return HX_ERR_ARG;
    HxhPackJob J;
    const int rc = hxh_pack_begin(J, arena, offsets, n_records, rec_flags);
    if (rc != HX_OK || J.end == J.begin) return rc;
    const char *nt_env = getenv("HYPEREX_PACK_NT");  // A/B switch: 0 keeps ordinary stores
    // large outputs bypass the cache: no read-for-ownership of the packed lines, and the next reader is far away
    const bool stream_stores = nt_env ? atoi(nt_env) != 0 : (J.end - J.begin) >= ((uint64_t)8 << 20);
    hxh_pack_piece(J, J.begin, J.end, (uint8_t *)((uintptr_t)packed - (uintptr_t)(packed_base / 2)), n_threads, stream_stores);
    return HX_OK;
}

// ------------------------------------------------------------------------------------
// byte sources: niffler 2.7.0 magic-number sniffing (call site utils.rs:148-154)
// ------------------------------------------------------------------------------------
class HxhByteSource {
  public:
    virtual ~HxhByteSource() {}
    virtual long read(uint8_t *dst, size_t n) = 0;  // bytes read, 0 at end, -1 on error
};

namespace {

class PlainSource : public HxhByteSource {
  public:
    explicit PlainSource(FILE *f) : f_(f) {}
    ~PlainSource() override { if (f_) fclose(f_); }
    long read(uint8_t *dst, size_t n) override {
        size_t r = fread(dst, 1, n, f_);
        if (r == 0 && ferror(f_)) return -1;
        return (long)r;
    }
  private:
    FILE *f_;
};

class GzSource : public HxhByteSource {  // zlib; handles concatenated members
  public:
    explicit GzSource(const char *path) { g_ = gzopen(path, "rb"); if (g_) gzbuffer(g_, 1 << 20); }
    ~GzSource() override { if (g_) gzclose(g_); }
    bool ok() const { return g_ != nullptr; }
    long read(uint8_t *dst, size_t n) override {
        int r = gzread(g_, dst, (unsigned)std::min<size_t>(n, 1u << 30));
        return r < 0 ? -1 : r;
    }
  private:
    gzFile g_ = nullptr;
};

// The image ships libbz2 / liblzma / libzstd only as runtime libraries (no headers): bind the
// few entry points by hand.
class Bz2Source : public HxhByteSource {
    typedef void *(*open_t)(int *, FILE *, int, int, void *, int);
    typedef int (*read_t)(int *, void *, void *, int);
    typedef void (*close_t)(int *, void *);
    typedef void (*unused_t)(int *, void *, void **, int *);
  public:
    explicit Bz2Source(FILE *f) : f_(f) {
        h_ = dlopen("libbz2.so.1", RTLD_NOW);
        if (!h_) h_ = dlopen("libbz2.so.1.0", RTLD_NOW);
        if (!h_) return;
        open_ = (open_t)dlsym(h_, "BZ2_bzReadOpen");
        read_ = (read_t)dlsym(h_, "BZ2_bzRead");
        close_ = (close_t)dlsym(h_, "BZ2_bzReadClose");
        unused_ = (unused_t)dlsym(h_, "BZ2_bzReadGetUnused");
        if (open_ && read_ && close_ && unused_) {
            int err = 0;
            bz_ = open_(&err, f_, 0, 0, nullptr, 0);
            if (err != 0) bz_ = nullptr;
        }
    }
    ~Bz2Source() override {
        int err = 0;
        if (bz_) close_(&err, bz_);
        if (f_) fclose(f_);
        if (h_) dlclose(h_);
    }
    bool ok() const { return bz_ != nullptr; }
    long read(uint8_t *dst, size_t n) override {
        size_t got = 0;
        while (got < n && bz_) {
            int err = 0;
            int r = read_(&err, bz_, dst + got, (int)std::min<size_t>(n - got, 1u << 30));
            if (err != 0 && err != 4) return -1;
            got += (size_t)r;
            if (err == 4) {  // BZ_STREAM_END: continue with a concatenated stream if any
                void *un = nullptr;
                int nun = 0;
                unused_(&err, bz_, &un, &nun);
                std::vector<char> keep((char *)un, (char *)un + nun);
                close_(&err, bz_);
                bz_ = nullptr;
                int c = EOF;
                if (keep.empty() && (c = fgetc(f_)) == EOF) break;
                if (keep.empty()) ungetc(c, f_);
                bz_ = open_(&err, f_, 0, 0, keep.empty() ? nullptr : keep.data(), (int)keep.size());
                if (err != 0) { bz_ = nullptr; break; }
            }
        }
        return (long)got;
    }
  private:
    FILE *f_;
    void *h_ = nullptr, *bz_ = nullptr;
    open_t open_ = nullptr; read_t read_ = nullptr; close_t close_ = nullptr; unused_t unused_ = nullptr;
};

struct LzmaStream {  // lzma_stream of liblzma 5.x
    const uint8_t *next_in; size_t avail_in; uint64_t total_in;
    uint8_t *next_out; size_t avail_out; uint64_t total_out;
    const void *allocator; void *internal;
    void *rp1, *rp2, *rp3, *rp4;
    uint64_t ri1, ri2; size_t ri3, ri4;
    int re1, re2;
};

class XzSource : public HxhByteSource {
    typedef int (*dec_t)(LzmaStream *, uint64_t, uint32_t);
    typedef int (*code_t)(LzmaStream *, int);
    typedef void (*end_t)(LzmaStream *);
  public:
    explicit XzSource(FILE *f) : f_(f), in_(1 << 20) {
        memset(&s_, 0, sizeof s_);
        h_ = dlopen("liblzma.so.5", RTLD_NOW);
        if (!h_) return;
        dec_t dec = (dec_t)dlsym(h_, "lzma_stream_decoder");
        code_ = (code_t)dlsym(h_, "lzma_code");
        end_ = (end_t)dlsym(h_, "lzma_end");
        if (dec && code_ && end_ && dec(&s_, UINT64_MAX, 0x08 /*LZMA_CONCATENATED*/) == 0) ok_ = true;
    }
    ~XzSource() override {
        if (ok_) end_(&s_);
        if (f_) fclose(f_);
        if (h_) dlclose(h_);
    }
    bool ok() const { return ok_; }
    long read(uint8_t *dst, size_t n) override {
        if (done_) return 0;
        s_.next_out = dst;
        s_.avail_out = n;
        while (s_.avail_out > 0) {
            if (s_.avail_in == 0 && !feof_) {
                size_t r = fread(in_.data(), 1, in_.size(), f_);
                if (r < in_.size()) feof_ = true;
                s_.next_in = in_.data();
                s_.avail_in = r;
            }
            int rc = code_(&s_, feof_ ? 3 /*LZMA_FINISH*/ : 0 /*LZMA_RUN*/);
            if (rc == 1) { done_ = true; break; }
            if (rc != 0) return -1;
        }
        return (long)(n - s_.avail_out);
    }
  private:
    FILE *f_;
    std::vector<uint8_t> in_;
    LzmaStream s_;
    void *h_ = nullptr;
    code_t code_ = nullptr; end_t end_ = nullptr;
    bool ok_ = false, feof_ = false, done_ = false;
};

struct ZBuf { void *p; size_t size, pos; };

class ZstdSource : public HxhByteSource {
    typedef void *(*create_t)();
    typedef size_t (*free_t)(void *);
    typedef size_t (*dec_t)(void *, ZBuf *, ZBuf *);
    typedef unsigned (*iserr_t)(size_t);
  public:
    explicit ZstdSource(FILE *f) : f_(f), in_(1 << 20) {
        h_ = dlopen("libzstd.so.1", RTLD_NOW);
        if (!h_) return;
        create_t cr = (create_t)dlsym(h_, "ZSTD_createDStream");
        free_ = (free_t)dlsym(h_, "ZSTD_freeDStream");
        dec_ = (dec_t)dlsym(h_, "ZSTD_decompressStream");
        iserr_ = (iserr_t)dlsym(h_, "ZSTD_isError");
        if (cr && free_ && dec_ && iserr_) z_ = cr();
        ib_.p = in_.data(); ib_.size = 0; ib_.pos = 0;
    }
    ~ZstdSource() override {
        if (z_) free_(z_);
        if (f_) fclose(f_);
        if (h_) dlclose(h_);
    }
    bool ok() const { return z_ != nullptr; }
    long read(uint8_t *dst, size_t n) override {
        ZBuf ob = {dst, n, 0};
        while (ob.pos < ob.size) {
            if (ib_.pos == ib_.size) {
                if (feof_) break;
                size_t r = fread(in_.data(), 1, in_.size(), f_);
                if (r < in_.size()) feof_ = true;
                ib_.size = r; ib_.pos = 0;
                if (r == 0) break;
            }
            size_t rc = dec_(z_, &ob, &ib_);
            if (iserr_(rc)) return -1;
        }
        return (long)ob.pos;
    }
  private:
    FILE *f_;
    std::vector<uint8_t> in_;
    ZBuf ib_;
    void *h_ = nullptr, *z_ = nullptr;
    free_t free_ = nullptr; dec_t dec_ = nullptr; iserr_t iserr_ = nullptr;
    bool feof_ = false;
};

}  // namespace


// ------------------------------------------------------------------------------------
// FASTA ingest: rust-bio 1.6.0 fasta::Reader::read + Records, restated over whole
// chunks of text so that chunks parse independently on parallel threads.
// ------------------------------------------------------------------------------------
void HxhRecordBatch::clear() {
    offsets.assign(1, 0);
    ids.clear();
    id_off.assign(1, 0);
    descs.clear();
    has_desc.clear();
    stop = false;
}

bool HxhRecordBatch::reserve(size_t bytes) {  // contents are not preserved: callers reserve before they fill
    if (bytes <= arena_cap) return true;
    if (arena) {
        if (pinned) hx_free_pinned(arena);
        else free(arena);
        arena = nullptr;
        arena_cap = 0;
    }
    const size_t want = std::max(bytes + bytes / 16, (size_t)1 << 20);  // chunks of one input are all about the same size
    uint8_t *p = (uint8_t *)hx_alloc_pinned(want);
    pinned = p != nullptr;
    if (!p) p = (uint8_t *)malloc(want);  // no CUDA device (CPU-only reader tests): plain memory
    if (!p) return false;
    arena = p;
    arena_cap = want;
    return true;
}

HxhRecordBatch::~HxhRecordBatch() {
    if (arena) {
        if (pinned) hx_free_pinned(arena);
        else free(arena);
    }
}

namespace {
inline bool is_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
}  // namespace

int hxh_parse_fasta_text(const uint8_t *p, size_t n, bool drop_last, bool keep_desc, HxhRecordBatch &b) {
    b.clear();
    if (!b.reserve(n + 64)) return HX_ERR_NOMEM;  // the sequence bytes of a text are fewer than the text
    uint8_t *const dst = b.arena;
    size_t used = 0;
    const uint8_t *cur = p, *const end = p + n;
    while (cur < end) {
        // header line: "Expected > at record start." ends the iteration
        if (*cur != '>') {
            b.stop = true;
            break;
        }
        const uint8_t *nl = (const uint8_t *)memchr(cur, '\n', (size_t)(end - cur));
        const uint8_t *le = nl ? nl : end;
        // line[1..].trim_end().splitn(2, char::is_whitespace): id = up to the first whitespace, desc = the rest
        const uint8_t *he = le;
        while (he > cur + 1 && is_ws(he[-1])) --he;
        const uint8_t *ie = cur + 1;
        while (ie < he && !is_ws(*ie)) ++ie;
        const bool has_desc = ie < he;
        const uint8_t *const id_s = cur + 1;
        cur = nl ? nl + 1 : end;
        const size_t seq_start = used;
        while (cur < end && *cur != '>') {  // a line that starts with '>' starts the next record
            nl = (const uint8_t *)memchr(cur, '\n', (size_t)(end - cur));
            le = nl ? nl : end;
            const uint8_t *te = le;
            while (te > cur && is_ws(te[-1])) --te;  // trim_end(): "\r", trailing blanks; interior bytes are kept
            memcpy(dst + used, cur, (size_t)(te - cur));
            used += (size_t)(te - cur);
            cur = nl ? nl + 1 : end;
        }
        if (ie == id_s && !has_desc && used == seq_start) {  // record.is_empty() ends the iteration
            b.stop = true;
            break;
        }
        b.ids.insert(b.ids.end(), (const char *)id_s, (const char *)ie);
        b.id_off.push_back(b.ids.size());
        b.offsets.push_back(used);
        if (keep_desc) {
            b.descs.emplace_back(has_desc ? std::string((const char *)ie + 1, (const char *)he) : std::string());
            b.has_desc.push_back(has_desc ? 1 : 0);
        }
    }
    if (drop_last && !b.stop) {  // the read that failed takes the record in progress with it and ends the iteration
        if (b.offsets.size() > 1) {
            b.offsets.pop_back();
            b.id_off.pop_back();
            b.ids.resize(b.id_off.back());
            if (keep_desc) {
                b.descs.pop_back();
                b.has_desc.pop_back();
            }
        }
        b.stop = true;
    }
    return HX_OK;
}

HxhChunker::HxhChunker() {}
HxhChunker::~HxhChunker() {
    delete src_;
    if (map_) munmap((void *)map_, map_len_);
}

int HxhChunker::open(const char *path, std::string &err) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        err = std::string("Failed to open file: ") + path;
        return HX_ERR_IO;
    }
    uint8_t magic[6] = {0, 0, 0, 0, 0, 0};
    const size_t n = fread(magic, 1, 5, f);
    if (n < 5) {  // niffler: FileTooShort
        fclose(f);
        err = std::string("Failed to determine compression for: ") + path;
        return HX_ERR_IO;
    }
    rewind(f);
    bool ok = true;
    if (magic[0] == 0x1f && magic[1] == 0x8b) {
        fclose(f);
        GzSource *g = new GzSource(path);
        ok = g->ok();
        src_ = g;
    } else if (magic[0] == 0x42 && magic[1] == 0x5a) {
        Bz2Source *b = new Bz2Source(f);
        ok = b->ok();
        src_ = b;
    } else if (magic[0] == 0xfd && magic[1] == 0x37 && magic[2] == 0x7a && magic[3] == 0x58 && magic[4] == 0x5a) {
        XzSource *x = new XzSource(f);
        ok = x->ok();
        src_ = x;
    } else if (magic[0] == 0x28 && magic[1] == 0xb5 && magic[2] == 0x2f && magic[3] == 0xfd) {
        ZstdSource *z = new ZstdSource(f);
        ok = z->ok();
        src_ = z;
    } else {
        // plain text: map the file, chunks are views into the mapping (no read() copy at all)
        struct stat st;
        if (fstat(fileno(f), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(f), 0);
            if (m != MAP_FAILED) {
                map_ = (const uint8_t *)m;
                map_len_ = (size_t)st.st_size;
                madvise(m, map_len_, MADV_SEQUENTIAL);
                madvise(m, map_len_, MADV_WILLNEED);
                fclose(f);
                return HX_OK;
            }
        }
        src_ = new PlainSource(f);
    }
    if (!ok) {
        err = std::string("Failed to determine compression for: ") + path + " (decoder library unavailable)";
        return HX_ERR_IO;
    }
    return HX_OK;
}

bool HxhChunker::next(size_t target, std::vector<uint8_t> &buf, const uint8_t *&p, size_t &n, bool &ended_by_error) {
    ended_by_error = false;
    if (target < 1) target = 1;
    if (map_) {
        if (map_pos_ >= map_len_) return false;
        const size_t s = map_pos_;
        size_t e = map_len_ - s > target ? s + target : map_len_;
        if (e < map_len_) {  // first line start at or after e that begins with '>'
            const uint8_t *q = map_ + e, *const end = map_ + map_len_;
            for (;;) {
                q = (const uint8_t *)memchr(q, '>', (size_t)(end - q));
                if (!q) {
                    e = map_len_;
                    break;
                }
                if (q[-1] == '\n') {
                    e = (size_t)(q - map_);
                    break;
                }
                ++q;
            }
        }
        p = map_ + s;
        n = e - s;
        map_pos_ = e;
        return true;
    }
    if (!src_) return false;
    buf.clear();
    buf.swap(carry_);  // starts with what followed the previous chunk's last record boundary
    for (;;) {
        while (buf.size() < target && !eof_) {
            const size_t have = buf.size(), want = std::max<size_t>(target - have, (size_t)1 << 16);
            buf.resize(have + want);
            const long r = src_->read(buf.data() + have, want);
            buf.resize(have + (r > 0 ? (size_t)r : 0));
            if (r < 0) error_ = true;
            if (r <= 0) eof_ = true;
        }
        if (eof_) break;
        // cut at the last line that starts with '>' (not the chunk's own first line)
        size_t cut = 0;
        for (size_t len = buf.size(); len > 1;) {
            const uint8_t *q = (const uint8_t *)memrchr(buf.data() + 1, '>', len - 1);
            if (!q) break;
            if (q[-1] == '\n') {
                cut = (size_t)(q - buf.data());
                break;
            }
            len = (size_t)(q - buf.data());
        }
        if (cut) {
            carry_.assign(buf.begin() + (ptrdiff_t)cut, buf.end());
            buf.resize(cut);
            break;
        }
        target = buf.size() * 2;  // one record larger than the target: keep reading
    }
    if (buf.empty()) return false;
    ended_by_error = eof_ && error_ && carry_.empty();
    p = buf.data();
    n = buf.size();
    return true;
}

// ------------------------------------------------------------------------------------
// The reader behind hx_fasta_open / hx_fasta_next / hx_fasta_next_packed.  With one thread (the default) a call cuts the
// next chunk, parses it and returns.  With more (hx_fasta_configure) the chunks are parsed — and, when asked, packed to
// 4 bits per base — ahead of the caller by background threads: a thread takes a free slot and cuts the next chunk in
// one step (one cutter at a time), so slots go to chunks in input order and the chunk the caller waits for always owns
// one; batches are handed out in input order.  The iteration ends with the first chunk whose parse set `stop`
// (rust-bio's Records stops at the first malformed / empty record): what was parsed ahead of it is dropped.
// ------------------------------------------------------------------------------------
struct hx_fasta {
    struct Slot {
        HxhRecordBatch batch;
        std::vector<uint8_t> buf;  // decoded text of a compressed input
        uint8_t *packed = nullptr;
        size_t packed_cap = 0;
        bool packed_pinned = false, is_packed = false;
        std::vector<uint8_t> rec_flags;
        std::vector<char> id_cstr;  // ids with terminators, for the C strings handed out
        std::vector<const char *> id_ptrs, desc_ptrs;
        int rc = HX_OK;
        ~Slot() {
            if (packed) {
                if (packed_pinned) hx_free_pinned(packed);
                else free(packed);
            }
        }
    };
    HxhChunker chunker;
    int threads = 1;
    bool pack_ahead = false, started = false;
    std::vector<std::unique_ptr<Slot>> slots;
    std::vector<Slot *> free_slots;
    std::map<uint64_t, Slot *> ready;
    Slot *held = nullptr;  // the batch in the caller's hands (valid until the next call)
    Slot *held_view = nullptr;  // ... as the accessors see it (NULL after the end / an error)
    std::mutex mu, cut_mu;
    std::condition_variable cv;
    std::vector<std::thread> workers;
    uint64_t n_cut = 0, next_out = 0, last_index = UINT64_MAX;
    size_t chunk_bytes = (size_t)256 << 20;
    bool cut_done = false, quit = false;

    static void finish(Slot &s) {  // C strings of the batch
        const HxhRecordBatch &b = s.batch;
        const size_t n = b.n_records();
        s.id_cstr.clear();
        s.id_cstr.reserve(b.ids.size() + n);
        s.id_ptrs.resize(n);
        s.desc_ptrs.resize(n);
        for (size_t i = 0; i < n; ++i) {
            s.id_cstr.insert(s.id_cstr.end(), b.ids.begin() + (ptrdiff_t)b.id_off[i], b.ids.begin() + (ptrdiff_t)b.id_off[i + 1]);
            s.id_cstr.push_back('\0');
        }
        size_t at = 0;
        for (size_t i = 0; i < n; ++i) {
            s.id_ptrs[i] = s.id_cstr.data() + at;
            at += (size_t)(b.id_off[i + 1] - b.id_off[i]) + 1;
            s.desc_ptrs[i] = b.has_desc[i] ? b.descs[i].c_str() : nullptr;
        }
    }

    static int pack(Slot &s, int n_threads) {
        const HxhRecordBatch &b = s.batch;
        const size_t n = b.n_records();
        const size_t want = (size_t)(b.offsets[n] / 2 + 1);
        if (want > s.packed_cap) {
            if (s.packed) {
                if (s.packed_pinned) hx_free_pinned(s.packed);
                else free(s.packed);
                s.packed = nullptr;
                s.packed_cap = 0;
            }
            const size_t cap = std::max(want + want / 16, (size_t)1 << 19);
            uint8_t *p = (uint8_t *)hx_alloc_pinned(cap);
            s.packed_pinned = p != nullptr;
            if (!p) p = (uint8_t *)malloc(cap);
            if (!p) return HX_ERR_NOMEM;
            s.packed = p;
            s.packed_cap = cap;
        }
        s.rec_flags.assign(std::max<size_t>(n, 1), 0);
        s.packed[want - 1] = 0;  // the last byte may be half used
        const int rc = hxh_pack_slab(b.arena, b.offsets.data(), (uint32_t)n, s.packed, 0, s.rec_flags.data(), n_threads);
        s.is_packed = rc == HX_OK;
        return rc;
    }

    // cuts the next chunk into `s` and parses it; false at the end of the input
    bool cut_and_parse(Slot &s, size_t target, bool do_pack, int pack_threads, uint64_t *index, std::unique_lock<std::mutex> *cut_lock) {
        const uint8_t *p = nullptr;
        size_t n = 0;
        bool err = false;
        const bool got = chunker.next(target, s.buf, p, n, err);
        if (got && index) {
            std::lock_guard<std::mutex> lk(mu);
            *index = n_cut++;
        }
        if (cut_lock) cut_lock->unlock();  // the next cutter may go on while this chunk is parsed
        if (!got) return false;
        s.is_packed = false;
        s.rc = hxh_parse_fasta_text(p, n, err, true, s.batch);
        if (s.rc == HX_OK && s.batch.n_records() > UINT32_MAX) s.rc = HX_ERR_ARG;
        if (s.rc == HX_OK && do_pack) s.rc = pack(s, pack_threads);
        if (s.rc == HX_OK) finish(s);
        return true;
    }

    void worker() {
        for (;;) {
            Slot *s = nullptr;
            size_t target = 0;
            uint64_t index = 0;
            std::unique_lock<std::mutex> cut(cut_mu);
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || cut_done || n_cut > last_index || !free_slots.empty(); });
                if (quit || cut_done || n_cut > last_index) return;
                s = free_slots.back();
                free_slots.pop_back();
                target = chunk_bytes;
            }
            if (!cut_and_parse(*s, target, pack_ahead, 1, &index, &cut)) {
                std::lock_guard<std::mutex> lk(mu);
                cut_done = true;
                free_slots.push_back(s);
                cv.notify_all();
                return;
            }
            std::lock_guard<std::mutex> lk(mu);
            if ((s->rc != HX_OK || s->batch.stop) && index < last_index) last_index = index;
            ready[index] = s;
            cv.notify_all();
        }
    }

    // Slots (pinned arenas, packed buffers, id tables) of closed readers wait here for the next reader of the process:
    // pinning memory costs about 1.5 ms per MB on the B200 hosts — more than parsing the text that goes into it — so a
    // caller that reads file after file pays for its slots once.  At most HYPEREX_READER_CACHE_MB (default 1024) are kept.
    struct SlotPool {
        std::mutex mu;
        std::vector<std::unique_ptr<Slot>> idle;
        size_t bytes = 0;
    };
    static SlotPool &pool() {
        static SlotPool *p = new SlotPool();  // never destroyed: cudaFreeHost after the runtime has shut down is an error
        return *p;
    }
    static size_t pool_cap() {
        const char *e = getenv("HYPEREX_READER_CACHE_MB");
        return (size_t)(e ? std::max(0L, atol(e)) : 1024L) << 20;
    }

    void start() {
        started = true;
        const int n_slots = threads > 1 ? threads + 2 : 1;
        {
            SlotPool &P = pool();
            std::lock_guard<std::mutex> lk(P.mu);
            while ((int)slots.size() < n_slots && !P.idle.empty()) {
                P.bytes -= P.idle.back()->batch.arena_cap + P.idle.back()->packed_cap;
                slots.push_back(std::move(P.idle.back()));
                P.idle.pop_back();
            }
        }
        while ((int)slots.size() < n_slots) slots.emplace_back(new Slot());
        for (auto &s : slots) {
            s->rc = HX_OK;
            s->is_packed = false;
            free_slots.push_back(s.get());
        }
        if (threads > 1)
            for (int i = 0; i < threads; ++i) workers.emplace_back([this] { worker(); });
    }

    // the next batch in input order: > 0 records and *out, 0 at the end of the iteration, < 0 hx_status
    int64_t next(size_t max_bytes, Slot **out) {
        *out = nullptr;
        if (!started) {
            chunk_bytes = max_bytes;
            start();
        }
        if (threads <= 1) {
            Slot &s = *slots[0];
            if (cut_done) return 0;
            if (!cut_and_parse(s, max_bytes, false, 1, nullptr, nullptr)) {
                cut_done = true;
                return 0;
            }
            if (s.rc != HX_OK || s.batch.stop) cut_done = true;
            if (s.rc != HX_OK) return s.rc;
            *out = &s;
            return (int64_t)s.batch.n_records();
        }
        std::unique_lock<std::mutex> lk(mu);
        chunk_bytes = max_bytes;
        if (held) {
            free_slots.push_back(held);
            held = nullptr;
            cv.notify_all();
        }
        Slot *s = nullptr;
        for (;;) {
            if (next_out > last_index) return 0;
            auto it = ready.find(next_out);
            if (it != ready.end()) {
                s = it->second;
                ready.erase(it);
                break;
            }
            if (cut_done && next_out >= n_cut) return 0;
            cv.wait(lk);
        }
        ++next_out;
        held = s;
        if (s->rc != HX_OK) return s->rc;
        *out = s;
        return (int64_t)s->batch.n_records();
    }

    ~hx_fasta() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            cv.notify_all();
        }
        for (std::thread &t : workers) t.join();
        SlotPool &P = pool();
        const size_t cap = pool_cap();
        std::lock_guard<std::mutex> lk(P.mu);
        for (auto &s : slots) {
            const size_t b = s->batch.arena_cap + s->packed_cap;
            if (b == 0 || P.bytes + b > cap) continue;  // freed with the reader
            std::vector<uint8_t>().swap(s->buf);
            P.bytes += b;
            P.idle.push_back(std::move(s));
        }
    }
};

extern "C" int hx_fasta_open(const char *path, hx_fasta **out) {
    if (!path || !out) return HX_ERR_ARG;
    *out = nullptr;
    hx_fasta *r = new (std::nothrow) hx_fasta();
    if (!r) return HX_ERR_NOMEM;
    std::string err;
    int rc = r->chunker.open(path, err);
    if (rc != HX_OK) {
        delete r;
        return rc;
    }
    *out = r;
    return HX_OK;
}

extern "C" int hx_fasta_configure(hx_fasta *r, int n_threads, int pack_ahead) {
    if (!r) return HX_ERR_ARG;
    if (r->started) return HX_ERR_STATE;
    // automatic: a parser thread moves 3-6 GB/s of text, eight of them more than one PCIe link takes
    if (n_threads <= 0) n_threads = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() - 1));
    r->threads = std::min(n_threads, 256);
    r->pack_ahead = pack_ahead != 0;
    return HX_OK;
}

extern "C" int64_t hx_fasta_next(hx_fasta *r, size_t max_bytes, const uint8_t **arena, const uint64_t **offsets,
                                 const char *const **ids) {
    if (!r) return HX_ERR_ARG;
    hx_fasta::Slot *s = nullptr;
    const int64_t n = r->next(max_bytes ? max_bytes : ((size_t)256 << 20), &s);
    if (n <= 0 || !s) {
        r->held_view = nullptr;
        return n;
    }
    r->held_view = s;
    if (arena) *arena = s->batch.arena;
    if (offsets) *offsets = s->batch.offsets.data();
    if (ids) *ids = s->id_ptrs.data();
    return n;
}

extern "C" int64_t hx_fasta_next_packed(hx_fasta *r, size_t max_bytes, const uint8_t **packed, const uint8_t **rec_flags,
                                        const uint64_t **offsets, const char *const **ids, const uint8_t **arena) {
    if (!r) return HX_ERR_ARG;
    const uint8_t *a = nullptr;
    const int64_t n = hx_fasta_next(r, max_bytes, &a, offsets, ids);
    if (n <= 0) return n;
    hx_fasta::Slot *s = r->held_view;
    if (!s->is_packed) {  // not packed ahead: all the reader's threads pack it now
        const int rc = hx_fasta::pack(*s, r->threads);
        if (rc != HX_OK) return rc;
    }
    if (packed) *packed = s->packed;
    if (rec_flags) *rec_flags = s->rec_flags.data();
    if (arena) *arena = a;
    return n;
}

extern "C" const char *const *hx_fasta_descs(hx_fasta *r) { return r && r->held_view ? r->held_view->desc_ptrs.data() : nullptr; }

extern "C" void hx_fasta_close(hx_fasta *r) { delete r; }

// ------------------------------------------------------------------------------------
// utils.rs:231-414 get_hypervar_regions, as a pipeline of host threads around the device path:
//
//   chunker  (calling thread)   cuts the input at record boundaries (mapped file: no copy; compressed: the decoder)
//   parsers  (P threads)        chunk -> record batch in a pinned arena            [fasta::Reader, utils.rs:238,270]
//   lanes    (2 per device)     batch -> H2D, scan, on-device FASTA / GFF3 text, D2H through pinned staging
//   writers  (W threads)        pwrite() the staged text at its final file offset  [utils.rs:365-369, 378-385]
//
// Batches finish out of order on different lanes and devices, the files do not: every batch learns the byte totals
// of its two texts on the device before any text is copied, takes its turn (batch index order) to reserve its byte
// range in both files, and then writes into that range from any thread.  The reference's warn!/error! lines are
// emitted by the lane threads, again in batch index order.
// ------------------------------------------------------------------------------------
namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

inline size_t put_u32(char *dst, uint32_t v) {  // decimal, no terminator
    char tmp[12];
    int i = 0;
    do {
        tmp[i++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    for (int j = 0; j < i; ++j) dst[j] = tmp[i - 1 - j];
    return (size_t)i;
}

bool pwrite_all(int fd, const char *p, size_t n, uint64_t off) {
    while (n) {
        const ssize_t w = pwrite(fd, p, n, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += w;
        n -= (size_t)w;
        off += (uint64_t)w;
    }
    return true;
}

// An output file and, when it can be mapped, windows of it mapped MAP_SHARED for the whole call: mmap() / munmap()
// take the process-wide mmap lock exclusively and stall every thread that is faulting pages in, so the file is not
// mapped per block.  A window covers HX_WINDOW bytes of file offsets (mapping past the end of the file is allowed;
// the bytes are only touched after ftruncate made them exist) and is created when the first block lands in it.
struct OutFile {
    static constexpr uint64_t HX_WINDOW = (uint64_t)1 << 36;
    int fd = -1;
    bool mappable = true;
    std::mutex mu;
    std::map<uint64_t, char *> windows;  // window index -> base address
    // address of file offset `off` (the caller never crosses a window boundary); NULL: use pwrite
    char *at(uint64_t off) {
        if (!mappable) return nullptr;
        const uint64_t w = off / HX_WINDOW;
        std::lock_guard<std::mutex> lk(mu);
        auto it = windows.find(w);
        if (it == windows.end()) {
            void *m = mmap(nullptr, HX_WINDOW, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)(w * HX_WINDOW));
            if (m == MAP_FAILED) {
                mappable = false;
                return nullptr;
            }
            it = windows.emplace(w, (char *)m).first;
        }
        return it->second + (off - w * HX_WINDOW);
    }
    // The bytes [off, off + n) are written: drop their page-table entries now (the pages stay in the page cache).  For a
    // shared file mapping MADV_DONTNEED only takes the mmap lock shared, so lanes release their own batches concurrently
    // instead of leaving 1.7 M entries to one munmap at the end (0.24 s of a 1.1 s call for 7 GB of output).
    void release(uint64_t off, uint64_t n) {
        if (!mappable || n < 2 * 4096) return;
        const uint64_t a = (off + 4095) & ~(uint64_t)4095, e = (off + n) & ~(uint64_t)4095;
        for (uint64_t p = a; p < e;) {
            const uint64_t w = p / HX_WINDOW, stop = std::min(e, (w + 1) * HX_WINDOW);
            char *base = nullptr;
            {
                std::lock_guard<std::mutex> lk(mu);
                auto it = windows.find(w);
                if (it != windows.end()) base = it->second;
            }
            if (base) madvise(base + (p - w * HX_WINDOW), (size_t)(stop - p), MADV_DONTNEED);
            p = stop;
        }
    }
    int close_all() {
        for (auto &kv : windows) munmap(kv.second, HX_WINDOW);
        windows.clear();
        const int rc = fd >= 0 ? close(fd) : 0;
        fd = -1;
        return rc;
    }
};

// Writes one block of bytes at a file offset on several threads at once.  Two ways:
//   mapped (default)  the threads copy into the mapped file, each one first pre-faulting its slice with
//                     madvise(MADV_POPULATE_WRITE) — which reports "disk full" as an error where a plain store
//                     would raise SIGBUS.  Page-cache pages are added under per-page locks, so the copies of
//                     different threads run in parallel (measurements: profiles/r02_file_e2e.md).
//   pwrite            buffered write() takes the inode lock exclusively: any number of threads writing to ONE file move
//                     what one thread moves (4.3 GB/s measured on tmpfs).  Used when the file cannot be mapped and
//                     with HYPEREX_WRITE=pwrite.
// The file must already be large enough (the byte range was reserved with ftruncate).  The caller copies the first
// slice itself and returns when all slices are written.
class WritePool {
  public:
    WritePool(int n_threads, bool mapped) : mapped_(mapped) {
        for (int i = 0; i < n_threads; ++i) th_.emplace_back([this] { run(); });
    }
    ~WritePool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : th_) t.join();
    }
    bool mapped() const { return mapped_; }
    bool write(OutFile &f, const char *p, size_t n, uint64_t off) {
        while (n) {  // a block that crosses a mapping window is written window by window
            const size_t part = (size_t)std::min<uint64_t>(n, OutFile::HX_WINDOW - off % OutFile::HX_WINDOW);
            if (!write_slices(f.fd, p, part, off, mapped_ ? f.at(off) : nullptr)) return false;
            p += part;
            n -= part;
            off += part;
        }
        return true;
    }

  private:
    struct Call {
        int pending = 0;
        bool ok = true;
    };
    struct Task {
        int fd;
        const char *p;
        size_t n;
        uint64_t off;
        char *dst;  // mapped destination of this slice, or NULL: pwrite
        Call *call;
    };
    static bool do_task(const Task &t) {
        if (!t.dst) return pwrite_all(t.fd, t.p, t.n, t.off);
        // pre-fault whole pages of the slice; EINVAL = a kernel without MADV_POPULATE_WRITE (< 5.14): plain stores then
        char *a = (char *)((uintptr_t)t.dst & ~(uintptr_t)4095);
        const size_t len = (size_t)(t.dst + t.n - a);
        while (madvise(a, len, MADV_POPULATE_WRITE) != 0) {
            if (errno == EAGAIN || errno == EINTR) continue;
            if (errno == EINVAL) break;
            return false;  // EFAULT / ENOMEM: the store would have raised SIGBUS (no space left on the device)
        }
        memcpy(t.dst, t.p, t.n);
        return true;
    }
    bool write_slices(int fd, const char *p, size_t n, uint64_t off, char *dst) {
        const size_t min_slice = (size_t)1 << 20;
        const size_t parts = std::min<size_t>(th_.size() + 1, (n + min_slice - 1) / min_slice);
        if (parts <= 1) return do_task(Task{fd, p, n, off, dst, nullptr});
        // slices end on page boundaries of the FILE so that two threads never fault the same page
        const size_t slice = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
        const size_t first = slice - (size_t)(off & 4095);
        Call call;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t s = first; s < n; s += slice) {
                q_.push_back(Task{fd, p + s, std::min(slice, n - s), off + s, dst ? dst + s : nullptr, &call});
                ++call.pending;
            }
        }
        cv_.notify_all();
        const bool ok = do_task(Task{fd, p, std::min(first, n), off, dst, nullptr});
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return call.pending == 0; });
        return ok && call.ok;
    }
    void run() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            const bool ok = do_task(t);
            std::lock_guard<std::mutex> lk(mu_);
            if (!ok) t.call->ok = false;
            if (--t.call->pending == 0) done_cv_.notify_all();
        }
    }
    const bool mapped_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Task> q_;
    bool stop_ = false;
};

// Host-side state kept in the context across calls (hxh_host_cache): pinned memory is expensive to allocate
// (tens of milliseconds per 100 MB), so record batches and the lanes' result tables are reused.
struct PinnedTable {
    hx_result *p = nullptr;
    size_t cap = 0;  // entries
    bool pinned = false;
    hx_result *get(size_t n) {
        if (n <= cap) return p;
        release();
        const size_t want = n + n / 4 + 1024;
        p = (hx_result *)hx_alloc_pinned(want * sizeof(hx_result));
        pinned = p != nullptr;
        if (!p) p = (hx_result *)malloc(want * sizeof(hx_result));
        cap = p ? want : 0;
        return p;
    }
    void release() {
        if (p) {
            if (pinned) hx_free_pinned(p);
            else free(p);
        }
        p = nullptr;
        cap = 0;
    }
};

struct HostCache {
    std::mutex mu;
    std::vector<HxhRecordBatch *> batches;
    std::vector<PinnedTable> tables;  // per lane
    double stats[HX_FILE_STATS_N] = {0};
    ~HostCache() {
        for (HxhRecordBatch *b : batches) delete b;
        for (PinnedTable &t : tables) t.release();
    }
};
void free_host_cache(void *p) { delete (HostCache *)p; }

HostCache *host_cache_of(hx_ctx *ctx) {
    void **slot = hxh_host_cache(ctx, free_host_cache);
    if (!*slot) *slot = new HostCache();
    return (HostCache *)*slot;
}

struct Chunk {
    uint64_t index;
    const uint8_t *p;
    size_t n;
    bool drop_last;
    std::vector<uint8_t> *buf;  // decoded bytes of a compressed input (NULL: view into the mapped file)
};

struct FileJob {
    hx_ctx *ctx = nullptr;
    uint32_t n_pairs = 0;
    const char *const *fwd = nullptr, *const *rev = nullptr;
    hx_log_fn log_fn = nullptr;
    void *user = nullptr;
    bool device_text = true;
    OutFile out[2];  // FASTA, GFF3
    std::vector<std::string> labels, fa_tail, gff_note;
    std::string tails_blob, gtails_blob;
    std::vector<uint32_t> tail_off, gtail_off;
    WritePool *writers = nullptr;
    HostCache *cache = nullptr;

    std::mutex mu;
    std::condition_variable cv;
    std::deque<Chunk> chunks;
    std::vector<std::vector<uint8_t> *> free_bufs;
    std::vector<HxhRecordBatch *> free_batches;
    std::map<uint64_t, HxhRecordBatch *> ready;
    uint64_t parsed_prefix = 0;        // every chunk with a smaller index has been parsed (its `stop` is known)
    std::set<uint64_t> parsed_ahead;   // parsed indices beyond the prefix
    uint64_t n_chunks = 0;             // chunks cut so far (final once chunker_done)
    bool chunker_done = false;
    uint64_t last_index = UINT64_MAX;  // the batch on which the reader's iteration ended (none after it counts)
    uint64_t next_claim = 0;           // next batch index a lane takes
    uint64_t out_turn = 0, log_turn = 0;
    uint64_t file_pos[2] = {0, 0};     // next free byte of each output file
    int rc = HX_OK;
    bool abort = false;
    std::string err;
    // statistics (seconds are summed over the threads of a stage)
    double t_parse = 0, t_device = 0, t_write = 0, t_log = 0;
    uint64_t n_bases = 0, n_records = 0, n_batches = 0;

    void fail(int code, const std::string &msg) {
        std::lock_guard<std::mutex> lk(mu);
        if (rc == HX_OK) {
            rc = code;
            err = msg;
        }
        abort = true;
        cv.notify_all();
    }
};

void parser_thread(FileJob *J) {
    for (;;) {
        Chunk c;
        HxhRecordBatch *b = nullptr;
        {
            // a chunk and a free batch are taken together, so batches go to chunks in index order and the lowest
            // unparsed index can never starve behind later ones
            std::unique_lock<std::mutex> lk(J->mu);
            for (;;) {
                J->cv.wait(lk, [&] {
                    return J->abort || (J->chunker_done && J->chunks.empty()) ||
                           (!J->chunks.empty() && (J->chunks.front().index > J->last_index || !J->free_batches.empty()));
                });
                if (J->abort || J->chunks.empty()) return;
                c = J->chunks.front();
                if (c.index <= J->last_index) break;
                J->chunks.pop_front();  // cut before the reader's iteration was known to have ended: dropped unparsed
                if (c.buf) J->free_bufs.push_back(c.buf);
                J->cv.notify_all();
            }
            J->chunks.pop_front();
            b = J->free_batches.back();
            J->free_batches.pop_back();
        }
        const double t0 = now_s();
        const int rc = hxh_parse_fasta_text(c.p, c.n, c.drop_last, false, *b);
        const double dt = now_s() - t0;
        if (rc != HX_OK) {
            J->fail(rc, "out of memory while reading the input");
            return;
        }
        std::lock_guard<std::mutex> lk(J->mu);
        J->t_parse += dt;
        if (c.buf) J->free_bufs.push_back(c.buf);
        if (b->stop && c.index < J->last_index) J->last_index = c.index;
        J->ready[c.index] = b;
        J->parsed_ahead.insert(c.index);
        while (J->parsed_ahead.erase(J->parsed_prefix)) ++J->parsed_prefix;
        J->cv.notify_all();
    }
}

struct LaneSink {  // HxhTextSink of one batch
    FileJob *J;
    uint64_t index;
    uint64_t base[2], total[2];
    bool begun;
    double t_write;
};

// takes the batch's turn to reserve its byte range in both files; false when the job was aborted
bool reserve_output(FileJob *J, uint64_t index, uint64_t fa_total, uint64_t gff_total, uint64_t base[2]) {
    std::unique_lock<std::mutex> lk(J->mu);
    J->cv.wait(lk, [&] { return J->abort || J->out_turn == index; });
    if (J->abort) return false;
    base[0] = J->file_pos[0];
    base[1] = J->file_pos[1];
    J->file_pos[0] += fa_total;
    J->file_pos[1] += gff_total;
    if (J->writers->mapped()) {  // the mapped copies need the bytes to exist; sizes only ever grow, in batch order
        if ((fa_total && ftruncate(J->out[0].fd, (off_t)J->file_pos[0]) != 0) || (gff_total && ftruncate(J->out[1].fd, (off_t)J->file_pos[1]) != 0)) {
            if (J->rc == HX_OK) {
                J->rc = HX_ERR_IO;
                J->err = "cannot extend the output files";
            }
            J->abort = true;
            J->cv.notify_all();
            return false;
        }
    }
    J->out_turn = index + 1;
    J->cv.notify_all();
    return true;
}

int sink_begin(void *user, uint64_t fa_total, uint64_t gff_total) {
    LaneSink *s = (LaneSink *)user;
    s->begun = true;
    s->total[0] = fa_total;
    s->total[1] = gff_total;
    return reserve_output(s->J, s->index, fa_total, gff_total, s->base) ? 0 : 1;
}

int sink_chunk(void *user, int kind, uint64_t off, const char *bytes, size_t n) {
    LaneSink *s = (LaneSink *)user;
    const double t0 = now_s();
    const bool ok = s->J->writers->write(s->J->out[kind], bytes, n, s->base[kind] + off);
    s->t_write += now_s() - t0;
    return ok ? 0 : 1;
}

// the reference's per-record / per-pair log lines of one batch (utils.rs:276-288, 387-408), in record x pair order
void emit_logs(FileJob *J, const HxhRecordBatch &b, const hx_result *res) {
    std::string msg, id;
    char num[32];
    const uint32_t n = (uint32_t)b.n_records(), np = J->n_pairs;
    for (uint32_t r = 0; r < n; ++r) {
        const hx_result *rr = res + (size_t)r * np;
        const size_t len = (size_t)(b.offsets[r + 1] - b.offsets[r]);
        bool need = (rr[0].flags & HX_ALPHABET_NONE) || len <= 1500;
        for (uint32_t p = 0; p < np && !need; ++p) need = (rr[p].flags & 7u) != 7u;
        if (!need) continue;
        id.assign(b.ids.data() + b.id_off[r], b.ids.data() + b.id_off[r + 1]);
        if (rr[0].flags & HX_ALPHABET_NONE) {  // utils.rs:276-279
            msg = "Unrecognized sequence type in record: " + id;
            J->log_fn(2, msg.c_str(), J->user);
            continue;
        }
        if (len <= 1500) {  // utils.rs:282-288
            snprintf(num, sizeof num, "%zu", len);
            msg = "Sequence " + id + " is short (" + num + " bp). Some regions may not be found.";
            J->log_fn(1, msg.c_str(), J->user);
        }
        for (uint32_t p = 0; p < np; ++p) {
            const hx_result &x = rr[p];
            const bool f_any = x.flags & HX_FWD_ANY, r_any = x.flags & HX_REV_ANY, ok = x.flags & HX_PAIR_VALID;
            if (f_any && r_any && ok) continue;
            if (f_any && r_any)  // utils.rs:387-408
                msg = "Forward and reverse primers for region " + J->labels[p] + " were both found in sequence " + id +
                      ", but not in a valid order/orientation to define a region";
            else if (!f_any && r_any)
                msg = std::string("Forward primer ") + J->fwd[p] + " not found for region " + J->labels[p] + " in sequence " + id;
            else if (f_any && !r_any)
                msg = std::string("Reverse primer ") + J->rev[p] + " not found for region " + J->labels[p] + " in sequence " + id;
            else
                msg = "Both primers not found for region " + J->labels[p] + " in sequence " + id;
            J->log_fn(1, msg.c_str(), J->user);
        }
    }
}

// host formatter (HYPEREX_HOST_FORMAT=1): the bytes of utils.rs:354-385 for one batch from its result table
void format_on_host(const FileJob *J, const HxhRecordBatch &b, const hx_result *res, std::vector<char> &fa, std::vector<char> &gff) {
    fa.clear();
    gff.clear();
    const uint32_t n = (uint32_t)b.n_records(), np = J->n_pairs;
    for (uint32_t r = 0; r < n; ++r) {
        const hx_result *rr = res + (size_t)r * np;
        if (rr[0].flags & HX_ALPHABET_NONE) continue;
        const char *id = b.ids.data() + b.id_off[r];
        const size_t idn = (size_t)(b.id_off[r + 1] - b.id_off[r]);
        const uint8_t *seq = b.arena + b.offsets[r];
        for (uint32_t p = 0; p < np; ++p) {
            const hx_result &x = rr[p];
            if ((x.flags & 7u) != 7u) continue;
            const size_t slen = (size_t)x.r_end + 1 - x.f_start;
            const std::string &ft = J->fa_tail[p], &gt = J->gff_note[p];
            size_t o = fa.size();
            fa.resize(o + 1 + idn + ft.size() + slen + 1);
            fa[o++] = '>';
            memcpy(&fa[o], id, idn);
            o += idn;
            memcpy(&fa[o], ft.data(), ft.size());
            o += ft.size();
            memcpy(&fa[o], seq + x.f_start, slen);  // original case (utils.rs:368)
            fa[o + slen] = '\n';
            size_t g = gff.size();
            gff.resize(g + idn + 16 + 24 + gt.size());
            memcpy(&gff[g], id, idn);
            g += idn;
            memcpy(&gff[g], "\thyperex\tregion\t", 16);
            g += 16;
            g += put_u32(&gff[g], x.f_start + 1);  // 1-based (utils.rs:382-383)
            gff[g++] = '\t';
            g += put_u32(&gff[g], x.r_end + 1);
            memcpy(&gff[g], gt.data(), gt.size());
            gff.resize(g + gt.size());
        }
    }
}

void lane_thread(FileJob *J, int lane) {
    std::vector<char> fa_host, gff_host;
    for (;;) {
        uint64_t index;
        HxhRecordBatch *b = nullptr;
        {
            std::unique_lock<std::mutex> lk(J->mu);
            index = J->next_claim++;
            // a batch may only start once every earlier chunk is parsed: one of them may end the reader's iteration
            J->cv.wait(lk, [&] {
                return J->abort || (J->ready.count(index) && J->parsed_prefix >= index) || index > J->last_index ||
                       (J->chunker_done && index >= J->n_chunks);
            });
            if (J->abort || index > J->last_index || !J->ready.count(index)) return;
            b = J->ready[index];
            J->ready.erase(index);
        }
        const uint32_t n = (uint32_t)b->n_records();
        const size_t n_entries = (size_t)n * J->n_pairs;
        const bool want_table = J->log_fn != nullptr || !J->device_text;
        hx_result *res = nullptr;
        if (want_table && n) {
            res = J->cache->tables[lane].get(n_entries);
            if (!res) {
                J->fail(HX_ERR_NOMEM, "out of memory for the result table");
                return;
            }
        }
        const double t0 = now_s();
        double t_write = 0;
        int rc = HX_OK;
        if (n == 0) {
            uint64_t base[2];
            if (!reserve_output(J, index, 0, 0, base)) return;
        } else if (J->device_text) {
            LaneSink sink{J, index, {0, 0}, {0, 0}, false, 0.0};
            const HxhTextSink hs{sink_begin, sink_chunk, &sink};
            rc = hxh_run_text_lane(J->ctx, lane, b->arena, b->offsets.data(), n, res, b->ids.data(), b->id_off.data(),
                                   J->tails_blob.data(), J->tail_off.data(), J->gtails_blob.data(), J->gtail_off.data(), &hs);
            t_write = sink.t_write;
            if (rc != HX_OK) {
                J->fail(rc, hx_last_error(J->ctx));
                return;
            }
            if (sink.begun) {
                const double tr = now_s();
                J->out[0].release(sink.base[0], sink.total[0]);
                J->out[1].release(sink.base[1], sink.total[1]);
                t_write += now_s() - tr;
            }
        } else {
            rc = hxh_run_batch_lane(J->ctx, lane, b->arena, b->offsets.data(), n, res);
            if (rc != HX_OK) {
                J->fail(rc, hx_last_error(J->ctx));
                return;
            }
            const double tw = now_s();
            format_on_host(J, *b, res, fa_host, gff_host);
            uint64_t base[2];
            if (!reserve_output(J, index, fa_host.size(), gff_host.size(), base)) return;
            const bool ok = J->writers->write(J->out[0], fa_host.data(), fa_host.size(), base[0]) &&
                            J->writers->write(J->out[1], gff_host.data(), gff_host.size(), base[1]);
            J->out[0].release(base[0], fa_host.size());
            J->out[1].release(base[1], gff_host.size());
            t_write = now_s() - tw;
            if (!ok) {
                J->fail(HX_ERR_IO, "write error on the output files");
                return;
            }
        }
        const double t1 = now_s();
        double t_log = 0;
        if (J->log_fn) {  // in batch order, like the single loop of the reference
            std::unique_lock<std::mutex> lk(J->mu);
            J->cv.wait(lk, [&] { return J->abort || J->log_turn == index; });
            if (J->abort) return;
            lk.unlock();
            const double tl = now_s();
            if (n) emit_logs(J, *b, res);
            t_log = now_s() - tl;
            lk.lock();
            J->log_turn = index + 1;
            J->cv.notify_all();
        }
        std::lock_guard<std::mutex> lk(J->mu);
        J->t_device += (t1 - t0) - t_write;
        J->t_write += t_write;
        J->t_log += t_log;
        J->n_bases += b->offsets[n];
        J->n_records += n;
        J->n_batches += 1;
        J->free_batches.push_back(b);
        J->cv.notify_all();
    }
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

}  // namespace

extern "C" int hx_get_hypervar_regions(hx_ctx *ctx, const char *file, uint32_t n_pairs, const char *const *fwd,
                                       const char *const *rev, const char *prefix, uint8_t mismatch, hx_log_fn log_fn,
                                       void *user) {
    if (!ctx || !file || !fwd || !rev || !prefix || n_pairs == 0) return HX_ERR_ARG;
    const double t_start = now_s();
    int rc = hx_set_primers(ctx, n_pairs, fwd, nullptr, rev, nullptr, mismatch);
    if (rc != HX_OK) return rc;
    HxhChunker chunker;
    std::string err;
    rc = chunker.open(file, err);
    if (rc != HX_OK) {
        if (log_fn) log_fn(2, err.c_str(), user);
        hxh_set_error(ctx, err.c_str());
        return rc;
    }
    FileJob J;
    J.ctx = ctx;
    J.n_pairs = n_pairs;
    J.fwd = fwd;
    J.rev = rev;
    J.log_fn = log_fn;
    J.user = user;
    // On-device extraction is the default on every device of the context: the FASTA / GFF3 text itself
    // is assembled on the GPU and the host only writes it; HYPEREX_HOST_FORMAT=1 keeps the host formatter.
    J.device_text = env_int("HYPEREX_HOST_FORMAT", 0) == 0;
    const std::string fa_path = std::string(prefix) + ".fa", gff_path = std::string(prefix) + ".gff";
    J.out[0].fd = open(fa_path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);  // fasta::Writer::to_file truncates (utils.rs:241)
    J.out[1].fd = open(gff_path.c_str(), O_RDWR | O_CREAT, 0644);           // OpenOptions create+append (utils.rs:242-245)
    if (J.out[0].fd < 0 || J.out[1].fd < 0) {
        J.out[0].close_all();
        J.out[1].close_all();
        hxh_set_error(ctx, "cannot create the output files");
        return HX_ERR_IO;
    }
    {
        struct stat st;  // append: the GFF3 text starts at the current end of the file
        J.file_pos[1] = fstat(J.out[1].fd, &st) == 0 ? (uint64_t)st.st_size : 0;
        if (!pwrite_all(J.out[1].fd, "##gff-version 3\n", 16, J.file_pos[1])) rc = HX_ERR_IO;  // utils.rs:247
        J.file_pos[1] += 16;
    }
    // per pair: everything of the FASTA header / GFF line that does not depend on the record
    J.labels.resize(n_pairs);
    J.fa_tail.resize(n_pairs);
    J.gff_note.resize(n_pairs);
    J.tail_off.assign(1, 0);
    J.gtail_off.assign(1, 0);
    for (uint32_t p = 0; p < n_pairs; ++p) {
        char lab[16];
        hx_primers_to_region(fwd[p], rev[p], lab);
        J.labels[p] = lab;
        const std::string fr = std::string("forward=") + fwd[p] + " reverse=" + rev[p];
        J.fa_tail[p] = J.labels[p].empty() ? " " + fr + "\n" : " region=" + J.labels[p] + " " + fr + "\n";  // utils.rs:356-363
        J.gff_note[p] = "\t.\t.\t.\tNote=" + (J.labels[p].empty() ? fr : "Hypervariable region " + J.labels[p]) + "\n";  // 372-385
        J.tails_blob += J.fa_tail[p];
        J.tail_off.push_back((uint32_t)J.tails_blob.size());
        J.gtails_blob += J.gff_note[p];
        J.gtail_off.push_back((uint32_t)J.gtails_blob.size());
    }
    // threads: HYPEREX_THREADS caps the parser and the writer pools (default: the machine's hardware threads)
    const int n_lanes = hxh_lane_count(ctx);
    const int hw = std::max(1, env_int("HYPEREX_THREADS", (int)std::thread::hardware_concurrency()));
    // (measured: parsing 1.5 GB of FASTA costs 0.3 thread-seconds in total, so a few parsers keep any number of lanes
    // fed; every extra parser is one more pinned batch to allocate on the first call)
    const int n_parsers = std::max(1, std::min(4, hw / 4));
    const int n_writers = std::max(1, std::min(24, env_int("HYPEREX_WRITERS", hw - 1)));
    // chunk size: large enough to amortise the per-batch device round trip, small enough that every lane gets
    // several batches and the first / last batch (which nothing overlaps) stay short
    size_t chunk_bytes = (size_t)16 << 20;
    if (chunker.mapped()) {
        const size_t share = chunker.mapped_bytes() / ((size_t)n_lanes * 4) + 1;
        chunk_bytes = std::min(chunk_bytes, std::max<size_t>(share, (size_t)4 << 20));
    }
    if (env_int("HYPEREX_CHUNK_BYTES", 0) > 0) chunk_bytes = (size_t)env_int("HYPEREX_CHUNK_BYTES", 0);  // tests: many tiny batches
    HostCache *cache = host_cache_of(ctx);
    J.cache = cache;
    const char *wmode = getenv("HYPEREX_WRITE");
    WritePool writers(n_writers, !(wmode && strcmp(wmode, "pwrite") == 0));
    J.writers = &writers;
    const int n_batches = n_parsers + n_lanes + 1;
    {
        std::lock_guard<std::mutex> lk(cache->mu);
        if ((int)cache->tables.size() < n_lanes) cache->tables.resize(n_lanes);
        while ((int)cache->batches.size() < n_batches) cache->batches.push_back(new HxhRecordBatch());
        J.free_batches.assign(cache->batches.begin(), cache->batches.begin() + n_batches);
    }
    std::vector<std::vector<uint8_t>> bufs(chunker.mapped() ? 0 : n_parsers + 2);
    for (auto &b : bufs) J.free_bufs.push_back(&b);

    const double t_setup = now_s() - t_start;
    std::vector<std::thread> threads;
    if (rc == HX_OK) {
        for (int i = 0; i < n_parsers; ++i) threads.emplace_back(parser_thread, &J);
        for (int l = 0; l < n_lanes; ++l) threads.emplace_back(lane_thread, &J, l);
        // the chunker runs here, on the calling thread
        double t_chunk = 0;
        for (;;) {
            std::vector<uint8_t> *buf = nullptr;
            std::vector<uint8_t> none;
            {
                // bound the look-ahead: a decode buffer must be free (compressed input), and at most the batch pool's
                // worth of chunks wait in the queue
                std::unique_lock<std::mutex> lk(J.mu);
                J.cv.wait(lk, [&] {
                    return J.abort || J.last_index != UINT64_MAX ||
                           (chunker.mapped() ? (int)J.chunks.size() < n_batches : !J.free_bufs.empty());
                });
                if (J.abort || J.last_index != UINT64_MAX) break;
                if (!chunker.mapped()) {
                    buf = J.free_bufs.back();
                    J.free_bufs.pop_back();
                }
            }
            Chunk c{0, nullptr, 0, false, buf};
            const double t0 = now_s();
            const bool more = chunker.next(chunk_bytes, buf ? *buf : none, c.p, c.n, c.drop_last);
            t_chunk += now_s() - t0;
            if (!more) break;
            std::lock_guard<std::mutex> lk(J.mu);
            c.index = J.n_chunks++;
            J.chunks.push_back(c);
            J.cv.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(J.mu);
            J.chunker_done = true;
            J.cv.notify_all();
        }
        for (std::thread &t : threads) t.join();
        cache->stats[HX_FS_CHUNK_S] = t_chunk;
    }
    if (rc == HX_OK) rc = J.rc;
    if (rc != HX_OK && !J.err.empty()) hxh_set_error(ctx, J.err.c_str());
    const double t_close0 = now_s();
    if (J.out[0].close_all() != 0 && rc == HX_OK) rc = HX_ERR_IO;
    if (J.out[1].close_all() != 0 && rc == HX_OK) rc = HX_ERR_IO;
    const double wall = now_s() - t_start;
    double *s = cache->stats;
    s[HX_FS_SETUP_S] = t_setup;
    s[HX_FS_CLOSE_S] = now_s() - t_close0;
    s[HX_FS_WALL_S] = wall;
    s[HX_FS_BASES] = (double)J.n_bases;
    s[HX_FS_RECORDS] = (double)J.n_records;
    s[HX_FS_BATCHES] = (double)J.n_batches;
    s[HX_FS_PARSE_S] = J.t_parse;
    s[HX_FS_DEVICE_S] = J.t_device;
    s[HX_FS_WRITE_S] = J.t_write;
    s[HX_FS_LOG_S] = J.t_log;
    s[HX_FS_FASTA_BYTES] = (double)J.file_pos[0];
    s[HX_FS_GFF_BYTES] = (double)J.file_pos[1];
    s[HX_FS_PARSERS] = n_parsers;
    s[HX_FS_LANES] = n_lanes;
    s[HX_FS_WRITERS] = n_writers;
    if (getenv("HYPEREX_TIMING"))  // where the wall time of the file-level path goes (stage seconds are summed over threads)
        fprintf(stderr, "[hyperex timing] wall=%.3fs bases=%llu records=%llu batches=%llu | chunk=%.3fs parse=%.3fs/%d thr "
                        "device=%.3fs/%d lanes write=%.3fs/%d thr log=%.3fs | setup=%.3fs unmap+close=%.3fs | fasta=%llu B gff=%llu B\n",
                wall, (unsigned long long)J.n_bases, (unsigned long long)J.n_records, (unsigned long long)J.n_batches,
                s[HX_FS_CHUNK_S], J.t_parse, n_parsers, J.t_device, n_lanes, J.t_write, n_writers, J.t_log, t_setup,
                s[HX_FS_CLOSE_S], (unsigned long long)J.file_pos[0], (unsigned long long)J.file_pos[1]);
    return rc;
}

extern "C" int hx_file_stats(hx_ctx *ctx, double *out, int n) {
    if (!ctx || !out || n < 0) return HX_ERR_ARG;
    HostCache *cache = host_cache_of(ctx);
    for (int i = 0; i < n; ++i) out[i] = i < HX_FILE_STATS_N ? cache->stats[i] : 0.0;
    return HX_OK;
}
