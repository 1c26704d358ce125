/*
 * hx_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see hx_oracle.h).
 *
 * Every function cites the reference lines it restates (Ebedthan/hyperex v0.2.0,
 * /root/reference/src/utils.rs) or the rust-bio 1.6.0 item whose published algorithm
 * it restates (the crate is not vendored in the reference tree).
 * PARITY STATUS: "parity unpinned" by reference tests; pinned by rust-bio doc vectors,
 * the independent DP below and a third-party fuzzy-regex engine (tests/test_oracle.py).
 */
#include "hx_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------
 * IUPAC ambiguity table handed to MyersBuilder::ambig — utils.rs:251-263.
 * A pattern symbol `code` matches itself and every byte of `equiv`; asymmetric.
 * ------------------------------------------------------------------------------- */
static const uint8_t HX_AMBIG_CODES[11] = {'M', 'R', 'W', 'S', 'Y', 'K', 'V', 'H', 'D', 'B', 'N'};
static const char *const HX_AMBIG_EQUIV[11] = {"AC",     "AG",     "AT",     "CG",     "CT",          "GT",
                                               "ACGMRS", "ACTMWY", "AGTRWK", "CGTSYK", "ACGTMRWSYKVHDB"};

/* rust-bio 1.6.0 MyersBuilder::build_64 (call sites utils.rs:301-302):
 * peq[a] |= 1<<i for the pattern byte itself and for each registered equivalent. */
int hxo_myers_build_custom(hxo_myers *my, const uint8_t *pattern, uint32_t m, uint32_t n_ambig,
                           const uint8_t *ambig_codes, const char *const *ambig_equiv) {
    if (m == 0 || m > 64) return -1; /* rust-bio asserts 0 < m <= 64 */
    memset(my->peq, 0, sizeof(my->peq));
    for (uint32_t i = 0; i < m; ++i) {
        uint8_t a = pattern[i];
        uint64_t bit = 1ull << i;
        my->peq[a] |= bit;
        for (uint32_t c = 0; c < n_ambig; ++c) {
            if (ambig_codes[c] != a) continue;
            for (const char *e = ambig_equiv[c]; *e; ++e) my->peq[(uint8_t)*e] |= bit;
        }
    }
    my->bound = 1ull << (m - 1);
    my->m = m;
    return 0;
}

int hxo_myers_build(hxo_myers *my, const uint8_t *pattern, uint32_t m, int use_hyperex_ambigs) {
    return hxo_myers_build_custom(my, pattern, m, use_hyperex_ambigs ? 11 : 0, HX_AMBIG_CODES, HX_AMBIG_EQUIV);
}

/* rust-bio 1.6.0 Myers<u64>::find_all_lazy (call sites utils.rs:305-309): State::init
 * (pv = low m bits, mv = 0, dist = m), then one step() per text byte; a position is
 * yielded iff dist <= k.  The lazy iterator's traceback bookkeeping is never queried by
 * hyperex and is omitted. */
size_t hxo_find_all_end(const hxo_myers *my, const uint8_t *text, size_t n, uint8_t k, uint64_t *ends,
                        uint8_t *dists, size_t cap) {
    uint64_t pv = (my->m == 64) ? ~0ull : ((1ull << my->m) - 1);
    uint64_t mv = 0;
    uint32_t dist = my->m;
    const uint64_t bound = my->bound;
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i) {
        uint64_t eq = my->peq[text[i]];
        uint64_t xv = eq | mv;
        uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
        uint64_t ph = mv | ~(xh | pv);
        uint64_t mh = pv & xh;
        if (ph & bound)
            dist += 1;
        else if (mh & bound)
            dist -= 1;
        ph <<= 1;
        mh <<= 1;
        pv = mh | ~(xv | ph);
        mv = ph & xv;
        if (dist <= k) {
            if (cnt < cap) {
                ends[cnt] = i;
                dists[cnt] = (uint8_t)dist;
            }
            ++cnt;
        }
    }
    return cnt;
}

/* match predicate implied by utils.rs:251-263 + build_64: pattern symbol p matches text
 * byte b iff b == p or b is listed for p. */
static int hxo_sym_match(uint8_t p, uint8_t b) {
    if (p == b) return 1;
    for (int c = 0; c < 11; ++c)
        if (HX_AMBIG_CODES[c] == p) return strchr(HX_AMBIG_EQUIV[c], (int)b) != NULL && b != 0;
    return 0;
}

/* Independent cross-check: D[i][0]=i, D[0][j]=0,
 * D[i][j]=min(D[i-1][j-1]+!match, D[i-1][j]+1, D[i][j-1]+1); hit at j-1 iff D[m][j]<=k. */
size_t hxo_dp_find_all_end(const uint8_t *pattern, uint32_t m, const uint8_t *text, size_t n, uint8_t k,
                           uint64_t *ends, uint8_t *dists, size_t cap) {
    uint32_t *col = (uint32_t *)malloc((m + 1) * sizeof(uint32_t));
    for (uint32_t i = 0; i <= m; ++i) col[i] = i;
    size_t cnt = 0;
    for (size_t j = 0; j < n; ++j) {
        uint32_t diag = col[0]; /* D[0][j] = 0 */
        col[0] = 0;
        for (uint32_t i = 1; i <= m; ++i) {
            uint32_t up = col[i - 1] + 1;       /* D[i-1][j+1] + 1 */
            uint32_t left = col[i] + 1;         /* D[i][j] + 1     */
            uint32_t dg = diag + (hxo_sym_match(pattern[i - 1], text[j]) ? 0u : 1u);
            diag = col[i];
            uint32_t v = up < left ? up : left;
            col[i] = dg < v ? dg : v;
        }
        if (col[m] <= k) {
            if (cnt < cap) {
                ends[cnt] = j;
                dists[cnt] = (uint8_t)col[m];
            }
            ++cnt;
        }
    }
    free(col);
    return cnt;
}

static inline uint8_t hxo_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

/* utils.rs:207-228 sequence_type */
int hxo_sequence_type(const uint8_t *seq, size_t n) {
    int has_u = 0, has_t = 0, all_iupac = 1;
    for (size_t i = 0; i < n; ++i) {
        uint8_t c = hxo_upper(seq[i]);
        if (c == 'U') has_u = 1;
        if (c == 'T') has_t = 1;
        if (c == 0 || !strchr("ACGTRYSWKMBDHVN", (int)c)) all_iupac = 0;
    }
    if (has_u && !has_t) return HXO_RNA;
    if (has_t && !has_u) return HXO_DNA;
    if (!has_u && !has_t) return all_iupac ? HXO_DNA : HXO_NONE;
    return HXO_NONE;
}

/* utils.rs:169-194 to_complement */
void hxo_to_complement(const uint8_t *primer, size_t n, int alphabet, uint8_t *out) {
    for (size_t i = 0; i < n; ++i) {
        uint8_t c = primer[i], o = c;
        if (alphabet == HXO_DNA) {
            if (c == 'A') o = 'T';
            else if (c == 'T') o = 'A';
            else if (c == 'C') o = 'G';
            else if (c == 'G') o = 'C';
        } else if (alphabet == HXO_RNA) {
            if (c == 'A') o = 'U';
            else if (c == 'U') o = 'A';
            else if (c == 'C') o = 'G';
            else if (c == 'G') o = 'C';
        }
        if (o == c) {
            switch (c) {
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
        }
        out[i] = o;
    }
}

/* utils.rs:197-199 to_reverse_complement */
void hxo_to_reverse_complement(const uint8_t *primer, size_t n, int alphabet, uint8_t *out) {
    uint8_t *tmp = (uint8_t *)malloc(n ? n : 1);
    hxo_to_complement(primer, n, alphabet, tmp);
    for (size_t i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
    free(tmp);
}

/* utils.rs:29-45 PRIMER_TO_REGION */
static const char *const HX_P2R[15][2] = {
    {"AGAGTTTGATCMTGGCTCAG", "v1"},     {"CCTACGGGNGGCWGCAG", "v3"},      {"GTGCCAGCMGCCGCGGTAA", "v4"},
    {"GTGYCAGCMGCCGCGGTAA", "v4"},      {"AACMGGATTAGATACCCKG", "v5"},    {"TAAAACTYAAAKGAATTGACGGGG", "v6"},
    {"YAACGAGCGCAACCC", "v7"},          {"ACTGCTGCSYCCCGTAGGAGTCT", "v2"}, {"ATTACCGCGGCTGCTGG", "v3"},
    {"GACTACHVGGGTATCTAATCC", "v4"},    {"CCGTCAATTYMTTTRAGT", "v5"},     {"GGACTACHVGGGTWTCTAAT", "v4"},
    {"CCCCGYCAATTCMTTTRAGT", "v5"},     {"ACGTCATCCCCACCTTCC", "v7"},     {"TACGGYTACCTTGTTAYGACTT", "v9"}};

static const char *hxo_p2r(const char *p) {
    for (int i = 0; i < 15; ++i)
        if (strcmp(HX_P2R[i][0], p) == 0) return HX_P2R[i][1];
    return "";
}

/* utils.rs:157-166 primers_to_region */
void hxo_primers_to_region(const char *fwd, const char *rev, char *out8) {
    const char *a = hxo_p2r(fwd), *b = hxo_p2r(rev);
    if (strcmp(a, "v4") == 0 && strcmp(b, "v4") == 0)
        snprintf(out8, 8, "%s", a);
    else
        snprintf(out8, 8, "%s%s", a, b);
}

/* utils.rs:23-25 */
static const char *const HX_REGIONS[10] = {"v1v2", "v1v3", "v1v9", "v3v4", "v3v5",
                                           "v4",   "v4v5", "v5v7", "v6v9", "v7v9"};
const char *const *hxo_supported_regions(uint32_t *n) {
    *n = 10;
    return HX_REGIONS;
}

/* utils.rs:47-66 + 112-131 region_to_primer */
int hxo_region_to_primer(const char *region, const char **fwd, const char **rev) {
    static const char *F27 = "AGAGTTTGATCMTGGCTCAG", *F341 = "CCTACGGGNGGCWGCAG", *F515 = "GTGCCAGCMGCCGCGGTAA",
                      *F515Y = "GTGYCAGCMGCCGCGGTAA", *F799 = "AACMGGATTAGATACCCKG",
                      *F928 = "TAAAACTYAAAKGAATTGACGGGG", *F1100 = "YAACGAGCGCAACCC";
    static const char *R336 = "ACTGCTGCSYCCCGTAGGAGTCT", *R534 = "ATTACCGCGGCTGCTGG", *R805 = "GACTACHVGGGTATCTAATCC",
                      *R926b = "CCGTCAATTYMTTTRAGT", *R806 = "GGACTACHVGGGTWTCTAAT",
                      *R909 = "CCCCGYCAATTCMTTTRAGT", *R1193 = "ACGTCATCCCCACCTTCC",
                      *R1492 = "TACGGYTACCTTGTTAYGACTT";
    struct { const char *name, *f, *r; } tab[10] = {
        {"v1v2", F27, R336},   {"v1v3", F27, R534},    {"v1v9", F27, R1492},  {"v3v4", F341, R805},
        {"v3v5", F341, R926b}, {"v4", F515, R806},     {"v4v5", F515Y, R909}, {"v5v7", F799, R1193},
        {"v6v9", F928, R1492}, {"v7v9", F1100, R1492}};
    for (int i = 0; i < 10; ++i)
        if (strcmp(tab[i].name, region) == 0) {
            *fwd = tab[i].f;
            *rev = tab[i].r;
            return 0;
        }
    return -1;
}

/* utils.rs:327-351, literally */
int hxo_pair_literal(const uint64_t *f_end, const uint8_t *f_dist, size_t nf, const uint64_t *r_end,
                     const uint8_t *r_dist, size_t nr, size_t forward_len, uint64_t *f_start_out,
                     uint64_t *r_end_out, uint8_t *combined_out) {
    int have = 0;
    uint64_t bfs = 0, bre = 0;
    uint8_t bd = 0;
    size_t sub = forward_len ? forward_len - 1 : 0; /* saturating_sub(1) */
    for (size_t a = 0; a < nf; ++a) {
        if (f_end[a] < sub) continue; /* checked_sub -> None */
        uint64_t f_start = f_end[a] - sub;
        for (size_t b = 0; b < nr; ++b) {
            if (r_end[b] <= f_end[a]) continue;
            unsigned c = (unsigned)f_dist[a] + (unsigned)r_dist[b];
            uint8_t combined = c > 255 ? 255 : (uint8_t)c; /* saturating_add */
            if (!have || combined < bd) {
                have = 1;
                bfs = f_start;
                bre = r_end[b];
                bd = combined;
            }
        }
    }
    if (have) {
        *f_start_out = bfs;
        *r_end_out = bre;
        *combined_out = bd;
    }
    return have;
}

/* --------------------------------------------------------------------------------- */
typedef struct {
    uint64_t *ends;
    uint8_t *dists;
    size_t cap;
} hxo_hits;

static size_t hxo_collect(const hxo_myers *my, const uint8_t *text, size_t n, uint8_t k, hxo_hits *h) {
    size_t c = hxo_find_all_end(my, text, n, k, h->ends, h->dists, h->cap);
    if (c > h->cap) {
        h->cap = c + c / 2 + 16;
        h->ends = (uint64_t *)realloc(h->ends, h->cap * sizeof(uint64_t));
        h->dists = (uint8_t *)realloc(h->dists, h->cap);
        c = hxo_find_all_end(my, text, n, k, h->ends, h->dists, h->cap);
    }
    return c;
}

typedef struct {
    const uint8_t *arena;
    const uint64_t *offsets;
    uint32_t r0, r1, n_pairs;
    const char *const *fwd;
    const char *const *rev;
    uint8_t k;
    hxo_result *out;
    int rebuild;
    int rc;
} hxo_job;

/* utils.rs:270-351 for records [r0, r1) */
static void *hxo_worker(void *arg) {
    hxo_job *j = (hxo_job *)arg;
    uint32_t np = j->n_pairs;
    hxo_myers *fm = (hxo_myers *)malloc(sizeof(hxo_myers) * np);
    hxo_myers *rm = (hxo_myers *)malloc(sizeof(hxo_myers) * np * 2); /* [pair][alphabet] */
    uint8_t rc[64];
    j->rc = 0;
    for (uint32_t p = 0; p < np; ++p) {
        size_t lf = strlen(j->fwd[p]), lr = strlen(j->rev[p]);
        if (lf == 0 || lf > 64 || lr == 0 || lr > 64) {
            j->rc = -1;
            free(fm);
            free(rm);
            return NULL;
        }
        hxo_myers_build(&fm[p], (const uint8_t *)j->fwd[p], (uint32_t)lf, 1);
        for (int a = 0; a < 2; ++a) {
            hxo_to_reverse_complement((const uint8_t *)j->rev[p], lr, a, rc);
            hxo_myers_build(&rm[2 * p + a], rc, (uint32_t)lr, 1);
        }
    }
    hxo_hits fh = {NULL, NULL, 0}, rh = {NULL, NULL, 0};
    uint8_t *upper = NULL;
    size_t upper_cap = 0;
    for (uint32_t r = j->r0; r < j->r1; ++r) {
        const uint8_t *seq = j->arena + j->offsets[r];
        size_t n = (size_t)(j->offsets[r + 1] - j->offsets[r]);
        hxo_result *res = j->out + (size_t)r * np;
        int alpha = hxo_sequence_type(seq, n); /* utils.rs:273-280 */
        if (alpha == HXO_NONE) {
            for (uint32_t p = 0; p < np; ++p) {
                memset(&res[p], 0, sizeof(hxo_result));
                res[p].flags = 8;
            }
            continue;
        }
        if (n > upper_cap) {
            upper_cap = n + n / 2 + 64;
            upper = (uint8_t *)realloc(upper, upper_cap);
        }
        for (size_t i = 0; i < n; ++i) upper[i] = hxo_upper(seq[i]); /* utils.rs:295 */
        for (uint32_t p = 0; p < np; ++p) {
            hxo_myers fl, rl;
            const hxo_myers *F = &fm[p], *R = &rm[2 * p + alpha];
            size_t lf = strlen(j->fwd[p]);
            if (j->rebuild) { /* utils.rs:299-302 rebuilds per record x pair */
                size_t lr = strlen(j->rev[p]);
                hxo_to_reverse_complement((const uint8_t *)j->rev[p], lr, alpha, rc);
                hxo_myers_build(&fl, (const uint8_t *)j->fwd[p], (uint32_t)lf, 1);
                hxo_myers_build(&rl, rc, (uint32_t)lr, 1);
                F = &fl;
                R = &rl;
            }
            size_t nf = hxo_collect(F, upper, n, j->k, &fh); /* utils.rs:305-309 */
            size_t nr = hxo_collect(R, upper, n, j->k, &rh);
            uint64_t fs = 0, re = 0;
            uint8_t comb = 0;
            int ok = hxo_pair_literal(fh.ends, fh.dists, nf, rh.ends, rh.dists, nr, lf, &fs, &re, &comb);
            memset(&res[p], 0, sizeof(hxo_result));
            res[p].flags = (uint8_t)((nf ? 1 : 0) | (nr ? 2 : 0) | (ok ? 4 : 0) | (alpha == HXO_RNA ? 16 : 0));
            if (ok) {
                res[p].f_start = (uint32_t)fs;
                res[p].r_end = (uint32_t)re;
                res[p].combined = comb;
            }
        }
    }
    free(fh.ends);
    free(fh.dists);
    free(rh.ends);
    free(rh.dists);
    free(upper);
    free(fm);
    free(rm);
    return NULL;
}

int hxo_run_batch(const uint8_t *arena, const uint64_t *offsets, uint32_t n_records, uint32_t n_pairs,
                  const char *const *fwd, const char *const *rev, uint8_t k, hxo_result *out, int nthreads,
                  int rebuild_per_record) {
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n_records && n_records) nthreads = (int)n_records;
    if (n_records == 0) return 0;
    hxo_job *jobs = (hxo_job *)calloc((size_t)nthreads, sizeof(hxo_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    /* split by bases so threads get equal work */
    uint64_t total = offsets[n_records] - offsets[0];
    uint32_t r = 0;
    for (int t = 0; t < nthreads; ++t) {
        uint64_t target = offsets[0] + total * (uint64_t)(t + 1) / (uint64_t)nthreads;
        uint32_t r1 = r;
        while (r1 < n_records && offsets[r1 + 1] <= target) ++r1;
        if (t == nthreads - 1) r1 = n_records;
        jobs[t] = (hxo_job){arena, offsets, r, r1, n_pairs, fwd, rev, k, out, rebuild_per_record, 0};
        r = r1;
    }
    int rc = 0;
    if (nthreads == 1) {
        hxo_worker(&jobs[0]);
        rc = jobs[0].rc;
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, hxo_worker, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) {
            pthread_join(th[t], NULL);
            if (jobs[t].rc) rc = jobs[t].rc;
        }
    }
    free(jobs);
    free(th);
    return rc;
}

/* --------------------------------------------------------------------------------- */
typedef struct {
    uint8_t *p;
    size_t len, cap;
} hxo_buf;

static void hxo_put(hxo_buf *b, const void *src, size_t n) {
    if (b->len + n > b->cap) {
        b->cap = (b->len + n) * 2 + 4096;
        b->p = (uint8_t *)realloc(b->p, b->cap);
    }
    memcpy(b->p + b->len, src, n);
    b->len += n;
}
static void hxo_puts(hxo_buf *b, const char *s) { hxo_put(b, s, strlen(s)); }

/* utils.rs:353-385 + rust-bio fasta::Writer::write_record (">id desc\nseq\n", no wrapping) */
int hxo_format_outputs(const uint8_t *arena, const uint64_t *offsets, const char *const *ids, uint32_t n_records,
                       uint32_t n_pairs, const char *const *fwd, const char *const *rev, const hxo_result *res,
                       uint8_t **fa, size_t *fa_len, uint8_t **gff, size_t *gff_len) {
    hxo_buf f = {NULL, 0, 0}, g = {NULL, 0, 0};
    char num[64], label[8];
    hxo_puts(&g, "##gff-version 3\n"); /* utils.rs:247 */
    for (uint32_t r = 0; r < n_records; ++r) {
        const uint8_t *seq = arena + offsets[r];
        for (uint32_t p = 0; p < n_pairs; ++p) {
            const hxo_result *x = &res[(size_t)r * n_pairs + p];
            if ((x->flags & 7) != 7) continue;
            hxo_primers_to_region(fwd[p], rev[p], label);
            hxo_puts(&f, ">");
            hxo_puts(&f, ids[r]);
            hxo_puts(&f, " ");
            if (label[0]) {
                hxo_puts(&f, "region=");
                hxo_puts(&f, label);
                hxo_puts(&f, " ");
            }
            hxo_puts(&f, "forward=");
            hxo_puts(&f, fwd[p]);
            hxo_puts(&f, " reverse=");
            hxo_puts(&f, rev[p]);
            hxo_puts(&f, "\n");
            hxo_put(&f, seq + x->f_start, (size_t)x->r_end + 1 - x->f_start);
            hxo_puts(&f, "\n");
            hxo_puts(&g, ids[r]);
            snprintf(num, sizeof num, "\thyperex\tregion\t%u\t%u\t.\t.\t.\tNote=", x->f_start + 1, x->r_end + 1);
            hxo_puts(&g, num);
            if (label[0]) {
                hxo_puts(&g, "Hypervariable region ");
                hxo_puts(&g, label);
            } else {
                hxo_puts(&g, "forward=");
                hxo_puts(&g, fwd[p]);
                hxo_puts(&g, " reverse=");
                hxo_puts(&g, rev[p]);
            }
            hxo_puts(&g, "\n");
        }
    }
    *fa = f.p ? f.p : (uint8_t *)malloc(1);
    *fa_len = f.len;
    *gff = g.p;
    *gff_len = g.len;
    return 0;
}

void hxo_free(void *p) { free(p); }

/* --------------------------------------------------------------------------------- */
static int hxo_is_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

/* rust-bio 1.6.0 fasta::Reader::read + Records iterator (call site utils.rs:238,270):
 * a record starts with '>' (else error => iteration ends); id = header up to the first
 * whitespace of line[1..].trim_end(); sequence lines appended with trim_end(); an empty
 * record (empty id, no desc, empty seq) ends iteration. */
int64_t hxo_parse_fasta(const uint8_t *buf, size_t len, uint8_t **arena_out, uint64_t **offsets_out, char **ids_out,
                        uint64_t **id_off_out) {
    hxo_buf arena = {NULL, 0, 0}, ids = {NULL, 0, 0}, offs = {NULL, 0, 0}, idoffs = {NULL, 0, 0};
    uint64_t zero = 0;
    hxo_put(&offs, &zero, 8);
    int64_t nrec = 0;
    size_t pos = 0;
    while (pos < len) {
        /* header line */
        size_t e = pos;
        while (e < len && buf[e] != '\n') ++e;
        size_t line_end = e; /* excl. '\n' */
        if (buf[pos] != '>') break; /* "Expected > at record start." */
        size_t hs = pos + 1, he = line_end;
        while (he > hs && hxo_is_ws(buf[he - 1])) --he;
        size_t ie = hs;
        while (ie < he && !hxo_is_ws(buf[ie])) ++ie;
        int has_desc = ie < he;
        pos = (e < len) ? e + 1 : len;
        size_t seq_start = arena.len;
        while (pos < len && buf[pos] != '>') {
            size_t le = pos;
            while (le < len && buf[le] != '\n') ++le;
            size_t te = le;
            while (te > pos && hxo_is_ws(buf[te - 1])) --te;
            hxo_put(&arena, buf + pos, te - pos);
            pos = (le < len) ? le + 1 : len;
        }
        if (ie == hs && !has_desc && arena.len == seq_start) break; /* record.is_empty() */
        uint64_t io = ids.len;
        hxo_put(&idoffs, &io, 8);
        hxo_put(&ids, buf + hs, ie - hs);
        hxo_put(&ids, "", 1);
        uint64_t ao = arena.len;
        hxo_put(&offs, &ao, 8);
        ++nrec;
    }
    if (!arena.p) arena.p = (uint8_t *)malloc(16);
    if (!ids.p) ids.p = (uint8_t *)malloc(16);
    if (!idoffs.p) idoffs.p = (uint8_t *)malloc(16);
    *arena_out = arena.p;
    *offsets_out = (uint64_t *)offs.p;
    *ids_out = (char *)ids.p;
    *id_off_out = (uint64_t *)idoffs.p;
    return nrec;
}
