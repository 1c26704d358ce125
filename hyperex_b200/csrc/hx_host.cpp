// hx_host.cpp — host services of the product around the device path: the reference's
// primer/region tables and string helpers (restated 1:1 because they shape the output
// bytes), FASTA ingest with compression sniffing, the FASTA/GFF3 writers and the file-level
// driver that replaces utils::get_hypervar_regions (src/utils.rs:231-414).
#include "hx_host.h"

#include "../../include/hyperex_b200.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <zlib.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
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
// FASTA reader: rust-bio 1.6.0 fasta::Reader::read + Records
// ------------------------------------------------------------------------------------
void HxhRecordBatch::clear() {
    offsets.assign(1, 0);
    ids.clear();
}

HxhRecordBatch::~HxhRecordBatch() {
    if (arena) {
        if (pinned) hx_free_pinned(arena);
        else free(arena);
    }
}

namespace {
bool grow_arena(HxhRecordBatch &b, size_t need, size_t used) {
    if (need <= b.arena_cap) return true;
    size_t want = std::max(need + need / 2, (size_t)1 << 20);
    uint8_t *p = (uint8_t *)hx_alloc_pinned(want);
    bool pinned = p != nullptr;
    if (!p) p = (uint8_t *)malloc(want);
    if (!p) return false;
    if (b.arena) {
        memcpy(p, b.arena, used);
        if (b.pinned) hx_free_pinned(b.arena);
        else free(b.arena);
    }
    b.arena = p;
    b.arena_cap = want;
    b.pinned = pinned;
    return true;
}
inline bool is_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
}  // namespace

HxhFastaReader::HxhFastaReader() : buf_(8u << 20) {}
HxhFastaReader::~HxhFastaReader() { delete src_; }

int HxhFastaReader::open(const char *path, std::string &err) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        err = std::string("Failed to open file: ") + path;
        return HX_ERR_IO;
    }
    uint8_t magic[6] = {0, 0, 0, 0, 0, 0};
    size_t n = fread(magic, 1, 5, f);
    if (n < 5) {  // niffler: FileTooShort
        fclose(f);
        err = std::string("Failed to determine compression for: ") + path;
        return HX_ERR_IO;
    }
    rewind(f);
    {
        struct stat st;
        size_hint_ = (fstat(fileno(f), &st) == 0 && st.st_size > 0) ? (size_t)st.st_size : 0;
        const bool plain = !((magic[0] == 0x1f && magic[1] == 0x8b) || (magic[0] == 0x42 && magic[1] == 0x5a) ||
                             (magic[0] == 0xfd && magic[1] == 0x37) || (magic[0] == 0x28 && magic[1] == 0xb5));
        if (!plain) size_hint_ *= 5;  // typical nucleotide compression ratio; grown on demand beyond that
    }
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
        src_ = new PlainSource(f);
    }
    if (!ok) {
        err = std::string("Failed to determine compression for: ") + path + " (decoder library unavailable)";
        return HX_ERR_IO;
    }
    return HX_OK;
}

// Makes the next complete line available inside buf_ (no copy): [line, line + len) excludes the '\n'.
// The buffer is compacted / grown / refilled as needed; the last line of the input may lack the '\n'.
bool HxhFastaReader::next_line(const uint8_t *&line, size_t &len) {
    for (;;) {
        const uint8_t *s0 = buf_.data() + pos_;
        const uint8_t *nl = pos_ < end_ ? (const uint8_t *)memchr(s0, '\n', end_ - pos_) : nullptr;
        if (nl) {
            line = s0;
            len = (size_t)(nl - s0);
            pos_ += len + 1;
            return true;
        }
        if (eof_) {
            if (pos_ == end_) return false;
            line = s0;
            len = end_ - pos_;
            pos_ = end_;
            return true;
        }
        // need more bytes: move the partial line to the front, grow if it fills the buffer, refill
        if (pos_ > 0) {
            memmove(buf_.data(), buf_.data() + pos_, end_ - pos_);
            end_ -= pos_;
            pos_ = 0;
        }
        if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);
        long r = src_->read(buf_.data() + end_, buf_.size() - end_);
        if (r <= 0) eof_ = true;
        else end_ += (size_t)r;
    }
}

size_t HxhFastaReader::next_batch(HxhRecordBatch &b, size_t max_bytes, size_t max_records) {
    b.clear();
    if (finished_ || !src_) return 0;
    size_t used = 0;
    if (!b.arena) {
        // one allocation for the common case: a plain file cannot hold more sequence than its size
        size_t est = size_hint_ ? std::min(max_bytes, size_hint_) : (size_t)1 << 20;
        if (!grow_arena(b, est + 4096, 0)) { finished_ = true; return 0; }
    }
    const uint8_t *line = nullptr;
    size_t len = 0;
    while (b.ids.size() < max_records && used < max_bytes) {
        if (!have_line_) {
            if (!next_line(line, len)) { finished_ = true; break; }  // clean EOF
            header_.assign((const char *)line, len);
            have_line_ = true;
        }
        if (header_.empty() || header_[0] != '>') { finished_ = true; break; }  // "Expected > at record start."
        // header: line[1..].trim_end(), id = up to the first whitespace
        size_t he = header_.size();
        while (he > 1 && is_ws((uint8_t)header_[he - 1])) --he;
        size_t ie = 1;
        while (ie < he && !is_ws((uint8_t)header_[ie])) ++ie;
        const bool has_desc = ie < he;
        std::string id = header_.substr(1, ie - 1);
        const size_t seq_start = used;
        have_line_ = false;
        for (;;) {
            if (!next_line(line, len)) break;  // EOF ends the record
            if (len && line[0] == '>') {
                header_.assign((const char *)line, len);
                have_line_ = true;
                break;
            }
            size_t te = len;
            while (te > 0 && is_ws(line[te - 1])) --te;  // trim_end()
            if (used + te + 64 > b.arena_cap && !grow_arena(b, used + te + 64, used)) {
                finished_ = true;
                return b.ids.size();
            }
            memcpy(b.arena + used, line, te);
            used += te;
        }
        if (id.empty() && !has_desc && used == seq_start) { finished_ = true; break; }  // record.is_empty()
        b.ids.push_back(std::move(id));
        b.offsets.push_back(used);
    }
    return b.ids.size();
}

struct hx_fasta {
    HxhFastaReader reader;
    HxhRecordBatch batch;
    std::vector<const char *> id_ptrs;
};

extern "C" int hx_fasta_open(const char *path, hx_fasta **out) {
    if (!path || !out) return HX_ERR_ARG;
    *out = nullptr;
    hx_fasta *r = new (std::nothrow) hx_fasta();
    if (!r) return HX_ERR_NOMEM;
    std::string err;
    int rc = r->reader.open(path, err);
    if (rc != HX_OK) {
        delete r;
        return rc;
    }
    *out = r;
    return HX_OK;
}

extern "C" int64_t hx_fasta_next(hx_fasta *r, size_t max_bytes, const uint8_t **arena, const uint64_t **offsets,
                                 const char *const **ids) {
    if (!r) return HX_ERR_ARG;
    size_t n = r->reader.next_batch(r->batch, max_bytes ? max_bytes : ((size_t)256 << 20), (size_t)0xFFFFFF00u);
    r->id_ptrs.clear();
    for (const std::string &s : r->batch.ids) r->id_ptrs.push_back(s.c_str());
    if (arena) *arena = r->batch.arena;
    if (offsets) *offsets = r->batch.offsets.data();
    if (ids) *ids = r->id_ptrs.data();
    return (int64_t)n;
}

extern "C" void hx_fasta_close(hx_fasta *r) { delete r; }

// ------------------------------------------------------------------------------------
// utils.rs:231-414 get_hypervar_regions
//   parser thread: FASTA -> pinned arenas (batches of ~64 MB of sequence, three in flight)
//   caller thread: hx_run_batch on the GPU(s), then FASTA / GFF3 formatting into large write buffers
// ------------------------------------------------------------------------------------
namespace {

struct OutBuf {  // append-only write buffer in front of a FILE*
    std::vector<char> v;
    size_t n = 0;
    FILE *f = nullptr;
    bool ok = true;
    explicit OutBuf(FILE *file) : v((size_t)16 << 20), f(file) {}
    void flush() {
        if (n && fwrite(v.data(), 1, n, f) != n) ok = false;
        n = 0;
    }
    char *reserve(size_t k) {
        if (n + k > v.size()) {
            flush();
            if (k > v.size()) v.resize(k + (k >> 1));
        }
        return v.data() + n;
    }
    void put(const void *p, size_t k) {
        memcpy(reserve(k), p, k);
        n += k;
    }
    void put(const std::string &x) { put(x.data(), x.size()); }
};

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

struct TextSink {  // hx_text_sink of the file-level path: kind 0 -> FASTA file, kind 1 -> GFF3 file
    FILE *fa, *gff;
    double seconds;
    static int write(void *user, int kind, const char *bytes, size_t n) {
        TextSink *t = (TextSink *)user;
        const auto t0 = std::chrono::steady_clock::now();
        const bool ok = fwrite(bytes, 1, n, kind ? t->gff : t->fa) == n;
        t->seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return ok ? 0 : 1;
    }
};

struct BatchQueue {  // parser -> consumer hand-off with recycling
    std::mutex mu;
    std::condition_variable cv;
    std::vector<HxhRecordBatch *> ready, free_list;
    bool done = false;
};

}  // namespace

extern "C" int hx_get_hypervar_regions(hx_ctx *ctx, const char *file, uint32_t n_pairs, const char *const *fwd,
                                       const char *const *rev, const char *prefix, uint8_t mismatch, hx_log_fn log_fn,
                                       void *user) {
    if (!ctx || !file || !fwd || !rev || !prefix || n_pairs == 0) return HX_ERR_ARG;
    int rc = hx_set_primers(ctx, n_pairs, fwd, nullptr, rev, nullptr, mismatch);
    if (rc != HX_OK) return rc;
    HxhFastaReader reader;
    std::string err;
    rc = reader.open(file, err);
    if (rc != HX_OK) {
        if (log_fn) log_fn(2, err.c_str(), user);
        return rc;
    }
    const std::string fa_path = std::string(prefix) + ".fa", gff_path = std::string(prefix) + ".gff";
    FILE *fa = fopen(fa_path.c_str(), "wb");   // fasta::Writer::to_file truncates (utils.rs:241)
    FILE *gff = fopen(gff_path.c_str(), "ab"); // OpenOptions create+append (utils.rs:242-245)
    if (!fa || !gff) {
        if (fa) fclose(fa);
        if (gff) fclose(gff);
        return HX_ERR_IO;
    }
    OutBuf fo(fa), go(gff);
    go.put("##gff-version 3\n", 16);  // utils.rs:247
    // per pair: everything of the FASTA header / GFF line that does not depend on the record
    std::vector<std::string> labels(n_pairs), fa_tail(n_pairs), gff_note(n_pairs);
    for (uint32_t p = 0; p < n_pairs; ++p) {
        char lab[16];
        hx_primers_to_region(fwd[p], rev[p], lab);
        labels[p] = lab;
        const std::string fr = std::string("forward=") + fwd[p] + " reverse=" + rev[p];
        fa_tail[p] = labels[p].empty() ? " " + fr + "\n" : " region=" + labels[p] + " " + fr + "\n";      // utils.rs:356-363
        gff_note[p] = "\t.\t.\t.\tNote=" + (labels[p].empty() ? fr : "Hypervariable region " + labels[p]) + "\n";  // 372-385
    }
    // On-device extraction: with one GPU the FASTA text itself is assembled on the device and the
    // host only write()s it; HYPEREX_HOST_FORMAT=1 (or a multi-GPU context) keeps the host formatter.
    const bool device_fasta = hxh_device_count(ctx) == 1 && getenv("HYPEREX_HOST_FORMAT") == nullptr;
    std::string tails_blob, gtails_blob, ids_blob;
    std::vector<uint32_t> tail_off(1, 0), gtail_off(1, 0);
    std::vector<uint64_t> id_off;
    for (uint32_t p = 0; p < n_pairs; ++p) {
        tails_blob += fa_tail[p];
        tail_off.push_back((uint32_t)tails_blob.size());
        gtails_blob += gff_note[p];
        gtail_off.push_back((uint32_t)gtails_blob.size());
    }
    // HYPEREX_TIMING=1 prints where the wall time of the file-level path goes
    const bool timing = getenv("HYPEREX_TIMING") != nullptr;
    double t_parse = 0, t_gpu = 0, t_write = 0, t_wait = 0;
    uint64_t n_bases = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };

    HxhRecordBatch pool[3];
    BatchQueue q;
    for (HxhRecordBatch &b : pool) q.free_list.push_back(&b);
    const size_t batch_bytes = (size_t)64 << 20;
    std::thread parser([&] {
        for (;;) {
            HxhRecordBatch *b = nullptr;
            {
                std::unique_lock<std::mutex> lk(q.mu);
                q.cv.wait(lk, [&] { return !q.free_list.empty() || q.done; });
                if (q.done) return;
                b = q.free_list.back();
                q.free_list.pop_back();
            }
            const double t0 = now();
            const size_t n = reader.next_batch(*b, batch_bytes, (size_t)8 << 20);
            t_parse += now() - t0;
            std::lock_guard<std::mutex> lk(q.mu);
            if (n == 0) {
                q.done = true;
                q.cv.notify_all();
                return;
            }
            q.ready.push_back(b);
            q.cv.notify_all();
        }
    });

    std::vector<hx_result> res;
    std::string msg;
    char num[96];
    for (;;) {
        HxhRecordBatch *bp = nullptr;
        {
            const double t0 = now();
            std::unique_lock<std::mutex> lk(q.mu);
            q.cv.wait(lk, [&] { return !q.ready.empty() || q.done; });
            t_wait += now() - t0;
            if (q.ready.empty()) break;
            bp = q.ready.front();
            q.ready.erase(q.ready.begin());
        }
        HxhRecordBatch &batch = *bp;
        const double t1 = now();
        const uint32_t n = (uint32_t)batch.ids.size();
        n_bases += batch.offsets[n];
        res.resize((size_t)n * n_pairs);
        if (device_fasta) {
            ids_blob.clear();
            id_off.assign(1, 0);
            for (const std::string &id : batch.ids) {
                ids_blob += id;
                id_off.push_back(ids_blob.size());
            }
            // the GPU produces the FASTA and GFF3 bytes of this batch; they are written as they arrive
            fo.flush();
            go.flush();
            TextSink sink{fa, gff, 0.0};
            rc = hx_run_batch_text_stream(ctx, batch.arena, batch.offsets.data(), n, res.data(), ids_blob.data(),
                                          id_off.data(), tails_blob.data(), tail_off.data(), gtails_blob.data(),
                                          gtail_off.data(), &TextSink::write, &sink, nullptr, nullptr);
            t_write += sink.seconds;
            t_gpu -= sink.seconds;
        } else {
            rc = hx_run_batch(ctx, batch.arena, batch.offsets.data(), n, res.data());
        }
        const double t2 = now();
        t_gpu += t2 - t1;
        if (rc == HX_OK) {
            for (uint32_t r = 0; r < n; ++r) {
                const std::string &id = batch.ids[r];
                const uint8_t *seq = batch.arena + batch.offsets[r];
                const size_t len = (size_t)(batch.offsets[r + 1] - batch.offsets[r]);
                const hx_result *rr = &res[(size_t)r * n_pairs];
                if (rr[0].flags & HX_ALPHABET_NONE) {  // utils.rs:276-279
                    if (log_fn) {
                        msg = "Unrecognized sequence type in record: " + id;
                        log_fn(2, msg.c_str(), user);
                    }
                    continue;
                }
                if (len <= 1500 && log_fn) {  // utils.rs:282-288
                    snprintf(num, sizeof num, "%zu", len);
                    msg = "Sequence " + id + " is short (" + num + " bp). Some regions may not be found.";
                    log_fn(1, msg.c_str(), user);
                }
                for (uint32_t p = 0; p < n_pairs; ++p) {
                    const hx_result &x = rr[p];
                    const bool f_any = x.flags & HX_FWD_ANY, r_any = x.flags & HX_REV_ANY, ok = x.flags & HX_PAIR_VALID;
                    if (f_any && r_any && ok) {  // utils.rs:354-386
                        if (!device_fasta) {
                            const size_t slen = (size_t)x.r_end + 1 - x.f_start;
                            char *o = fo.reserve(1 + id.size() + fa_tail[p].size() + slen + 1);
                            *o++ = '>';
                            memcpy(o, id.data(), id.size());
                            o += id.size();
                            memcpy(o, fa_tail[p].data(), fa_tail[p].size());
                            o += fa_tail[p].size();
                            memcpy(o, seq + x.f_start, slen);  // original case (utils.rs:368)
                            o += slen;
                            *o++ = '\n';
                            fo.n = (size_t)(o - fo.v.data());
                        }
                        if (!device_fasta) {
                            char *g = go.reserve(id.size() + 48 + gff_note[p].size());
                            memcpy(g, id.data(), id.size());
                            g += id.size();
                            memcpy(g, "\thyperex\tregion\t", 16);
                            g += 16;
                            g += put_u32(g, x.f_start + 1);  // 1-based (utils.rs:382-383)
                            *g++ = '\t';
                            g += put_u32(g, x.r_end + 1);
                            memcpy(g, gff_note[p].data(), gff_note[p].size());
                            g += gff_note[p].size();
                            go.n = (size_t)(g - go.v.data());
                        }
                    } else if (log_fn) {  // utils.rs:387-408
                        if (f_any && r_any)
                            msg = "Forward and reverse primers for region " + labels[p] + " were both found in sequence " +
                                  id + ", but not in a valid order/orientation to define a region";
                        else if (!f_any && r_any)
                            msg = std::string("Forward primer ") + fwd[p] + " not found for region " + labels[p] +
                                  " in sequence " + id;
                        else if (f_any && !r_any)
                            msg = std::string("Reverse primer ") + rev[p] + " not found for region " + labels[p] +
                                  " in sequence " + id;
                        else
                            msg = "Both primers not found for region " + labels[p] + " in sequence " + id;
                        log_fn(1, msg.c_str(), user);
                    }
                }
            }
            if (!fo.ok || !go.ok) rc = HX_ERR_IO;
        }
        t_write += now() - t2;
        {
            std::lock_guard<std::mutex> lk(q.mu);
            q.free_list.push_back(bp);
            if (rc != HX_OK) q.done = true;
            q.cv.notify_all();
        }
        if (rc != HX_OK) break;
    }
    {
        std::lock_guard<std::mutex> lk(q.mu);
        q.done = true;
        q.cv.notify_all();
    }
    parser.join();
    fo.flush();
    go.flush();
    if ((!fo.ok || !go.ok) && rc == HX_OK) rc = HX_ERR_IO;
    if (fclose(fa) != 0 && rc == HX_OK) rc = HX_ERR_IO;
    if (fclose(gff) != 0 && rc == HX_OK) rc = HX_ERR_IO;
    if (timing)
        fprintf(stderr, "[hyperex timing] bases=%llu parse=%.3fs (own thread) wait_for_parser=%.3fs device=%.3fs "
                        "format+write=%.3fs\n",
                (unsigned long long)n_bases, t_parse, t_wait, t_gpu, t_write);
    return rc;
}
