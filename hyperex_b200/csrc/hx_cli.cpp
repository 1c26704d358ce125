// hx_cli.cpp — `hyperex` command line: a drop-in for the reference binary (src/cli.rs:8-83,
// src/main.rs:17-46) whose core step runs on the GPU through the C ABI.
//
//   hyperex [options] [<FILE>]
//     -f, --forward <STR>   forward primer (requires -r, conflicts with --region)     cli.rs:24-32
//     -r, --reverse <STR>   reverse primer                                            cli.rs:35-43
//         --region <STR>... region names v1v2 .. v7v9, or a comma-separated primer file cli.rs:46-53, utils.rs:494
//     -m, --mismatch <N>    allowed edit distance, default 0                           cli.rs:56-64
//     -p, --prefix <PATH>   output prefix, default hyperex_out                         cli.rs:67-74
//         --force           overwrite outputs                                          cli.rs:77-78
//     -q, --quiet           only warnings and errors                                   cli.rs:81-82
//         --gpus <N>        (extension) number of GPUs to shard records over, default 1
#include "../../include/hyperex_b200.h"

#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

namespace {

const char *kVersion = "0.2.0";
bool g_quiet = false;
FILE *g_logfile = nullptr;

enum Level { INFO = 0, WARN = 1, ERROR = 2 };

// utils.rs:69-109 setup_logging: "[HH:MM:SS][LEVEL] msg" on stdout (level coloured like fern's
// ColoredLevelConfig) and "[date][time][target][LEVEL] msg" appended to ./hyperex.log
void log_msg(Level lv, const char *target, const std::string &msg) {
    if (g_quiet && lv == INFO) return;  // quiet => LevelFilter::Warn (utils.rs:76-80)
    static const char *names[] = {"INFO", "WARN", "ERROR"};
    static const int colors[] = {32, 33, 31};
    // a run over a million reads prints a million lines: the time stamp is formatted once per second, and both
    // streams are fully buffered (main sets 1 MB buffers; they are flushed at exit and before anything goes to stderr)
    static time_t last = (time_t)-1;
    static char hms[16], date[16];
    const time_t t = time(nullptr);
    if (t != last) {
        struct tm tmv;
        localtime_r(&t, &tmv);
        strftime(hms, sizeof hms, "%H:%M:%S", &tmv);
        strftime(date, sizeof date, "%Y-%m-%d", &tmv);
        last = t;
    }
    printf("[%s][\x1b[%dm%s\x1b[0m] %s\n", hms, colors[lv], names[lv], msg.c_str());
    if (g_logfile) fprintf(g_logfile, "[%s][%s][%s][%s] %s\n", date, hms, target, names[lv], msg.c_str());
}

void lib_log(int level, const char *msg, void *) { log_msg(level >= 2 ? ERROR : WARN, "hyperex::utils", msg); }

bool is_file(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
bool exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

void usage_error(const std::string &msg) {
    fflush(stdout);
    fprintf(stderr, "error: %s\n\nUsage: hyperex [options] [<FILE>]\n\nFor more information, try '--help'.\n", msg.c_str());
    exit(2);  // clap exits with status 2 on usage errors
}

void print_help() {
    puts("Hypervariable region primer-based extractor\n\n"
         "Usage: hyperex [options] [<FILE>]\n\n"
         "Arguments:\n  [FILE]  Input fasta file or stdin\n\n"
         "Options:\n"
         "  -f, --forward <STR>  Forward primer sequence\n"
         "  -r, --reverse <STR>  Reverse primer sequence\n"
         "      --region <STR>...  Hypervariable region name\n"
         "  -m, --mismatch <N>   Number of allowed mismatch [default: 0]\n"
         "  -p, --prefix <PATH>  Prefix of output files [default: hyperex_out]\n"
         "      --force          Overwrite output\n"
         "  -q, --quiet          Decreases program verbosity\n"
         "      --gpus <N>       Number of GPUs to shard records over [default: 1]\n"
         "  -h, --help           Print help (see more with '--help')\n"
         "  -V, --version        Print version");
}

// utils.rs:444-453: spool stdin (plain FASTA) to ./infile.fa, re-written record by record
std::string read_stdin_to_temp_file() {
    const char *tmp = "infile.fa";
    FILE *raw = fopen("infile.fa.stdin", "wb");
    if (!raw) return tmp;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) fwrite(buf, 1, n, raw);
    fclose(raw);
    // fasta::Writer::write_record: ">id desc\nSEQ\n" — normalise through the library reader
    hx_fasta *r = nullptr;
    FILE *out = fopen(tmp, "wb");
    if (out && hx_fasta_open("infile.fa.stdin", &r) == HX_OK) {
        const uint8_t *arena;
        const uint64_t *off;
        const char *const *ids;
        int64_t cnt;
        while ((cnt = hx_fasta_next(r, 0, &arena, &off, &ids)) > 0) {
            const char *const *descs = hx_fasta_descs(r);
            for (int64_t i = 0; i < cnt; ++i) {
                if (descs[i]) fprintf(out, ">%s %s\n", ids[i], descs[i]);
                else fprintf(out, ">%s\n", ids[i]);
                fwrite(arena + off[i], 1, (size_t)(off[i + 1] - off[i]), out);
                fputc('\n', out);
            }
        }
        hx_fasta_close(r);
    }
    if (out) fclose(out);
    unlink("infile.fa.stdin");
    return tmp;
}

}  // namespace

int main(int argc, char **argv) {
    const auto start = std::chrono::steady_clock::now();
    std::string file, forward, reverse, prefix = "hyperex_out";
    bool has_file = false, has_f = false, has_r = false, has_region = false, force = false;
    std::vector<std::string> regions;
    long mismatch = 0;
    int gpus = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&](const char *name) -> std::string {
            size_t eq = a.find('=');
            if (a.rfind("--", 0) == 0 && eq != std::string::npos) return a.substr(eq + 1);
            if (a.rfind("--", 0) != 0 && a.size() > 2) return a.substr(2);  // -m2
            if (i + 1 >= argc) usage_error(std::string("a value is required for '") + name + "' but none was supplied");
            return argv[++i];
        };
        auto is_opt = [&](const char *s, const char *l) {
            if (s && a.rfind(s, 0) == 0 && a.rfind("--", 0) != 0) return true;
            return l && (a == l || a.rfind(std::string(l) + "=", 0) == 0);
        };
        if (a == "-h" || a == "--help") { print_help(); return 0; }
        if (a == "-V" || a == "--version") { printf("hyperex %s\n", kVersion); return 0; }
        if (a == "--force") force = true;
        else if (a == "-q" || a == "--quiet") g_quiet = true;
        else if (is_opt("-f", "--forward")) { forward = value("--forward <STR>"); has_f = true; }
        else if (is_opt("-r", "--reverse")) { reverse = value("--reverse <STR>"); has_r = true; }
        else if (is_opt("-p", "--prefix")) prefix = value("--prefix <PATH>");
        else if (is_opt("-m", "--mismatch")) {
            std::string v = value("--mismatch <N>");
            char *end = nullptr;
            mismatch = strtol(v.c_str(), &end, 10);
            if (v.empty() || *end || mismatch < 0 || mismatch > 255)
                usage_error("invalid value '" + v + "' for '--mismatch <N>': not a u8");
        } else if (is_opt(nullptr, "--gpus")) {
            const std::string v = value("--gpus <N>");
            char *end = nullptr;
            const long g = strtol(v.c_str(), &end, 10);
            if (v.empty() || *end || g < 1 || g > 64) usage_error("invalid value '" + v + "' for '--gpus <N>': expected 1..64");
            gpus = (int)g;
        }
        else if (is_opt(nullptr, "--region")) {
            has_region = true;
            if (a.find('=') != std::string::npos) regions.push_back(value("--region"));
            // num_args = 1..: every following non-option word belongs to --region (cli.rs:51)
            while (i + 1 < argc && argv[i + 1][0] != '-') regions.push_back(argv[++i]);
            if (regions.empty()) usage_error("a value is required for '--region <STR>...' but none was supplied");
        } else if (a == "-") { file = a; has_file = true; }
        else if (a[0] == '-') usage_error("unexpected argument '" + a + "' found");
        else if (!has_file) { file = a; has_file = true; }
        else usage_error("unexpected argument '" + a + "' found");
    }
    if ((has_f || has_r) && has_region) usage_error("the argument '--forward <STR>' cannot be used with '--region <STR>...'");
    if (has_f != has_r) usage_error("the following required arguments were not provided: " +
                                    std::string(has_f ? "--reverse <STR>" : "--forward <STR>"));
    g_logfile = fopen("hyperex.log", "ab");
    setvbuf(stdout, nullptr, _IOFBF, 1 << 20);
    if (g_logfile) setvbuf(g_logfile, nullptr, _IOFBF, 1 << 20);

    // handle_input (utils.rs:416-442)
    std::string infile;
    if (!has_file || file == "-") infile = read_stdin_to_temp_file();
    else {
        if (!is_file(file)) {
            fprintf(stderr, "error: '%s' is not a valid file. Is the path correct? Do you have permission to read it?\n",
                    file.c_str());
            return 1;
        }
        infile = file;
    }
    // handle_output_files (utils.rs:455-481)
    for (const char *ext : {".fa", ".gff"}) {
        const std::string p = prefix + ext;
        if (!exists(p)) continue;
        if (!force) {
            fprintf(stderr, "error: file %s already exists. Please change it using --prefix option or use --force to "
                            "overwrite it\n", p.c_str());
            return 1;
        }
        unlink(p.c_str());
    }
    // process_primers (utils.rs:483-520)
    std::vector<std::string> F, R;
    char **csv_f = nullptr, **csv_r = nullptr;
    int csv_n = 0;
    size_t csv_longest = 0;
    auto add_region = [&](const char *name) {
        const char *f = nullptr, *r = nullptr;
        if (hx_region_to_primer(name, &f, &r) != HX_OK) return false;
        F.push_back(f);
        R.push_back(r);
        return true;
    };
    if (has_f && has_r) {
        F.push_back(forward);
        R.push_back(reverse);
    } else if (has_region) {
        if (is_file(regions[0])) {  // utils.rs:494-495
            csv_n = hx_file_to_vec(regions[0].c_str(), &csv_f, &csv_r);
            if (csv_n < 0) {
                fprintf(stderr, "Error: %s\n", csv_n == HX_ERR_ARG ? "Primer file must be comma-separated"
                                                                    : "cannot read the primer file");
                return 1;
            }
            for (int i = 0; i < csv_n; ++i) {
                F.push_back(csv_f[i]);
                R.push_back(csv_r[i]);
            }
            hx_free_primer_file(csv_f, csv_r, csv_n);
            // validate_mismatch flattens EVERY field of every line (utils.rs:523-528), not just the two primers
            if (FILE *cf = fopen(regions[0].c_str(), "rb")) {
                size_t cur = 0;
                for (int ch; (ch = fgetc(cf)) != EOF;) {
                    if (ch == ',' || ch == '\n') {
                        csv_longest = std::max(csv_longest, cur);
                        cur = 0;
                    } else if (ch != '\r') {
                        ++cur;
                    }
                }
                csv_longest = std::max(csv_longest, cur);
                fclose(cf);
            }
        } else {
            bool ok = true;
            for (const std::string &r : regions) ok = ok && add_region(r.c_str());
            if (!ok) {
                fprintf(stderr, "Supplied region is not a correct file name nor a supported region name\n");
                return 1;
            }
        }
    } else {
        uint32_t n = 0;
        const char *const *names = hx_supported_regions(&n);
        for (uint32_t i = 0; i < n; ++i) add_region(names[i]);
    }
    // validate_mismatch (utils.rs:522-540)
    size_t longest = csv_longest;
    for (size_t i = 0; i < F.size(); ++i) longest = std::max(longest, std::max(F[i].size(), R[i].size()));
    if (F.empty()) {
        fprintf(stderr, "Error: No primer sequences detected\n");
        return 1;
    }
    if ((size_t)mismatch > longest) {
        log_msg(ERROR, "hyperex::utils", "Supplied mismatch (" + std::to_string(mismatch) +
                                             ") is greater than length of longest primer (" + std::to_string(longest) + ")");
        log_msg(ERROR, "hyperex::utils", "Aborting...");
        return 1;
    }
    // log_program_info (utils.rs:542-558)
    log_msg(INFO, "hyperex::utils", std::string("This is hyperex v") + kVersion);
    log_msg(INFO, "hyperex::utils", "Written by Anicet Ebou");
    log_msg(INFO, "hyperex::utils", "Available at https://github.com/Ebedthan/hyperex.git");
    {
        time_t t = time(nullptr);
        struct tm tmv;
        localtime_r(&t, &tmv);
        char hms[16];
        strftime(hms, sizeof hms, "%H:%M:%S", &tmv);
        log_msg(INFO, "hyperex::utils", std::string("Localtime is ") + hms);
    }
    if (mismatch != 0)
        log_msg(WARN, "hyperex::utils", "You have allowed " + std::to_string(mismatch) + " mismatch in the primer sequence");
    if (force) log_msg(WARN, "hyperex::utils", "Overwriting " + prefix + ".fa and " + prefix + ".gff files");

    // get_hypervar_regions (utils.rs:231-414) on the GPU(s)
    hx_ctx *ctx = nullptr;
    const bool timing = getenv("HYPEREX_TIMING") != nullptr;
    auto since_start = [&] {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    };
    const double t_before_create = since_start();
    int rc = hx_create(&ctx, nullptr, gpus < 1 ? 1 : gpus);
    const double t_after_create = since_start();
    if (rc != HX_OK) {
        fprintf(stderr, "Error: Failed to extract hypervariable regions\n\nCaused by:\n    %s\n", hx_last_error(nullptr));
        return 1;
    }
    std::vector<const char *> fp, rp;
    for (size_t i = 0; i < F.size(); ++i) {
        fp.push_back(F[i].c_str());
        rp.push_back(R[i].c_str());
    }
    rc = hx_get_hypervar_regions(ctx, infile.c_str(), (uint32_t)F.size(), fp.data(), rp.data(), prefix.c_str(),
                                 (uint8_t)mismatch, lib_log, nullptr);
    if (rc != HX_OK) {
        fprintf(stderr, "Error: Failed to extract hypervariable regions\n\nCaused by:\n    %s\n", hx_last_error(ctx));
        hx_destroy(ctx);
        return 1;
    }
    const double t_after_run = since_start();
    // The context is NOT destroyed on this path: both output files are complete and closed, and the process leaves
    // through _exit below; freeing pinned batches and device blocks one by one first cost 0.35 s of the wall time of a
    // 1 M-read run for nothing (HYPEREX_DESTROY=1 keeps the orderly teardown, e.g. under a leak checker).
    if (getenv("HYPEREX_DESTROY")) hx_destroy(ctx);
    if (timing)
        fprintf(stderr, "[hyperex timing] startup=%.3fs hx_create=%.3fs get_hypervar_regions=%.3fs hx_destroy=%.3fs\n",
                t_before_create, t_after_create - t_before_create, t_after_run - t_after_create,
                since_start() - t_after_run);
    log_msg(INFO, "hyperex", "Done getting hypervariable regions");
    // cleanup_and_log (utils.rs:560-582)
    if (exists("infile.fa") && (!has_file || file == "-")) unlink("infile.fa");
    const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - start).count();
    char wt[96];
    snprintf(wt, sizeof wt, "Walltime: %lldh:%lldm:%llds.%lldms", (long long)(ms / 3600000), (long long)(ms % 3600000 / 60000),
             (long long)(ms % 60000 / 1000), (long long)(ms % 1000));
    log_msg(INFO, "hyperex::utils", wt);
    log_msg(INFO, "hyperex::utils", "Enjoy. Share. Come back again!");
    if (g_logfile) fclose(g_logfile);
    if (timing) fprintf(stderr, "[hyperex timing] main total=%.3fs\n", since_start());
    // Everything is released and every file is closed: leave without the CUDA runtime's exit-time teardown of the
    // primary context, which costs more than the whole search on a small input.
    fflush(stdout);
    fflush(stderr);
    _exit(0);
}
