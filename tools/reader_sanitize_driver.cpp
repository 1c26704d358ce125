#include "../include/hyperex_b200.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cstdlib>
int main(int argc, char **argv) {
    // write a FASTA with a stop in the middle optionally
    for (int variant = 0; variant < 3; ++variant) {
        std::string buf;
        srand(variant);
        for (int i = 0; i < 20000; ++i) {
            buf += ">r" + std::to_string(i) + (i % 3 ? "" : " desc") + "\n";
            int n = rand() % 300;
            for (int j = 0; j < n; ++j) { buf += "ACGTUNacgt"[rand() % 10]; if (j % 70 == 69) buf += '\n'; }
            buf += '\n';
            if (variant == 1 && i == 15000) buf += ">\n";
        }
        FILE *f = fopen("/tmp/hx_reader_sanitize.fa", "wb"); fwrite(buf.data(), 1, buf.size(), f); fclose(f);
        unsigned long long sums[4] = {0, 0, 0, 0};
        int cfg = 0;
        for (int threads : {1, 3, 6, 0}) {
            hx_fasta *r = nullptr;
            if (hx_fasta_open("/tmp/hx_reader_sanitize.fa", &r)) return 1;
            hx_fasta_configure(r, threads, threads != 1);
            const uint8_t *pk, *fl, *ar; const uint64_t *off; const char *const *ids;
            int64_t n; unsigned long long sum = 0; int batches = 0;
            while ((n = hx_fasta_next_packed(r, 50000, &pk, &fl, &off, &ids, &ar)) > 0) {
                for (int64_t i = 0; i < n; ++i) sum = sum * 31 + fl[i] + strlen(ids[i]) + (off[i + 1] - off[i]);
                for (uint64_t b = 0; b < off[n] / 2; ++b) sum = sum * 131 + pk[b];
                for (uint64_t b = 0; b < off[n]; ++b) sum += ar[b];
                ++batches;
                if (variant == 2 && batches == 5) break;  // early close with work in flight
            }
            hx_fasta_close(r);
            sums[cfg++] = sum;
            printf("variant %d threads %d batches %d sum %llx rc %lld\n", variant, threads, batches, sum, (long long)n);
        }
        if (sums[0] != sums[1] || sums[0] != sums[2] || sums[0] != sums[3]) { printf("MISMATCH\n"); return 2; }
    }
    return 0;
}
