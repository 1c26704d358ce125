// write_bench.cpp — what can this host's cores put into ONE file of the page cache (tmpfs /dev/shm by default)?
// The ceiling of the file-level path's write stage (profiles/r02_file_e2e.md).  T threads each copy 1 MiB slices of a
// 16 MiB source buffer to consecutive file offsets:
//   pwrite     buffered write() (serialises on the inode lock)
//   map        one MAP_SHARED mapping, madvise(MADV_POPULATE_WRITE) + memcpy per slice (what the library does)
//   map-plain  the same without the pre-fault (page faults taken by the stores)
//   anon       memcpy into anonymous memory already faulted in: the memory-bandwidth ceiling, no page cache involved
// g++ -O2 -std=c++17 tools/write_bench.cpp -o /tmp/write_bench -lpthread && /tmp/write_bench [GiB] [dir]
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    const size_t total = (size_t)(argc > 1 ? atoi(argv[1]) : 4) << 30, SL = 1 << 20, SRC = 16 << 20;
    const std::string path = std::string(argc > 2 ? argv[2] : "/dev/shm") + "/hx_write_bench.bin";
    char *src = (char *)malloc(SRC);
    memset(src, 'A', SRC);
    char *anon = (char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    memset(anon, 1, total);
    const int hw = (int)std::thread::hardware_concurrency();
    printf("{\"host_threads\": %d, \"GiB\": %zu, \"file\": \"%s\", \"GBps\": {", hw, total >> 30, path.c_str());
    bool first_mode = true;
    for (const char *mode : {"pwrite", "map", "map-plain", "anon"}) {
        printf("%s\"%s\": {", first_mode ? "" : ", ", mode);
        first_mode = false;
        bool first = true;
        for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
            if (T > hw) break;
            int fd = open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
            if (fd < 0) return 1;
            char *map = nullptr;
            if (mode[0] == 'm') {
                if (ftruncate(fd, (off_t)total) != 0) return 1;
                map = (char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                if (map == MAP_FAILED) return 1;
            }
            std::atomic<size_t> next{0};
            const double t0 = now_s();
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([&] {
                    for (;;) {
                        const size_t off = next.fetch_add(SL);
                        if (off >= total) break;
                        const char *s = src + off % SRC;
                        if (!strcmp(mode, "pwrite")) {
                            if (pwrite(fd, s, SL, (off_t)off) != (ssize_t)SL) abort();
                        } else if (!strcmp(mode, "anon")) {
                            memcpy(anon + off, s, SL);
                        } else {
                            if (!strcmp(mode, "map") && madvise(map + off, SL, MADV_POPULATE_WRITE) != 0) abort();
                            memcpy(map + off, s, SL);
                        }
                    }
                });
            for (auto &t : th) t.join();
            const double dt = now_s() - t0;
            if (map) munmap(map, total);
            close(fd);
            unlink(path.c_str());
            printf("%s\"%d\": %.2f", first ? "" : ", ", T, total / dt / 1e9);
            first = false;
            fflush(stdout);
        }
        printf("}");
    }
    printf("}}\n");
    return 0;
}
