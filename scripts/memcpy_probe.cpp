// Host memcpy ceiling of the box (developer probe): N threads copy disjoint 2 MB blocks of a 128 MB pageable source into a second buffer.
// g++ -O2 -pthread scripts/memcpy_probe.cpp -o /tmp/memcpy_probe && /tmp/memcpy_probe
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main() {
    const size_t n = 128u << 20, block = 2u << 20;
    char* src = (char*)malloc(n);
    char* dst = (char*)malloc(n);
    memset(src, 1, n);
    memset(dst, 2, n);
    for (int threads : {1, 2, 4, 6, 8, 12, 16}) {
        double best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t)
                pool.emplace_back([&, t]() {
                    for (size_t off = (size_t)t * block; off < n; off += (size_t)threads * block) memcpy(dst + off, src + off, block);
                });
            for (auto& th : pool) th.join();
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        printf("threads %2d: %.2f ms for 128 MB = %.1f GB/s\n", threads, best * 1e3, n / best / 1e9);
    }
    return 0;
}
