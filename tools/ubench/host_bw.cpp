// host_bw.cpp — what the host of a GPU box can feed a packer with: multi-threaded read / copy bandwidth over a buffer the
// size of the C4 register matrix, and the rate of the 8-bit -> 4-bit register packer of csrc/hostpack.h if present.
// g++ -O3 -march=native -fopenmp tools/ubench/host_bw.cpp -o /tmp/host_bw && /tmp/host_bw
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>
#ifdef USE_PINNED
#include <cuda_runtime.h>
#endif
#include <vector>
#include "../../cuda_selection_criteria_b200/csrc/hostpack.h"
int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? atoll(argv[1]) : 1638400000ll);
#ifdef USE_PINNED      // nvcc -DUSE_PINNED: source and staging in cudaMallocHost memory, as the loader sees them
    uint8_t *src = nullptr, *dst = nullptr;
    if (cudaMallocHost(&src, bytes) != cudaSuccess || cudaMallocHost(&dst, bytes) != cudaSuccess) { printf("cudaMallocHost failed\n"); return 1; }
    printf("pinned source and staging\n");
#else
    uint8_t* src = (uint8_t*)aligned_alloc(4096, bytes);
    uint8_t* dst = (uint8_t*)aligned_alloc(4096, bytes);
#endif
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)bytes; i += 4096) { memset(src + i, (int)(i >> 12) & 15, 4096); memset(dst + i, 0, 4096); }
    const int max_threads = omp_get_max_threads();
    {   // realistic register bytes: geometric values above a per-genome minimum
        unsigned long long x = 88172645463325252ull;
        for (size_t i = 0; i < bytes; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            src[i] = (uint8_t)(7 + __builtin_ctzll(x | (1ull << 20)));
        }
    }
    const size_t m = 16384;
    const long long rows = (long long)(bytes / m);
    std::vector<uint32_t> exc((size_t)rows * selb::NIB4_EXC_CAP);
    std::vector<selb::Nib4Hdr> hdr((size_t)rows);
    for (int threads : {1, 2, 4, 8, 16, 32, 64}) {
        if (threads > max_threads) break;
        omp_set_num_threads(threads);
        double best_cp = 1e9, best_rd = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < (long long)bytes; i += 1 << 20) memcpy(dst + i, src + i, std::min<size_t>(1 << 20, bytes - i));
            auto t1 = std::chrono::steady_clock::now();
            unsigned long long acc = 0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
            for (long long i = 0; i < (long long)bytes; i += 1 << 20) {
                const uint64_t* p = (const uint64_t*)(src + i);
                unsigned long long a = 0;
                const size_t nn = std::min<size_t>(1 << 20, bytes - i) / 8;
                for (size_t j = 0; j < nn; ++j) a += p[j];
                acc += a;
            }
            auto t2 = std::chrono::steady_clock::now();
            best_cp = std::min(best_cp, std::chrono::duration<double>(t1 - t0).count());
            best_rd = std::min(best_rd, std::chrono::duration<double>(t2 - t1).count());
            if (acc == 42) printf("!");
        }
        double best_pk = 1e9;
        long long raw = 0;
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            // chunks of 1024 genomes into a ring of four 8 MiB slots, like the loader
            for (long long g0 = 0; g0 < rows; g0 += 1024) {
                const long long r = std::min<long long>(1024, rows - g0);
                raw += selb::nib4_pack(src + (size_t)g0 * m, r, m, dst + (size_t)((g0 >> 10) & 3) * 1024 * (m / 2), exc.data() + (size_t)g0 * selb::NIB4_EXC_CAP, hdr.data() + g0, threads);
            }
            auto t1 = std::chrono::steady_clock::now();
            best_pk = std::min(best_pk, std::chrono::duration<double>(t1 - t0).count());
        }
        printf("threads %2d: nib4 pack (%s) %.1f GB/s of register bytes (%lld raw genomes)\n", threads, selb::nib4_impl(), bytes / best_pk / 1e9, raw);
        printf("threads %2d: memcpy %.1f GB/s (read+write %.1f)   read %.1f GB/s\n", threads, bytes / best_cp / 1e9, 2 * bytes / best_cp / 1e9, bytes / best_rd / 1e9);
    }
    return 0;
}
