// Integer-pipe microbenchmark for sm_100a: lane-ops per clock per SM of the instructions the
// selection kernels are built from (LOP3, IADD3, POPC, PRMT, VIADDMNMX.U16x2, IMAD, LDS/STS RMW).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int_pipes int_pipes.cu ; run on one GPU.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int UNROLL = 8;   // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(256) k_op(uint32_t* out, uint32_t seed) {
    uint32_t x[UNROLL], y = seed ^ threadIdx.x, z = seed * 3 + blockIdx.x;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) x[u] = seed + u * 0x9e3779b9u + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[u]) : "r"(y), "r"(z));
            if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[u]) : "r"(y));
            if (OP == 2) asm volatile("popc.b32 %0, %0;" : "+r"(x[u]));
            if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x[u]) : "r"(y));
            if (OP == 4) x[u] = __viaddmin_u16x2(x[u], y, z);
            if (OP == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[u]) : "r"(y), "r"(z));
            if (OP == 6) { uint32_t t; asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(x[u])); x[u] += t; }   // POPC + IADD
            if (OP == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(x[u]) : "r"(y));
            if (OP == 8) { x[u] = __vmaxu4(x[u], y); }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s ^= x[u];
    if (s == 0x12345678u) out[0] = s;
}

// private-counter shared-memory RMW (the k_pair_hist inner step): LDS + IADD + STS per lane-op
__global__ void __launch_bounds__(64) k_rmw(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t h[52 * 64];
    for (int b = 0; b < 52; ++b) h[b * 64 + threadIdx.x] = 0;
    uint32_t v = seed + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            v = v * 1664525u + 1013904223u;
            const uint32_t bin = (v >> 26) % 52u;
            h[bin * 64 + threadIdx.x] += 1;
        }
    }
    uint32_t s = 0;
    for (int b = 0; b < 52; ++b) s += h[b * 64 + threadIdx.x];
    if (s == 0x12345678u) out[0] = s;
}

template <class F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    uint32_t* d;
    cudaMalloc(&d, 64);
    const int sms = pr.multiProcessorCount;
    const char* names[] = {"LOP3", "IADD", "POPC", "PRMT", "VIADDMNMX.U16x2", "IMAD", "POPC+IADD (per pair of instr)", "SHF", "__vmaxu4 (emulated)"};
    printf("%s, %d SMs, %.0f MHz nominal\n", pr.name, sms, clk / 1e6);
#define RUN(OP)                                                                                         \
    {                                                                                                   \
        const int grid = sms * 8;                                                                       \
        float ms = time_ms([&] { k_op<OP><<<grid, 256>>>(d, 12345u); });                                \
        const double ops = (double)grid * 256 * ITERS * UNROLL;                                         \
        printf("%-32s %8.3f ms  %7.1f lane-ops/clk/SM\n", names[OP], ms, ops / (ms * 1e-3) / clk / sms); \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    for (int per_sm : {8, 16, 24}) {
        const int grid = sms * per_sm;
        cudaFuncSetAttribute(k_rmw, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        float ms = time_ms([&] { k_rmw<<<grid, 64>>>(d, 777u); });
        const double ops = (double)grid * 64 * ITERS * UNROLL;
        printf("smem RMW (LDS+IADD+STS), %2d CTAs/SM %8.3f ms  %7.1f lane-RMW/clk/SM\n", per_sm, ms, ops / (ms * 1e-3) / clk / sms);
    }
    return 0;
}
