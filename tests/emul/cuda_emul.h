// cuda_emul.h — just enough of the CUDA execution model to run a ONE-WARP-PER-CTA kernel of this repo on the
// CPU, from the same .inl source the GPU build compiles (test infrastructure; used by tests/emul/*.cpp only).
//
// A warp is 32 std::threads in lock step at every warp-wide operation: a shuffle / vote is "everybody writes
// its value, barrier, everybody reads, barrier".  That is valid for kernels that call warp-wide operations
// only from converged code with the full mask, which the emulated kernels do.  Bulk copies (TMA) complete
// immediately at issue, an mbarrier wait is a warp barrier (so the issuing lane's copy is visible), shared
// memory is one global array (CTAs run one after the other).  Timing, phases and asynchrony are NOT modelled:
// this checks arithmetic, index math and the producer/consumer walk, not the hardware protocol.
#pragma once
#include <barrier>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__
#define __align__(x)

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct dim3e { unsigned x = 1, y = 1, z = 1; };
static thread_local dim3e threadIdx, blockIdx;
static dim3e gridDim, blockDim;

namespace emul {
static std::barrier<>* bar = nullptr;
static uint64_t xchg[32];
static inline void sync() { bar->arrive_and_wait(); }
static inline uint64_t exchange(uint64_t mine, int src) {
    xchg[threadIdx.x & 31] = mine;
    sync();
    const uint64_t v = xchg[src & 31];
    sync();
    return v;
}
// run `body` as one CTA of 32 threads (one warp) for every block index of the grid
static inline void launch(unsigned grid, const std::function<void()>& body) {
    gridDim.x = grid;
    blockDim.x = 32;
    for (unsigned b = 0; b < grid; ++b) {
        std::barrier<> br(32);
        bar = &br;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < 32; ++t)
            th.emplace_back([&, t] {
                threadIdx.x = t;
                blockIdx.x = b;
                body();
                br.arrive_and_drop();     // a lane that returns early must not block the others
            });
        for (auto& x : th) x.join();
    }
}
}  // namespace emul

template <class T> static inline T __shfl_sync(uint32_t, T v, int src) { return (T)emul::exchange((uint64_t)v, src); }
template <class T> static inline T __shfl_xor_sync(uint32_t, T v, int m) { return (T)emul::exchange((uint64_t)v, (int)(threadIdx.x ^ (unsigned)m)); }
template <class T> static inline T __shfl_up_sync(uint32_t, T v, int d) {
    const int lane = (int)threadIdx.x;
    return (T)emul::exchange((uint64_t)v, lane >= d ? lane - d : lane);
}
static inline uint32_t __ballot_sync(uint32_t, bool p) {
    emul::xchg[threadIdx.x] = p ? 1u : 0u;
    emul::sync();
    uint32_t m = 0;
    for (int l = 0; l < 32; ++l) m |= (uint32_t)emul::xchg[l] << l;
    emul::sync();
    return m;
}
static inline bool __any_sync(uint32_t m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline void __syncwarp() { emul::sync(); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a < b ? b : a; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// LOP3: bit i of the result = LUT[(a_i << 2) | (b_i << 1) | c_i]
template <int LUT> static inline uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d = 0;
    for (int i = 0; i < 32; ++i) {
        const int idx = (int)(((a >> i) & 1u) << 2 | ((b >> i) & 1u) << 1 | ((c >> i) & 1u));
        d |= (uint32_t)((LUT >> idx) & 1) << i;
    }
    return d;
}

// shared memory: one array; "shared addresses" are offsets from SMEM_BASE
alignas(128) static uint8_t pl_smem[1 << 16];
constexpr uint32_t SMEM_BASE = 0x400u;
static inline uint32_t __cvta_generic_to_shared(const void* p) { return SMEM_BASE + (uint32_t)((const uint8_t*)p - pl_smem); }
static uint64_t emul_tma_bytes = 0, emul_tma_expected = 0;
static inline void mbar_init(uint32_t, uint32_t) {}
static inline void mbar_expect_tx(uint32_t, uint32_t bytes) { emul_tma_expected += bytes; }
static inline void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t) {
    if ((dst & 15u) || ((uintptr_t)src & 15u) || (bytes & 15u) || bytes == 0) { fprintf(stderr, "bulk copy: misaligned or empty (%u bytes)\n", bytes); abort(); }
    std::memcpy(pl_smem + (dst - SMEM_BASE), src, bytes);
    emul_tma_bytes += bytes;
}
static inline void mbar_wait(uint32_t, uint32_t) { emul::sync(); }
