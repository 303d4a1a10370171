// cuda_emul.h — just enough of the CUDA execution model to run the kernels of this repo on the CPU, from the same
// .inl sources the GPU build compiles (test infrastructure; used by tests/emul/*.cpp only).
//
// A CTA is blockDim.x std::threads; a warp is 32 of them in lock step at every warp-wide operation: a shuffle /
// vote is "everybody writes its value, warp barrier, everybody reads, warp barrier"; __syncthreads is a CTA
// barrier.  That is valid for kernels that call warp-wide operations only from converged code with the full
// mask, which the emulated kernels do (warp_claim, which uses __activemask in divergent code, is replaced by a
// plain atomic claim: same set of slots, another order).  cp.async (per thread) and bulk copies (TMA) on
// mbarriers take the latest legal completion: the destination is poisoned at issue and filled at the wait, and
// every wait is checked for its phase parity and expect_tx balance (see mbar_wait below) — single-warp CTAs, one
// outstanding phase per barrier, which is what the union kernels use.  __shared__ variables are statics and the
// dynamic shared memory one global array (CTAs run one after the other).  Timing and real concurrency of the copy
// engine are NOT modelled: this checks arithmetic, index math, control flow and the ordering rules of the ring, not
// the hardware itself.
#pragma once
#include <barrier>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x)

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int2 { int x, y; };
struct ulonglong2 { unsigned long long x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3e { unsigned x = 1, y = 1, z = 1; };
static thread_local dim3e threadIdx, blockIdx;
static dim3e gridDim, blockDim;

static uint32_t* emul_hists[4] = {nullptr, nullptr, nullptr, nullptr};   // see hist_bias below

static inline void emul_thread_exit();      // a kernel must not end with cp.async copies in flight
namespace emul {
// instruction counts as the source writes them (per thread, added up when a thread of a launch ends): lets a harness
// compare two forms of a kernel by LOP3 / POPC per work item without a GPU
static thread_local uint64_t t_lop3 = 0, t_popc = 0;
static uint64_t n_lop3 = 0, n_popc = 0;
struct Warp {
    std::barrier<> bar;
    uint64_t xchg[32];
    explicit Warp(int lanes) : bar(lanes) {}
};
static thread_local Warp* warp = nullptr;
static std::barrier<>* cta_bar = nullptr;
static inline void sync() { warp->bar.arrive_and_wait(); }
static inline uint64_t exchange(uint64_t mine, int src) {
    warp->xchg[threadIdx.x & 31] = mine;
    sync();
    const uint64_t v = warp->xchg[src & 31];
    sync();
    return v;
}
// run `body` as CTAs of `block` threads (a multiple of 32) for every block index of the grid, one CTA at a time
static inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    gridDim.x = grid;
    blockDim.x = block;
    for (auto& h : emul_hists) h = nullptr;
    for (unsigned b = 0; b < grid; ++b) {
        std::barrier<> cb((std::ptrdiff_t)block);
        cta_bar = &cb;
        std::vector<std::unique_ptr<Warp>> warps;
        for (unsigned w = 0; w < block / 32; ++w) warps.emplace_back(new Warp(32));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t] {
                threadIdx.x = t;
                blockIdx.x = b;
                warp = warps[t / 32].get();
                t_lop3 = t_popc = 0;
                body();
                emul_thread_exit();
                __atomic_fetch_add(&n_lop3, t_lop3, __ATOMIC_RELAXED);
                __atomic_fetch_add(&n_popc, t_popc, __ATOMIC_RELAXED);
                warp->bar.arrive_and_drop();     // a thread that returns early must not block the others
                cb.arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
    cta_bar = nullptr;
}
static inline void launch(unsigned grid, const std::function<void()>& body) { launch(grid, 32, body); }
// two-dimensional grid: blockIdx.y runs over [0, grid_y), one x-sweep after the other
static thread_local unsigned block_y = 0;
static unsigned grid_y_now = 0;
static inline void launch2(unsigned grid_x, unsigned grid_y, unsigned block, const std::function<void()>& body) {
    for (unsigned y = 0; y < grid_y; ++y) {
        grid_y_now = y;
        launch(grid_x, block, [&] { blockIdx.y = grid_y_now; body(); });
    }
    gridDim.y = grid_y;
}
}  // namespace emul

// values travel as bit patterns (floating-point operands keep their bits, as on the device)
namespace emul {
template <class T> static inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8);
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> static inline T from_bits(uint64_t b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
}  // namespace emul
template <class T> static inline T __shfl_sync(uint32_t, T v, int src) { return emul::from_bits<T>(emul::exchange(emul::to_bits(v), src)); }
template <class T> static inline T __shfl_xor_sync(uint32_t, T v, int m) {
    return emul::from_bits<T>(emul::exchange(emul::to_bits(v), (int)((threadIdx.x & 31) ^ (unsigned)m)));
}
template <class T> static inline T __shfl_up_sync(uint32_t, T v, int d) {
    const int lane = (int)(threadIdx.x & 31);
    return emul::from_bits<T>(emul::exchange(emul::to_bits(v), lane >= d ? lane - d : lane));
}
template <class T> static inline T __shfl_down_sync(uint32_t, T v, int d) {
    const int lane = (int)(threadIdx.x & 31);
    return emul::from_bits<T>(emul::exchange(emul::to_bits(v), lane + d < 32 ? lane + d : lane));
}
static inline uint32_t __ballot_sync(uint32_t, bool p) {
    emul::warp->xchg[threadIdx.x & 31] = p ? 1u : 0u;
    emul::sync();
    uint32_t m = 0;
    for (int l = 0; l < 32; ++l) m |= (uint32_t)emul::warp->xchg[l] << l;
    emul::sync();
    return m;
}
static inline bool __any_sync(uint32_t m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline void __syncwarp() { emul::sync(); }
static inline void __syncthreads() { emul::cta_bar->arrive_and_wait(); }
static inline int __popc(uint32_t v) { ++emul::t_popc; return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
// PRMT, default mode: result byte i = byte (selector nibble i) of the eight bytes {y, x}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((src >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
// SIMD-in-word integer intrinsics of the smh filter (per unsigned 16-bit half)
static inline uint32_t emul_half_op(uint32_t a, uint32_t b, uint32_t (*f)(uint32_t, uint32_t)) {
    return (f(a & 0xffffu, b & 0xffffu) & 0xffffu) | (f(a >> 16, b >> 16) << 16);
}
static inline uint32_t __vminu2(uint32_t a, uint32_t b) {
    return emul_half_op(a, b, [](uint32_t x, uint32_t y) { return x < y ? x : y; });
}
static inline uint32_t __vimin3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vminu2(__vminu2(a, b), c); }
static inline uint32_t __viaddmin_u16x2(uint32_t a, uint32_t b, uint32_t c) {      // min(a + b, c) per half, sum mod 2^16
    return __vminu2(emul_half_op(a, b, [](uint32_t x, uint32_t y) { return (x + y) & 0xffffu; }), c);
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a < b ? b : a; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned int atomicAdd_system(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd_system(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline unsigned long long gtime_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
constexpr unsigned long long GATHER_TIMEOUT_NS = 300ull * 1000 * 1000;     // 0.3 s instead of the device's 20 s
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
// helpers.inl's warp_claim aggregates over __activemask(); here every thread claims its own slot
static inline unsigned long long warp_claim(unsigned long long* counter) { return atomicAdd(counter, 1ull); }
// Byte-histogram helpers of helpers.inl (inline PTX there: PRMT-built shared addresses).  Same contract here:
// hist_bias(h) names a [bin][64 threads] counter array, the bias rides in every byte of a packed register word
// (registers are <= 63, so value + 64 * index has no carries), hist_inc* bump the counter [value][thread].
static inline uint32_t hist_bias(const void* hist) {
    for (uint32_t k = 0; k < 4; ++k) {
        uint32_t* expect = nullptr;
        if (__atomic_load_n(&emul_hists[k], __ATOMIC_SEQ_CST) == hist ||
            __atomic_compare_exchange_n(&emul_hists[k], &expect, (uint32_t*)hist, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST) ||
            expect == hist)
            return (k << 6) * 0x01010101u;
    }
    fprintf(stderr, "hist_bias: more than 4 histograms\n");
    abort();
}
static inline void emul_hist_bump(uint32_t byte, uint32_t tb) { emul_hists[byte >> 6][(byte & 63u) * 64u + tb / 4u]++; }
template <int B0, int B1> static inline void hist_inc2(uint32_t wb, uint32_t tb) {
    emul_hist_bump((wb >> (8 * B0)) & 0xffu, tb);
    emul_hist_bump((wb >> (8 * B1)) & 0xffu, tb);
}
template <int B> static inline void hist_inc_dual(uint32_t wb0, uint32_t wb1, uint32_t tb) {
    emul_hist_bump((wb0 >> (8 * B)) & 0xffu, tb);
    emul_hist_bump((wb1 >> (8 * B)) & 0xffu, tb);
}
static inline uint32_t max4_lt128(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t x = (a >> (8 * i)) & 0xffu, y = (b >> (8 * i)) & 0xffu;
        r |= (x > y ? x : y) << (8 * i);
    }
    return r;
}
static inline uint32_t __vmaxu4(uint32_t a, uint32_t b) { return max4_lt128(a, b); }    // bytes of any value
alignas(16) static uint8_t smem_raw[160 << 10];           // the dynamic shared memory of the sketch builder
alignas(1024) static uint32_t hist_dyn[2 * 64 * 64];     // the dynamic shared memory of the hll filters
// cp.async, per thread as on the hardware, with the latest legal completion: the 16 destination bytes are poisoned
// when the copy is queued and written when the thread's wait_group retires the copy's group (all but the newest N
// groups); a reader that has not passed that wait plus the CTA barrier behind it sees the poison.
struct EmulCpAsync { void* dst; const void* src; };
static thread_local std::vector<std::vector<EmulCpAsync>> emul_cp_groups;     // committed groups, oldest first
static thread_local std::vector<EmulCpAsync> emul_cp_open;                    // copies queued since the last commit
static inline void cp_async16(void* smem_dst, const void* gsrc) {
    std::memset(smem_dst, 0xEE, 16);
    emul_cp_open.push_back(EmulCpAsync{smem_dst, gsrc});
}
static inline void cp_async_commit() {
    emul_cp_groups.push_back(std::move(emul_cp_open));
    emul_cp_open.clear();
}
static inline void emul_thread_exit() {
    if (!emul_cp_groups.empty() || !emul_cp_open.empty()) { fprintf(stderr, "cp.async model: thread ends with copies in flight\n"); abort(); }
}
template <int N> static inline void cp_async_wait() {
    while (emul_cp_groups.size() > (size_t)N) {
        for (const EmulCpAsync& c : emul_cp_groups.front()) std::memcpy(c.dst, c.src, 16);
        emul_cp_groups.erase(emul_cp_groups.begin());
    }
}

// LOP3: bit i of the result = LUT[(a_i << 2) | (b_i << 1) | c_i]
template <int LUT> static inline uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d = 0;
    ++emul::t_lop3;
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
// Bulk copies and mbarriers, modelled as the LATEST legal completion: a copy poisons its destination when it is issued
// (the hardware may start writing at once, so nobody may still be reading that stage) and delivers the data only when
// the barrier it completes on is waited for (nobody may read a stage before its wait).  Per barrier: the number of
// completed phases, the bytes announced by expect_tx and the copies in flight.  A wait must name the parity of the
// phase in flight (anything else spins forever or falls through on hardware), a phase's copies must add up to its
// expect_tx, and a barrier takes a new phase only after the previous one was waited for.
struct EmulBulkCopy { uint32_t dst; const void* src; uint32_t bytes; };
struct EmulMbar {
    uint32_t addr = 0;
    uint64_t phases_done = 0, expected = 0;
    bool armed = false;
    std::vector<EmulBulkCopy> inflight;
};
static std::vector<EmulMbar> emul_mbars;
[[noreturn]] static inline void emul_mbar_die(const char* what, uint32_t bar) {
    fprintf(stderr, "mbarrier model: %s (barrier at shared offset 0x%x)\n", what, bar);
    abort();
}
static inline EmulMbar& emul_mbar(uint32_t bar) {
    for (EmulMbar& b : emul_mbars)
        if (b.addr == bar) return b;
    emul_mbar_die("barrier used before mbarrier.init", bar);
}
static inline void mbar_init(uint32_t bar, uint32_t) {
    for (EmulMbar& b : emul_mbars)
        if (b.addr == bar) { b = EmulMbar{}; b.addr = bar; return; }
    emul_mbars.push_back(EmulMbar{});
    emul_mbars.back().addr = bar;
}
static inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    EmulMbar& b = emul_mbar(bar);
    if (b.armed) emul_mbar_die("expect_tx on a barrier whose previous phase was never waited for", bar);
    b.armed = true;
    b.expected = bytes;
    emul_tma_expected += bytes;
}
static inline void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    if ((dst & 15u) || ((uintptr_t)src & 15u) || (bytes & 15u) || bytes == 0) { fprintf(stderr, "bulk copy: misaligned or empty (%u bytes)\n", bytes); abort(); }
    EmulMbar& b = emul_mbar(bar);
    if (!b.armed) emul_mbar_die("bulk copy completing on a barrier without expect_tx", bar);
    std::memset(pl_smem + (dst - SMEM_BASE), 0xEE, bytes);          // in flight: the old content is gone, the new not there
    b.inflight.push_back(EmulBulkCopy{dst, src, bytes});
    emul_tma_bytes += bytes;
}
static inline void mbar_wait(uint32_t bar, uint32_t parity) {
    emul::sync();                                                     // the issuing lane is past its issue
    if ((threadIdx.x & 31) == 0) {
        EmulMbar& b = emul_mbar(bar);
        if (!b.armed) emul_mbar_die("wait on a barrier with no phase in flight", bar);
        if ((parity & 1u) != (b.phases_done & 1u)) emul_mbar_die("wait names the wrong phase parity", bar);
        uint64_t got = 0;
        for (const EmulBulkCopy& c : b.inflight) { std::memcpy(pl_smem + (c.dst - SMEM_BASE), c.src, c.bytes); got += c.bytes; }
        if (got != b.expected) emul_mbar_die("the phase's copies do not add up to its expect_tx", bar);
        b.inflight.clear();
        b.armed = false;
        ++b.phases_done;
    }
    emul::sync();                                                     // the data is there for every lane
}
