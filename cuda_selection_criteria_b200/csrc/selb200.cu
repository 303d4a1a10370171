// selb200.cu — B200 (sm_100a) kernels + C-ABI for the all-pairs genome selection path.
//
// Replaces, behind include/selb200.h, the hot loop of the reference
// (src/selection.cpp:241-291 and its CUDA restatement src/selection_kernels.cu:13-117):
//
//   load    : chunked H2D overlapped with k_max_byte (validation), k_pair_hist (per-genome histograms),
//             k_genome_cards (Ertl MLE) and k_planes_from_bytes (bit planes of the registers); device
//             radix sort by cardinality, with the host std::sort of selection.cpp:251-256 as the exact
//             fallback when two cardinalities tie; auxiliary sketches re-laid out in sorted order
//             (k_gather_rows / k_aux_transpose / k_aux_planes)
//   run     : K2  k_cb_bounds        CB band [lo(i),hi(i)] per sorted row (binary search, fp64 div)
//                 k_rowblock_span + scan + k_tile_table   the band's 128x128 tile list, built on the device
//             K3  k_smh_signatures   16-bit signature per (genome, LSH band), two bands per word, transposed
//             K4  k_tile_filter_smh  8x8 register micro-tiles: one VIADDMNMX.U16x2 per two bands, cp.async ring
//                 k_smh_verify       exact uint64 compare of the signature-matching band(s)
//                 k_tile_filter_hll_planes  hll_a / hll_an: thread-per-pair aux-HLL union histogram on bit
//                                    planes + MLE (k_tile_filter_hll: the byte form, p_aux < 6)
//                 k_tile_enum        CB-only: every pair of the band
//             K5  k_pair_hist_planes warp-per-pair HLL-14 register max + histogram on bit planes (LOP3 carry-
//                                    save logic, TMA staging); k_pair_hist: the byte form (load, wide pairs)
//             K6  k_estimate_emit    Ertl MLE of the union, Jaccard, tau test, warp-aggregated emit
//                 k_gather_*         multi-GPU: the list goes straight into the root GPU's memory (peer stores)
//             K7  k_rowsort_*        (i,k) order of the reference's stdout (bucket by row; radix sort if dense)
//   Every kernel after K2 reads its work count from device memory, so a run has one host sync before
//   the sort and one at the end.
//
// No tensor cores: the path is byte/integer work bounded by the integer ALU pipe, shared-memory
// wavefronts and L2/HBM bandwidth.  Compile with -fmad=false (see estimators.cuh).
#include "../../include/selb200.h"

#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "estimators.cuh"

// ============================================================================
// error plumbing
// ============================================================================
namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(SELB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

// SELB200_DEBUG_SYNC=1: synchronise after every stage of a run and name the stage that failed
#define DBG_SYNC(c, what)                                                                               \
    do {                                                                                                \
        static const bool dbg__ = getenv("SELB200_DEBUG_SYNC") != nullptr;                              \
        if (dbg__) {                                                                                    \
            cudaError_t e__ = cudaStreamSynchronize((c)->stream);                                       \
            if (e__ != cudaSuccess)                                                                     \
                return fail(SELB200_ECUDA, "stage '%s' failed: %s (%s:%d)", what, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)

#define CKR(call)                      \
    do {                               \
        int r__ = (call);              \
        if (r__ != SELB200_OK) return r__; \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap && p) return SELB200_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes < 256 ? 256 : bytes;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(SELB200_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return SELB200_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int TILE = 128;          // pair tile edge (rows x cols of the sorted order)
constexpr int SIG_CHUNK = 8;       // signature words (2 LSH bands each) staged per shared-memory item
constexpr int64_t PAIR_CAP = 8ll << 20;   // pairs per filter->union pass (list 64 MB, histograms 2 GB)
constexpr int SNAP_MAX = 4096;            // tile ranges per run
#ifndef FILTER_CTAS_PER_SM
#define FILTER_CTAS_PER_SM 2
#endif

}  // namespace

struct LoadState {            // one load in progress (begin -> chunks -> end)
    bool active = false;
    bool regs_borrowed = false;
    bool have_stored = false;
    size_t aux_row_bytes = 0;
    const void* d_aux = nullptr;      // raw aux rows in file-list order (borrowed or scratch)
    int64_t rows_per_chunk = 0;
    int64_t rows_done = 0;
    size_t ev_i = 0;
    // streaming: pinned staging slots the caller decodes into
    int64_t acq_g0 = -1, acq_rows = 0;
    int acq_slot = -1;
};

struct StageSlot {            // pinned host staging of one chunk
    uint8_t* regs = nullptr;
    uint8_t* aux = nullptr;
    double* stored = nullptr;
    size_t regs_cap = 0, aux_cap = 0, stored_cap = 0;
    cudaEvent_t free_ev = nullptr;    // recorded on the copy stream after the slot's H2D copies
    bool in_flight = false;
};

struct selb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;          // H2D staging of load_host, overlapped with the run stream
    std::vector<cudaEvent_t> copy_events;
    int sm_count = 148;

    // loaded sketches
    bool loaded = false;
    int64_t n = 0;
    int p = 0;
    size_t m = 0;
    const uint8_t* d_regs = nullptr;     // [n][m], file-list order (owned copy or borrowed)
    DevBuf regs_own;
    int aux_kind = SELB200_AUX_NONE;
    int aux_len = 0;                     // smh: buckets m_aux; hll: p_aux
    int64_t npad = 0;                    // n rounded up to TILE
    DevBuf aux_sorted;                   // smh: uint64 [n][m_aux] in sorted order
    DevBuf auxT;                         // hll: uint32 [2^p_aux/4][npad], sorted order, transposed
    DevBuf cards_in;                     // double [n] file-list order
    DevBuf e_sorted;                     // uint64 [n] truncated cardinalities, sorted order
    DevBuf order_dev;                    // int32 [n] sorted pos -> file-list index
    std::vector<double> h_cards_sorted;
    std::vector<int32_t> h_order;
    std::vector<uint64_t> h_e;

    // run scratch (grow-only)
    DevBuf lo, hi, tile_prefix, tile_cb0, tile_rc, sigT, cand, pairs, hist, counters, cub_tmp;
    DevBuf out_keys, out_j, out_keys2, out_j2, near_keys, near_j;
    int64_t out_count = 0, near_count = 0;
    int64_t hist_cap_pairs = 0, out_cap = 0;      // grow-only capacities of the sync-free run pipeline
    LoadState ld;
    StageSlot slots[3];
    int next_slot = 0;
    bool _order_cache_valid = false;
    unsigned long long* h_snap = nullptr;         // pinned: per-range counter snapshots
    const uint64_t* res_keys = nullptr;
    const double* res_j = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    // device-built tile list
    DevBuf tile_nt, rb_pairs;
    int64_t tile_cap = 0;
    std::vector<int32_t> h_tprefix;
    std::vector<unsigned long long> h_rb_pairs;
    // peer-memory gather (selb200_gather_*)
    struct Gather {
        bool attached = false, is_root = false, mapped = false;
        int rank = 0, world = 1;
        void* zone = nullptr;            // root: own allocation; others: cudaIpcOpenMemHandle mapping
        int64_t cap = 0, near_cap = 0;
        uint32_t epoch = 0;              // gather runs completed so far (all ranks advance together)
        unsigned long long* h_merged = nullptr;   // pinned [4]
    } g;
    DevBuf g_push, g_merged;
    DevBuf row_cnt, row_off, sort_tmp;
    DevBuf auxP, agrange;                // bit planes / register ranges of the auxiliary HLLs (sorted order)
    DevBuf planes, grange, wide_list;    // bit-plane copy of the primary registers (file-list order)
    int chunk_regs = 0;
    void* h_res = nullptr;               // pinned host copy of the result lists (params.host_results)
    size_t h_res_cap = 0;                // in pairs: keys at [0, cap), Jaccards at [cap, 2 cap)
    int64_t host_count = -1;

    cudaEvent_t ev() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        cudaEvent_t e = ev_pool[ev_used++];
        cudaEventRecord(e, stream);
        return e;
    }
};

// ============================================================================
// device helpers
// ============================================================================
namespace {

// SWAR byte-wise max for bytes < 128 (HLL registers are <= 64-p+1 <= 63):
// the top bit of each byte of (a|0x80..)-b is set iff a>=b, with no borrow between bytes;
// PRMT in sign-replicate mode turns those bits into byte masks.  4 instructions per 4 registers
// (__vmaxu4 is a 7-instruction emulation on sm_100a).
__device__ __forceinline__ uint32_t max4_lt128(uint32_t a, uint32_t b) {
    const uint32_t d = (a | 0x80808080u) - b;
    uint32_t msk;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(msk) : "r"(d), "r"(0u), "r"(0xba98u));
    return (a & msk) | (b & ~msk);
}

// Histogram addressing.  Counters are laid out [bin][64 threads] uint32 in the CTA's static
// shared memory, so the counter of thread t for register value v lives at shared address
//   base + (v << 8) + t*4 .
// `base` is 256-aligned and small, so adding (base >> 8) to every byte of the packed
// register word (no carries: v <= 63) lets ONE PRMT build the complete address from the word
// and tb = t*4 — no per-byte add, and the bank is t mod 32: conflict-free.
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t hist_bias(const void* hist) {
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(hist);
    if ((sbase & 0xffu) || sbase > 0x8000u) __trap();   // must fit: (63 + bias) < 256 and address < 64 KiB
    return (sbase >> 8) * 0x01010101u;
}
template <int B>
__device__ __forceinline__ uint32_t hist_addr(uint32_t wb, uint32_t tb) {
    return __byte_perm(wb, tb, 0x5504 | (B << 4));
}

// Two register values per step into ONE histogram: both counters are loaded before either is
// stored (two LDS in flight instead of a serial LDS->ADD->STS chain); if both hit the same
// counter the second store carries the first increment (select), and stores stay in order.
template <int B0, int B1>
__device__ __forceinline__ void hist_inc2(uint32_t wb, uint32_t tb) {
    const uint32_t o0 = hist_addr<B0>(wb, tb), o1 = hist_addr<B1>(wb, tb);
    const uint32_t c0 = lds_u32(o0) + 1;
    uint32_t c1 = lds_u32(o1);
    c1 = (o1 == o0) ? c0 : c1;
    sts_u32(o0, c0);
    sts_u32(o1, c1 + 1);
}

__device__ __forceinline__ void hist_inc_max16(const uint4& x, const uint4& y, uint32_t bias, uint32_t tb) {
    uint32_t w;
    w = max4_lt128(x.x, y.x) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
    w = max4_lt128(x.y, y.y) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
    w = max4_lt128(x.z, y.z) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
    w = max4_lt128(x.w, y.w) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
}

// One register value into each of TWO different histograms (never alias): both loads first.
template <int B>
__device__ __forceinline__ void hist_inc_dual(uint32_t wb0, uint32_t wb1, uint32_t tb) {
    const uint32_t o0 = hist_addr<B>(wb0, tb), o1 = hist_addr<B>(wb1, tb);
    const uint32_t c0 = lds_u32(o0), c1 = lds_u32(o1);
    sts_u32(o0, c0 + 1);
    sts_u32(o1, c1 + 1);
}

// Warp-aggregated slot claim: one atomicAdd per warp per call site, lanes get consecutive slots.
__device__ __forceinline__ unsigned long long warp_claim(unsigned long long* counter) {
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

// ============================================================================
// K0: register validation — max byte over a buffer (values must be <= 64-p+1)
// ============================================================================
__global__ void __launch_bounds__(256) k_max_byte(const uint4* __restrict__ data, size_t n16, uint32_t* out) {
    uint32_t mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(data + i);
        mx = __vmaxu4(mx, __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w)));
    }
    uint32_t b = max(max(mx & 0xff, (mx >> 8) & 0xff), max((mx >> 16) & 0xff, mx >> 24));
    for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(out, b);
}

// ============================================================================
// K5: warp-per-pair register max + histogram (primary HLL, m >= 512)
//   reference: sketch/include/sketch/hll.h:1188-1206 (union_size: _mm_max_epu8 + 64-bin counts)
//   CTA = 2 warps, each warp owns its own pairs; NB bins x 64 threads x 4 B static smem.
//   Src  : where pair (row a, row b) number pi comes from, and how many there are
//   Epi  : what happens to the finished histogram (lane L holds bins L and L+32)
// ============================================================================
struct SrcPairs {            // pair list of the selection path: sorted positions, mapped through `order`
    const uint2* pairs;
    const int32_t* order;    // nullptr: entries are row indices already
    long long n;             // count, or the capacity when n_dev is given
    const unsigned long long* n_dev;   // optional: the count lives in device memory (no host sync)
    __device__ __forceinline__ long long count() const {
        return n_dev ? (long long)min((unsigned long long)n, *n_dev) : n;
    }
    __device__ __forceinline__ uint2 rows(long long pi, uint2& id) const {
        id = pairs[pi];
        return order ? make_uint2((uint32_t)order[id.x], (uint32_t)order[id.y]) : id;
    }
    __device__ __forceinline__ long long slot(long long pi) const { return pi; }   // histogram row of pair pi
};

struct SrcSelf {             // rows g0..g0+n-1 against themselves: per-genome histograms (max(a,a) = a)
    long long g0, n;
    const uint32_t* max_seen;   // written by k_max_byte earlier on the stream: a register above
    uint32_t max_ok;            // 64-p+1 would index past the histogram, so nothing is processed
    __device__ __forceinline__ long long count() const { return *max_seen > max_ok ? 0 : n; }
    __device__ __forceinline__ uint2 rows(long long pi, uint2& id) const {
        id = make_uint2((uint32_t)(g0 + pi), (uint32_t)(g0 + pi));
        return id;
    }
    __device__ __forceinline__ long long slot(long long pi) const { return pi; }
};

struct EpiWriteHist {        // histogram rows for k_estimate_emit
    uint32_t* out;
    __device__ __forceinline__ void operator()(long long pi, uint2, uint32_t s0, uint32_t s1, uint32_t lane) const {
        out[pi * 64 + lane] = s0;
        out[pi * 64 + 32 + lane] = s1;
    }
};

template <int NB, class Src, class Epi>
__global__ void __launch_bounds__(64)
k_pair_hist(const uint8_t* __restrict__ regs, size_t row_stride, size_t m, Src src, Epi epi) {
    __shared__ __align__(1024) uint32_t hist[NB * 64];
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, tb = t * 4;
    const uint32_t bias = hist_bias(hist);
#pragma unroll 4
    for (int b = 0; b < NB; ++b) hist[b * 64 + t] = 0;
    __syncwarp();
    const int nchunk = (int)(m >> 9);   // 512 B per warp-wide 128-bit load
    const int ngroups = nchunk >> 2;    // software pipeline works on groups of 4 chunks
    const long long nw = (long long)gridDim.x * 2;
    const long long npairs = src.count();
    uint32_t prev0 = 0, prev1 = 0;
    for (long long pi = (long long)blockIdx.x * 2 + w; pi < npairs; pi += nw) {
        uint2 id;
        const uint2 rw = src.rows(pi, id);
        const uint4* a = reinterpret_cast<const uint4*>(regs + (size_t)rw.x * row_stride) + lane;
        const uint4* b = reinterpret_cast<const uint4*>(regs + (size_t)rw.y * row_stride) + lane;
        if (ngroups) {
            // two chunks being histogrammed while the next two are in flight (no register rotation)
            uint4 ax0 = __ldg(a), ay0 = __ldg(b), ax1 = __ldg(a + 32), ay1 = __ldg(b + 32);
            for (int g = 0; g < ngroups; ++g) {
                const uint4 bx0 = __ldg(a + 64), by0 = __ldg(b + 64), bx1 = __ldg(a + 96), by1 = __ldg(b + 96);
                hist_inc_max16(ax0, ay0, bias, tb);
                hist_inc_max16(ax1, ay1, bias, tb);
                a += 128; b += 128;
                if (g + 1 < ngroups) { ax0 = __ldg(a); ay0 = __ldg(b); ax1 = __ldg(a + 32); ay1 = __ldg(b + 32); }
                hist_inc_max16(bx0, by0, bias, tb);
                hist_inc_max16(bx1, by1, bias, tb);
            }
        }
        for (int c = ngroups * 4; c < nchunk; ++c, a += 32, b += 32) hist_inc_max16(__ldg(a), __ldg(b), bias, tb);
        __syncwarp();
        // transposed, conflict-free column sums: lane L totals bins L and L+32.  Counters are
        // never cleared: they run cumulatively (mod 2^32) and the pair's histogram is the
        // difference to the previous totals, which saves the 64 clearing stores per pair.
        uint32_t s0 = 0, s1 = 0;
        const uint32_t cb = w * 32;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const uint32_t col = cb + ((lane + r) & 31);
            s0 += hist[lane * 64 + col];
            if (lane + 32 < NB) s1 += hist[(lane + 32) * 64 + col];
        }
        __syncwarp();
        epi(src.slot(pi), id, s0 - prev0, s1 - prev1, lane);
        prev0 = s0;
        prev1 = s1;
    }
}

// ============================================================================
// K5 (bit-plane form): the same union histogram computed on BIT PLANES of the registers.
//
// The byte kernel above is bound by shared-memory read-modify-writes (measured ~9 registers per clock
// per SM: two wavefronts per 32 registers, tools/ubench/int_pipes.cu).  HLL registers are 6-bit numbers,
// so a genome can also be stored as 6 planes of 2^p bits (12 KiB instead of 16 KiB at p=14), and one
// 32-bit logic instruction then handles 32 registers at once:
//   max(a,b)   : borrow chain of a-b over the planes (1 LOP3 per plane) -> mask "a<b", then one select
//                per plane                                                            12 LOP3 / 32 regs
//   decode     : 8 masks of the low 3 bits + 4 masks of the high 3 bits (the pair's values lie in a
//                window of 32 consecutive values starting at a multiple of 8)         12 LOP3
//   count      : per value, mask = high & low, accumulated with carry-save adders over 4 words
//                (2 CSA = 4 LOP3, 2 POPC, 1 IADD3 per value and 4 words)              ~2 LOP3 / value / word
// i.e. ~2.3-2.8 ALU-pipe operations per register instead of two shared-memory wavefronts per 32.
// POPC issues at 16 lanes/clk/SM on B200 (LOP3: 63), hence the carry-save adders.
// The planes are staged into shared memory by cp.async.bulk (TMA) copies completing on mbarriers;
// one warp per CTA, ~9 CTAs per SM.
// Pairs whose value range does not fit a 32-value window (never seen on real sketches) go to a
// "wide" list and through the byte kernel.
// layout: genome g at planes + g * 6 * m/8 bytes; chunk c (PL_CHUNK_REGS registers, or m if smaller) holds its
//         6 planes back to back: [chunk][plane][chunk_regs/32 words], bit r of word w = register 32w+r
// ============================================================================
// chunk / ring geometry, measured at n=100k (511 521 pairs): 2048 regs x 4 stages x 16 CTAs/SM 1.94 ms,
// 4096 x 3 x 12: 2.02 ms, 8192 x 2 x 9: 1.87 ms (fewer, longer steps: less pipeline control per register)
#ifndef PL_CHUNK_REGS_V
#define PL_CHUNK_REGS_V 8192
#endif
constexpr int PL_CHUNK_REGS = PL_CHUNK_REGS_V;
constexpr int PL_NQ = PL_CHUNK_REGS / 64;   // uint2 per plane of a full chunk

__global__ void __launch_bounds__(256)
k_planes_from_bytes(const uint8_t* __restrict__ regs, long long rows, size_t m, int chunk_regs,
                    uint32_t* __restrict__ planes) {
    const int lane = threadIdx.x & 31;
    const long long nblk = rows * (long long)(m >> 9);          // 512-register blocks
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int blk_per_genome = (int)(m >> 9), blk_per_chunk = chunk_regs >> 9;
    const size_t chunk_words = (size_t)6 * (chunk_regs >> 5);
    for (long long blk = warp0; blk < nblk; blk += nwarps) {
        const long long g = blk / blk_per_genome;
        const int bg = (int)(blk - g * blk_per_genome);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(regs + (size_t)g * m + (size_t)bg * 512) + lane);
        const int chunk = bg / blk_per_chunk, bc = bg - chunk * blk_per_chunk;
        uint32_t* dst = planes + (size_t)g * 6 * (m >> 5) + (size_t)chunk * chunk_words + (size_t)bc * 16;
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            // bit b of the lane's 16 registers -> 16-bit mask (multiply gathers the 4 byte-bits of a word)
            const uint32_t nx = ((((v.x >> b) & 0x01010101u) * 0x10204080u) >> 28);
            const uint32_t ny = ((((v.y >> b) & 0x01010101u) * 0x10204080u) >> 28);
            const uint32_t nz = ((((v.z >> b) & 0x01010101u) * 0x10204080u) >> 28);
            const uint32_t nw = ((((v.w >> b) & 0x01010101u) * 0x10204080u) >> 28);
            const uint32_t h = nx | (ny << 4) | (nz << 8) | (nw << 12);
            const uint32_t lo = __shfl_sync(0xffffffffu, h, 2 * (lane & 15));
            const uint32_t hi = __shfl_sync(0xffffffffu, h, 2 * (lane & 15) + 1);
            if (lane < 16) dst[(size_t)b * (chunk_regs >> 5) + lane] = lo | (hi << 16);
        }
    }
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(phase)
        : "memory");
}

// One chunk (<= PL_CHUNK_REGS registers) against the running carry-save state.  Per step, lane q holds
// two consecutive words of every plane (LDS.64).  Written stage by stage over the 8 values of a group so
// that eight independent dependency chains are in flight (LOP3 latency 4 at one issue per 2 clocks).
template <int G0, int NQ>   // NQ > 0: uint2 per plane known at compile time (full 2048-register chunks)
__device__ __forceinline__ void plane_chunk(const uint2* __restrict__ sA, const uint2* __restrict__ sB, int nq_rt,
                                            int lane, uint32_t gmask, uint32_t (&S)[32], uint32_t (&C2)[32]) {
    const int nq = NQ > 0 ? NQ : nq_rt;
    // nq = uint2 (2 x 32 registers) per plane; plane b of genome X at sX + b*nq.
    // Window 0 holds values below 32 only: plane 5 is all zero there and is neither copied nor read.
    constexpr int NP = (G0 == 0) ? 5 : 6;
#pragma unroll 1
    for (int q = lane; q < nq; q += 32) {
        uint32_t M[2][6];
        {
            uint2 a[NP], b[NP];
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) { a[pl] = sA[pl * nq + q]; b[pl] = sB[pl * nq + q]; }
            uint32_t lt0 = 0u, lt1 = 0u;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {      // borrow of a - b, plane by plane: ends as the mask a < b
                lt0 = lop3<0x8E>(a[pl].x, b[pl].x, lt0);
                lt1 = lop3<0x8E>(a[pl].y, b[pl].y, lt1);
            }
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {      // max = a < b ? b : a
                M[0][pl] = lop3<0xCA>(lt0, b[pl].x, a[pl].x);
                M[1][pl] = lop3<0xCA>(lt1, b[pl].y, a[pl].y);
            }
            if (NP == 5) { M[0][5] = 0u; M[1][5] = 0u; }
        }
        uint32_t L[2][8];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            L[w][0] = lop3<0x01>(M[w][2], M[w][1], M[w][0]);
            L[w][1] = lop3<0x02>(M[w][2], M[w][1], M[w][0]);
            L[w][2] = lop3<0x04>(M[w][2], M[w][1], M[w][0]);
            L[w][3] = lop3<0x08>(M[w][2], M[w][1], M[w][0]);
            L[w][4] = lop3<0x10>(M[w][2], M[w][1], M[w][0]);
            L[w][5] = lop3<0x20>(M[w][2], M[w][1], M[w][0]);
            L[w][6] = lop3<0x40>(M[w][2], M[w][1], M[w][0]);
            L[w][7] = lop3<0x80>(M[w][2], M[w][1], M[w][0]);
        }
        // SPARSE: the top group of a window holds a handful of registers per genome, so most warp-wide
        // steps see none of them and skip the group's 40 instructions after one vote
#define SELB_PLANE_GROUP(T, SPARSE)                                                                       \
        if (gmask & (1u << T)) {                                                                          \
            const uint32_t H0 = lop3<(1 << (G0 + T))>(M[0][5], M[0][4], M[0][3]);                         \
            const uint32_t H1 = lop3<(1 << (G0 + T))>(M[1][5], M[1][4], M[1][3]);                         \
            if (!SPARSE || __any_sync(0xffffffffu, (H0 | H1) != 0u)) {                                    \
                uint32_t m0[8], m1[8], kk[8];                                                             \
                _Pragma("unroll") for (int j = 0; j < 8; ++j) { m0[j] = H0 & L[0][j]; m1[j] = H1 & L[1][j]; } \
                _Pragma("unroll") for (int j = 0; j < 8; ++j) kk[j] = lop3<0xE8>(S[T * 8 + j], m0[j], m1[j]); \
                _Pragma("unroll") for (int j = 0; j < 8; ++j) S[T * 8 + j] = lop3<0x96>(S[T * 8 + j], m0[j], m1[j]); \
                _Pragma("unroll") for (int j = 0; j < 8; ++j) C2[T * 8 + j] += __popc(kk[j]);             \
            }                                                                                             \
        }
        SELB_PLANE_GROUP(0, false)
        SELB_PLANE_GROUP(1, false)
        SELB_PLANE_GROUP(2, false)
        SELB_PLANE_GROUP(3, false)
#undef SELB_PLANE_GROUP
    }
}

// pair list -> histogram rows, bit-plane form.  grange[g] = min | max<<8 of genome g's registers.
// One warp per CTA.  Work comes in batches of 32 consecutive pairs claimed from a device counter (dynamic
// balance, no tail): each lane fetches the descriptor of one pair of the batch (rows through `order`, value
// window from grange), so the dependent global loads are paid once per 32 pairs and the warp then reads
// descriptors with shuffles.  The pairs' planes flow chunk by chunk (8192 registers = 2 x 6 KiB) through a
// ring of PL_STAGES shared-memory stages: lane 0 keeps PL_STAGES-1 bulk copies (TMA) in flight ahead of
// the chunk being counted, across pair and batch boundaries.
#ifndef PL_STAGES_V
#define PL_STAGES_V 2
#endif
constexpr int PL_STAGES = PL_STAGES_V;
#ifndef PL_MIN_CTAS
#define PL_MIN_CTAS 9
#endif

template <class Epi>
__global__ void __launch_bounds__(32, PL_MIN_CTAS)
k_pair_hist_planes(const uint32_t* __restrict__ planes, size_t m, int chunk_regs, const uint16_t* __restrict__ grange,
                   SrcPairs src, Epi epi, uint32_t* __restrict__ wide_list, unsigned long long* __restrict__ wide_count,
                   unsigned long long* __restrict__ batch_counter) {
    extern __shared__ __align__(128) uint8_t pl_smem[];
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int nchunks = (int)(m / (size_t)chunk_regs);
    const uint32_t chunk_bytes = (uint32_t)(6 * (chunk_regs >> 3));
    const int nq = chunk_regs >> 6;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(pl_smem);
    const uint32_t bar0 = smem0 + PL_STAGES * 2 * chunk_bytes;
    if (lane == 0)
        for (int st = 0; st < PL_STAGES; ++st) mbar_init(bar0 + 8 * st, 1);
    __syncwarp();
    const size_t genome_bytes = (size_t)6 * (m >> 3);
    const long long npairs = src.count();
    // batch size: 32 pairs when there is plenty of work, fewer (down to 4) when the list is short, so that
    // every warp still gets several batches and the dynamic claiming can balance the tail
    int bsz = 32;
    while (bsz > 4 && npairs < (long long)bsz * gridDim.x * 4) bsz >>= 1;

    // two descriptor sets (batch k lives in set k&1); per lane: one pair of the batch
    uint32_t d_rx0 = 0, d_ry0 = 0, d_ix0 = 0, d_iy0 = 0, d_gm0 = 0, d_rx1 = 0, d_ry1 = 0, d_ix1 = 0, d_iy1 = 0, d_gm1 = 0;
    uint32_t mask0 = 0, mask1 = 0;          // lanes of the set holding a pair to do (warp-uniform)
    long long base0 = 0, base1 = 0;         // first pair index of the batch
    bool end0 = false, end1 = false;        // the batch starts past the end of the list: nothing follows
    int filled = -1;

    auto fill = [&](int k) {
        long long bidx = 0;
        if (lane == 0) bidx = (long long)atomicAdd(batch_counter, 1ull);
        bidx = __shfl_sync(FULL, bidx, 0);
        const long long pi = bidx * bsz + lane;
        bool ok = lane < bsz && pi < npairs;
        uint2 id = make_uint2(0u, 0u), rw = id;
        uint32_t gm = 0;
        if (ok) {
            rw = src.rows(pi, id);
            const uint32_t ra = grange[rw.x], rb = grange[rw.y];
            const int lo = max((int)(ra & 0xff), (int)(rb & 0xff)), hi = max((int)(ra >> 8), (int)(rb >> 8));
            const int g0 = min(lo >> 3, 4);
            if ((hi >> 3) > g0 + 3) {        // value range wider than the window: the byte kernel does this pair
                wide_list[atomicAdd(wide_count, 1ull)] = (uint32_t)pi;
                ok = false;
            } else {
                uint32_t gmask = 0;
                for (int t = 0; t < 4; ++t)
                    if ((g0 + t) >= (lo >> 3) && (g0 + t) <= (hi >> 3)) gmask |= 1u << t;
                gm = (uint32_t)g0 | (gmask << 8);
            }
        }
        const uint32_t msk = __ballot_sync(FULL, ok);
        if (k & 1) { d_rx1 = rw.x; d_ry1 = rw.y; d_ix1 = id.x; d_iy1 = id.y; d_gm1 = gm; mask1 = msk; base1 = bidx * bsz; end1 = bidx * bsz >= npairs; }
        else       { d_rx0 = rw.x; d_ry0 = rw.y; d_ix0 = id.x; d_iy0 = id.y; d_gm0 = gm; mask0 = msk; base0 = bidx * bsz; end0 = bidx * bsz >= npairs; }
        filled = k;
    };

    struct Cur {               // position in the warp's sequence of (pair, chunk) items; warp-uniform
        int k;                 // batch number
        uint32_t mask;         // pairs of the batch not started yet
        bool valid, done;
        int ch;
        uint32_t rx, ry, ix, iy, gm;
        long long pi;
        const uint8_t* ga;     // planes of the two genomes (producer side)
        const uint8_t* gb;
    };
    Cur cons, prod;
    auto next_pair = [&](Cur& c, bool is_cons) {
        c.valid = false;
        for (;;) {
            // A batch can be empty without being the end (all its pairs wide: one odd genome, many consecutive
            // pairs).  The consumer then walks through it in one go and recycles its set, so a producer
            // still parked on that batch number must not read the set any more: it rejoins the consumer,
            // whose batch it has not touched yet.
            if (!is_cons && c.k < cons.k) {
                c.k = cons.k;
                c.mask = (c.k & 1) ? mask1 : mask0;
            }
            if (c.mask) {
                const int j = __ffs((int)c.mask) - 1;
                c.mask &= c.mask - 1;
                const bool odd = c.k & 1;
                c.rx = __shfl_sync(FULL, odd ? d_rx1 : d_rx0, j);
                c.ry = __shfl_sync(FULL, odd ? d_ry1 : d_ry0, j);
                c.ix = __shfl_sync(FULL, odd ? d_ix1 : d_ix0, j);
                c.iy = __shfl_sync(FULL, odd ? d_iy1 : d_iy0, j);
                c.gm = __shfl_sync(FULL, odd ? d_gm1 : d_gm0, j);
                c.pi = (odd ? base1 : base0) + j;
                if (!is_cons) {
                    c.ga = reinterpret_cast<const uint8_t*>(planes) + (size_t)c.rx * genome_bytes;
                    c.gb = reinterpret_cast<const uint8_t*>(planes) + (size_t)c.ry * genome_bytes;
                }
                c.ch = 0;
                c.valid = true;
                return;
            }
            if ((c.k & 1) ? end1 : end0) { c.done = true; return; }
            if (c.k + 1 > filled) return;                 // producer only: the next batch is not there yet
            ++c.k;
            c.mask = (c.k & 1) ? mask1 : mask0;
            // the consumer has left batch k-1: its set is free for batch k+1
            if (is_cons && !(((filled & 1) ? end1 : end0))) fill(c.k + 1);
        }
    };
    auto issue = [&](const Cur& c, uint32_t n_issued) {      // lane 0: two bulk copies into the next stage
        const uint32_t st = n_issued % PL_STAGES;
        const uint32_t dst = smem0 + st * 2 * chunk_bytes, bar = bar0 + 8 * st;
        const uint8_t* ga = c.ga + (uint32_t)c.ch * chunk_bytes;
        const uint8_t* gb = c.gb + (uint32_t)c.ch * chunk_bytes;
        // window 0 (values < 32): plane 5 is zero and stays behind — 5/6 of the bytes
        const uint32_t nbytes = (c.gm & 0xffu) == 0u ? chunk_bytes / 6u * 5u : chunk_bytes;
        mbar_expect_tx(bar, 2 * nbytes);
        tma_bulk_g2s(dst, ga, nbytes, bar);
        tma_bulk_g2s(dst + chunk_bytes, gb, nbytes, bar);
    };

    fill(0);
    if (!end0) fill(1);
    cons.k = 0; cons.mask = mask0; cons.valid = false; cons.done = false; cons.ch = 0;
    cons.rx = cons.ry = cons.ix = cons.iy = cons.gm = 0; cons.pi = 0;
    cons.ga = cons.gb = nullptr;
    prod = cons;
    next_pair(cons, true);
    next_pair(prod, false);
    uint32_t n_issued = 0, n_done = 0;
    for (int k = 0; k < PL_STAGES - 1 && prod.valid; ++k) {
        if (lane == 0) issue(prod, n_issued);
        ++n_issued;
        if (++prod.ch >= nchunks) next_pair(prod, false);
    }
    uint32_t S[32], C2[32];
#pragma unroll
    for (int v = 0; v < 32; ++v) { S[v] = 0; C2[v] = 0; }
    while (cons.valid) {
        __syncwarp();                          // every lane has finished reading the stage about to be refilled
        if (!prod.valid && !prod.done) next_pair(prod, false);
        if (prod.valid) {
            if (lane == 0) issue(prod, n_issued);
            ++n_issued;
            if (++prod.ch >= nchunks) next_pair(prod, false);
        }
        if (n_done == n_issued) {              // cannot happen: the consumer never overtakes the producer
            if (lane == 0)
                atomicExch(batch_counter + 1, 0xBA00000000000000ull | ((unsigned long long)prod.valid << 55) |
                                                  ((unsigned long long)prod.done << 54) | ((unsigned long long)end0 << 53) |
                                                  ((unsigned long long)end1 << 52) | ((unsigned long long)(prod.k & 0xfff) << 40) |
                                                  ((unsigned long long)(cons.k & 0xfff) << 28) |
                                                  ((unsigned long long)(filled & 0xfff) << 16) | (n_issued & 0xffffu));
            return;
        }
        const uint32_t st = n_done % PL_STAGES;
        mbar_wait(bar0 + 8 * st, (n_done / PL_STAGES) & 1u);
        ++n_done;
        const uint2* pa = reinterpret_cast<const uint2*>(pl_smem + (size_t)st * 2 * chunk_bytes);
        const uint2* pb = reinterpret_cast<const uint2*>(pl_smem + (size_t)st * 2 * chunk_bytes + chunk_bytes);
        const int g0 = (int)(cons.gm & 0xffu);
        const uint32_t gmask = cons.gm >> 8;
        if (nq == PL_NQ) {
            switch (g0) {
                case 0: plane_chunk<0, PL_NQ>(pa, pb, nq, lane, gmask, S, C2); break;
                case 1: plane_chunk<1, PL_NQ>(pa, pb, nq, lane, gmask, S, C2); break;
                case 2: plane_chunk<2, PL_NQ>(pa, pb, nq, lane, gmask, S, C2); break;
                case 3: plane_chunk<3, PL_NQ>(pa, pb, nq, lane, gmask, S, C2); break;
                default: plane_chunk<4, PL_NQ>(pa, pb, nq, lane, gmask, S, C2); break;
            }
        } else {
            switch (g0) {
                case 0: plane_chunk<0, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                case 1: plane_chunk<1, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                case 2: plane_chunk<2, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                case 3: plane_chunk<3, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                default: plane_chunk<4, 0>(pa, pb, nq, lane, gmask, S, C2); break;
            }
        }
        if (cons.ch == nchunks - 1) {
            // per-lane totals, then a transposing butterfly: lane L ends with the warp total of value 8*g0 + L
            uint32_t x[32];
#pragma unroll
            for (int v = 0; v < 32; ++v) { x[v] = 2u * C2[v] + (uint32_t)__popc(S[v]); S[v] = 0; C2[v] = 0; }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const bool upper = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < o; ++i) {
                    const uint32_t send = upper ? x[i] : x[i + o];
                    const uint32_t keep = upper ? x[i + o] : x[i];
                    x[i] = keep + __shfl_xor_sync(FULL, send, o);
                }
            }
            const uint32_t tot = __shfl_sync(FULL, x[0], (lane - 8 * g0) & 31);
            const bool first = lane >= 8 * g0;     // bin `lane` lies inside the window; else bin lane+32 does
            epi(src.slot(cons.pi), make_uint2(cons.ix, cons.iy), first ? tot : 0u, first ? 0u : tot, (uint32_t)lane);
        }
        if (++cons.ch >= nchunks) next_pair(cons, true);
    }
}

// the wide list as a pair source for the byte kernel (histogram row = the pair's own slot)
struct SrcWide {
    const uint2* pairs;
    const int32_t* order;
    const uint32_t* wide_list;
    const unsigned long long* n_dev;
    __device__ __forceinline__ long long count() const { return (long long)*n_dev; }
    __device__ __forceinline__ uint2 rows(long long wi, uint2& id) const {
        id = pairs[wide_list[wi]];
        return order ? make_uint2((uint32_t)order[id.x], (uint32_t)order[id.y]) : id;
    }
    __device__ __forceinline__ long long slot(long long wi) const { return (long long)wide_list[wi]; }
};

__global__ void k_iota_i32(int32_t* v, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

// sorted cardinalities -> truncated e (size_t e = card, selection.cpp:275,280) + tie detection
__global__ void k_sorted_prep(const double* __restrict__ cards_sorted, long long n, unsigned long long* __restrict__ e,
                              uint32_t* __restrict__ tie_flag) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cd = cards_sorted[i];
    e[i] = (unsigned long long)cd;
    if (i + 1 < n && !(cd < cards_sorted[i + 1])) *tie_flag = 1;
}

// per-genome cardinality: hll.h:834-837 (sum) / :1138-1141 (trusted stored value)
__global__ void k_genome_cards(const uint32_t* __restrict__ hist, const double* __restrict__ stored, long long n,
                               int p, double* __restrict__ cards, const uint32_t* __restrict__ max_seen,
                               uint32_t max_ok, uint16_t* __restrict__ grange) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (*max_seen > max_ok) { cards[i] = 0.; grange[i] = 0; return; }   // malformed input: the load fails after the sync
    {   // smallest and largest register value of the genome (window choice of the bit-plane union kernel)
        int vmin = 63, vmax = 0;
        for (int b = 0; b < 64; ++b)
            if (hist[i * 64 + b]) { vmin = min(vmin, b); vmax = max(vmax, b); }
        grange[i] = (uint16_t)(min(vmin, vmax) | (vmax << 8));
    }
    if (stored && stored[i] >= 0.) { cards[i] = stored[i]; return; }
    cards[i] = selb::ertl_mle(hist + i * 64, p);
}

__global__ void k_mle_only(const uint32_t* __restrict__ hist, long long n, int p, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = selb::ertl_mle(hist + i * 64, p);
}

// ============================================================================
// load-time re-layout
// ============================================================================
// dst[i][:] = src[order[i]][:], rows of row_words uint32
__global__ void k_gather_rows(const uint32_t* __restrict__ src, const int32_t* __restrict__ order, long long n,
                              int row_words, uint32_t* __restrict__ dst) {
    const long long total = n * row_words;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / row_words;
        const int j = (int)(idx - i * row_words);
        dst[idx] = src[(size_t)order[i] * row_words + j];
    }
}

// auxT[j][g] = word j of the aux HLL of the g-th genome in sorted order (pad columns stay 0)
__global__ void k_aux_transpose(const uint32_t* __restrict__ src, const int32_t* __restrict__ order, long long n,
                                long long npad, int row_words, uint32_t* __restrict__ dst) {
    const long long total = n * row_words;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx / n);
        const long long g = idx - (long long)j * n;
        dst[(size_t)j * npad + g] = src[(size_t)order[g] * row_words + j];
    }
}

// ============================================================================
// K2: CB band per sorted row
//   reference: src/selection.cpp:278-283 — skip e2==0, break at the first CB failure.
//   Sorted ascending + correctly-rounded fp64 division => the passing set of row i is the
//   contiguous range [lo(i), hi(i)], lo = max(i+1, first index with e>0).
// ============================================================================
__global__ void k_cb_bounds(const unsigned long long* __restrict__ e, int n, int zeros, double tau,
                            int32_t* __restrict__ lo, int32_t* __restrict__ hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e1 = e[i];
    const int l = max(i + 1, zeros);
    int a = l, b = n;   // first k in [l,n) failing CB
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (selb::crit_cb(tau, e1, e[mid])) a = mid + 1; else b = mid;
    }
    lo[i] = l;
    hi[i] = a - 1;
}

// ============================================================================
// tile list of the CB band, built on the device (no host round trip):
//   k_rowblock_span : per 128-row block, the column-block span of its band and its pair count
//   cub exclusive scan over the spans -> first tile index of every row block
//   k_tile_table    : (row block, column block) of every tile, so that a filter CTA finds its
//                     tile with one 8-byte load
// meta[] (unsigned long long, device): [0] candidates [1] pairs [2] out [3] near of the current
// range, [4] pairs inside the CB band, [5] tiles of the band, [6] gather: pushed flag
// ============================================================================
enum { M_CAND = 0, M_PAIRS = 1, M_OUT = 2, M_NEAR = 3, M_PAIRS_CB = 4, M_TILES = 5, M_PUSHED = 6, M_WIDE = 7, M_BATCH = 8, M_KERR = 9, M_UNIT = 10, M_WORDS = 16 };

__global__ void __launch_bounds__(128)
k_rowblock_span(const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n, int nrb,
                int32_t* __restrict__ nt, int32_t* __restrict__ cb0, unsigned long long* __restrict__ rb_pairs,
                unsigned long long* __restrict__ meta) {
    const int rb = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rb > nrb) return;
    if (rb == nrb) { if (lane == 0) nt[nrb] = 0; return; }   // scan sentinel: prefix[nrb] = total
    int cmin = INT32_MAX, cmax = -1;
    unsigned long long cnt = 0;
    for (int i = rb * TILE + lane; i < min(n, (rb + 1) * TILE); i += 32) {
        const int l = lo[i], h = hi[i];
        if (h < l) continue;
        cnt += (unsigned long long)(h - l + 1);
        cmin = min(cmin, l);
        cmax = max(cmax, h);
    }
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    if (lane == 0) {
        nt[rb] = cnt ? cmax / TILE - cmin / TILE + 1 : 0;
        cb0[rb] = cnt ? cmin / TILE : 0;
        rb_pairs[rb] = cnt;
        if (cnt) atomicAdd(meta + M_PAIRS_CB, cnt);
    }
}

__global__ void __launch_bounds__(128)
k_tile_table(const int32_t* __restrict__ tile_prefix, const int32_t* __restrict__ cb0, int nrb, long long tile_cap,
             int2* __restrict__ tile_rc, unsigned long long* __restrict__ meta) {
    const int rb = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rb >= nrb) return;
    const int a = tile_prefix[rb], cnt = tile_prefix[rb + 1] - a, c0 = cb0[rb];
    for (int t = lane; t < cnt; t += 32)
        if (a + t < tile_cap) tile_rc[a + t] = make_int2(rb, c0 + t);
    if (rb == 0 && lane == 0) meta[M_TILES] = (unsigned long long)tile_prefix[nrb];
}

// tiles owned by one shard: tile = shard + j * n_shards for j in [0, count)
struct TileWalk {
    const int2* tile_rc;
    const unsigned long long* meta;
    long long tile_cap;
    int shard, n_shards;
    int j0, j1;          // this launch covers j in [j0, j1) (clipped to the shard's tile count)
    __device__ __forceinline__ int count() const {
        const long long total = (long long)min((unsigned long long)tile_cap, meta[M_TILES]);
        const long long mine = total > shard ? (total - shard + n_shards - 1) / n_shards : 0;
        return (int)min((long long)j1, mine);
    }
    __device__ __forceinline__ int2 tile(int j) const { return __ldg(tile_rc + shard + (long long)j * n_shards); }
};

// ============================================================================
// K3: LSH band signatures.  For the g-th sorted genome and band b, sig(b,g) = 16 bits of a mix
// of the band's n_rows buckets.  Equal bands => equal signatures, so "some band equal"
// (criteria_sketch.hpp:71-79) implies "some signature equal"; the converse is checked exactly
// by k_smh_verify.  Two bands share one 32-bit word: word w holds bands 2w (low half) and 2w+1.
//   sigR[w][g] =  sig              (row operand)
//   sigC[w][g] = -sig per half     (column operand), so  r + c == 0 (mod 2^16)  <=>  equal
// An odd band count leaves a pad half that can never match (row 0, column 1); pad genomes hold
// row 0 / column 0x0101.
// ============================================================================
__device__ __forceinline__ uint32_t band_sig16(const uint64_t* v, int n_rows) {
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int r = 0; r < n_rows; ++r) h = mix64(h ^ v[r]);
    return (uint32_t)(h >> 48);
}

__global__ void __launch_bounds__(256)
k_smh_signatures(const uint64_t* __restrict__ aux_sorted, long long n, long long npad, int m_aux,
                 int n_rows, int n_bands, uint32_t* __restrict__ sigR, uint32_t* __restrict__ sigC) {
    // thread = (genome, band), band fastest: a warp reads consecutive bands of one genome, i.e. one contiguous
    // run of its sketch; lane pairs then pack two bands into a word
    const int nw = (n_bands + 1) >> 1;
    const int nb2 = nw * 2;                                    // bands rounded up to even (pad band never matches)
    const long long total = npad * nb2;                        // pad genomes included: row halves 0, column halves 0x0101
    const long long stride = (long long)gridDim.x * blockDim.x;          // even: lane pairs stay together
    for (long long idx0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx0 - (threadIdx.x & 31) < total;
         idx0 += stride) {
        const bool live = idx0 < total;
        const long long g = live ? idx0 / nb2 : 0;
        const int b = live ? (int)(idx0 - g * nb2) : 0;
        uint32_t sig = 0;
        const bool real = live && b < n_bands && g < n;
        if (real) sig = band_sig16(aux_sorted + (size_t)g * m_aux + (size_t)b * n_rows, n_rows);
        const uint32_t other = __shfl_down_sync(0xffffffffu, sig, 1);
        const bool other_real = __shfl_down_sync(0xffffffffu, (int)real, 1) != 0;
        if (live && !(b & 1) && g >= n) {                      // pad genome: never matches anything
            const int w = b >> 1;
            sigR[(size_t)w * npad + g] = 0u;
            sigC[(size_t)w * npad + g] = 0x01010101u;
        } else if (live && !(b & 1)) {
            const uint32_t r1 = other_real ? other : 0u;
            const uint32_t c1 = other_real ? ((0u - other) & 0xffffu) : 1u;   // pad half: row 0, column 1
            const int w = b >> 1;
            sigR[(size_t)w * npad + g] = sig | (r1 << 16);
            sigC[(size_t)w * npad + g] = ((0u - sig) & 0xffffu) | (c1 << 16);
        }
    }
}

// ============================================================================
// K4: smh_a tile pre-filter.  One CTA = one 128x128 tile of the sorted pair space,
// 256 threads, each an 8x8 register micro-tile.  Per signature word (two bands): 4 LDS.128 and
// 64 x VIADDMNMX.U16x2 (acc = min(acc, r + c) per 16-bit half); a zero half at the end
// <=> some band signature matched.  Candidates (rare) leave through warp-aggregated atomics.
// ============================================================================
// Accumulate: acc = min(acc, r + c) per 16-bit half in ONE instruction (VIADDMNMX.U16x2); the column
// operand holds the negated halves, so a half reaches 0 exactly when the two signatures are equal.
// Measured alternatives on B200 (n=100k, 4.66e8 CB pairs): XOR+MIN on 32-bit signatures 1.22 ms;
// (LOP3, IADD, LOP3) zero-half test on packed halves 1.07 ms; this form 0.86 ms.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, FILTER_CTAS_PER_SM)
k_tile_filter_smh(const uint32_t* __restrict__ sigR, const uint32_t* __restrict__ sigC, long long npad, int n_words,
                  TileWalk tw, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
                  uint2* __restrict__ cand, unsigned long long* __restrict__ cand_count,
                  unsigned long long cand_cap) {
    // three buffers: the signature words of the next two (tile, chunk) items stream in with cp.async while
    // the current one is being compared — a tile's 8 KiB arrive in about the time its 512 instructions per
    // thread take, so without the overlap the ALU pipe idles half the time; with three buffers ONE barrier
    // per item both publishes the item's copies and frees the buffer of the item before it
    constexpr int NBUF = 3;
    __shared__ __align__(16) uint32_t sR[NBUF][SIG_CHUNK][TILE];
    __shared__ __align__(16) uint32_t sC[NBUF][SIG_CHUNK][TILE];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int jend = tw.count();
    const int nchunk = (n_words + SIG_CHUNK - 1) / SIG_CHUNK;
    // persistent CTAs: the shard's tile count lives in device memory, so no host sync sizes the grid
    const int j_first = tw.j0 + (int)blockIdx.x;
    if (j_first >= jend) return;
    const int my_tiles = (jend - 1 - j_first) / (int)gridDim.x + 1;
    const int n_items = my_tiles * nchunk;                 // items = (tile, chunk of SIG_CHUNK words), no divisions below

    auto tile_at = [&](int t) -> int2 { return t < my_tiles ? tw.tile(j_first + t * (int)gridDim.x) : make_int2(0, 0); };
    auto stage = [&](int ch, int2 rc, int buf) {           // queue the loads of one item
        const int b0 = ch * SIG_CHUNK;
        const int nb = min(SIG_CHUNK, n_words - b0);
        const int r0 = rc.x * TILE, c0 = rc.y * TILE;
        for (int idx = tid; idx < nb * 64; idx += 256) {   // nb words x (128 row + 128 col) / 4 per copy
            const int bb = idx >> 6, part = idx & 63, x = (part & 31) * 4;
            if (part < 32) cp_async16(&sR[buf][bb][x], sigR + (size_t)(b0 + bb) * npad + r0 + x);
            else cp_async16(&sC[buf][bb][x], sigC + (size_t)(b0 + bb) * npad + c0 + x);
        }
        cp_async_commit();
    };
    // zero 16-bit half somewhere in x
    auto has_zero_half = [](uint32_t x) { return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u; };

    uint32_t acc[8][8];
    // (t, ch) = item being compared; (ts, chs) = the next item to copy, two items ahead.  The coordinates of the
    // copy cursor's tile and of the tile after it are fetched ahead of use (rc_s, rc_sn).
    int t = 0, ch = 0, buf = 0;
    int ts = 0, chs = 0, bufs = 0;
    int2 rc_s = tile_at(0), rc_sn = tile_at(1);
    __shared__ int2 s_rc[4];                               // coordinates of the tiles in flight, by tile number & 3
    auto stage_next = [&]() {
        if (chs == 0 && tid == 0) s_rc[ts & 3] = rc_s;
        stage(chs, rc_s, bufs);
        bufs = bufs + 1 == NBUF ? 0 : bufs + 1;
        if (++chs == nchunk) {
            chs = 0;
            ++ts;
            rc_s = rc_sn;
            rc_sn = tile_at(ts + 1);
        }
    };
    int staged = 0;
    for (; staged < 2 && staged < n_items; ++staged) stage_next();
    for (int item = 0; item < n_items; ++item) {
        if (item + 1 < staged) cp_async_wait<1>();          // everything but the newest group has landed
        else cp_async_wait<0>();
        __syncthreads();
        if (staged < n_items) { stage_next(); ++staged; }   // into the buffer item-1 was compared from
        const int nb = min(SIG_CHUNK, n_words - ch * SIG_CHUNK);
        int bb = 0;
        if (ch == 0) {                                      // first word of a tile: no accumulator to read
            const uint4 ra = *reinterpret_cast<const uint4*>(&sR[buf][0][ty * 8]);
            const uint4 rb = *reinterpret_cast<const uint4*>(&sR[buf][0][ty * 8 + 4]);
            const uint4 ca = *reinterpret_cast<const uint4*>(&sC[buf][0][tx * 4]);
            const uint4 cb = *reinterpret_cast<const uint4*>(&sC[buf][0][64 + tx * 4]);
            const uint32_t rs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const uint32_t cs[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = __viaddmin_u16x2(rs[a], cs[b], 0xffffffffu);
            bb = 1;
        }
        for (; bb < nb; ++bb) {
            const uint4 ra = *reinterpret_cast<const uint4*>(&sR[buf][bb][ty * 8]);
            const uint4 rb = *reinterpret_cast<const uint4*>(&sR[buf][bb][ty * 8 + 4]);
            const uint4 ca = *reinterpret_cast<const uint4*>(&sC[buf][bb][tx * 4]);
            const uint4 cb = *reinterpret_cast<const uint4*>(&sC[buf][bb][64 + tx * 4]);
            const uint32_t rs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const uint32_t cs[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = __viaddmin_u16x2(rs[a], cs[b], acc[a][b]);
        }
        buf = buf + 1 == NBUF ? 0 : buf + 1;
        if (++ch < nchunk) continue;
        const int t_done = t;
        ch = 0;
        ++t;
        // per-row minima first: a 16-bit signature collides by chance once per ~64 thread-tiles, so four
        // warps in ten come here with ONE row to look at, not 64 cells
        uint32_t rowmin[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const uint32_t m0 = __vimin3_u16x2(acc[a][0], acc[a][1], acc[a][2]);
            const uint32_t m1 = __vimin3_u16x2(acc[a][3], acc[a][4], acc[a][5]);
            rowmin[a] = __vimin3_u16x2(m0, m1, __vminu2(acc[a][6], acc[a][7]));
        }
        const uint32_t any = __vimin3_u16x2(__vimin3_u16x2(rowmin[0], rowmin[1], rowmin[2]),
                                            __vimin3_u16x2(rowmin[3], rowmin[4], rowmin[5]),
                                            __vminu2(rowmin[6], rowmin[7]));
        if (!has_zero_half(any)) continue;
        const int2 rc = s_rc[t_done & 3];                   // written when the tile was queued, barriers ago
        const int r0 = rc.x * TILE, c0 = rc.y * TILE;
        unsigned long long cells = 0ull;                    // bit a*8+b: cell (a,b) has a matching band signature
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            if (!has_zero_half(rowmin[a])) continue;
            uint32_t rowbits = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) rowbits |= has_zero_half(acc[a][b]) ? (1u << b) : 0u;
            cells |= (unsigned long long)rowbits << (a * 8);
        }
        while (cells) {
            const int bit = __ffsll((long long)cells) - 1;
            cells &= cells - 1;
            const int a = bit >> 3, b = bit & 7;
            const int i = r0 + ty * 8 + a;
            const int k = c0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
            if (i >= n || k < lo[i] || k > hi[i]) continue;
            const unsigned long long slot = warp_claim(cand_count);
            if (slot < cand_cap) cand[slot] = make_uint2((uint32_t)i, (uint32_t)k);
        }
    }
}

// exact smh_a on the candidates: include/criteria_sketch.hpp:66-81.  Thread per candidate.
// A band can only be equal if its 16-bit signatures are, so the thread re-reads the (L2-resident)
// signature words of both genomes, and compares bucket by bucket only the bands whose signatures
// match, stopping at the first band that is really equal: ~2 x 64 B of auxiliary sketch per
// candidate instead of 2 x 8m B.
__global__ void __launch_bounds__(256)
k_smh_verify(const uint64_t* __restrict__ aux_sorted, const uint32_t* __restrict__ sigR, long long npad, int m_aux,
             int n_rows, int n_bands, const uint2* __restrict__ cand,
             const unsigned long long* __restrict__ ncand_dev, unsigned long long cand_cap,
             uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count, unsigned long long pair_cap) {
    const long long ncand = (long long)min(*ncand_dev, cand_cap);
    const int n_words = (n_bands + 1) >> 1;
    for (long long ci = blockIdx.x * (long long)blockDim.x + threadIdx.x; ci < ncand;
         ci += (long long)gridDim.x * blockDim.x) {
        const uint2 pr = cand[ci];
        const uint64_t* v1 = aux_sorted + (size_t)pr.x * m_aux;
        const uint64_t* v2 = aux_sorted + (size_t)pr.y * m_aux;
        bool hit = false;
        for (int w = 0; w < n_words && !hit; ++w) {
            const uint32_t x = __ldg(sigR + (size_t)w * npad + pr.x) ^ __ldg(sigR + (size_t)w * npad + pr.y);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int b = 2 * w + half;
                if (hit || b >= n_bands || ((x >> (16 * half)) & 0xffffu) != 0) continue;
                bool eq = true;
                for (int r = 0; r < n_rows; ++r)
                    if (__ldg(v1 + (size_t)b * n_rows + r) != __ldg(v2 + (size_t)b * n_rows + r)) { eq = false; break; }
                hit = eq;
            }
        }
        if (hit) {
            const unsigned long long slot = warp_claim(pair_count);
            if (slot < pair_cap) pairs[slot] = pr;
        }
    }
}

// CB only: every pair of the band inside this tile
__global__ void __launch_bounds__(256)
k_tile_enum(TileWalk tw, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
            uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count, unsigned long long pair_cap) {
    const int jend = tw.count();
    for (int j = tw.j0 + (int)blockIdx.x; j < jend; j += (int)gridDim.x) {
        const int2 rc = tw.tile(j);
        const int r0 = rc.x * TILE, c0 = rc.y * TILE;
        for (int idx = threadIdx.x; idx < TILE * TILE; idx += 256) {
            const int i = r0 + (idx >> 7), k = c0 + (idx & (TILE - 1));
            if (i >= n || k >= n) continue;
            if (k < lo[i] || k > hi[i]) continue;
            const unsigned long long slot = warp_claim(pair_count);
            if (slot < pair_cap) pairs[slot] = make_uint2((uint32_t)i, (uint32_t)k);
        }
    }
}

// ============================================================================
// K4': hll_a / hll_an tile filter.  Thread per pair; lanes = 32 consecutive columns of one
// row pair (R=2 rows share each column word).  Aux registers come transposed
// (auxT[word][genome]) so a warp's column load is one coalesced 128 B line and the row word
// is a broadcast.  Each thread keeps R private histograms [bin][64 threads] in static smem
// (same PRMT addressing as k_pair_hist), then runs the Ertl MLE on its own columns and the
// criterion:
//   hll_a  include/criteria_sketch.hpp:60-64,36-43   hll_an  :52-58,22-34
// One CTA (2 warps) handles a 32-row x 128-col quarter of a tile.
// ============================================================================
struct StopHll {      // early exit of the MLE: the criterion already fails at the lower bound
    double tau;
    unsigned long long e1, e2;
    float zs;
    int order_n;
    int an;
    __device__ __forceinline__ bool crit(double t) const {
        return an ? selb::crit_hll_an(tau, e1, e2, t, zs, order_n) : selb::crit_hll_a(tau, e1, e2, t, zs);
    }
    // both criteria are non-increasing in t only for Z*sigma >= 0 (the reference hard-codes Z = 1.96)
    __device__ __forceinline__ bool operator()(double t_lb) const { return zs >= 0.f && !crit(t_lb); }
};

template <int AN>
__global__ void __launch_bounds__(64)
k_tile_filter_hll(const uint32_t* __restrict__ auxT, long long npad, int p_aux, TileWalk tw,
                  const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
                  const unsigned long long* __restrict__ e, double tau, float zs, int order_n,
                  uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count,
                  unsigned long long pair_cap, unsigned long long* __restrict__ unit_counter) {
    extern __shared__ __align__(1024) uint32_t hist_dyn[];   // 2 x [nbins][64 threads]
    __shared__ int s_unit;
    const int nbins = 64 - p_aux + 2;
    uint32_t* hist0 = hist_dyn;
    uint32_t* hist1 = hist_dyn + nbins * 64;
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, tb = t * 4;
    const int words = (1 << p_aux) >> 2;
    const uint32_t bias0 = hist_bias(hist0), bias1 = hist_bias(hist1);
    for (int b = 0; b < nbins; ++b) { hist0[b * 64 + t] = 0; hist1[b * 64 + t] = 0; }
    __syncwarp();
    const int uend = tw.count() * 4;
    // persistent CTAs claim (tile, quarter) units from a device counter: units on the edge of the band hold
    // few pairs, full ones 4096, so a static deal leaves a long tail
    for (;;) {
        __syncthreads();
        if (t == 0) s_unit = tw.j0 * 4 + (int)atomicAdd(unit_counter, 1ull);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= uend) break;
        const int2 rc = tw.tile(unit >> 2);
        const int r0 = rc.x * TILE + (unit & 3) * 32, c0 = rc.y * TILE;
        // 16 row pairs x 4 column groups = 64 items, split over the 2 warps
        for (int item = w; item < 64; item += 2) {
            const int i0 = r0 + (item >> 2) * 2, i1 = i0 + 1;
            const int k = c0 + (item & 3) * 32 + (int)lane;
            const bool v0 = i0 < n && k < n && k >= lo[min(i0, n - 1)] && k <= hi[min(i0, n - 1)];
            const bool v1 = i1 < n && k < n && k >= lo[min(i1, n - 1)] && k <= hi[min(i1, n - 1)];
            if (!__any_sync(0xffffffffu, v0 || v1)) continue;
            const uint32_t* colp = auxT + min((long long)k, npad - 1);
            const uint32_t* row0 = auxT + min(i0, n - 1);
            const uint32_t* row1 = auxT + min(i1, n - 1);
#pragma unroll 2
            for (int j = 0; j < words; ++j) {
                const uint32_t cw = __ldg(colp + (size_t)j * npad);
                const uint32_t a0 = __ldg(row0 + (size_t)j * npad);
                const uint32_t a1 = __ldg(row1 + (size_t)j * npad);
                const uint32_t m0 = max4_lt128(a0, cw) + bias0, m1 = max4_lt128(a1, cw) + bias1;
                hist_inc_dual<0>(m0, m1, tb);
                hist_inc_dual<1>(m0, m1, tb);
                hist_inc_dual<2>(m0, m1, tb);
                hist_inc_dual<3>(m0, m1, tb);
            }
            bool pass0 = false, pass1 = false;
            if (v0) {
                bool stopped = false;
                const StopHll stop{tau, e[i0], e[k], zs, order_n, AN};
                const double tu = selb::ertl_mle(hist0 + t, p_aux, 64, stop, &stopped);
                pass0 = !stopped && stop.crit(tu);
            }
            if (v1) {
                bool stopped = false;
                const StopHll stop{tau, e[i1], e[k], zs, order_n, AN};
                const double tu = selb::ertl_mle(hist1 + t, p_aux, 64, stop, &stopped);
                pass1 = !stopped && stop.crit(tu);
            }
            for (int b = 0; b < nbins; ++b) { hist0[b * 64 + t] = 0; hist1[b * 64 + t] = 0; }
            if (pass0) {
                const unsigned long long slot = warp_claim(pair_count);
                if (slot < pair_cap) pairs[slot] = make_uint2((uint32_t)i0, (uint32_t)k);
            }
            if (pass1) {
                const unsigned long long slot = warp_claim(pair_count);
                if (slot < pair_cap) pairs[slot] = make_uint2((uint32_t)i1, (uint32_t)k);
            }
        }
    }
}

// ============================================================================
// K4'' : hll_a / hll_an tile filter on BIT PLANES of the auxiliary sketches (p_aux >= 6).
// Same tile walk, same thread-per-pair shape (lane = column, row word = broadcast), same MLE + criterion
// as k_tile_filter_hll; the union histogram of a pair is built with the logic of k_pair_hist_planes
// (LOP3 borrow-chain max, 3+3-bit decode, carry-save counting) instead of 2^p_aux shared-memory
// read-modify-writes, and written once into the thread's shared-memory column for the estimator.
//   auxP[(plane*nw + w)*npad + g] : word w (32 registers) of a plane of the g-th sorted genome
//   agrange[g]                    : min | max<<8 of that genome's auxiliary registers
// The 32 pairs of a warp step share one 32-value window (their genomes sit within the CB band of each
// other, so their register ranges coincide); a step whose pairs do not fit one window takes the byte
// path of k_tile_filter_hll for its pairs.
// ============================================================================
#ifndef HLLP_MIN_CTAS
#define HLLP_MIN_CTAS 8
#endif

__global__ void __launch_bounds__(256)
k_aux_planes(const uint8_t* __restrict__ aux, const int32_t* __restrict__ order, long long n, long long npad,
             int p_aux, uint32_t* __restrict__ auxP) {
    const int nw = (1 << p_aux) >> 5;
    const long long total = n * nw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(idx / n);
        const long long g = idx - (long long)w * n;
        const uint4* src = reinterpret_cast<const uint4*>(aux + ((size_t)order[g] << p_aux) + (size_t)w * 32);
        const uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
        const uint32_t wd[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) m |= ((((wd[q] >> b) & 0x01010101u) * 0x10204080u) >> 28) << (4 * q);
            auxP[((size_t)b * nw + w) * npad + g] = m;
        }
    }
}

__global__ void __launch_bounds__(256)
k_aux_range(const uint8_t* __restrict__ aux, const int32_t* __restrict__ order, long long n, int p_aux,
            uint16_t* __restrict__ agrange) {
    // one warp per genome: smallest / largest register value
    const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= n) return;
    const uint8_t* row = aux + ((size_t)order[g] << p_aux);
    int vmin = 255, vmax = 0;
    for (int j = lane; j < (1 << p_aux); j += 32) { const int v = row[j]; vmin = min(vmin, v); vmax = max(vmax, v); }
    for (int o = 16; o; o >>= 1) {
        vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if (lane == 0) agrange[g] = (uint16_t)(min(vmin, vmax) | (vmax << 8));
}

// all word pairs of one (row, 32 columns) step for window G0: S / C2 end as the per-value carry-save state
template <int G0>
__device__ __forceinline__ void aux_plane_pairs(const uint32_t* __restrict__ rowp, const uint32_t* __restrict__ colp,
                                                long long npad, int nw, uint32_t gmask, uint32_t (&x)[32]) {
    uint32_t S[32], C2[32];
#pragma unroll
    for (int v = 0; v < 32; ++v) { S[v] = 0; C2[v] = 0; }
#pragma unroll 1
    for (int w = 0; w < nw; w += 2) {
        uint32_t M[2][6];
        {
            uint32_t a[2][6], b[2][6];
#pragma unroll
            for (int pl = 0; pl < 6; ++pl) {
                const size_t o0 = ((size_t)pl * nw + w) * (size_t)npad, o1 = o0 + (size_t)npad;
                a[0][pl] = __ldg(rowp + o0); a[1][pl] = __ldg(rowp + o1);
                b[0][pl] = __ldg(colp + o0); b[1][pl] = __ldg(colp + o1);
            }
            uint32_t lt0 = 0u, lt1 = 0u;
#pragma unroll
            for (int pl = 0; pl < 6; ++pl) {
                lt0 = lop3<0x8E>(a[0][pl], b[0][pl], lt0);
                lt1 = lop3<0x8E>(a[1][pl], b[1][pl], lt1);
            }
#pragma unroll
            for (int pl = 0; pl < 6; ++pl) {
                M[0][pl] = lop3<0xCA>(lt0, b[0][pl], a[0][pl]);
                M[1][pl] = lop3<0xCA>(lt1, b[1][pl], a[1][pl]);
            }
        }
        uint32_t L[2][8];
#pragma unroll
        for (int ws = 0; ws < 2; ++ws) {
            L[ws][0] = lop3<0x01>(M[ws][2], M[ws][1], M[ws][0]);
            L[ws][1] = lop3<0x02>(M[ws][2], M[ws][1], M[ws][0]);
            L[ws][2] = lop3<0x04>(M[ws][2], M[ws][1], M[ws][0]);
            L[ws][3] = lop3<0x08>(M[ws][2], M[ws][1], M[ws][0]);
            L[ws][4] = lop3<0x10>(M[ws][2], M[ws][1], M[ws][0]);
            L[ws][5] = lop3<0x20>(M[ws][2], M[ws][1], M[ws][0]);
            L[ws][6] = lop3<0x40>(M[ws][2], M[ws][1], M[ws][0]);
            L[ws][7] = lop3<0x80>(M[ws][2], M[ws][1], M[ws][0]);
        }
#define SELB_AUX_GROUP(T)                                                                                 \
        if (gmask & (1u << T)) {                                                                          \
            const uint32_t H0 = lop3<(1 << (G0 + T))>(M[0][5], M[0][4], M[0][3]);                         \
            const uint32_t H1 = lop3<(1 << (G0 + T))>(M[1][5], M[1][4], M[1][3]);                         \
            uint32_t m0[8], m1[8], kk[8];                                                                 \
            _Pragma("unroll") for (int j = 0; j < 8; ++j) { m0[j] = H0 & L[0][j]; m1[j] = H1 & L[1][j]; } \
            _Pragma("unroll") for (int j = 0; j < 8; ++j) kk[j] = lop3<0xE8>(S[T * 8 + j], m0[j], m1[j]); \
            _Pragma("unroll") for (int j = 0; j < 8; ++j) S[T * 8 + j] = lop3<0x96>(S[T * 8 + j], m0[j], m1[j]); \
            _Pragma("unroll") for (int j = 0; j < 8; ++j) C2[T * 8 + j] += __popc(kk[j]);                 \
        }
        SELB_AUX_GROUP(0)
        SELB_AUX_GROUP(1)
        SELB_AUX_GROUP(2)
        SELB_AUX_GROUP(3)
#undef SELB_AUX_GROUP
    }
#pragma unroll
    for (int v = 0; v < 32; ++v) x[v] = 2u * C2[v] + (uint32_t)__popc(S[v]);
}

template <int G0>
__device__ __forceinline__ void aux_plane_hist(const uint32_t* __restrict__ rowp, const uint32_t* __restrict__ colp,
                                               long long npad, int nw, uint32_t gmask, uint32_t* __restrict__ hcol,
                                               int nbins) {
    uint32_t x[32];
    aux_plane_pairs<G0>(rowp, colp, npad, nw, gmask, x);
    // the thread's histogram column: zeros outside the window, the counts inside
    for (int b = 0; b < 8 * G0; ++b) hcol[b * 64] = 0u;
#pragma unroll
    for (int v = 0; v < 32; ++v)
        if (8 * G0 + v < nbins) hcol[(8 * G0 + v) * 64] = x[v];
    for (int b = 8 * G0 + 32; b < nbins; ++b) hcol[b * 64] = 0u;
}

template <int AN>
__global__ void __launch_bounds__(64, HLLP_MIN_CTAS)
k_tile_filter_hll_planes(const uint32_t* __restrict__ auxP, const uint16_t* __restrict__ agrange,
                         const uint32_t* __restrict__ auxT, long long npad, int p_aux, TileWalk tw,
                         const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
                         const unsigned long long* __restrict__ e, double tau, float zs, int order_n,
                         uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count,
                         unsigned long long pair_cap, unsigned long long* __restrict__ unit_counter) {
    extern __shared__ __align__(1024) uint32_t hist_dyn[];   // [nbins][64 threads]
    __shared__ int s_unit;
    const int nbins = 64 - p_aux + 2;
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, tb = t * 4;
    const int nw = (1 << p_aux) >> 5;
    const int words = (1 << p_aux) >> 2;
    uint32_t* hcol = hist_dyn + t;
    const uint32_t bias0 = hist_bias(hist_dyn);
    const int uend = tw.count() * 4;
    for (;;) {
        __syncthreads();
        if (t == 0) s_unit = tw.j0 * 4 + (int)atomicAdd(unit_counter, 1ull);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= uend) break;
        const int2 rc = tw.tile(unit >> 2);
        const int r0 = rc.x * TILE + (unit & 3) * 32, c0 = rc.y * TILE;
        // 32 rows x 4 column groups = 128 steps, split over the 2 warps
        for (int item = (int)w; item < 128; item += 2) {
            const int i = r0 + (item >> 2);
            const int k = c0 + (item & 3) * 32 + (int)lane;
            if (i >= n) continue;
            const bool v = k < n && k >= lo[i] && k <= hi[i];
            if (!__any_sync(0xffffffffu, v)) continue;
            const int kc = (int)min((long long)k, npad - 1);
            // common value window of the step's pairs
            const uint32_t ra = agrange[i], rb = agrange[min(kc, n - 1)];
            int vlo = v ? max((int)(ra & 0xff), (int)(rb & 0xff)) : 255;
            int vhi = v ? max((int)(ra >> 8), (int)(rb >> 8)) : 0;
            for (int o = 16; o; o >>= 1) {
                vlo = min(vlo, __shfl_xor_sync(0xffffffffu, vlo, o));
                vhi = max(vhi, __shfl_xor_sync(0xffffffffu, vhi, o));
            }
            const int g0 = min(vlo >> 3, 4);
            if ((vhi >> 3) <= g0 + 3) {
                uint32_t gmask = 0;
                for (int tt = 0; tt < 4; ++tt)
                    if ((g0 + tt) >= (vlo >> 3) && (g0 + tt) <= (vhi >> 3)) gmask |= 1u << tt;
                const uint32_t* rowp = auxP + i;
                const uint32_t* colp = auxP + kc;
                switch (g0) {
                    case 0: aux_plane_hist<0>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    case 1: aux_plane_hist<1>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    case 2: aux_plane_hist<2>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    case 3: aux_plane_hist<3>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    default: aux_plane_hist<4>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                }
            } else {
                // register ranges too far apart for one window: byte path (shared-memory counters)
                for (int b = 0; b < nbins; ++b) hcol[b * 64] = 0u;
                const uint32_t* colp = auxT + kc;
                const uint32_t* row0 = auxT + i;
                for (int j = 0; j < words; ++j) {
                    const uint32_t m0 = max4_lt128(__ldg(row0 + (size_t)j * npad), __ldg(colp + (size_t)j * npad)) + bias0;
                    hist_inc2<0, 1>(m0, tb);
                    hist_inc2<2, 3>(m0, tb);
                }
            }
            bool pass = false;
            if (v) {
                bool stopped = false;
                const StopHll stop{tau, e[i], e[k], zs, order_n, AN};
                const double tu = selb::ertl_mle(hcol, p_aux, 64, stop, &stopped);
                pass = !stopped && stop.crit(tu);
            }
            if (pass) {
                const unsigned long long slot = warp_claim(pair_count);
                if (slot < pair_cap) pairs[slot] = make_uint2((uint32_t)i, (uint32_t)k);
            }
        }
    }
}

// ============================================================================
// K6: union estimate -> Jaccard -> tau test -> emit
//   reference: hll.h:1206 (calculate_estimate(counts, ERTL_MLE...)), selection.cpp:286-288
// ============================================================================
__global__ void __launch_bounds__(128)
k_estimate_emit(const uint32_t* __restrict__ hist, const uint2* __restrict__ pairs,
                const unsigned long long* __restrict__ npairs_dev, unsigned long long npairs_cap,
                const unsigned long long* __restrict__ e, int p, double tau,
                uint64_t* __restrict__ out_keys, double* __restrict__ out_j,
                unsigned long long* __restrict__ out_count, unsigned long long out_cap,
                uint64_t* __restrict__ near_keys, double* __restrict__ near_j,
                unsigned long long* __restrict__ near_count, unsigned long long near_cap) {
    // J is non-increasing in t: once it is below tau (and outside the near-tau window) at the
    // MLE's lower bound the pair can neither be emitted nor listed as near
    struct StopJ {
        double tau, slack;
        unsigned long long e1, e2;
        __device__ __forceinline__ bool operator()(double t_lb) const {
            return selb::jaccard(e1, e2, t_lb) < tau - slack;
        }
    };
    const long long npairs = (long long)min(*npairs_dev, npairs_cap);
    for (long long pi = blockIdx.x * (long long)blockDim.x + threadIdx.x; pi < npairs;
         pi += (long long)gridDim.x * blockDim.x) {
        const uint2 pr = pairs[pi];
        const unsigned long long e1 = e[pr.x], e2 = e[pr.y];
        bool stopped = false;
        const double t = selb::ertl_mle(hist + pi * 64, p, 1, StopJ{tau, 1e-6 * fabs(tau), e1, e2}, &stopped);
        if (stopped) continue;
        const double jac = selb::jaccard(e1, e2, t);
        const uint64_t key = ((uint64_t)pr.x << 32) | pr.y;
        if (jac >= tau) {
            const unsigned long long slot = warp_claim(out_count);
            if (slot < out_cap) { out_keys[slot] = key; out_j[slot] = jac; }
        }
        if (fabs(jac - tau) <= 1e-6 * fabs(tau)) {
            const unsigned long long slot = warp_claim(near_count);
            if (slot < near_cap) { near_keys[slot] = key; near_j[slot] = jac; }
        }
    }
}

// ============================================================================
// K7: (i,k) print order of the reference (selection.cpp:297-300) for SPARSE outputs: bucket by row
// (count -> scan -> scatter), then every element finds its place inside its row by counting the
// smaller columns.  Four small launches instead of the ~9 of a 49-bit radix sort; rows hold a
// handful of pairs (cluster mates), so the quadratic in-row step is a few loads per element.
// ============================================================================
__global__ void __launch_bounds__(256)
k_rowsort_count(const uint64_t* __restrict__ keys, long long cnt, int32_t* __restrict__ rowcnt) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e < cnt) atomicAdd(rowcnt + (keys[e] >> 32), 1);
}

__global__ void __launch_bounds__(256)
k_rowsort_scatter(const uint64_t* __restrict__ keys, const double* __restrict__ jac, long long cnt,
                  int32_t* __restrict__ rowcnt, const int32_t* __restrict__ rowoff,
                  uint64_t* __restrict__ tkeys, double* __restrict__ tj) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= cnt) return;
    const uint64_t key = keys[e];
    const uint32_t i = (uint32_t)(key >> 32);
    const int pos = rowoff[i] + atomicSub(rowcnt + i, 1) - 1;    // counts back down to zero
    tkeys[pos] = key;
    tj[pos] = jac[e];
}

__global__ void __launch_bounds__(256)
k_rowsort_rank(const uint64_t* __restrict__ tkeys, const double* __restrict__ tj, long long cnt,
               const int32_t* __restrict__ rowoff, uint64_t* __restrict__ out_keys, double* __restrict__ out_j) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= cnt) return;
    const uint64_t key = tkeys[e];
    const uint32_t i = (uint32_t)(key >> 32);
    const int a = rowoff[i], b = rowoff[i + 1];
    int r = 0;
    for (int t = a; t < b; ++t) r += tkeys[t] < key;      // keys are unique
    out_keys[a + r] = key;
    out_j[a + r] = tj[e];
}

// ============================================================================
// Peer-memory gather (multi-GPU, one process per GPU): every rank pushes its emitted (key, J)
// list straight into the ROOT GPU's landing zone with plain stores over NVLink/NVSwitch (the zone
// is mapped into each process with CUDA IPC).  One system-scope atomicAdd claims a contiguous
// block per rank, a second one signals completion; the root spins on its own memory until all
// ranks have signalled, then sorts the merged list.  No NCCL call and no host round trip between
// the emit kernel and the merged result (SURVEY.md §8e "gather of (i,k,J) lists to GPU 0").
//   landing zone: GatherHdr | keys[2][cap] | jac[2][cap] | near_keys[2][ncap] | near_j[2][ncap]
//   two buffers (epoch parity) so that a fast rank may already push run e+1 while the root still
//   merges run e; `consumed` stops it from getting two runs ahead.
// ============================================================================
struct GatherHdr {
    unsigned long long count[2];        // slots claimed per parity
    unsigned long long near_count[2];
    unsigned int done[2];               // ranks whose push is complete, per parity
    unsigned int consumed;              // runs the root has merged (monotone)
    unsigned int error;                 // 1: a wait timed out
    unsigned long long pad[26];
};
static_assert(sizeof(GatherHdr) == 256, "landing-zone header is 256 bytes");

struct GatherPush {                     // local to each rank
    unsigned long long base, near_base;
    unsigned int go, blocks_done;
};

struct GatherZone {                     // pointers into the (local or IPC-mapped) landing zone
    GatherHdr* hdr;
    uint64_t* keys;                     // [2][cap]
    double* jac;
    uint64_t* near_keys;                // [2][near_cap]
    double* near_j;
    unsigned long long cap, near_cap;
};

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long GATHER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

// one thread: (optionally) make sure the pass did not overflow, wait until the buffer of this parity
// has been merged by the root two runs ago, claim the rank's block in the root's lists
__global__ void k_gather_claim(GatherZone z, unsigned int epoch, unsigned long long* __restrict__ meta, int check,
                               unsigned long long cand_cap, unsigned long long pair_lim, unsigned long long out_cap,
                               unsigned long long tile_cap, unsigned long long near_cap_local,
                               GatherPush* __restrict__ st) {
    if (threadIdx.x | blockIdx.x) return;
    st->go = 0;
    st->blocks_done = 0;
    if (check && (meta[M_CAND] > cand_cap || meta[M_PAIRS] > pair_lim || meta[M_OUT] > out_cap ||
                  meta[M_TILES] > tile_cap))
        return;                              // the host redoes the pass and pushes afterwards
    if (epoch >= 2) {
        const unsigned long long t0 = gtime_ns();
        while (*(volatile unsigned int*)&z.hdr->consumed + 1u < epoch) {
            if (gtime_ns() - t0 > GATHER_TIMEOUT_NS) { meta[M_PUSHED] = 2; return; }
            __nanosleep(200);
        }
    }
    const unsigned int b = epoch & 1u;
    st->base = atomicAdd_system(&z.hdr->count[b], meta[M_OUT]);
    st->near_base = atomicAdd_system(&z.hdr->near_count[b], min(meta[M_NEAR], near_cap_local));
    st->go = 1;
    meta[M_PUSHED] = 1;
}

// all CTAs: copy the rank's lists into its block of the root's lists; the last CTA signals
__global__ void __launch_bounds__(256)
k_gather_copy(GatherZone z, unsigned int epoch, const unsigned long long* __restrict__ meta,
              unsigned long long near_cap_local, const uint64_t* __restrict__ keys, const double* __restrict__ jac,
              const uint64_t* __restrict__ near_keys, const double* __restrict__ near_j, GatherPush* __restrict__ st) {
    if (!st->go) return;
    const unsigned int b = epoch & 1u;
    const unsigned long long cnt = meta[M_OUT], ncnt = min(meta[M_NEAR], near_cap_local);
    const unsigned long long base = st->base, nbase = st->near_base;
    uint64_t* dk = z.keys + b * z.cap;
    double* dj = z.jac + b * z.cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cnt;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long slot = base + i;
        if (slot < z.cap) { dk[slot] = keys[i]; dj[slot] = jac[i]; }
    }
    uint64_t* nk = z.near_keys + b * z.near_cap;
    double* nj = z.near_j + b * z.near_cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < ncnt;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long slot = nbase + i;
        if (slot < z.near_cap) { nk[slot] = near_keys[i]; nj[slot] = near_j[i]; }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1) {
            st->blocks_done = 0;
            __threadfence_system();
            atomicAdd_system(&z.hdr->done[b], 1u);
        }
    }
}

// root: wait for every rank's signal, then publish the merged counts where the host can read them
__global__ void k_gather_wait(GatherZone z, unsigned int epoch, unsigned int world, const GatherPush* __restrict__ st,
                              unsigned long long* __restrict__ merged /* [count, near_count, error] */) {
    if (threadIdx.x | blockIdx.x) return;
    if (!st->go) { merged[2] = 2; return; }      // the root's own pass is being redone: nothing to wait for yet
    const unsigned int b = epoch & 1u;
    const unsigned long long t0 = gtime_ns();
    unsigned long long err = 0;
    while (*(volatile unsigned int*)&z.hdr->done[b] < world) {
        if (gtime_ns() - t0 > GATHER_TIMEOUT_NS) { err = 1; z.hdr->error = 1; break; }
        __nanosleep(100);
    }
    __threadfence_system();
    merged[0] = *(volatile unsigned long long*)&z.hdr->count[b];
    merged[1] = *(volatile unsigned long long*)&z.hdr->near_count[b];
    merged[2] = err;
}

// root, after the merge of this parity has been copied out: hand the buffer back
__global__ void k_gather_release(GatherZone z, unsigned int epoch) {
    if (threadIdx.x | blockIdx.x) return;
    const unsigned int b = epoch & 1u;
    z.hdr->count[b] = 0;
    z.hdr->near_count[b] = 0;
    z.hdr->done[b] = 0;
    __threadfence_system();
    *(volatile unsigned int*)&z.hdr->consumed = epoch + 1u;
}

// ============================================================================
// host side
// ============================================================================
template <typename T>
int upload(DevBuf& buf, const std::vector<T>& v, cudaStream_t s) {
    CKR(buf.ensure(v.size() * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return SELB200_OK;
}

// resident CTAs per SM with the maximum shared-memory carve-out (queried once per kernel)
template <typename K>
int resident_ctas(K kernel, int threads) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 8;
    }
    return per_sm;
}

template <class Src, class Epi>
int launch_pair_hist_t(cudaStream_t stream, int sm_count, const uint8_t* regs, size_t row_stride, size_t m, int p,
                       int64_t max_pairs, Src src, Epi epi) {
    if (max_pairs <= 0) return SELB200_OK;
    const int64_t ctas_needed = (max_pairs + 1) / 2;
    const int nbins = 64 - p + 2;
    if (nbins <= 52) {
        static const int per_sm = resident_ctas(k_pair_hist<52, Src, Epi>, 64);
        const int grid = (int)std::min<int64_t>(ctas_needed, (int64_t)sm_count * per_sm);
        k_pair_hist<52, Src, Epi><<<grid, 64, 0, stream>>>(regs, row_stride, m, src, epi);
    } else {
        static const int per_sm = resident_ctas(k_pair_hist<64, Src, Epi>, 64);
        const int grid = (int)std::min<int64_t>(ctas_needed, (int64_t)sm_count * per_sm);
        k_pair_hist<64, Src, Epi><<<grid, 64, 0, stream>>>(regs, row_stride, m, src, epi);
    }
    CK(cudaGetLastError());
    return SELB200_OK;
}

int launch_pair_hist(selb200_ctx* c, const uint8_t* regs, size_t m, int p, const int32_t* order,
                     const uint2* pairs, int64_t npairs, uint32_t* hist_out,
                     const unsigned long long* npairs_dev = nullptr) {
    if (npairs <= 0) return SELB200_OK;
    if (m < 512) return fail(SELB200_EINVAL, "primary sketches below 512 registers are not supported");
    SrcPairs src{pairs, order, (long long)npairs, npairs_dev};
    EpiWriteHist epi{hist_out};
    return launch_pair_hist_t(c->stream, c->sm_count, regs, m, m, p, npairs, src, epi);
}

// bit-plane union pass over the run's pair list (+ the byte kernel on whatever landed in the wide list)
int launch_pair_hist_planes(selb200_ctx* c, const uint2* pairs, int64_t max_pairs, uint32_t* hist_out,
                            const unsigned long long* npairs_dev, unsigned long long* wide_count, int* launches,
                            bool counters_are_zero) {
    if (max_pairs <= 0) return SELB200_OK;
    cudaStream_t s = c->stream;
    CKR(c->wide_list.ensure((size_t)max_pairs * 4));
    SrcPairs src{pairs, c->order_dev.as<int32_t>(), (long long)max_pairs, npairs_dev};
    EpiWriteHist epi{hist_out};
    const size_t smem = (size_t)PL_STAGES * 2 * 6 * (c->chunk_regs >> 3) + 8 * PL_STAGES;
    static int per_sm = 0;
    static size_t per_sm_smem = 0;
    if (!per_sm || per_sm_smem != smem) {
        cudaFuncSetAttribute(k_pair_hist_planes<EpiWriteHist>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pair_hist_planes<EpiWriteHist>, 32, smem) != cudaSuccess ||
            per_sm < 1) {
            cudaGetLastError();
            per_sm = 4;
        }
        per_sm_smem = smem;
    }
    // wide count, batch counter, kernel error word (adjacent words of meta[])
    if (!counters_are_zero) CK(cudaMemsetAsync(wide_count, 0, 24, s));
    const int grid = (int)std::min<int64_t>((max_pairs + 3) / 4, (int64_t)c->sm_count * per_sm);
    k_pair_hist_planes<EpiWriteHist><<<grid, 32, smem, s>>>(c->planes.as<uint32_t>(), c->m, c->chunk_regs,
                                                           c->grange.as<uint16_t>(), src, epi,
                                                           c->wide_list.as<uint32_t>(), wide_count, wide_count + 1);
    CK(cudaGetLastError());
    // pairs whose value range exceeds the 32-value window: byte kernel, small persistent grid
    SrcWide wsrc{pairs, c->order_dev.as<int32_t>(), c->wide_list.as<uint32_t>(), wide_count};
    if (c->m >= 512) {
        const int nbins = 64 - c->p + 2;
        const int wgrid = (int)std::min<int64_t>((max_pairs + 1) / 2, (int64_t)c->sm_count * 2);
        if (nbins <= 52) k_pair_hist<52, SrcWide, EpiWriteHist><<<wgrid, 64, 0, s>>>(c->d_regs, c->m, c->m, wsrc, epi);
        else k_pair_hist<64, SrcWide, EpiWriteHist><<<wgrid, 64, 0, s>>>(c->d_regs, c->m, c->m, wsrc, epi);
        CK(cudaGetLastError());
    }
    if (launches) *launches += 2;
    return SELB200_OK;
}

// ---------------------------------------------------------------------------------------------
// load = begin -> chunks -> end.  selb200_load_host / _device run all three over caller memory;
// the streaming entry points (selb200_load_begin / acquire / commit / end) let the caller decode
// sketch files straight into pinned staging slots while earlier chunks are already on the device.
// ---------------------------------------------------------------------------------------------
int load_begin(selb200_ctx* c, int64_t n, int p, int aux_kind, int aux_len, const uint8_t* d_regs_borrowed,
               const void* d_aux_borrowed) {
    if (!c) return fail(SELB200_EINVAL, "null context");
    c->loaded = false;
    c->ld.active = false;
    if (n < 0 || n > 0x7fffff00ll) return fail(SELB200_EINVAL, "n=%lld out of range", (long long)n);
    if (p < 9 || p > 20) return fail(SELB200_EINVAL, "primary HLL precision p=%d unsupported (9..20)", p);
    size_t aux_row_bytes = 0;
    if (aux_kind == SELB200_AUX_SMH) {
        if (aux_len < 1 || aux_len > 65536) return fail(SELB200_EINVAL, "smh bucket count %d out of range", aux_len);
        aux_row_bytes = (size_t)aux_len * 8;
    } else if (aux_kind == SELB200_AUX_HLL) {
        if (aux_len < 4 || aux_len > 14) return fail(SELB200_EINVAL, "aux HLL precision %d unsupported (4..14)", aux_len);
        aux_row_bytes = (size_t)1 << aux_len;
    } else if (aux_kind != SELB200_AUX_NONE) {
        return fail(SELB200_EINVAL, "unknown aux kind %d", aux_kind);
    }
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    c->n = n; c->p = p; c->m = (size_t)1 << p;
    c->aux_kind = aux_kind; c->aux_len = aux_len;
    c->npad = (n + TILE - 1) / TILE * TILE;
    c->h_cards_sorted.assign((size_t)n, 0.);
    c->h_order.assign((size_t)n, 0);
    c->h_e.assign((size_t)n, 0);
    c->out_count = c->near_count = 0;

    LoadState& L = c->ld;
    L = LoadState();
    L.aux_row_bytes = aux_row_bytes;
    L.regs_borrowed = d_regs_borrowed != nullptr;
    L.rows_per_chunk = std::max<int64_t>(1, (int64_t)(64u << 20) / (int64_t)c->m);
    L.active = true;
    if (n == 0) return SELB200_OK;
    if (L.regs_borrowed) {
        c->d_regs = d_regs_borrowed;
    } else {
        CKR(c->regs_own.ensure((size_t)n * c->m));
        c->d_regs = c->regs_own.as<uint8_t>();
    }
    if (aux_kind != SELB200_AUX_NONE) {
        if (d_aux_borrowed) {
            L.d_aux = d_aux_borrowed;
        } else {
            CKR(c->cand.ensure((size_t)n * aux_row_bytes));   // raw aux rows in file-list order (scratch)
            L.d_aux = c->cand.p;
        }
    }
    c->chunk_regs = (int)std::min<size_t>(c->m, (size_t)PL_CHUNK_REGS);
    CKR(c->planes.ensure((size_t)n * 6 * (c->m >> 3)));
    CKR(c->grange.ensure((size_t)n * sizeof(uint16_t)));
    CKR(c->hist.ensure((size_t)n * 64 * sizeof(uint32_t)));
    CKR(c->cards_in.ensure((size_t)n * sizeof(double)));
    CKR(c->out_j.ensure((size_t)n * sizeof(double)));        // stored value_ of each header (-1 = recompute)
    CKR(c->counters.ensure(64));
    CK(cudaMemsetAsync(c->counters.p, 0, 64, s));
    return SELB200_OK;
}

// the run stream waits for everything queued on the copy stream so far
int load_join_copies(selb200_ctx* c) {
    LoadState& L = c->ld;
    if (L.ev_i == c->copy_events.size()) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->copy_events.push_back(e);
    }
    CK(cudaEventRecord(c->copy_events[L.ev_i], c->copy_stream));
    CK(cudaStreamWaitEvent(c->stream, c->copy_events[L.ev_i], 0));
    ++L.ev_i;
    return SELB200_OK;
}

// rows [g0, g0+rows): optional H2D from host pointers (copy stream), then validation, per-genome
// histogram and cardinality MLE on the run stream
int load_chunk(selb200_ctx* c, int64_t g0, int64_t rows, const uint8_t* h_regs, const double* h_stored,
               const void* h_aux) {
    LoadState& L = c->ld;
    cudaStream_t s = c->stream;
    const int p = c->p;
    if (rows <= 0) return SELB200_OK;
    bool copied = false;
    if (h_aux && c->aux_kind != SELB200_AUX_NONE) {
        CK(cudaMemcpyAsync((uint8_t*)const_cast<void*>(L.d_aux) + (size_t)g0 * L.aux_row_bytes, h_aux,
                           (size_t)rows * L.aux_row_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        copied = true;
    }
    if (h_stored) {
        CK(cudaMemcpyAsync(c->out_j.as<double>() + g0, h_stored, (size_t)rows * 8, cudaMemcpyHostToDevice,
                           c->copy_stream));
        L.have_stored = true;
        copied = true;
    }
    if (h_regs) {
        CK(cudaMemcpyAsync(c->regs_own.as<uint8_t>() + (size_t)g0 * c->m, h_regs, (size_t)rows * c->m,
                           cudaMemcpyHostToDevice, c->copy_stream));
        copied = true;
    }
    if (copied) CKR(load_join_copies(c));
    const size_t n16 = (size_t)rows * c->m / 16;
    k_max_byte<<<(int)std::min<size_t>((n16 + 255) / 256, (size_t)c->sm_count * 8), 256, 0, s>>>(
        reinterpret_cast<const uint4*>(c->d_regs + (size_t)g0 * c->m), n16, c->counters.as<uint32_t>());
    CK(cudaGetLastError());
    SrcSelf src{(long long)g0, (long long)rows, c->counters.as<uint32_t>(), (uint32_t)(64 - p + 1)};
    EpiWriteHist epi{c->hist.as<uint32_t>() + (size_t)g0 * 64};
    CKR(launch_pair_hist_t(s, c->sm_count, c->d_regs, c->m, c->m, p, rows, src, epi));
    k_genome_cards<<<(unsigned)((rows + 127) / 128), 128, 0, s>>>(
        c->hist.as<uint32_t>() + (size_t)g0 * 64, h_stored ? c->out_j.as<double>() + g0 : nullptr, rows, p,
        c->cards_in.as<double>() + g0, c->counters.as<uint32_t>(), (uint32_t)(64 - p + 1),
        c->grange.as<uint16_t>() + g0);
    CK(cudaGetLastError());
    {   // bit-plane copy of the chunk for the union kernel
        const long long nblk = rows * (long long)(c->m >> 9);
        const int grid = (int)std::min<long long>((nblk + 7) / 8, (long long)c->sm_count * 16);
        k_planes_from_bytes<<<grid, 256, 0, s>>>(c->d_regs + (size_t)g0 * c->m, rows, c->m, c->chunk_regs,
                                                 c->planes.as<uint32_t>() + (size_t)g0 * 6 * (c->m >> 5));
        CK(cudaGetLastError());
    }
    L.rows_done += rows;
    return SELB200_OK;
}

int load_end(selb200_ctx* c) {
    LoadState& L = c->ld;
    if (!L.active) return fail(SELB200_ESTATE, "selb200_load_end without selb200_load_begin");
    const int64_t n = c->n;
    const int p = c->p, aux_kind = c->aux_kind, aux_len = c->aux_len;
    cudaStream_t s = c->stream;
    if (L.rows_done != n) {
        L.active = false;
        return fail(SELB200_ESTATE, "load ended after %lld of %lld rows", (long long)L.rows_done, (long long)n);
    }
    L.active = false;
    if (n == 0) { c->loaded = true; return SELB200_OK; }
    const void* d_aux = L.d_aux;
    const size_t aux_row_bytes = L.aux_row_bytes;
    if (aux_kind == SELB200_AUX_HLL) {
        const size_t n16 = (size_t)n * aux_row_bytes / 16;
        k_max_byte<<<(int)std::min<size_t>((n16 + 255) / 256, (size_t)c->sm_count * 8), 256, 0, s>>>(
            reinterpret_cast<const uint4*>(d_aux), n16, c->counters.as<uint32_t>() + 1);
        CK(cudaGetLastError());
    }
    // ---- sort by cardinality.  Distinct keys have ONE sorted order, so a device radix sort then
    // equals the reference's std::sort; any tie falls back to that exact std::sort on the host
    // (its unstable tie order depends on the whole sequence, selection.cpp:251-256). --------------
    CKR(c->order_dev.ensure((size_t)n * 4));
    CKR(c->e_sorted.ensure((size_t)n * 8));
    CKR(c->pairs.ensure((size_t)n * 4));          // iota scratch
    CKR(c->out_keys.ensure((size_t)n * 8));       // sorted cardinalities
    k_iota_i32<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->pairs.as<int32_t>(), n);
    CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->cards_in.as<double>(), c->out_keys.as<double>(),
                                       c->pairs.as<int32_t>(), c->order_dev.as<int32_t>(), (int)n, 0, 64, s));
    CKR(c->cub_tmp.ensure(tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp_bytes, c->cards_in.as<double>(), c->out_keys.as<double>(),
                                       c->pairs.as<int32_t>(), c->order_dev.as<int32_t>(), (int)n, 0, 64, s));
    k_sorted_prep<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->out_keys.as<double>(), n,
                                                               c->e_sorted.as<unsigned long long>(),
                                                               c->counters.as<uint32_t>() + 2);
    CK(cudaGetLastError());
    uint32_t h_flags[4] = {0, 0, 0, 0};   // max primary register, max aux register, tie flag
    CK(cudaMemcpyAsync(h_flags, c->counters.p, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c->h_cards_sorted.data(), c->out_keys.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c->h_order.data(), c->order_dev.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h_flags[0] > (uint32_t)(64 - p + 1))
        return fail(SELB200_EINVAL, "primary sketch holds register value %u > %u (= 64-p+1, p=%d): not an HLL of that precision",
                    h_flags[0], (uint32_t)(64 - p + 1), p);
    if (aux_kind == SELB200_AUX_HLL && h_flags[1] > (uint32_t)(64 - aux_len + 1))
        return fail(SELB200_EINVAL, "auxiliary sketch holds register value %u > %u (= 64-p+1, p=%d)", h_flags[1],
                    (uint32_t)(64 - aux_len + 1), aux_len);
    if (h_flags[2]) {
        std::vector<double> cards((size_t)n);
        CK(cudaMemcpyAsync(cards.data(), c->cards_in.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        selb200_sort_order(n, cards.data(), c->h_order.data());
        for (int64_t i = 0; i < n; ++i) c->h_cards_sorted[(size_t)i] = cards[(size_t)c->h_order[(size_t)i]];
        CKR(upload(c->order_dev, c->h_order, s));
    }
    for (int64_t i = 0; i < n; ++i)
        c->h_e[(size_t)i] = (uint64_t)(size_t)c->h_cards_sorted[(size_t)i];   // size_t e = card (selection.cpp:275,280)
    if (h_flags[2]) CKR(upload(c->e_sorted, c->h_e, s));

    if (aux_kind == SELB200_AUX_SMH) {
        CKR(c->aux_sorted.ensure((size_t)n * aux_row_bytes));
        const int row_words = (int)(aux_row_bytes / 4);
        const int grid = (int)std::min<int64_t>((n * row_words + 255) / 256, (int64_t)c->sm_count * 16);
        k_gather_rows<<<grid, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(d_aux), c->order_dev.as<int32_t>(), n,
                                           row_words, c->aux_sorted.as<uint32_t>());
        CK(cudaGetLastError());
    } else if (aux_kind == SELB200_AUX_HLL) {
        const int row_words = (int)(aux_row_bytes / 4);
        CKR(c->auxT.ensure((size_t)row_words * c->npad * 4));
        CK(cudaMemsetAsync(c->auxT.p, 0, (size_t)row_words * c->npad * 4, s));
        const int grid = (int)std::min<int64_t>((n * row_words + 255) / 256, (int64_t)c->sm_count * 16);
        k_aux_transpose<<<grid, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(d_aux), c->order_dev.as<int32_t>(), n,
                                             c->npad, row_words, c->auxT.as<uint32_t>());
        CK(cudaGetLastError());
        if (aux_len >= 6) {       // bit planes for k_tile_filter_hll_planes (two 32-register words per step)
            const int nw = (1 << aux_len) >> 5;
            CKR(c->auxP.ensure((size_t)6 * nw * c->npad * 4));
            CKR(c->agrange.ensure((size_t)c->npad * sizeof(uint16_t)));
            CK(cudaMemsetAsync(c->auxP.p, 0, (size_t)6 * nw * c->npad * 4, s));
            CK(cudaMemsetAsync(c->agrange.p, 0, (size_t)c->npad * sizeof(uint16_t), s));
            const int g2 = (int)std::min<int64_t>((n * nw + 255) / 256, (int64_t)c->sm_count * 16);
            k_aux_planes<<<g2, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_aux), c->order_dev.as<int32_t>(), n, c->npad,
                                            aux_len, c->auxP.as<uint32_t>());
            CK(cudaGetLastError());
            k_aux_range<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_aux),
                                                                         c->order_dev.as<int32_t>(), n, aux_len,
                                                                         c->agrange.as<uint16_t>());
            CK(cudaGetLastError());
        }
    }
    CK(cudaStreamSynchronize(s));
    c->loaded = true;
    return SELB200_OK;
}

int do_load(selb200_ctx* c, int64_t n, int p, const uint8_t* regs, bool on_device, const double* stored,
            int aux_kind, int aux_len, const void* aux) {
    if (n > 0 && !regs) return fail(SELB200_EINVAL, "null register matrix");
    if (aux_kind != SELB200_AUX_NONE && n > 0 && !aux) return fail(SELB200_EINVAL, "null aux matrix");
    CKR(load_begin(c, n, p, aux_kind, aux_len, on_device ? regs : nullptr, on_device ? aux : nullptr));
    LoadState& L = c->ld;
    const int64_t step = on_device ? std::max<int64_t>(n, 1) : L.rows_per_chunk;
    for (int64_t g0 = 0; g0 < n; g0 += step) {
        const int64_t rows = std::min(step, n - g0);
        CKR(load_chunk(c, g0, rows, on_device ? nullptr : regs + (size_t)g0 * c->m, stored ? stored + g0 : nullptr,
                       (on_device || !aux) ? nullptr : (const uint8_t*)aux + (size_t)g0 * L.aux_row_bytes));
    }
    return load_end(c);
}

}  // namespace

// ============================================================================
// C-ABI
// ============================================================================
extern "C" {

int selb200_abi_version(void) { return SELB200_ABI_VERSION; }
const char* selb200_last_error(void) { return g_err.c_str(); }

int selb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int selb200_create(int device, void* stream, selb200_ctx** out) {
    if (!out) return fail(SELB200_EINVAL, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SELB200_ECUDA, "no CUDA device available (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return fail(SELB200_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    CK(cudaSetDevice(device));
    selb200_ctx* c = new selb200_ctx();
    c->device = device;
    if (stream) {
        c->stream = reinterpret_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return fail(SELB200_ECUDA, "cudaStreamCreate failed");
        }
        c->own_stream = true;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return fail(SELB200_ECUDA, "cudaStreamCreate failed");
    }
    *out = c;
    return SELB200_OK;
}

void selb200_destroy(selb200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->regs_own, &c->aux_sorted, &c->auxT, &c->cards_in, &c->e_sorted, &c->order_dev,
                      &c->lo, &c->hi, &c->tile_prefix, &c->tile_cb0, &c->tile_rc, &c->sigT, &c->cand, &c->pairs, &c->hist,
                      &c->counters, &c->cub_tmp, &c->out_keys, &c->out_j, &c->out_keys2, &c->out_j2,
                      &c->near_keys, &c->near_j, &c->tile_nt, &c->rb_pairs, &c->g_push, &c->g_merged, &c->row_cnt, &c->row_off, &c->sort_tmp, &c->planes, &c->grange, &c->wide_list, &c->auxP, &c->agrange};
    for (DevBuf* b : bufs) b->release();
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    selb200_gather_close(c);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->h_snap) cudaFreeHost(c->h_snap);
    for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
    for (StageSlot& sl : c->slots) {
        if (sl.regs) cudaFreeHost(sl.regs);
        if (sl.aux) cudaFreeHost(sl.aux);
        if (sl.stored) cudaFreeHost(sl.stored);
        if (sl.free_ev) cudaEventDestroy(sl.free_ev);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int selb200_load_host(selb200_ctx* ctx, int64_t n, int p, const uint8_t* regs, const double* stored,
                      int aux_kind, int aux_len, const void* aux) {
    return do_load(ctx, n, p, regs, false, stored, aux_kind, aux_len, aux);
}

int selb200_load_device(selb200_ctx* ctx, int64_t n, int p, const uint8_t* d_regs, const double* stored_host,
                        int aux_kind, int aux_len, const void* d_aux) {
    return do_load(ctx, n, p, d_regs, true, stored_host, aux_kind, aux_len, d_aux);
}

// device-resident matrices that become complete piece by piece (in the order of the context's stream)
int selb200_load_device_begin(selb200_ctx* ctx, int64_t n, int p, const uint8_t* d_regs, int aux_kind, int aux_len,
                              const void* d_aux) {
    if (n > 0 && !d_regs) return fail(SELB200_EINVAL, "null register matrix");
    if (aux_kind != SELB200_AUX_NONE && n > 0 && !d_aux) return fail(SELB200_EINVAL, "null aux matrix");
    return load_begin(ctx, n, p, aux_kind, aux_len, d_regs, d_aux);
}

int selb200_load_device_rows(selb200_ctx* c, int64_t g0, int64_t count) {
    if (!c || !c->ld.active || !c->ld.regs_borrowed) return fail(SELB200_ESTATE, "selb200_load_device_rows outside load_device_begin/load_end");
    if (g0 < 0 || count < 0 || g0 + count > c->n) return fail(SELB200_EINVAL, "rows [%lld,+%lld) outside the matrix", (long long)g0, (long long)count);
    CK(cudaSetDevice(c->device));
    return load_chunk(c, g0, count, nullptr, nullptr, nullptr);
}

int selb200_load_begin(selb200_ctx* ctx, int64_t n, int p, int aux_kind, int aux_len, int64_t* rows_per_chunk) {
    CKR(load_begin(ctx, n, p, aux_kind, aux_len, nullptr, nullptr));
    if (rows_per_chunk) *rows_per_chunk = ctx->ld.rows_per_chunk;
    return SELB200_OK;
}

static int pinned_ensure(void** ptr, size_t* cap, size_t bytes) {
    if (bytes <= *cap && *ptr) return SELB200_OK;
    if (*ptr) cudaFreeHost(*ptr);
    *ptr = nullptr;
    *cap = 0;
    CK(cudaMallocHost(ptr, bytes ? bytes : 16));
    *cap = bytes;
    return SELB200_OK;
}

int selb200_load_acquire(selb200_ctx* c, int64_t g0, int64_t count, uint8_t** regs, double** stored, void** aux) {
    if (!c || !c->ld.active) return fail(SELB200_ESTATE, "selb200_load_acquire outside begin/end");
    LoadState& L = c->ld;
    if (L.acq_slot >= 0) return fail(SELB200_ESTATE, "previous staging slot not committed");
    if (g0 != L.rows_done || count < 1 || count > L.rows_per_chunk || g0 + count > c->n)
        return fail(SELB200_EINVAL, "rows [%lld,+%lld) out of sequence (next row %lld, chunk limit %lld)",
                    (long long)g0, (long long)count, (long long)L.rows_done, (long long)L.rows_per_chunk);
    CK(cudaSetDevice(c->device));
    StageSlot& sl = c->slots[c->next_slot];
    if (!sl.free_ev) CK(cudaEventCreateWithFlags(&sl.free_ev, cudaEventDisableTiming));
    if (sl.in_flight) { CK(cudaEventSynchronize(sl.free_ev)); sl.in_flight = false; }   // its last H2D finished
    CKR(pinned_ensure((void**)&sl.regs, &sl.regs_cap, (size_t)L.rows_per_chunk * c->m));
    CKR(pinned_ensure((void**)&sl.stored, &sl.stored_cap, (size_t)L.rows_per_chunk * 8));
    if (L.aux_row_bytes) CKR(pinned_ensure((void**)&sl.aux, &sl.aux_cap, (size_t)L.rows_per_chunk * L.aux_row_bytes));
    L.acq_slot = c->next_slot;
    L.acq_g0 = g0;
    L.acq_rows = count;
    if (regs) *regs = sl.regs;
    if (stored) *stored = sl.stored;
    if (aux) *aux = L.aux_row_bytes ? (void*)sl.aux : nullptr;
    return SELB200_OK;
}

int selb200_load_commit(selb200_ctx* c) {
    if (!c || !c->ld.active || c->ld.acq_slot < 0) return fail(SELB200_ESTATE, "nothing acquired to commit");
    LoadState& L = c->ld;
    StageSlot& sl = c->slots[L.acq_slot];
    CK(cudaSetDevice(c->device));
    CKR(load_chunk(c, L.acq_g0, L.acq_rows, sl.regs, sl.stored, L.aux_row_bytes ? sl.aux : nullptr));
    CK(cudaEventRecord(sl.free_ev, c->copy_stream));
    sl.in_flight = true;
    c->next_slot = (c->next_slot + 1) % 3;
    L.acq_slot = -1;
    return SELB200_OK;
}

int selb200_load_end(selb200_ctx* c) {
    if (!c) return fail(SELB200_EINVAL, "null context");
    if (c->ld.acq_slot >= 0) return fail(SELB200_ESTATE, "a staging slot is still acquired");
    return load_end(c);
}

int selb200_get_order(selb200_ctx* c, double* cards_sorted, int32_t* order) {
    if (!c || !c->loaded) return fail(SELB200_ESTATE, "no sketches loaded");
    if (cards_sorted) std::memcpy(cards_sorted, c->h_cards_sorted.data(), (size_t)c->n * sizeof(double));
    if (order) std::memcpy(order, c->h_order.data(), (size_t)c->n * sizeof(int32_t));
    return SELB200_OK;
}

void selb200_default_params(selb200_params* p) {
    std::memset(p, 0, sizeof *p);
    p->tau = 0.9f;
    p->criterion = SELB200_CRIT_SMH_A;
    p->z_score = 1.96f;
    p->order_n = 1;
    p->n_shards = 1;
    p->sort_output = 1;
}

int selb200_band_params(int m, float tau, int cpu_variant, int* n_bands, int* n_rows) {
    if (m < 1 || !n_bands || !n_rows) return fail(SELB200_EINVAL, "bad band-search arguments");
    int nb = 1, nr = 1;
    for (int band = 1; band <= m; ++band) {
        if (m % band != 0) continue;
        if (cpu_variant) { nb = band; nr = m / band; }
        const double inner = std::pow((double)tau, (double)((float)m / (float)band));
        const float P_r = (float)(1.0 - std::pow(1.0 - inner, (double)(float)band));
        if (P_r >= 0.95) {
            if (!cpu_variant) { nb = band; nr = m / band; }
            break;
        }
    }
    *n_bands = nb;
    *n_rows = nr;
    return SELB200_OK;
}

int selb200_sort_order(int64_t n, const double* cards, int32_t* order) {
    if (n < 0 || (n && (!cards || !order))) return fail(SELB200_EINVAL, "bad sort arguments");
    std::vector<std::pair<int32_t, double>> v((size_t)n);
    for (int64_t i = 0; i < n; ++i) v[(size_t)i] = {(int32_t)i, cards[i]};
    std::sort(v.begin(), v.end(),
              [](const std::pair<int32_t, double>& x, const std::pair<int32_t, double>& y) { return x.second < y.second; });
    for (int64_t i = 0; i < n; ++i) order[i] = v[(size_t)i].first;
    return SELB200_OK;
}

int selb200_run(selb200_ctx* c, const selb200_params* prm, selb200_stats* st_out) {
    if (!c || !prm) return fail(SELB200_EINVAL, "null argument");
    if (!c->loaded) return fail(SELB200_ESTATE, "selb200_run before a successful load");
    const int crit = prm->criterion;
    if (crit < SELB200_CRIT_CB || crit > SELB200_CRIT_HLL_AN) return fail(SELB200_EINVAL, "unknown criterion %d", crit);
    if (crit == SELB200_CRIT_SMH_A && c->aux_kind != SELB200_AUX_SMH)
        return fail(SELB200_ESTATE, "criterion smh_a needs SuperMinHash auxiliary sketches");
    if ((crit == SELB200_CRIT_HLL_A || crit == SELB200_CRIT_HLL_AN) && c->aux_kind != SELB200_AUX_HLL)
        return fail(SELB200_ESTATE, "criterion hll_a/hll_an needs auxiliary HLL sketches");
    const int n_shards = prm->n_shards > 0 ? prm->n_shards : 1;
    if (prm->shard < 0 || prm->shard >= n_shards) return fail(SELB200_EINVAL, "shard %d of %d", prm->shard, n_shards);
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    selb200_stats st;
    std::memset(&st, 0, sizeof st);
    const int n = (int)c->n;
    st.n = n;
    st.pairs_total = (int64_t)n * (n - 1) / 2;
    c->out_count = c->near_count = 0;
    c->res_keys = nullptr; c->res_j = nullptr;
    c->ev_used = 0;
    if (n < 2) { if (st_out) *st_out = st; return SELB200_OK; }
    const double tau = (double)prm->tau;

    int n_rows = prm->n_rows, n_bands = prm->n_bands;
    bool smh_shape_ok = true;
    if (crit == SELB200_CRIT_SMH_A) {
        if (n_rows <= 0 || n_bands <= 0) selb200_band_params(c->aux_len, prm->tau, 1, &n_bands, &n_rows);
        // criteria_sketch.hpp:67-70: a shape that does not tile the sketch selects nothing
        smh_shape_ok = (int64_t)n_rows * n_bands == c->aux_len;
        st.n_rows = n_rows; st.n_bands = n_bands;
    }

    // SELB200_UNION=bytes selects the shared-memory byte kernel for the union pass (A/B measurements)
    static const bool union_bytes = [] { const char* e = getenv("SELB200_UNION"); return e && !strcmp(e, "bytes"); }();
    const bool gather = prm->gather != 0;
    if (gather && !c->g.attached) return fail(SELB200_ESTATE, "params.gather set without selb200_gather_attach");
    if (gather && (prm->shard != c->g.rank || n_shards != c->g.world))
        return fail(SELB200_EINVAL, "gather: shard %d/%d does not match the attached rank %d/%d", prm->shard, n_shards,
                    c->g.rank, c->g.world);

    // meta[]: see the M_* enum (device counters of the sync-free pipeline)
    CKR(c->counters.ensure(M_WORDS * 8));
    CK(cudaMemsetAsync(c->counters.p, 0, M_WORDS * 8, s));
    unsigned long long* d_cnt = c->counters.as<unsigned long long>();

    cudaEvent_t ev_begin = c->ev();
    // ---- K2: CB band + tile list, all on the device ------------------------------------------
    int zeros = 0;
    while (zeros < n && c->h_e[(size_t)zeros] == 0) ++zeros;
    const int nrb = (n + TILE - 1) / TILE;
    CKR(c->lo.ensure((size_t)n * 4));
    CKR(c->hi.ensure((size_t)n * 4));
    CKR(c->tile_nt.ensure(((size_t)nrb + 1) * 4));
    CKR(c->tile_prefix.ensure(((size_t)nrb + 1) * 4));
    CKR(c->tile_cb0.ensure((size_t)nrb * 4));
    CKR(c->rb_pairs.ensure((size_t)nrb * 8));
    // a band can never hold more tiles than the block triangle; 16 Mi tiles (128 MB) to start with
    // for inputs beyond n = 724k, grown on demand like every other list
    const int64_t tri = (int64_t)nrb * (nrb + 1) / 2;
    if (c->tile_cap < std::min<int64_t>(tri, 16ll << 20)) c->tile_cap = std::min<int64_t>(tri, 16ll << 20);
    // no_cb: every ratio passes a bound of -inf, so each row's range is (i, n-1] minus the e==0 columns
    const double tau_cb = prm->no_cb ? -__builtin_huge_val() : tau;
    k_cb_bounds<<<(n + 255) / 256, 256, 0, s>>>(c->e_sorted.as<unsigned long long>(), n, zeros, tau_cb,
                                                c->lo.as<int32_t>(), c->hi.as<int32_t>());
    CK(cudaGetLastError());
    k_rowblock_span<<<(nrb + 1 + 3) / 4, 128, 0, s>>>(c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, nrb,
                                                      c->tile_nt.as<int32_t>(), c->tile_cb0.as<int32_t>(),
                                                      c->rb_pairs.as<unsigned long long>(), d_cnt);
    CK(cudaGetLastError());
    {
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->tile_nt.as<int32_t>(), c->tile_prefix.as<int32_t>(),
                                         nrb + 1, s));
        CKR(c->cub_tmp.ensure(tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp_bytes, c->tile_nt.as<int32_t>(),
                                         c->tile_prefix.as<int32_t>(), nrb + 1, s));
    }
    st.launches += 3;
    DBG_SYNC(c, "cb bounds + row-block spans + scan");
    cudaEvent_t ev_bounds = nullptr;

    // ---- K3: signatures ---------------------------------------------------------------
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> t_filter, t_verify, t_union, t_est;
    const int n_words = (n_bands + 1) / 2;
    const bool use_smh = crit == SELB200_CRIT_SMH_A && smh_shape_ok;

    CKR(c->cand.ensure((size_t)PAIR_CAP * sizeof(uint2)));
    CKR(c->pairs.ensure((size_t)PAIR_CAP * sizeof(uint2)));
    CKR(c->near_keys.ensure((size_t)(1 << 16) * 8));
    CKR(c->near_j.ensure((size_t)(1 << 16) * 8));
    const unsigned long long near_cap = 1ull << 16;
    const float zs = prm->z_score * (crit >= SELB200_CRIT_HLL_A ? selb::sigma_p(c->aux_len) : 0.f);
    size_t hll_smem = 0;
    int hll_grid = 0;
    // SELB200_HLLFILTER=bytes keeps the shared-memory-counter filter (A/B measurements); sketches below 64
    // registers have no bit planes
    static const bool hll_bytes_env = [] { const char* e = getenv("SELB200_HLLFILTER"); return e && !strcmp(e, "bytes"); }();
    const bool hll_planes = crit >= SELB200_CRIT_HLL_A && c->aux_len >= 6 && !hll_bytes_env;
    if (crit >= SELB200_CRIT_HLL_A) {
        hll_smem = (size_t)(hll_planes ? 1 : 2) * (64 - c->aux_len + 2) * 64 * sizeof(uint32_t);
        static bool carve = false;
        if (!carve) {
            cudaFuncSetAttribute(k_tile_filter_hll<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_tile_filter_hll<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_tile_filter_hll_planes<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_tile_filter_hll_planes<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carve = true;
        }
        int per_sm = 0;
        const cudaError_t oe = hll_planes
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_filter_hll_planes<0>, 64, hll_smem)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_filter_hll<0>, 64, hll_smem);
        if (oe != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 4;
        }
        hll_grid = c->sm_count * per_sm;
    }
    static const int smh_grid_per_sm = [] {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_filter_smh, 256, 0) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = FILTER_CTAS_PER_SM;
        }
        return per_sm;
    }();

    // ---- filter -> union passes over tile ranges ---------------------------------------
    // Optimistic, sync-free pipeline: the tile list, its length and every later work count live in
    // device memory (persistent grids), list capacities are fixed up front, and the counters of
    // each range are snapshotted into pinned host memory.  ONE synchronisation at the end checks
    // the snapshots; an overflow (rare: a filter far less selective than the capacities assume)
    // grows the buffers / halves the offending range and the whole pass is redone.
    // Ranges index j: the shard's j-th tile is tile shard + j*n_shards of the row-major list — dealing
    // tiles round-robin gives every shard the same mix of band positions (contiguous slices measured
    // 1.6x imbalance in the union pass at 8 shards).  {0, INT32_MAX} = "all of them".
    std::vector<std::pair<int, int>> ranges;
    if (crit == SELB200_CRIT_CB) c->hist_cap_pairs = PAIR_CAP;
    if (c->hist_cap_pairs < (1ll << 20)) c->hist_cap_pairs = 1ll << 20;
    if (c->out_cap < (1ll << 21)) c->out_cap = 1ll << 21;
    if (!c->h_snap) CK(cudaMallocHost(&c->h_snap, (SNAP_MAX * 4 + M_WORDS) * sizeof(unsigned long long)));
    unsigned long long* h_fin = c->h_snap + SNAP_MAX * 4;     // the final meta[] block
    std::memset(h_fin, 0, M_WORDS * sizeof(unsigned long long));
    c->h_tprefix.resize((size_t)nrb + 1);
    c->h_rb_pairs.resize((size_t)nrb);
    const int launches_fixed = st.launches;
    GatherZone gz{};
    if (gather) {
        uint8_t* zb = (uint8_t*)c->g.zone;
        gz.hdr = (GatherHdr*)zb;
        gz.cap = (unsigned long long)c->g.cap;
        gz.near_cap = (unsigned long long)c->g.near_cap;
        gz.keys = (uint64_t*)(zb + sizeof(GatherHdr));
        gz.jac = (double*)(gz.keys + 2 * gz.cap);
        gz.near_keys = (uint64_t*)(gz.jac + 2 * gz.cap);
        gz.near_j = (double*)(gz.near_keys + 2 * gz.near_cap);
        CKR(c->g_push.ensure(sizeof(GatherPush)));
        CKR(c->g_merged.ensure(32));
    }
    auto launch_push = [&](int check, unsigned long long pair_lim) -> int {
        k_gather_claim<<<1, 32, 0, s>>>(gz, c->g.epoch, d_cnt, check, (unsigned long long)PAIR_CAP, pair_lim,
                                        (unsigned long long)c->out_cap, (unsigned long long)c->tile_cap, near_cap,
                                        c->g_push.as<GatherPush>());
        CK(cudaGetLastError());
        k_gather_copy<<<32, 256, 0, s>>>(gz, c->g.epoch, d_cnt, near_cap, c->out_keys.as<uint64_t>(),
                                         c->out_j.as<double>(), c->near_keys.as<uint64_t>(), c->near_j.as<double>(),
                                         c->g_push.as<GatherPush>());
        CK(cudaGetLastError());
        st.launches += 2;
        return SELB200_OK;
    };
    int64_t tiles_total = -1;      // host copy of meta[M_TILES] once known
    auto shard_tiles = [&](int64_t total) -> int {
        return total > prm->shard ? (int)((total - prm->shard + n_shards - 1) / n_shards) : 0;
    };
    bool pushed = false;
    for (int attempt = 0;; ++attempt) {
        if (attempt > 40) return fail(SELB200_ENOMEM, "candidate lists keep overflowing");
        CKR(c->tile_rc.ensure((size_t)std::max<int64_t>(c->tile_cap, 1) * sizeof(int2)));
        CKR(c->hist.ensure((size_t)c->hist_cap_pairs * 64 * sizeof(uint32_t)));
        CKR(c->out_keys.ensure((size_t)c->out_cap * 8));
        CKR(c->out_j.ensure((size_t)c->out_cap * 8));
        st.launches = launches_fixed;
        k_tile_table<<<(nrb + 3) / 4, 128, 0, s>>>(c->tile_prefix.as<int32_t>(), c->tile_cb0.as<int32_t>(), nrb,
                                                   (long long)c->tile_cap, c->tile_rc.as<int2>(), d_cnt);
        CK(cudaGetLastError());
        st.launches++;
        DBG_SYNC(c, "tile table");
        if (!ev_bounds) ev_bounds = c->ev();
        if (ranges.empty()) {
            if (crit == SELB200_CRIT_CB) {
                // CB-only fills whole tiles: its ranges are cut on the host to what the pair list is sure to hold
                CK(cudaMemcpyAsync(h_fin + M_TILES, d_cnt + M_TILES, 8, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                tiles_total = (int64_t)h_fin[M_TILES];
                const int t_end = shard_tiles(std::min<int64_t>(tiles_total, c->tile_cap));
                const int step = (int)(PAIR_CAP / (TILE * TILE));
                for (int a0 = 0; a0 < t_end; a0 += step) ranges.push_back({a0, std::min(t_end, a0 + step)});
            } else if (smh_shape_ok) {
                ranges.push_back({0, INT32_MAX});
            }
        }
        if ((int)ranges.size() > SNAP_MAX) return fail(SELB200_ENOMEM, "too many tile ranges (%zu)", ranges.size());
        t_filter.clear(); t_verify.clear(); t_union.clear(); t_est.clear();
        if (use_smh) {
            cudaEvent_t a0 = c->ev();
            const size_t sig_bytes = (size_t)n_words * c->npad * 4;
            CKR(c->sigT.ensure(2 * sig_bytes));
            const int grid = (int)std::min<int64_t>(((int64_t)c->npad * n_words * 2 + 255) / 256, (int64_t)c->sm_count * 16);
            k_smh_signatures<<<grid, 256, 0, s>>>(c->aux_sorted.as<uint64_t>(), n, c->npad, c->aux_len, n_rows, n_bands,
                                                  c->sigT.as<uint32_t>(), c->sigT.as<uint32_t>() + (size_t)n_words * c->npad);
            CK(cudaGetLastError());
            st.launches++;
            DBG_SYNC(c, "smh signatures");
            t_filter.push_back({a0, c->ev()});
        }
        if (attempt > 0) {          // the memset at the top of the run covers the first attempt
            CK(cudaMemsetAsync(d_cnt, 0, 32, s));
            CK(cudaMemsetAsync(d_cnt + M_PUSHED, 0, 8, s));
        }
        const unsigned long long pair_lim = (unsigned long long)std::min<int64_t>(PAIR_CAP, c->hist_cap_pairs);
        for (size_t ri = 0; ri < ranges.size(); ++ri) {
            const std::pair<int, int> rg = ranges[ri];
            const int64_t nt = (int64_t)rg.second - rg.first;     // upper bound when the end is open
            const TileWalk tw{c->tile_rc.as<int2>(), d_cnt, (long long)c->tile_cap, prm->shard, n_shards, rg.first, rg.second};
            if (attempt > 0 || ri > 0) {
                CK(cudaMemsetAsync(d_cnt, 0, 16, s));   // candidates + pairs of this range
                if (crit >= SELB200_CRIT_HLL_A) CK(cudaMemsetAsync(d_cnt + M_UNIT, 0, 8, s));
            }
            cudaEvent_t f0 = c->ev();
            if (crit == SELB200_CRIT_SMH_A) {
                const int grid = (int)std::min<int64_t>(nt, (int64_t)c->sm_count * smh_grid_per_sm);
                k_tile_filter_smh<<<grid, 256, 0, s>>>(
                    c->sigT.as<uint32_t>(), c->sigT.as<uint32_t>() + (size_t)n_words * c->npad, c->npad, n_words, tw,
                    c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->cand.as<uint2>(), d_cnt + M_CAND,
                    (unsigned long long)PAIR_CAP);
            } else if (crit == SELB200_CRIT_CB) {
                const int grid = (int)std::min<int64_t>(nt, (int64_t)c->sm_count * 8);
                k_tile_enum<<<grid, 256, 0, s>>>(tw, c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->pairs.as<uint2>(),
                                                 d_cnt + M_PAIRS, (unsigned long long)PAIR_CAP);
            } else if (hll_planes) {
                const int grid = (int)std::min<int64_t>(nt * 4, (int64_t)hll_grid);
                if (crit == SELB200_CRIT_HLL_A)
                    k_tile_filter_hll_planes<0><<<grid, 64, hll_smem, s>>>(
                        c->auxP.as<uint32_t>(), c->agrange.as<uint16_t>(), c->auxT.as<uint32_t>(), c->npad, c->aux_len, tw,
                        c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->e_sorted.as<unsigned long long>(), tau, zs,
                        prm->order_n, c->pairs.as<uint2>(), d_cnt + M_PAIRS, (unsigned long long)PAIR_CAP, d_cnt + M_UNIT);
                else
                    k_tile_filter_hll_planes<1><<<grid, 64, hll_smem, s>>>(
                        c->auxP.as<uint32_t>(), c->agrange.as<uint16_t>(), c->auxT.as<uint32_t>(), c->npad, c->aux_len, tw,
                        c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->e_sorted.as<unsigned long long>(), tau, zs,
                        prm->order_n, c->pairs.as<uint2>(), d_cnt + M_PAIRS, (unsigned long long)PAIR_CAP, d_cnt + M_UNIT);
            } else if (crit == SELB200_CRIT_HLL_A) {
                const int grid = (int)std::min<int64_t>(nt * 4, (int64_t)hll_grid);
                k_tile_filter_hll<0><<<grid, 64, hll_smem, s>>>(
                    c->auxT.as<uint32_t>(), c->npad, c->aux_len, tw, c->lo.as<int32_t>(), c->hi.as<int32_t>(), n,
                    c->e_sorted.as<unsigned long long>(), tau, zs, prm->order_n, c->pairs.as<uint2>(), d_cnt + M_PAIRS,
                    (unsigned long long)PAIR_CAP, d_cnt + M_UNIT);
            } else {
                const int grid = (int)std::min<int64_t>(nt * 4, (int64_t)hll_grid);
                k_tile_filter_hll<1><<<grid, 64, hll_smem, s>>>(
                    c->auxT.as<uint32_t>(), c->npad, c->aux_len, tw, c->lo.as<int32_t>(), c->hi.as<int32_t>(), n,
                    c->e_sorted.as<unsigned long long>(), tau, zs, prm->order_n, c->pairs.as<uint2>(), d_cnt + M_PAIRS,
                    (unsigned long long)PAIR_CAP, d_cnt + M_UNIT);
            }
            CK(cudaGetLastError());
            st.launches++;
            DBG_SYNC(c, "tile filter");
            cudaEvent_t f1 = c->ev();
            t_filter.push_back({f0, f1});
            if (crit == SELB200_CRIT_SMH_A) {
                k_smh_verify<<<c->sm_count * 8, 256, 0, s>>>(
                    c->aux_sorted.as<uint64_t>(), c->sigT.as<uint32_t>(), c->npad, c->aux_len, n_rows, n_bands,
                    c->cand.as<uint2>(), d_cnt + M_CAND, (unsigned long long)PAIR_CAP, c->pairs.as<uint2>(),
                    d_cnt + M_PAIRS, (unsigned long long)PAIR_CAP);
                CK(cudaGetLastError());
                st.launches++;
            }
            DBG_SYNC(c, "smh verify");
            // ---- K5 + K6 --------------------------------------------------------------
            cudaEvent_t u0 = c->ev();
            if (crit == SELB200_CRIT_SMH_A) t_verify.push_back({f1, u0});
            if (union_bytes) {
                CKR(launch_pair_hist(c, c->d_regs, c->m, c->p, c->order_dev.as<int32_t>(), c->pairs.as<uint2>(),
                                     (int64_t)pair_lim, c->hist.as<uint32_t>(), d_cnt + M_PAIRS));
                st.launches++;
            } else {
                CKR(launch_pair_hist_planes(c, c->pairs.as<uint2>(), (int64_t)pair_lim, c->hist.as<uint32_t>(),
                                            d_cnt + M_PAIRS, d_cnt + M_WIDE, &st.launches, attempt == 0 && ri == 0));
            }
            DBG_SYNC(c, "union histogram (planes + wide)");
            cudaEvent_t u1 = c->ev();
            k_estimate_emit<<<c->sm_count * 8, 128, 0, s>>>(
                c->hist.as<uint32_t>(), c->pairs.as<uint2>(), d_cnt + M_PAIRS, pair_lim,
                c->e_sorted.as<unsigned long long>(), c->p, tau, c->out_keys.as<uint64_t>(), c->out_j.as<double>(),
                d_cnt + M_OUT, (unsigned long long)c->out_cap, c->near_keys.as<uint64_t>(), c->near_j.as<double>(),
                d_cnt + M_NEAR, near_cap);
            CK(cudaGetLastError());
            st.launches++;
            DBG_SYNC(c, "estimate + emit");
            cudaEvent_t u2 = c->ev();
            t_union.push_back({u0, u1});
            t_est.push_back({u1, u2});
            CK(cudaMemcpyAsync(c->h_snap + ri * 4, d_cnt, 32, cudaMemcpyDeviceToHost, s));
        }
        // gather, optimistic form: the push is queued behind the only range and checks on the device
        // that nothing overflowed; the root's wait follows, so its single sync also covers the peers
        if (gather && ranges.size() <= 1) {
            CKR(launch_push(1, pair_lim));
            if (c->g.is_root) {
                k_gather_wait<<<1, 32, 0, s>>>(gz, c->g.epoch, (unsigned)c->g.world, c->g_push.as<GatherPush>(),
                                               c->g_merged.as<unsigned long long>());
                CK(cudaGetLastError());
                st.launches++;
                CK(cudaMemcpyAsync(c->g.h_merged, c->g_merged.p, 24, cudaMemcpyDeviceToHost, s));
            }
        }
        CK(cudaMemcpyAsync(h_fin, d_cnt, M_WORDS * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(c->h_tprefix.data(), c->tile_prefix.p, ((size_t)nrb + 1) * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(c->h_rb_pairs.data(), c->rb_pairs.p, (size_t)nrb * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        // ---- overflow check ------------------------------------------------------------------
        tiles_total = (int64_t)h_fin[M_TILES];
        if (h_fin[M_KERR]) return fail(SELB200_ECUDA, "internal: union kernel pipeline error %llx", h_fin[M_KERR]);
        pushed = h_fin[M_PUSHED] == 1;
        if (h_fin[M_PUSHED] == 2) return fail(SELB200_ECUDA, "gather: timed out waiting for the root to merge an earlier run");
        bool redo = false;
        if (tiles_total > c->tile_cap) {
            c->tile_cap = tiles_total;
            ranges.clear();           // built again against the full list
            redo = true;
        }
        std::vector<std::pair<int, int>> next;
        int64_t cand_sum = 0, pair_sum = 0;
        const int t_end = shard_tiles(tiles_total);
        for (size_t ri = 0; ri < ranges.size(); ++ri) {
            const unsigned long long* sn = c->h_snap + ri * 4;
            const int ra = ranges[ri].first, rb = std::min(ranges[ri].second, t_end);
            const bool too_many = sn[0] > (unsigned long long)PAIR_CAP || sn[1] > (unsigned long long)PAIR_CAP;
            if (too_many) {
                const int nt = rb - ra;
                if (nt <= 1) return fail(SELB200_ENOMEM, "a single tile produced %llu pairs", std::max(sn[0], sn[1]));
                const int mid = ra + nt / 2;
                next.push_back({ra, mid});
                next.push_back({mid, rb});
                redo = true;
                continue;
            }
            next.push_back({ra, rb});
            if ((int64_t)sn[1] > c->hist_cap_pairs) { c->hist_cap_pairs = (int64_t)sn[1]; redo = true; }
            cand_sum += crit == SELB200_CRIT_SMH_A ? (int64_t)sn[0] : (int64_t)sn[1];
            pair_sum += (int64_t)sn[1];
        }
        if ((int64_t)h_fin[M_OUT] > c->out_cap) { c->out_cap = (int64_t)h_fin[M_OUT] + (1 << 16); redo = true; }
        if (!redo) {
            st.pairs_cand = cand_sum;
            st.pairs_aux = pair_sum;
            st.batches = (int32_t)ranges.size();
            break;
        }
        if (pushed) return fail(SELB200_ECUDA, "internal: gather pushed a pass that overflowed");
        if (!ranges.empty()) ranges.swap(next);
    }
    st.pairs_cb = (int64_t)h_fin[M_PAIRS_CB];
    st.tiles_total = tiles_total;
    st.tiles_shard = shard_tiles(tiles_total);
    c->out_count = (int64_t)h_fin[M_OUT];
    c->near_count = (int64_t)std::min<unsigned long long>(h_fin[M_NEAR], near_cap);
    st.pairs_out = c->out_count;
    st.pairs_near = (int64_t)h_fin[M_NEAR];

    // ---- gather, second half ----------------------------------------------------------------
    const uint64_t* src_keys = c->out_keys.as<uint64_t>();
    const double* src_j = c->out_j.as<double>();
    if (gather) {
        if (!pushed) {           // several ranges (or a redone pass): push now that the pass is final
            CKR(launch_push(0, 0));
            if (c->g.is_root) {
                k_gather_wait<<<1, 32, 0, s>>>(gz, c->g.epoch, (unsigned)c->g.world, c->g_push.as<GatherPush>(),
                                               c->g_merged.as<unsigned long long>());
                CK(cudaGetLastError());
                st.launches++;
                CK(cudaMemcpyAsync(c->g.h_merged, c->g_merged.p, 24, cudaMemcpyDeviceToHost, s));
            }
            CK(cudaMemcpyAsync(h_fin + M_PUSHED, d_cnt + M_PUSHED, 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (h_fin[M_PUSHED] != 1) return fail(SELB200_ECUDA, "gather: timed out waiting for the root to merge an earlier run");
        }
        const unsigned b = c->g.epoch & 1u;
        if (c->g.is_root) {
            if (c->g.h_merged[2]) return fail(SELB200_ECUDA, "gather: timed out waiting for %d ranks to push their lists", c->g.world);
            if ((int64_t)c->g.h_merged[0] > c->g.cap)
                return fail(SELB200_ENOMEM, "gather: %llu pairs exceed the landing zone of %lld (selb200_gather_create)",
                            c->g.h_merged[0], (long long)c->g.cap);
            c->out_count = (int64_t)c->g.h_merged[0];
            c->near_count = (int64_t)std::min<unsigned long long>(c->g.h_merged[1], gz.near_cap);
            src_keys = gz.keys + b * gz.cap;
            src_j = gz.jac + b * gz.cap;
            // the merged near-tau list moves to the context's own buffers (the zone is handed back below)
            if (c->near_count) {
                CKR(c->near_keys.ensure((size_t)gz.near_cap * 8));
                CKR(c->near_j.ensure((size_t)gz.near_cap * 8));
                CK(cudaMemcpyAsync(c->near_keys.p, gz.near_keys + b * gz.near_cap, (size_t)c->near_count * 8, cudaMemcpyDeviceToDevice, s));
                CK(cudaMemcpyAsync(c->near_j.p, gz.near_j + b * gz.near_cap, (size_t)c->near_count * 8, cudaMemcpyDeviceToDevice, s));
            }
        } else {
            c->out_count = 0;     // this rank's pairs now live on the root
            c->near_count = 0;
        }
    }

    // ---- K7: reference print order -----------------------------------------------------------
    cudaEvent_t s0 = c->ev();
    c->res_keys = src_keys;
    c->res_j = src_j;
    const bool must_copy = gather && c->g.is_root;    // results leave the landing zone either way
    if ((prm->sort_output || must_copy) && c->out_count > 0) {
        const int64_t cnt = c->out_count;
        CKR(c->out_keys2.ensure((size_t)cnt * 8));
        CKR(c->out_j2.ensure((size_t)cnt * 8));
        if (prm->sort_output && cnt > 1 && cnt <= 4ll * n) {
            // sparse output (the selective criteria): bucket by row, rank inside the row
            CKR(c->row_cnt.ensure(((size_t)n + 1) * 4));
            CKR(c->row_off.ensure(((size_t)n + 1) * 4));
            CKR(c->sort_tmp.ensure((size_t)cnt * 16));
            uint64_t* tkeys = c->sort_tmp.as<uint64_t>();
            double* tj = reinterpret_cast<double*>(tkeys + cnt);
            CK(cudaMemsetAsync(c->row_cnt.p, 0, ((size_t)n + 1) * 4, s));
            const unsigned grid = (unsigned)((cnt + 255) / 256);
            k_rowsort_count<<<grid, 256, 0, s>>>(src_keys, cnt, c->row_cnt.as<int32_t>());
            CK(cudaGetLastError());
            size_t tmp_bytes = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->row_cnt.as<int32_t>(), c->row_off.as<int32_t>(), n + 1, s));
            CKR(c->cub_tmp.ensure(tmp_bytes));
            CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp_bytes, c->row_cnt.as<int32_t>(), c->row_off.as<int32_t>(),
                                             n + 1, s));
            k_rowsort_scatter<<<grid, 256, 0, s>>>(src_keys, src_j, cnt, c->row_cnt.as<int32_t>(), c->row_off.as<int32_t>(),
                                                   tkeys, tj);
            CK(cudaGetLastError());
            k_rowsort_rank<<<grid, 256, 0, s>>>(tkeys, tj, cnt, c->row_off.as<int32_t>(), c->out_keys2.as<uint64_t>(),
                                                c->out_j2.as<double>());
            CK(cudaGetLastError());
            st.launches += 4;
        } else if (prm->sort_output && cnt > 1) {
            size_t tmp_bytes = 0;
            int nbits = 1;
            while ((1ll << nbits) < (long long)n) ++nbits;       // key = i<<32 | k with i,k < n
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, src_keys, c->out_keys2.as<uint64_t>(), src_j,
                                               c->out_j2.as<double>(), (int)cnt, 0, 32 + nbits, s));
            CKR(c->cub_tmp.ensure(tmp_bytes));
            CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp_bytes, src_keys, c->out_keys2.as<uint64_t>(), src_j,
                                               c->out_j2.as<double>(), (int)cnt, 0, 32 + nbits, s));
        } else {
            CK(cudaMemcpyAsync(c->out_keys2.p, src_keys, (size_t)cnt * 8, cudaMemcpyDeviceToDevice, s));
            CK(cudaMemcpyAsync(c->out_j2.p, src_j, (size_t)cnt * 8, cudaMemcpyDeviceToDevice, s));
        }
        c->res_keys = c->out_keys2.as<uint64_t>();
        c->res_j = c->out_j2.as<double>();
    }
    c->host_count = -1;
    if (prm->host_results) {
        // the lists also land in pinned host memory before the run's last synchronisation
        const size_t cnt = (size_t)c->out_count;
        if (cnt > c->h_res_cap) {
            if (c->h_res) cudaFreeHost(c->h_res);
            c->h_res = nullptr;
            c->h_res_cap = 0;
            CK(cudaMallocHost(&c->h_res, (cnt + cnt / 4 + 1024) * 16));
            c->h_res_cap = cnt + cnt / 4 + 1024;
        }
        if (cnt) {
            CK(cudaMemcpyAsync(c->h_res, c->res_keys, cnt * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync((uint8_t*)c->h_res + c->h_res_cap * 8, c->res_j, cnt * 8, cudaMemcpyDeviceToHost, s));
        }
        c->host_count = (int64_t)cnt;
    }
    if (gather) {
        if (c->g.is_root) {
            k_gather_release<<<1, 32, 0, s>>>(gz, c->g.epoch);
            CK(cudaGetLastError());
            st.launches++;
        }
        c->g.epoch++;
    }
    cudaEvent_t ev_end = c->ev();
    CK(cudaStreamSynchronize(s));

    auto sum_ms = [](const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v) {
        float tot = 0.f;
        for (const auto& pr : v) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) tot += ms;
        }
        return tot;
    };
    cudaEventElapsedTime(&st.ms_bounds, ev_begin, ev_bounds);
    st.ms_filter = sum_ms(t_filter);
    st.ms_verify = sum_ms(t_verify);
    st.ms_union = sum_ms(t_union);
    st.ms_estimate = sum_ms(t_est);
    cudaEventElapsedTime(&st.ms_sort, s0, ev_end);
    cudaEventElapsedTime(&st.ms_total, ev_begin, ev_end);
    // shard share of the CB band: a row block's pairs are apportioned by how many of its tiles
    // the shard owns
    {
        double acc = 0.;
        for (int rb = 0; rb < nrb; ++rb) {
            const int a = c->h_tprefix[(size_t)rb], b = c->h_tprefix[(size_t)rb + 1];
            if (b <= a) continue;
            // tiles t in [a,b) with t % n_shards == shard
            const int first = a + ((prm->shard - a % n_shards) % n_shards + n_shards) % n_shards;
            const int mine = first < b ? (b - 1 - first) / n_shards + 1 : 0;
            acc += (double)c->h_rb_pairs[(size_t)rb] * (double)mine / (double)(b - a);
        }
        st.pairs_cb_shard = (int64_t)(acc + 0.5);
    }
    if (st_out) *st_out = st;
    return SELB200_OK;
}

// ---- peer-memory gather ------------------------------------------------------------------
namespace {
struct GatherHandle {            // what selb200_gather_create exports (SELB200_GATHER_HANDLE_BYTES)
    cudaIpcMemHandle_t ipc;      // 64 bytes
    int64_t cap, near_cap;
    int64_t pid;
    uint64_t raw;                // the pointer itself: used when root and peer share a process
    int32_t device, pad;
};
static_assert(sizeof(GatherHandle) <= SELB200_GATHER_HANDLE_BYTES, "handle blob too small");

size_t gather_zone_bytes(int64_t cap, int64_t near_cap) {
    return sizeof(GatherHdr) + (size_t)cap * 32 + (size_t)near_cap * 32;
}
}  // namespace

void selb200_gather_close(selb200_ctx* c) {
    if (!c || !c->g.zone) { if (c) c->g = selb200_ctx::Gather(); return; }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->g.mapped) cudaIpcCloseMemHandle(c->g.zone);
    else if (c->g.is_root) cudaFree(c->g.zone);
    if (c->g.h_merged) cudaFreeHost(c->g.h_merged);
    c->g = selb200_ctx::Gather();
}

int selb200_gather_create(selb200_ctx* c, int64_t cap_pairs, void* handle_out) {
    if (!c || !handle_out) return fail(SELB200_EINVAL, "null argument");
    if (cap_pairs < 1) return fail(SELB200_EINVAL, "gather capacity %lld", (long long)cap_pairs);
    selb200_gather_close(c);
    CK(cudaSetDevice(c->device));
    const int64_t near_cap = 1 << 16;
    void* zone = nullptr;
    if (cudaMalloc(&zone, gather_zone_bytes(cap_pairs, near_cap)) != cudaSuccess) {
        cudaGetLastError();
        return fail(SELB200_ENOMEM, "cudaMalloc of the %zu-byte landing zone failed", gather_zone_bytes(cap_pairs, near_cap));
    }
    CK(cudaMemset(zone, 0, sizeof(GatherHdr)));
    GatherHandle h;
    std::memset(&h, 0, sizeof h);
    CK(cudaIpcGetMemHandle(&h.ipc, zone));
    h.cap = cap_pairs; h.near_cap = near_cap; h.pid = (int64_t)getpid(); h.raw = (uint64_t)(uintptr_t)zone;
    h.device = c->device;
    std::memset(handle_out, 0, SELB200_GATHER_HANDLE_BYTES);
    std::memcpy(handle_out, &h, sizeof h);
    c->g.zone = zone; c->g.is_root = true; c->g.cap = cap_pairs; c->g.near_cap = near_cap;
    return SELB200_OK;
}

int selb200_gather_attach(selb200_ctx* c, int rank, int world, const void* root_handle) {
    if (!c || !root_handle) return fail(SELB200_EINVAL, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(SELB200_EINVAL, "rank %d of %d", rank, world);
    CK(cudaSetDevice(c->device));
    GatherHandle h;
    std::memcpy(&h, root_handle, sizeof h);
    if (c->g.is_root && c->g.zone) {
        if ((uint64_t)(uintptr_t)c->g.zone != h.raw || h.pid != (int64_t)getpid())
            return fail(SELB200_EINVAL, "gather_attach on the root needs the handle it created");
    } else {
        selb200_gather_close(c);
        if (h.pid == (int64_t)getpid()) {
            // same process (several contexts in one host program): the pointer itself is valid here
            if (h.device != c->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(SELB200_ECUDA, "no peer access from device %d to %d: %s", c->device, h.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            c->g.zone = (void*)(uintptr_t)h.raw;
        } else {
            void* zone = nullptr;
            CK(cudaIpcOpenMemHandle(&zone, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            c->g.zone = zone;
            c->g.mapped = true;
        }
        c->g.cap = h.cap; c->g.near_cap = h.near_cap;
    }
    if (!c->g.h_merged) CK(cudaMallocHost(&c->g.h_merged, 32));
    c->g.rank = rank; c->g.world = world; c->g.epoch = 0;
    c->g.attached = true;
    return SELB200_OK;
}

int64_t selb200_result_count(selb200_ctx* c) { return c ? c->out_count : 0; }
int64_t selb200_near_count(selb200_ctx* c) { return c ? c->near_count : 0; }

static int copy_list(selb200_ctx* c, const uint64_t* d_keys, const double* d_j, int64_t count, int64_t cap,
                     int32_t* i, int32_t* k, double* jac) {
    const int64_t cnt = std::min(count, cap);
    if (cnt <= 0) return SELB200_OK;
    CK(cudaSetDevice(c->device));
    std::vector<uint64_t> keys((size_t)cnt);
    CK(cudaMemcpyAsync(keys.data(), d_keys, (size_t)cnt * 8, cudaMemcpyDeviceToHost, c->stream));
    if (jac) CK(cudaMemcpyAsync(jac, d_j, (size_t)cnt * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int64_t t = 0; t < cnt; ++t) {
        if (i) i[t] = (int32_t)(keys[(size_t)t] >> 32);
        if (k) k[t] = (int32_t)(keys[(size_t)t] & 0xffffffffu);
    }
    return SELB200_OK;
}

int selb200_result_host(selb200_ctx* c, const uint64_t** keys, const double** jaccard) {
    if (!c) return fail(SELB200_EINVAL, "null context");
    if (c->host_count < 0) return fail(SELB200_ESTATE, "the last run did not set params.host_results");
    if (keys) *keys = (const uint64_t*)c->h_res;
    if (jaccard) *jaccard = (const double*)((const uint8_t*)c->h_res + c->h_res_cap * 8);
    return SELB200_OK;
}

int selb200_copy_results(selb200_ctx* c, int64_t cap, int32_t* i, int32_t* k, double* jaccard) {
    if (!c) return fail(SELB200_EINVAL, "null context");
    if (c->host_count >= 0) {       // already on the host (params.host_results)
        const int64_t cnt = std::min(c->host_count, cap);
        const uint64_t* hk = (const uint64_t*)c->h_res;
        const double* hj = (const double*)((const uint8_t*)c->h_res + c->h_res_cap * 8);
        for (int64_t t = 0; t < cnt; ++t) {
            if (i) i[t] = (int32_t)(hk[t] >> 32);
            if (k) k[t] = (int32_t)(hk[t] & 0xffffffffu);
        }
        if (jaccard && cnt > 0) std::memcpy(jaccard, hj, (size_t)cnt * 8);
        return SELB200_OK;
    }
    return copy_list(c, c->res_keys, c->res_j, c->out_count, cap, i, k, jaccard);
}

int selb200_copy_near(selb200_ctx* c, int64_t cap, int32_t* i, int32_t* k, double* jaccard) {
    if (!c) return fail(SELB200_EINVAL, "null context");
    return copy_list(c, c->near_keys.as<uint64_t>(), c->near_j.as<double>(), c->near_count, cap, i, k, jaccard);
}

int selb200_result_device(selb200_ctx* c, const uint64_t** d_keys, const double** d_jaccard) {
    if (!c) return fail(SELB200_EINVAL, "null context");
    if (d_keys) *d_keys = c->res_keys;
    if (d_jaccard) *d_jaccard = c->res_j;
    return SELB200_OK;
}

int selb200_debug_union(selb200_ctx* c, int which, int64_t count, const int32_t* a, const int32_t* b, double* t) {
    if (!c || !c->loaded) return fail(SELB200_ESTATE, "no sketches loaded");
    if (count <= 0) return SELB200_OK;
    if (which != 0) return fail(SELB200_EINVAL, "debug_union: only the primary sketches are addressable");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    std::vector<uint2> pr((size_t)count);
    for (int64_t x = 0; x < count; ++x) {
        if (a[x] < 0 || a[x] >= c->n || b[x] < 0 || b[x] >= c->n) return fail(SELB200_EINVAL, "pair index out of range");
        pr[(size_t)x] = make_uint2((uint32_t)a[x], (uint32_t)b[x]);
    }
    CKR(upload(c->pairs, pr, s));
    CKR(c->hist.ensure((size_t)count * 64 * 4));
    CKR(c->out_j.ensure((size_t)count * 8));
    CKR(launch_pair_hist(c, c->d_regs, c->m, c->p, nullptr, c->pairs.as<uint2>(), count, c->hist.as<uint32_t>()));
    k_mle_only<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(c->hist.as<uint32_t>(), count, c->p, c->out_j.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(t, c->out_j.p, (size_t)count * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    c->out_count = 0;
    c->host_count = -1;
    return SELB200_OK;
}

}  // extern "C"

#include "shims.inl"
