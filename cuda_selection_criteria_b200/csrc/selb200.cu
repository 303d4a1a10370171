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
#include <time.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <cooperative_groups.h>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "estimators.cuh"
#include "hostpack.h"

// ============================================================================
// error plumbing
// ============================================================================
namespace {

thread_local std::string g_err;

// SELB200_HLLFILTER (read once per process): "onepass" = exact MLE for every pair of the band on bit planes (the form the
// two-pass filter replaced), "bytes" = shared-memory-counter filter over the transposed registers; both for A/B measurements
int hll_filter_mode() {
    static const int mode = [] {
        const char* e = getenv("SELB200_HLLFILTER");
        return !e ? 0 : !strcmp(e, "onepass") ? 1 : !strcmp(e, "bytes") ? 2 : 0;
    }();
    return mode;
}

// SELB200_SMHFILTER=tiles | join (read once per process) forces the all-pairs tile filter + verify, or the equality join.
// Default: the join at every shard count.  Its keys and buckets are the same 0.06 ms on every shard, expansion and item walk
// divide (a shard expands its own rows); the tile filter divides as a whole (0.59 ms / shards + 0.05 for signatures and
// verify).  (While the expansion was replicated too — 0.15 ms on every shard — the tile filter won from four shards on.)
bool smh_join_enabled(int n_shards) {
    static const int mode = [] {
        const char* e = getenv("SELB200_SMHFILTER");
        return !e ? 0 : !strcmp(e, "tiles") ? 1 : !strcmp(e, "join") ? 2 : 0;
    }();
    (void)n_shards;
    return mode != 1;
}

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

struct PackSlot {             // pinned host staging of one packed piece (hostpack.h)
    uint8_t* buf = nullptr;
    size_t cap = 0;
    cudaEvent_t free_ev = nullptr;    // recorded on the copy stream after the slot's H2D copy
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
    DevBuf lo, hi, tile_prefix, tile_cb0, tile_rc, sigT, cand, pairs, hist, counters, cub_tmp, cub_tmp2;
    DevBuf out_keys, out_j, out_keys2, out_j2, near_keys, near_j;
    int64_t out_count = 0, near_count = 0;
    int64_t hist_cap_pairs = 0, out_cap = 0;      // grow-only capacities of the sync-free run pipeline
    int64_t near_cap = 1 << 16;                   // near-tau list: grown (and the pass redone) when a run overflows it
    LoadState ld;
    PackSlot pack_slots[4];
    cudaEvent_t h2d_evs[16] = {};        // one per register copy in flight during a packed load: how much work the link still holds
    int64_t ld_h2d_bytes = 0, ld_rows_packed = 0, ld_rows_raw = 0;     // register bytes the last host load moved, and how
    DevBuf pk_buf;                       // packed pieces as they land on the device, before k_unpack_nib4
    DevBuf join_buf, join_items;         // smh_a equality join: keys / values, unsorted and sorted, genome-major signatures; items
    int64_t join_item_cap = 0;           // grow-only, like every other list of the run
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
    // pinned: tile prefix (nrb + 1 x i32) and pairs per row block (nrb x u64) of a sharded run (the shard's share of the band)
    void* h_band = nullptr;
    size_t h_band_cap = 0;
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
    DevBuf row_cnt, row_off, sort_tmp, sort_blocksum;
    DevBuf auxP, agrange, atail;         // bit planes (quad layout) / register ranges / tail sums of the auxiliary HLLs (sorted order)
    bool auxp_quad = false;              // auxP holds the planes of the loaded auxiliary HLLs (k_aux_planes_quad)
    DevBuf planes, grange, gtop, wide_list, wide_flag;    // bit-plane copy of the primary registers (file-list order)
    uint32_t wide_epoch = 0;             // stamp of the current union pass in wide_flag
    int chunk_regs = 0;
    // counting step of the plane kernel: subset masks on groups of four values (k_pair_hist_planes<EpiSubsets<..>>,
    // default) or one-hot masks on groups of eight (SELB200_UNION=planes)
    bool union_subsets = true;
    bool union_tops = true;              // per-step group limit from the per-eighth maxima (SELB200_UNION_TOPS=0: the pair's range)
    void* h_res = nullptr;               // pinned host copy of the result lists (params.host_results)
    size_t h_res_cap = 0;                // in pairs: keys at [0, cap), Jaccards at [cap, 2 cap)
    int64_t host_count = -1;

    cudaEvent_t ev() { return ev_on(stream); }
    cudaEvent_t ev_on(cudaStream_t where) {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        cudaEvent_t e = ev_pool[ev_used++];
        cudaEventRecord(e, where);
        return e;
    }
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;     // side-stream work of a run (signatures next to the bounds)
};

// ============================================================================
// device code: one anonymous namespace, one translation unit; the kernels live in kernels/*.inl
// ============================================================================
namespace {

#include "kernels/helpers.inl"
#include "kernels/union_bytes.inl"
#include "kernels/union_planes.inl"
#include "kernels/load_kernels.inl"
#include "kernels/tiles.inl"
#include "kernels/filter_smh.inl"
#include "kernels/filter_hll.inl"
#include "kernels/estimate_sort.inl"
#include "kernels/gather.inl"

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

// cooperative launch (grid-wide barriers inside the kernel): the grid must be co-resident.  SELB200_FUSED=0 keeps the
// separate kernels (A/B measurements); a device without cooperative launch gets them too
bool fused_launches_enabled(int device) {
    static const bool env_on = [] { const char* e = getenv("SELB200_FUSED"); return !(e && !strcmp(e, "0")); }();
    if (!env_on) return false;
    int ok = 0;
    if (cudaDeviceGetAttribute(&ok, cudaDevAttrCooperativeLaunch, device) != cudaSuccess) { cudaGetLastError(); return false; }
    return ok != 0;
}
template <typename K>
int coop_grid(K kernel, int threads, int sm_count, long long wanted) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return (int)std::max<long long>(1, std::min<long long>(wanted, (long long)sm_count * std::min(per_sm, 4)));
}

template <class Src, class Epi>
int launch_pair_hist_t(cudaStream_t stream, int sm_count, const uint8_t* regs, size_t row_stride, size_t m, int p,
                       int64_t max_pairs, Src src, Epi epi) {
    if (max_pairs <= 0) return SELB200_OK;
    const int64_t ctas_needed = (max_pairs + 1) / 2;
    const int nbins = Src::kMask != 0xffffffffu ? 64 : 64 - p + 2;   // masked (unvalidated) bytes can take any 6-bit value
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
// The plane kernel on the run stream; the byte kernel for the pairs it hands to the wide list on `wide_stream` — by default
// the run stream itself, so that one estimate launch behind it sees every row.  With another stream (SELB200_WIDE=side) the
// caller estimates the wide pairs there too and joins afterwards; the plane kernel stamps wide_flag[pair] = c->wide_epoch so
// that the main estimate leaves those rows alone (measured slower: the side chain runs behind the main estimate, not beside it).
int launch_pair_hist_planes(selb200_ctx* c, const uint2* pairs, int64_t max_pairs, uint32_t* hist_out,
                            const unsigned long long* npairs_dev, unsigned long long* wide_count, int* launches,
                            bool counters_are_zero, cudaStream_t wide_stream = nullptr) {
    if (max_pairs <= 0) return SELB200_OK;
    cudaStream_t s = c->stream;
    if (!wide_stream) wide_stream = s;
    CKR(c->wide_list.ensure((size_t)max_pairs * 4));
    {
        const void* had = c->wide_flag.p;
        CKR(c->wide_flag.ensure((size_t)max_pairs * 4));
        if (c->wide_flag.p != had) {              // fresh memory: no stamp of an epoch to come
            CK(cudaMemsetAsync(c->wide_flag.p, 0, c->wide_flag.cap, s));
            c->wide_epoch = 0;
        }
        ++c->wide_epoch;
    }
    uint32_t* wflag = c->wide_flag.as<uint32_t>();
    const uint32_t wepoch = c->wide_epoch;
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
    // wide count and batch counter (adjacent words of meta[]); the kernel error word behind them is sticky for the
    // whole run: an error raised in an earlier range must still be there when the host reads meta[] at the end
    if (!counters_are_zero) CK(cudaMemsetAsync(wide_count, 0, 16, s));
    if (c->union_subsets) {
        static int u_per_sm = 0;
        static size_t u_per_sm_smem = 0;
        if (!u_per_sm || u_per_sm_smem != smem) {
            cudaFuncSetAttribute(k_pair_hist_planes<EpiSubsets<EpiWriteHist>>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&u_per_sm, k_pair_hist_planes<EpiSubsets<EpiWriteHist>>, 32, smem) != cudaSuccess ||
                u_per_sm < 1) {
                cudaGetLastError();
                u_per_sm = 4;
            }
            u_per_sm_smem = smem;
        }
        const int grid = (int)std::min<int64_t>((max_pairs + 3) / 4, (int64_t)c->sm_count * u_per_sm);
        k_pair_hist_planes<EpiSubsets<EpiWriteHist>><<<grid, 32, smem, s>>>(c->planes.as<uint32_t>(), c->m, c->chunk_regs,
                                                                       c->grange.as<uint16_t>(), src, EpiSubsets<EpiWriteHist>{epi},
                                                                       c->wide_list.as<uint32_t>(), wide_count, wide_count + 1, wflag, wepoch,
                                                                       c->union_tops && c->m >= 16384 ? c->gtop.as<uint32_t>() : nullptr);
    } else {
        const int grid = (int)std::min<int64_t>((max_pairs + 3) / 4, (int64_t)c->sm_count * per_sm);
        k_pair_hist_planes<EpiWriteHist><<<grid, 32, smem, s>>>(c->planes.as<uint32_t>(), c->m, c->chunk_regs,
                                                               c->grange.as<uint16_t>(), src, epi,
                                                               c->wide_list.as<uint32_t>(), wide_count, wide_count + 1, wflag, wepoch);
    }
    CK(cudaGetLastError());
    if (wide_stream != s) {
        CK(cudaEventRecord(c->fork_ev, s));
        CK(cudaStreamWaitEvent(wide_stream, c->fork_ev, 0));
    }
    // pairs whose value range exceeds the 32-value window: byte kernel, small persistent grid
    SrcWide wsrc{pairs, c->order_dev.as<int32_t>(), c->wide_list.as<uint32_t>(), wide_count};
    if (c->m >= 512) {
        const int nbins = 64 - c->p + 2;
        // (two, four or eight CTAs per SM: 22 us either way for the ~800 wide pairs of n = 100k — latency, not throughput)
        const int wgrid = (int)std::min<int64_t>((max_pairs + 1) / 2, (int64_t)c->sm_count * 2);
        if (nbins <= 52) k_pair_hist<52, SrcWide, EpiWriteHist><<<wgrid, 64, 0, wide_stream>>>(c->d_regs, c->m, c->m, wsrc, epi);
        else k_pair_hist<64, SrcWide, EpiWriteHist><<<wgrid, 64, 0, wide_stream>>>(c->d_regs, c->m, c->m, wsrc, epi);
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
    CKR(c->gtop.ensure((size_t)n * 8 * sizeof(uint32_t)));
    CKR(c->hist.ensure((size_t)n * 64 * sizeof(uint32_t)));
    CKR(c->cards_in.ensure((size_t)n * sizeof(double)));
    CKR(c->out_j.ensure((size_t)n * sizeof(double)));        // stored value_ of each header (-1 = recompute)
    CKR(c->counters.ensure(64));
    CK(cudaMemsetAsync(c->counters.p, 0, 64, s));
    return SELB200_OK;
}

int pinned_ensure(void** ptr, size_t* cap, size_t bytes) {
    if (bytes <= *cap && *ptr) return SELB200_OK;
    if (*ptr) cudaFreeHost(*ptr);
    *ptr = nullptr;
    *cap = 0;
    CK(cudaMallocHost(ptr, bytes ? bytes : 16));
    *cap = bytes;
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
        // per-eighth maxima next to the planes (the subset union kernel's per-step group limit)
        CK(cudaMemsetAsync(c->gtop.as<uint32_t>() + (size_t)g0 * 8, 0, (size_t)rows * 8 * sizeof(uint32_t), s));
        k_planes_from_bytes<<<grid, 256, 0, s>>>(c->d_regs + (size_t)g0 * c->m, rows, c->m, c->chunk_regs,
                                                 c->planes.as<uint32_t>() + (size_t)g0 * 6 * (c->m >> 5),
                                                 c->m >= 4096 ? c->gtop.as<uint32_t>() + (size_t)g0 * 8 : nullptr);
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
        c->auxp_quad = false;
        // bit planes for the plane filter, in the quad layout (a word pair = two 32-register words per step); the filter
        // addresses it with 32-bit uint4 offsets, a matrix beyond that keeps the byte form over auxT
        const int nw = (1 << aux_len) >> 5;
        if (aux_len >= 6 && (uint64_t)6 * nw * (uint64_t)c->npad < (1ull << 32)) {
            CKR(c->auxP.ensure((size_t)6 * nw * c->npad * 4));
            CKR(c->agrange.ensure((size_t)c->npad * sizeof(uint16_t)));
            CKR(c->atail.ensure((size_t)c->npad * sizeof(AuxTail)));
            CK(cudaMemsetAsync(c->auxP.p, 0, (size_t)6 * nw * c->npad * 4, s));
            CK(cudaMemsetAsync(c->agrange.p, 0, (size_t)c->npad * sizeof(uint16_t), s));
            CK(cudaMemsetAsync(c->atail.p, 0, (size_t)c->npad * sizeof(AuxTail), s));
            const int g2 = (int)std::min<int64_t>((n * nw + 255) / 256, (int64_t)c->sm_count * 16);
            k_aux_planes_quad<<<g2, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_aux), c->order_dev.as<int32_t>(), n,
                                                 c->npad, aux_len, c->auxP.as<uint32_t>());
            CK(cudaGetLastError());
            k_aux_range<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_aux),
                                                                         c->order_dev.as<int32_t>(), n, aux_len,
                                                                         c->agrange.as<uint16_t>(), c->atail.as<AuxTail>());
            CK(cudaGetLastError());
            c->auxp_quad = true;
        }
    }
    CK(cudaStreamSynchronize(s));
    c->loaded = true;
    return SELB200_OK;
}

// SELB200_H2D=raw: registers cross PCIe as the bytes they are (A/B measurements); default: packed (hostpack.h)
bool h2d_packed() {
    static const bool on = [] { const char* e = getenv("SELB200_H2D"); return !(e && !strcmp(e, "raw")); }();
    return on;
}
// threads of the host packer: SELB200_PACK_THREADS, else the machine's hardware threads shared among the processes of
// this node (LOCAL_WORLD_SIZE, as torchrun exports it; torchrun also pins OMP_NUM_THREADS to 1, which is why the OpenMP
// default is not used here)
int pack_threads() {
    static const int nt = [] {
        if (const char* e = getenv("SELB200_PACK_THREADS")) { const int v = atoi(e); if (v > 0) return std::min(v, 256); }
        long hw = sysconf(_SC_NPROCESSORS_ONLN);
        if (hw < 1) hw = 1;
        int local = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(e));
        return (int)std::max<long>(1, std::min<long>(hw / local, 64));
    }();
    return nt;
}

// packed pieces that sit in device memory (piece_rows rows each, the last one shorter, piece_stride bytes apart) -> `count`
// rows of the byte matrix starting at regs_rows (run stream)
int unpack_pieces(selb200_ctx* c, const uint8_t* d_pieces, size_t piece_stride, int64_t piece_rows, int64_t count,
                  bool maybe_raw, uint8_t* regs_rows) {
    if (count <= 0) return SELB200_OK;
    const unsigned n_pieces = (unsigned)((count + piece_rows - 1) / piece_rows);
    const Nib4Pieces a{d_pieces, piece_stride, (long long)piece_rows, (long long)count, c->p};
    const long long n16 = (long long)std::min(piece_rows, count) * (long long)(c->m >> 4);
    const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>((n16 + 255) / 256, (long long)c->sm_count * 16 / n_pieces + 1));
    k_unpack_nib4<<<dim3(gx, n_pieces), 256, 0, c->stream>>>(a, regs_rows);
    CK(cudaGetLastError());
    const long long ne = (long long)std::min(piece_rows, count) * selb::NIB4_EXC_CAP;
    const unsigned ex = (unsigned)std::max<long long>(1, std::min<long long>((ne + 255) / 256, (long long)c->sm_count * 8 / n_pieces + 1));
    k_apply_nib4_exc<<<dim3(ex, n_pieces), 256, 0, c->stream>>>(a, regs_rows);
    CK(cudaGetLastError());
    if (maybe_raw) {
        k_apply_nib4_raw<<<dim3(selb::NIB4_RAW_CAP, n_pieces), 256, 0, c->stream>>>(a, regs_rows);
        CK(cudaGetLastError());
    }
    return SELB200_OK;
}

// host registers -> device in packed form: 16 MiB of registers at a time are packed by the host threads into one of four
// pinned slots and copied (half the bytes) while the next piece is being packed; the run stream unpacks each piece as it
// lands and, every rows_per_chunk rows, digests what is there like any other chunk
int load_host_packed(selb200_ctx* c, int64_t n, const uint8_t* regs, const double* stored, const void* aux) {
    LoadState& L = c->ld;
    const size_t m = c->m;
    const int64_t pk_rows = std::max<int64_t>(1, (int64_t)(16u << 20) / (int64_t)m);
    const selb::Nib4Piece P = selb::nib4_piece(pk_rows, m);
    const int64_t n_pieces = (n + pk_rows - 1) / pk_rows;
    CKR(c->pk_buf.ensure((size_t)n_pieces * P.bytes));
    for (PackSlot& ps : c->pack_slots) {
        if (ps.cap < P.bytes) {
            if (ps.buf) cudaFreeHost(ps.buf);
            ps.buf = nullptr;
            ps.cap = 0;
            CK(cudaMallocHost(&ps.buf, P.bytes));
            ps.cap = P.bytes;
        }
        if (!ps.free_ev) CK(cudaEventCreateWithFlags(&ps.free_ev, cudaEventDisableTiming));
    }
    const int nt = pack_threads();
    int slot_i = 0;
    int64_t group0 = 0;                       // first row not yet digested
    static const bool trace = getenv("SELB200_PACK_TRACE") != nullptr;
    // SELB200_H2D=packed: every piece packed (A/B measurements); default: a piece goes RAW whenever the copy engine has
    // nothing left to do — packing is bounded by the host's memory system (47 - 146 GB/s of registers measured on this
    // pool's hosts, depending on the machine and its neighbours), the raw copy by PCIe (54 GB/s), and the two resources
    // work side by side: a fast host packs everything (copy-bound, half the bytes), a slow one lets the link carry
    // raw pieces while it packs
    static const bool always_pack = [] { const char* e = getenv("SELB200_H2D"); return e && !strcmp(e, "packed"); }();
    bool src_pinned = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, regs) == cudaSuccess) src_pinned = at.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    for (cudaEvent_t& e : c->h2d_evs)
        if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // copies in flight (they complete in order): ring of (event, estimated link time); the link is assumed to move 50 GB/s
    double q_cost[16];
    int q_head = 0, q_tail = 0;             // [q_tail, q_head) in flight, indices mod 16
    auto link_pending = [&]() -> double {
        while (q_tail != q_head && cudaEventQuery(c->h2d_evs[q_tail & 15]) == cudaSuccess) ++q_tail;
        cudaGetLastError();
        double t = 0.;
        for (int q = q_tail; q != q_head; ++q) t += q_cost[q & 15];
        return t;
    };
    auto link_push = [&](size_t bytes) -> int {
        if (q_head - q_tail == 16) { CK(cudaEventSynchronize(c->h2d_evs[q_tail & 15])); ++q_tail; }
        CK(cudaEventRecord(c->h2d_evs[q_head & 15], c->copy_stream));
        q_cost[q_head & 15] = (double)bytes / 50e9;
        ++q_head;
        return SELB200_OK;
    };
    double pack_est = (double)pk_rows * (double)m / 100e9;     // seconds to pack a piece: 100 GB/s to start with, then measured
    c->ld_h2d_bytes = 0; c->ld_rows_packed = 0; c->ld_rows_raw = 0;
    double t_pack = 0., t_wait = 0., t_queue = 0.;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    const double t_begin = now();
    for (int64_t g0 = 0; g0 < n; g0 += pk_rows) {
        const int64_t rows = std::min(pk_rows, n - g0);
        const selb::Nib4Piece Q = selb::nib4_piece(rows, m);          // the last piece may be shorter
        uint8_t* d_piece = c->pk_buf.as<uint8_t>() + (size_t)(g0 / pk_rows) * P.bytes;
        uint8_t* d_rows = c->regs_own.as<uint8_t>() + (size_t)g0 * m;
        const double t0 = now();
        bool raw_piece = false;
        if (src_pinned && !always_pack) {
            // the copies queued so far will be done before this piece is packed: the link would idle — the piece goes raw
            // and keeps it busy at no cost to the host
            raw_piece = link_pending() < pack_est;
        }
        double t2 = t0;
        if (!raw_piece) {
            PackSlot& ps = c->pack_slots[slot_i];
            slot_i = (slot_i + 1) & 3;
            if (ps.in_flight) { CK(cudaEventSynchronize(ps.free_ev)); ps.in_flight = false; }
            const double t1 = now();
            const int64_t n_raw = selb::nib4_pack_piece(regs + (size_t)g0 * m, rows, m, ps.buf, nt);
            t2 = now();
            t_wait += t1 - t0;
            t_pack += t2 - t1;
            pack_est = 0.75 * pack_est + 0.25 * (t2 - t1) * ((double)pk_rows / (double)rows);
            if (n_raw > selb::NIB4_RAW_CAP) {     // not the registers of HLLs of real sets: the piece travels as it is
                raw_piece = true;
            } else {
                const size_t nbytes = Q.off_raw + (size_t)n_raw * m;
                CK(cudaMemcpyAsync(d_piece, ps.buf, nbytes, cudaMemcpyHostToDevice, c->copy_stream));
                CK(cudaEventRecord(ps.free_ev, c->copy_stream));
                ps.in_flight = true;
                CKR(link_push(nbytes));
                CKR(load_join_copies(c));
                CKR(unpack_pieces(c, d_piece, P.bytes, rows, rows, n_raw > 0, d_rows));
                c->ld_h2d_bytes += (int64_t)nbytes;
                c->ld_rows_packed += rows;
            }
        }
        if (raw_piece) {
            CK(cudaMemcpyAsync(d_rows, regs + (size_t)g0 * m, (size_t)rows * m, cudaMemcpyHostToDevice, c->copy_stream));
            CKR(link_push((size_t)rows * m));
            CKR(load_join_copies(c));
            c->ld_h2d_bytes += (int64_t)rows * (int64_t)m;
            c->ld_rows_raw += rows;
        }
        const int64_t done = g0 + rows;
        if (done - group0 >= L.rows_per_chunk || done == n) {
            CKR(load_chunk(c, group0, done - group0, nullptr, stored ? stored + group0 : nullptr,
                           aux ? (const uint8_t*)aux + (size_t)group0 * L.aux_row_bytes : nullptr));
            group0 = done;
        }
        t_queue += now() - t2;
    }
    if (trace)
        fprintf(stderr, "selb200 packed load: %lld rows (%lld packed, %lld raw), %d threads: pack %.2f ms, waiting for a slot %.2f ms, "
                "queueing %.2f ms, loop %.2f ms\n", (long long)n, (long long)c->ld_rows_packed, (long long)c->ld_rows_raw, nt,
                t_pack * 1e3, t_wait * 1e3, t_queue * 1e3, (now() - t_begin) * 1e3);
    return SELB200_OK;
}

int do_load(selb200_ctx* c, int64_t n, int p, const uint8_t* regs, bool on_device, const double* stored,
            int aux_kind, int aux_len, const void* aux) {
    if (n > 0 && !regs) return fail(SELB200_EINVAL, "null register matrix");
    if (aux_kind != SELB200_AUX_NONE && n > 0 && !aux) return fail(SELB200_EINVAL, "null aux matrix");
    CKR(load_begin(c, n, p, aux_kind, aux_len, on_device ? regs : nullptr, on_device ? aux : nullptr));
    LoadState& L = c->ld;
    if (!on_device && h2d_packed() && n > 0) {
        CKR(load_host_packed(c, n, regs, stored, aux));
        return load_end(c);
    }
    c->ld_h2d_bytes = on_device ? 0 : (int64_t)n * (int64_t)c->m;
    c->ld_rows_packed = 0;
    c->ld_rows_raw = on_device ? 0 : n;
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
    {   // SELB200_UNION: subsets (default) | planes (one-hot counting) | bytes — form of the union pass, read per context
        const char* e = getenv("SELB200_UNION");
        c->union_subsets = !(e && !strcmp(e, "planes"));
        const char* t = getenv("SELB200_UNION_TOPS");
        c->union_tops = !(t && !strcmp(t, "0"));
    }
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
                      &c->counters, &c->cub_tmp, &c->cub_tmp2, &c->out_keys, &c->out_j, &c->out_keys2, &c->out_j2,
                      &c->near_keys, &c->near_j, &c->tile_nt, &c->rb_pairs, &c->g_push, &c->g_merged, &c->row_cnt, &c->row_off, &c->sort_tmp, &c->sort_blocksum, &c->planes, &c->grange, &c->gtop, &c->wide_list, &c->wide_flag, &c->auxP, &c->agrange, &c->atail, &c->pk_buf, &c->join_buf, &c->join_items};
    for (DevBuf* b : bufs) b->release();
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    selb200_gather_close(c);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->h_snap) cudaFreeHost(c->h_snap);
    if (c->h_band) cudaFreeHost(c->h_band);
    for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->h2d_evs) if (e) cudaEventDestroy(e);
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    if (c->join_ev) cudaEventDestroy(c->join_ev);
    for (PackSlot& ps : c->pack_slots) {
        if (ps.buf) cudaFreeHost(ps.buf);
        if (ps.free_ev) cudaEventDestroy(ps.free_ev);
    }
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

int selb200_load_info(selb200_ctx* c, int64_t* h2d_register_bytes, int64_t* rows_packed, int64_t* rows_raw) {
    if (!c) return fail(SELB200_EINVAL, "null context");
    if (h2d_register_bytes) *h2d_register_bytes = c->ld_h2d_bytes;
    if (rows_packed) *rows_packed = c->ld_rows_packed;
    if (rows_raw) *rows_raw = c->ld_rows_raw;
    return SELB200_OK;
}

int64_t selb200_nib4_piece_bytes(int64_t rows, int p) {
    if (rows < 0 || p < 9 || p > 20) { fail(SELB200_EINVAL, "nib4 piece: rows=%lld p=%d", (long long)rows, p); return -1; }
    return (int64_t)selb::nib4_piece(rows, (size_t)1 << p).bytes;
}

int64_t selb200_nib4_pack_piece(const uint8_t* regs, int64_t rows, int p, uint8_t* piece, int threads) {
    if (rows < 0 || p < 9 || p > 20 || (rows && (!regs || !piece))) { fail(SELB200_EINVAL, "nib4 pack: bad arguments"); return -1; }
    return selb::nib4_pack_piece(regs, rows, (size_t)1 << p, piece, threads > 0 ? threads : pack_threads());
}

// rows [g0, g0+count) arrive as packed pieces in device memory (selb200_nib4_pack_piece on some host, then any copy or
// collective): unpacked INTO the matrix given to selb200_load_device_begin, then digested like selb200_load_device_rows
int selb200_load_device_rows_packed(selb200_ctx* c, int64_t g0, int64_t count, const uint8_t* d_pieces, int64_t piece_rows) {
    if (!c || !c->ld.active || !c->ld.regs_borrowed) return fail(SELB200_ESTATE, "selb200_load_device_rows_packed outside load_device_begin/load_end");
    if (g0 < 0 || count < 0 || g0 + count > c->n) return fail(SELB200_EINVAL, "rows [%lld,+%lld) outside the matrix", (long long)g0, (long long)count);
    if (count && (!d_pieces || piece_rows < 1)) return fail(SELB200_EINVAL, "null pieces / piece_rows < 1");
    CK(cudaSetDevice(c->device));
    if (count)
        CKR(unpack_pieces(c, d_pieces, selb::nib4_piece(piece_rows, c->m).bytes, piece_rows, count, true,
                          const_cast<uint8_t*>(c->d_regs) + (size_t)g0 * c->m));
    return load_chunk(c, g0, count, nullptr, nullptr, nullptr);
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
    // smh_a: the signatures / bucket keys do not depend on the CB band, so they run on the side stream next to the bounds
    // kernel (fork here, join before the filter): 20 us of a run that is 0.6 ms at eight shards
    cudaStream_t side = c->copy_stream;
    if (!c->fork_ev) {
        CK(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(c->fork_ev, s));
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
    // tile capacity is known before the first pass (it only grows after an overflow), so the table can be part of the
    // fused launch; a redone pass rebuilds the table alone
    CKR(c->tile_rc.ensure((size_t)std::max<int64_t>(c->tile_cap, 1) * sizeof(int2)));
    const bool fused = fused_launches_enabled(c->device) && nrb <= 65536;
    if (fused) {
        static const int fgrid_cap = coop_grid(k_bounds_fused, 256, c->sm_count, 1 << 30);
        const int grid = std::max(1, std::min(fgrid_cap, (n + 255) / 256));
        const unsigned long long* a_e = c->e_sorted.as<unsigned long long>();
        int a_n = n, a_zeros = zeros, a_nrb = nrb;
        double a_tau = tau_cb;
        int32_t *a_lo = c->lo.as<int32_t>(), *a_hi = c->hi.as<int32_t>(), *a_nt = c->tile_nt.as<int32_t>(),
                *a_pre = c->tile_prefix.as<int32_t>(), *a_cb0 = c->tile_cb0.as<int32_t>();
        unsigned long long* a_rbp = c->rb_pairs.as<unsigned long long>();
        long long a_cap = (long long)c->tile_cap;
        int2* a_rc = c->tile_rc.as<int2>();
        unsigned long long* a_meta = d_cnt;
        void* args[] = {&a_e, &a_n, &a_zeros, &a_tau, &a_lo, &a_hi, &a_nrb, &a_nt, &a_pre, &a_cb0, &a_rbp, &a_cap, &a_rc, &a_meta};
        CK(cudaLaunchCooperativeKernel((void*)k_bounds_fused, dim3(grid), dim3(256), args, 0, s));
        st.launches += 1;
    } else {
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
    }
    DBG_SYNC(c, "cb bounds + row-block spans + scan");
    cudaEvent_t ev_bounds = nullptr;

    // ---- K3: signatures ---------------------------------------------------------------
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> t_filter, t_verify, t_union, t_est;
    const int n_words = (n_bands + 1) / 2;
    const bool use_smh = crit == SELB200_CRIT_SMH_A && smh_shape_ok;
    const bool smh_join = use_smh && smh_join_enabled(n_shards) && (int64_t)n * n_bands < (1ll << 31) && n_bands <= 65536;
    const long long jn_keys = (long long)n * n_bands;
    // join_buf: keys, ranks, bucket members (u32 per (genome, band)), genome-major signatures, bucket counters and offsets.
    // Key = band << sbits | top sbits of the signature: 16 bits up to 256 bands, fewer beyond (the table stays <= 2^24)
    int j_sbits = 16;
    while (j_sbits > 1 && ((long long)n_bands << j_sbits) > (1ll << 24)) --j_sbits;
    const long long j_buckets = (long long)std::max(n_bands, 1) << j_sbits;
    const long long jn_words = 3 * jn_keys + (long long)n * ((n_bands + 1) / 2) + 2 * (j_buckets + 1);

    CKR(c->cand.ensure((size_t)PAIR_CAP * sizeof(uint2)));
    CKR(c->pairs.ensure((size_t)PAIR_CAP * sizeof(uint2)));
    const float zs = prm->z_score * (crit >= SELB200_CRIT_HLL_A ? selb::sigma_p(c->aux_len) : 0.f);
    size_t hll_smem = 0;
    int hll_grid = 0;
    // hll_a / hll_an: two passes on bit planes (bound over the band, exact decision for the survivors) unless the
    // auxiliary sketches have no planes (below 64 registers, or a matrix beyond 32-bit offsets) or SELB200_HLLFILTER says so
    const bool hll_planes = crit >= SELB200_CRIT_HLL_A && c->auxp_quad && hll_filter_mode() != 2;
    const bool hll_twopass = hll_planes && hll_filter_mode() == 0;
    int hll_bound_grid = 0;
    if (crit >= SELB200_CRIT_HLL_A) {
        hll_smem = (size_t)(hll_planes ? 1 : 2) * (64 - c->aux_len + 2) * 64 * sizeof(uint32_t);
        static bool carve = false;
        if (!carve) {
            cudaFuncSetAttribute(k_tile_filter_hll<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_tile_filter_hll<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_tile_filter_hll_planes<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_tile_filter_hll_planes<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_hll_verify<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_hll_verify<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carve = true;
        }
        int per_sm = 0;
        const cudaError_t oe = hll_twopass ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hll_verify<0>, 64, hll_smem)
                               : hll_planes ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_filter_hll_planes<0>, 64, hll_smem)
                                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_filter_hll<0>, 64, hll_smem);
        if (oe != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 4;
        }
        hll_grid = c->sm_count * per_sm;
        static const int bound_per_sm = resident_ctas(k_tile_filter_hll_bound<0>, 64);
        hll_bound_grid = c->sm_count * bound_per_sm;
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
    // a sharded run reports its share of the CB band (stats.pairs_cb_shard) from two small tables of the bounds step; one
    // shard owns the whole band and copies nothing (the two copies went to pageable memory before: staged by the driver,
    // 15 us on the critical path of every run)
    const size_t band_off = (((size_t)nrb + 1) * 4 + 7) & ~(size_t)7;
    if (n_shards > 1) CKR(pinned_ensure(&c->h_band, &c->h_band_cap, band_off + (size_t)nrb * 8));
    int32_t* const h_tprefix = static_cast<int32_t*>(c->h_band);
    unsigned long long* const h_rb_pairs = n_shards > 1 ? reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(c->h_band) + band_off) : nullptr;
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
        k_gather_claim<<<1, 32, 0, s>>>(gz, c->g.epoch, d_cnt, check, smh_join ? ~0ull : (unsigned long long)PAIR_CAP, pair_lim,
                                        (unsigned long long)c->out_cap, (unsigned long long)c->tile_cap,
                                        (unsigned long long)c->near_cap, c->g_push.as<GatherPush>());
        CK(cudaGetLastError());
        k_gather_copy<<<32, 256, 0, s>>>(gz, c->g.epoch, d_cnt, (unsigned long long)c->near_cap, c->out_keys.as<uint64_t>(),
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
        CKR(c->near_keys.ensure((size_t)c->near_cap * 8));
        CKR(c->near_j.ensure((size_t)c->near_cap * 8));
        const unsigned long long near_cap = (unsigned long long)c->near_cap;
        st.launches = launches_fixed;
        if (!fused || attempt > 0) {
            k_tile_table<<<(nrb + 3) / 4, 128, 0, s>>>(c->tile_prefix.as<int32_t>(), c->tile_cb0.as<int32_t>(), nrb,
                                                       (long long)c->tile_cap, c->tile_rc.as<int2>(), d_cnt);
            CK(cudaGetLastError());
            st.launches++;
        }
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
        if (use_smh && smh_join) {
            // keys, ranks, members (u32 per (genome, band)), genome-major signatures, bucket counters and offsets
            CK(cudaStreamWaitEvent(side, c->fork_ev, 0));
            cudaEvent_t a0 = c->ev_on(side);
            CKR(c->join_buf.ensure((size_t)jn_words * 4));
            if (c->join_item_cap < (4ll << 20)) c->join_item_cap = 4ll << 20;
            CKR(c->join_items.ensure((size_t)c->join_item_cap * sizeof(uint4)));
            uint32_t* jk = c->join_buf.as<uint32_t>();
            uint32_t *j_keys = jk, *j_rank = jk + jn_keys, *j_memb = jk + 2 * jn_keys, *j_sig = jk + 3 * jn_keys;
            uint32_t *j_cnt = j_sig + (size_t)n * n_words, *j_off = j_cnt + (j_buckets + 1);
            CK(cudaMemsetAsync(j_cnt, 0, (size_t)(j_buckets + 1) * 4, side));
            const int grid = (int)std::min<int64_t>(((int64_t)n * n_words * 2 + 255) / 256, (int64_t)c->sm_count * 16);
            k_smh_sigkeys<<<grid, 256, 0, side>>>(c->aux_sorted.as<uint64_t>(), n, c->aux_len, n_rows, n_bands, n_words, j_sbits, j_keys,
                                               j_rank, j_cnt, j_sig);
            CK(cudaGetLastError());
            size_t tmp_bytes = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, j_cnt, j_off, (int)(j_buckets + 1), side));
            CKR(c->cub_tmp2.ensure(tmp_bytes));      // its own scratch: the main stream may be scanning too
            CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp2.p, tmp_bytes, j_cnt, j_off, (int)(j_buckets + 1), side));
            k_smh_scatter<<<(int)std::min<int64_t>((jn_keys + 255) / 256, (int64_t)c->sm_count * 16), 256, 0, side>>>(
                j_keys, j_rank, j_off, jn_keys, n_bands, j_memb);
            CK(cudaGetLastError());
            st.launches += 3;
            t_filter.push_back({a0, c->ev_on(side)});
            CK(cudaEventRecord(c->join_ev, side));
            CK(cudaStreamWaitEvent(s, c->join_ev, 0));
            DBG_SYNC(c, "smh signature keys + buckets");
        } else if (use_smh) {
            CK(cudaStreamWaitEvent(side, c->fork_ev, 0));
            cudaEvent_t a0 = c->ev_on(side);
            const size_t sig_bytes = (size_t)n_words * c->npad * 4;
            CKR(c->sigT.ensure(2 * sig_bytes));
            const int grid = (int)std::min<int64_t>(((int64_t)c->npad * n_words * 2 + 255) / 256, (int64_t)c->sm_count * 16);
            k_smh_signatures<<<grid, 256, 0, side>>>(c->aux_sorted.as<uint64_t>(), n, c->npad, c->aux_len, n_rows, n_bands,
                                                  c->sigT.as<uint32_t>(), c->sigT.as<uint32_t>() + (size_t)n_words * c->npad);
            CK(cudaGetLastError());
            st.launches++;
            t_filter.push_back({a0, c->ev_on(side)});
            CK(cudaEventRecord(c->join_ev, side));
            CK(cudaStreamWaitEvent(s, c->join_ev, 0));
            DBG_SYNC(c, "smh signatures");
        }
        if (attempt > 0) {          // the memset at the top of the run covers the first attempt
            CK(cudaMemsetAsync(d_cnt, 0, 32, s));
            CK(cudaMemsetAsync(d_cnt + M_PUSHED, 0, 8, s));
            CK(cudaMemsetAsync(d_cnt + M_STEPS, 0, 8, s));
            CK(cudaMemsetAsync(d_cnt + M_ITEMS_MAX, 0, 8, s));
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
            if (crit == SELB200_CRIT_SMH_A && smh_join) {
                // a tile range [j0, j1) of the shard's t_end tiles stands for the same fraction of the sorted keys (ranges
                // other than "everything" only exist after a pass whose lists overflowed, when t_end is known)
                long long s0 = 0, s1 = jn_keys;
                if (rg.second != INT32_MAX) {
                    const long long t_end = std::max<long long>(1, shard_tiles(tiles_total));
                    s0 = jn_keys * std::min<long long>(rg.first, t_end) / t_end;
                    s1 = jn_keys * std::min<long long>(rg.second, t_end) / t_end;
                }
                const uint32_t* jk = c->join_buf.as<uint32_t>();
                const uint32_t *j_keys = jk, *j_memb = jk + 2 * jn_keys, *j_sig = jk + 3 * jn_keys;
                const uint32_t* j_off = j_sig + (size_t)n * n_words + (j_buckets + 1);
                CK(cudaMemsetAsync(d_cnt + M_ITEMS, 0, 8, s));
                const long long j_elems = n_shards > 1 ? (jn_keys + n_shards - 1) / n_shards + n_bands : s1 - s0;
                const int grid = (int)std::max<long long>(1, std::min<long long>((j_elems + 255) / 256, (long long)c->sm_count * 16));
                k_smh_join_expand<<<grid, 256, 0, s>>>(j_keys, j_off, j_memb, s0, s1, n_bands, j_sbits, c->lo.as<int32_t>(),
                                                       c->hi.as<int32_t>(), c->join_items.as<uint4>(), d_cnt + M_ITEMS,
                                                       (unsigned long long)c->join_item_cap, prm->shard, n_shards, (long long)n);
                CK(cudaGetLastError());
                k_smh_join<<<c->sm_count * 16, 256, 0, s>>>(c->join_items.as<uint4>(), d_cnt + M_ITEMS, (unsigned long long)c->join_item_cap,
                                                            j_sig, n_words, j_sbits, c->aux_sorted.as<uint64_t>(), c->aux_len, n_rows,
                                                            n_bands, c->pairs.as<uint2>(), d_cnt + M_PAIRS,
                                                            (unsigned long long)PAIR_CAP, d_cnt + M_CAND, d_cnt + M_ITEMS_MAX);
                st.launches += 1;
            } else if (crit == SELB200_CRIT_SMH_A) {
                const int grid = (int)std::min<int64_t>(nt, (int64_t)c->sm_count * smh_grid_per_sm);
                k_tile_filter_smh<<<grid, 256, 0, s>>>(
                    c->sigT.as<uint32_t>(), c->sigT.as<uint32_t>() + (size_t)n_words * c->npad, c->npad, n_words, tw,
                    c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->cand.as<uint2>(), d_cnt + M_CAND,
                    (unsigned long long)PAIR_CAP);
            } else if (crit == SELB200_CRIT_CB) {
                const int grid = (int)std::min<int64_t>(nt, (int64_t)c->sm_count * 8);
                k_tile_enum<<<grid, 256, 0, s>>>(tw, c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->pairs.as<uint2>(),
                                                 d_cnt + M_PAIRS, (unsigned long long)PAIR_CAP);
            } else if (hll_twopass) {
                const int grid = (int)std::min<int64_t>(nt * 4, (int64_t)hll_bound_grid);
                if (crit == SELB200_CRIT_HLL_A)
                    k_tile_filter_hll_bound<0><<<grid, 64, 0, s>>>(
                        c->auxP.as<uint32_t>(), c->agrange.as<uint16_t>(), c->atail.as<AuxTail>(), c->npad, c->aux_len, tw,
                        c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->e_sorted.as<unsigned long long>(), (float)tau, zs,
                        prm->order_n, c->cand.as<uint2>(), d_cnt + M_CAND, (unsigned long long)PAIR_CAP, d_cnt + M_UNIT);
                else
                    k_tile_filter_hll_bound<1><<<grid, 64, 0, s>>>(
                        c->auxP.as<uint32_t>(), c->agrange.as<uint16_t>(), c->atail.as<AuxTail>(), c->npad, c->aux_len, tw,
                        c->lo.as<int32_t>(), c->hi.as<int32_t>(), n, c->e_sorted.as<unsigned long long>(), (float)tau, zs,
                        prm->order_n, c->cand.as<uint2>(), d_cnt + M_CAND, (unsigned long long)PAIR_CAP, d_cnt + M_UNIT);
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
            if (hll_twopass) {
                if (crit == SELB200_CRIT_HLL_A)
                    k_hll_verify<0><<<hll_grid, 64, hll_smem, s>>>(
                        c->auxP.as<uint32_t>(), c->agrange.as<uint16_t>(), c->auxT.as<uint32_t>(), c->npad, c->aux_len,
                        c->cand.as<uint2>(), d_cnt + M_CAND, (unsigned long long)PAIR_CAP, n, c->e_sorted.as<unsigned long long>(),
                        tau, zs, prm->order_n, c->pairs.as<uint2>(), d_cnt + M_PAIRS, (unsigned long long)PAIR_CAP);
                else
                    k_hll_verify<1><<<hll_grid, 64, hll_smem, s>>>(
                        c->auxP.as<uint32_t>(), c->agrange.as<uint16_t>(), c->auxT.as<uint32_t>(), c->npad, c->aux_len,
                        c->cand.as<uint2>(), d_cnt + M_CAND, (unsigned long long)PAIR_CAP, n, c->e_sorted.as<unsigned long long>(),
                        tau, zs, prm->order_n, c->pairs.as<uint2>(), d_cnt + M_PAIRS, (unsigned long long)PAIR_CAP);
                CK(cudaGetLastError());
                st.launches++;
            }
            if (crit == SELB200_CRIT_SMH_A && !smh_join) {
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
            bool wide_on_side = false;
            if ((crit == SELB200_CRIT_SMH_A && !smh_join) || hll_twopass) t_verify.push_back({f1, u0});
            if (union_bytes) {
                CKR(launch_pair_hist(c, c->d_regs, c->m, c->p, c->order_dev.as<int32_t>(), c->pairs.as<uint2>(),
                                     (int64_t)pair_lim, c->hist.as<uint32_t>(), d_cnt + M_PAIRS));
                st.launches++;
            } else {
                // SELB200_WIDE=side: the wide pairs' byte kernel and an estimate of their own on the side stream, the main
                // estimate skipping the rows the plane kernel stamped.  Measured: the side chain (22 + 17 us) does not run
                // beside the main estimate but behind it (0.138 against 0.099 ms for the estimate phase), so the default
                // is the byte kernel on the run stream, behind the plane kernel, and one estimate over every pair.
                static const bool wide_side = [] { const char* e = getenv("SELB200_WIDE"); return e && !strcmp(e, "side"); }();
                CKR(launch_pair_hist_planes(c, c->pairs.as<uint2>(), (int64_t)pair_lim, c->hist.as<uint32_t>(),
                                            d_cnt + M_PAIRS, d_cnt + M_WIDE, &st.launches, attempt == 0 && ri == 0,
                                            wide_side ? side : nullptr));
                if (wide_side) {
                    k_estimate_emit<<<c->sm_count, 128, 0, side>>>(
                        c->hist.as<uint32_t>(), c->pairs.as<uint2>(), c->wide_list.as<uint32_t>(), d_cnt + M_WIDE, pair_lim,
                        c->e_sorted.as<unsigned long long>(), c->p, tau, c->out_keys.as<uint64_t>(), c->out_j.as<double>(),
                        d_cnt + M_OUT, (unsigned long long)c->out_cap, c->near_keys.as<uint64_t>(), c->near_j.as<double>(),
                        d_cnt + M_NEAR, near_cap);
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(c->join_ev, side));
                    st.launches++;
                    wide_on_side = true;
                }
            }
            DBG_SYNC(c, "union histogram (planes + wide)");
            cudaEvent_t u1 = c->ev();
            if (crit == SELB200_CRIT_CB) {
                // unfiltered list: screen first (the survivor list borrows the candidate list's buffer, unused in this mode)
                if (attempt > 0 || ri > 0) CK(cudaMemsetAsync(d_cnt + M_SURV, 0, 8, s));
                k_estimate_screen<<<c->sm_count * 8, 128, 0, s>>>(
                    c->hist.as<uint32_t>(), c->pairs.as<uint2>(), d_cnt + M_PAIRS, pair_lim,
                    c->e_sorted.as<unsigned long long>(), c->p, tau, c->cand.as<uint32_t>(), d_cnt + M_SURV,
                    wide_on_side ? c->wide_flag.as<uint32_t>() : nullptr, c->wide_epoch);
                CK(cudaGetLastError());
                k_estimate_emit<<<c->sm_count * 4, 128, 0, s>>>(
                    c->hist.as<uint32_t>(), c->pairs.as<uint2>(), c->cand.as<uint32_t>(), d_cnt + M_SURV, pair_lim,
                    c->e_sorted.as<unsigned long long>(), c->p, tau, c->out_keys.as<uint64_t>(), c->out_j.as<double>(),
                    d_cnt + M_OUT, (unsigned long long)c->out_cap, c->near_keys.as<uint64_t>(), c->near_j.as<double>(),
                    d_cnt + M_NEAR, near_cap);
                CK(cudaGetLastError());
                st.launches += 2;
            } else {
                k_estimate_emit<<<c->sm_count * 8, 128, 0, s>>>(
                    c->hist.as<uint32_t>(), c->pairs.as<uint2>(), nullptr, d_cnt + M_PAIRS, pair_lim,
                    c->e_sorted.as<unsigned long long>(), c->p, tau, c->out_keys.as<uint64_t>(), c->out_j.as<double>(),
                    d_cnt + M_OUT, (unsigned long long)c->out_cap, c->near_keys.as<uint64_t>(), c->near_j.as<double>(),
                    d_cnt + M_NEAR, near_cap, wide_on_side ? c->wide_flag.as<uint32_t>() : nullptr, c->wide_epoch);
                CK(cudaGetLastError());
                st.launches++;
            }
            if (wide_on_side) CK(cudaStreamWaitEvent(s, c->join_ev, 0));     // the wide pairs' results are in the lists too
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
        if (n_shards > 1) {
            CK(cudaMemcpyAsync(h_tprefix, c->tile_prefix.p, ((size_t)nrb + 1) * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(h_rb_pairs, c->rb_pairs.p, (size_t)nrb * 8, cudaMemcpyDeviceToHost, s));
        }
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
            // the join only counts its candidates: no list to overflow
            const bool too_many = (!smh_join && sn[0] > (unsigned long long)PAIR_CAP) || sn[1] > (unsigned long long)PAIR_CAP;
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
            cand_sum += crit == SELB200_CRIT_SMH_A || hll_twopass ? (int64_t)sn[0] : (int64_t)sn[1];
            pair_sum += (int64_t)sn[1];
        }
        if ((int64_t)h_fin[M_OUT] > c->out_cap) { c->out_cap = (int64_t)h_fin[M_OUT] + (1 << 16); redo = true; }
        // the join's item list: the largest count of any range of the pass
        if (smh_join && (int64_t)h_fin[M_ITEMS_MAX] > c->join_item_cap) {
            c->join_item_cap = (int64_t)h_fin[M_ITEMS_MAX] + (int64_t)h_fin[M_ITEMS_MAX] / 4;
            redo = true;
        }
        // the near-tau list is part of the result (north_star: "listed separately"): never truncated, grown like the others
        if ((int64_t)h_fin[M_NEAR] > c->near_cap) { c->near_cap = (int64_t)h_fin[M_NEAR] + (1 << 12); redo = true; }
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
    st.filter_steps = (int64_t)h_fin[M_STEPS];
    st.tiles_total = tiles_total;
    st.tiles_shard = shard_tiles(tiles_total);
    c->out_count = (int64_t)h_fin[M_OUT];
    c->near_count = (int64_t)h_fin[M_NEAR];
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
            if (c->g.h_merged[1] > gz.near_cap)
                return fail(SELB200_ENOMEM, "gather: %llu near-tau pairs exceed the landing zone's %llu (selb200_gather_create)",
                            c->g.h_merged[1], gz.near_cap);
            c->out_count = (int64_t)c->g.h_merged[0];
            c->near_count = (int64_t)c->g.h_merged[1];
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
            if (fused) {
                static const int sgrid_cap = coop_grid(k_rowsort_fused, 256, c->sm_count, 1 << 30);
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(sgrid_cap, (std::max<int64_t>(cnt, n + 1) + 255) / 256));
                CKR(c->sort_blocksum.ensure((size_t)grid * 4));
                const uint64_t* a_keys = src_keys;
                const double* a_j = src_j;
                long long a_cnt = cnt;
                int a_n = n;
                int32_t *a_rc = c->row_cnt.as<int32_t>(), *a_ro = c->row_off.as<int32_t>(), *a_bs = c->sort_blocksum.as<int32_t>();
                uint64_t* a_ok = c->out_keys2.as<uint64_t>();
                double* a_oj = c->out_j2.as<double>();
                void* args[] = {&a_keys, &a_j, &a_cnt, &a_n, &a_rc, &a_ro, &a_bs, &tkeys, &tj, &a_ok, &a_oj};
                CK(cudaLaunchCooperativeKernel((void*)k_rowsort_fused, dim3(grid), dim3(256), args, 0, s));
                st.launches += 1;
            } else {
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
            }
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
    if (n_shards <= 1) {
        st.pairs_cb_shard = st.pairs_cb;
    } else {
        double acc = 0.;
        for (int rb = 0; rb < nrb; ++rb) {
            const int a = h_tprefix[(size_t)rb], b = h_tprefix[(size_t)rb + 1];
            if (b <= a) continue;
            // tiles t in [a,b) with t % n_shards == shard
            const int first = a + ((prm->shard - a % n_shards) % n_shards + n_shards) % n_shards;
            const int mine = first < b ? (b - 1 - first) / n_shards + 1 : 0;
            acc += (double)h_rb_pairs[(size_t)rb] * (double)mine / (double)(b - a);
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
static_assert(sizeof(selb200_stats) == 160, "selb200_stats is part of the C ABI: fields are added inside the reserved words");

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
