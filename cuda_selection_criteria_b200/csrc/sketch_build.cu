// sketch_build.cu — B200 build of the sketches `build_sketch` writes.
//
// Reference (paths relative to the reference tree):
//   src/build_sketch.cpp:26-39   canonical_kmer (k = 31, 2-bit bases, min of forward / reverse complement)
//   src/build_sketch.cpp:61-92   rolling k-mer, restarted by any non-ACGT character and by every record
//   sketch/include/sketch/hash.h:44-53    WangHash
//   sketch/include/sketch/hll.h:886-894   hll_t::add: index = top p bits, value = clz(low bits) + 1, max
//   sketch/include/sketch/bbmh.h:639-670  SuperMinHash::addh, RNG wy::WyHash<uint32_t,1>
//                                         (sketch/include/aesctr/wy.h:53-56,98-150)
//
// One CTA per genome.  The HLL registers (primary and optional auxiliary) live in shared memory as
// 32-bit cells updated with atomicMax; the SuperMinHash buckets as 64-bit cells updated with
// atomicMin.  Both reductions are order-independent, so the result equals the reference's
// sequential loop bit for bit:
//   * HLL: register = max over k-mers of the rank.
//   * SuperMinHash: step j of element x draws v_j = wyhash(x + (j+1)*C) (a counter-based stream:
//     k = low32 & (m-1), r = high32), swaps p[j] and p[k] of the element's own lazily reset
//     permutation and offers (j<<32 | r) to bucket p[j]; the bucket keeps the minimum.  The
//     reference stops an element at j > a, a = the largest step index stored in any bucket, because
//     later offers cannot win; any OLDER (larger) value of that bound is also safe, so the CTA
//     refreshes it once per tile instead of per element.
#include "../../include/selb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

namespace {

constexpr int KMER = 31;
constexpr int SK_THREADS = 512;
constexpr int SK_SPAN = 32;          // positions per thread per tile
constexpr int MAX_SMH = 1024;        // buckets the per-thread permutation scratch is sized for

__device__ __forceinline__ int base_code(uint8_t ch) {
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return -1;
}

__device__ __forceinline__ uint64_t canonical_kmer(uint64_t kmer) {
    uint64_t r = kmer;
    r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
    r = ((r >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((r & 0x0F0F0F0F0F0F0F0Full) << 4);
    r = ((r >> 8) & 0x00FF00FF00FF00FFull) | ((r & 0x00FF00FF00FF00FFull) << 8);
    r = ((r >> 16) & 0x0000FFFF0000FFFFull) | ((r & 0x0000FFFF0000FFFFull) << 16);
    r = (r >> 32) | (r << 32);
    const uint64_t rev = (~r) >> (64 - 2 * KMER);
    return kmer < rev ? kmer : rev;
}

__device__ __forceinline__ uint64_t wang_hash(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

__device__ __forceinline__ void hll_offer(uint32_t* regs, int p, uint64_t h) {
    const uint32_t index = (uint32_t)(h >> (64 - p));
    const uint32_t rank = (uint32_t)__clzll((long long)(((h << 1) | 1ull) << (p - 1))) + 1u;
    if (regs[index] < rank) atomicMax(regs + index, rank);
}

// One element into the shared SuperMinHash buckets.  `perm` is this thread's permutation scratch
// (identity between elements), `undo` records the touched slots so that it can be restored.
template <typename PermT>
__device__ __forceinline__ void smh_offer(unsigned long long* buckets, uint32_t m, uint32_t bound, uint64_t item,
                                          PermT* perm, PermT* undo) {
    uint64_t state = item ? item : 1337ull;
    uint32_t j = 0;
    while (j <= bound) {
        state += 0x60bee2bee120fc15ull;
        const uint64_t x = state ^ 0xe7037ed1a0b428dbull;
        const uint64_t v = (x * state) ^ __umul64hi(x, state);
        const uint32_t k = (uint32_t)v & (m - 1u);
        const uint32_t r = (uint32_t)(v >> 32);
        const PermT pj = perm[j], pk = perm[k];
        perm[j] = pk;
        perm[k] = pj;
        undo[j] = (PermT)k;
        const unsigned long long crj = ((unsigned long long)j << 32) | r;
        if (crj < buckets[pk]) atomicMin(buckets + pk, crj);
        ++j;
    }
    // restore the identity: slots 0..j-1 and every recorded partner
    for (uint32_t t = 0; t < j; ++t) {
        const uint32_t k = undo[t];
        perm[k] = (PermT)k;
        perm[t] = (PermT)t;
    }
}

template <typename PermT>
__global__ void __launch_bounds__(SK_THREADS)
k_sketch_build(const uint8_t* __restrict__ seq, const long long* __restrict__ offsets, int p, int aux_kind,
               int aux_len, uint8_t* __restrict__ out_hll, uint8_t* __restrict__ out_aux_hll,
               unsigned long long* __restrict__ out_smh, PermT* __restrict__ perm_scratch) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int g = blockIdx.x;
    const long long s0 = offsets[g], s1 = offsets[g + 1];
    const uint32_t m_hll = 1u << p;
    const uint32_t m_aux = aux_kind == SELB200_AUX_HLL ? (1u << aux_len) : 0u;
    const uint32_t m_smh = aux_kind == SELB200_AUX_SMH ? (uint32_t)aux_len : 0u;
    uint32_t* regs = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* regs_aux = regs + m_hll;
    unsigned long long* buckets = reinterpret_cast<unsigned long long*>(regs_aux + m_aux);
    __shared__ uint32_t s_bound;

    for (uint32_t i = threadIdx.x; i < m_hll; i += SK_THREADS) regs[i] = 0;
    for (uint32_t i = threadIdx.x; i < m_aux; i += SK_THREADS) regs_aux[i] = 0;
    for (uint32_t i = threadIdx.x; i < m_smh; i += SK_THREADS) buckets[i] = ~0ull;
    if (threadIdx.x == 0) s_bound = m_smh ? m_smh - 1 : 0;
    PermT* perm = nullptr;
    PermT* undo = nullptr;
    if (m_smh) {
        perm = perm_scratch + ((size_t)blockIdx.x * SK_THREADS + threadIdx.x) * 2 * m_smh;
        undo = perm + m_smh;
        for (uint32_t i = 0; i < m_smh; ++i) perm[i] = (PermT)i;
    }
    __syncthreads();

    const uint64_t kmask = (1ull << (2 * KMER)) - 1;
    for (long long tile = s0; tile < s1; tile += (long long)SK_THREADS * SK_SPAN) {
        const long long first = tile + (long long)threadIdx.x * SK_SPAN;     // first position of this thread
        if (first < s1) {
            // warm the rolling state on the 30 bases before `first` (never before the genome start)
            uint64_t kmer = 0;
            uint32_t run = 0;                                              // valid bases ending here, capped at 31
            const long long w0 = first - (KMER - 1) > s0 ? first - (KMER - 1) : s0;
            for (long long i = w0; i < first; ++i) {
                const int c = base_code(__ldg(seq + i));
                if (c < 0) { run = 0; kmer = 0; } else { kmer = ((kmer << 2) | (uint64_t)c) & kmask; run = run < KMER ? run + 1 : run; }
            }
            const long long last = first + SK_SPAN < s1 ? first + SK_SPAN : s1;
            const uint32_t bound = m_smh ? s_bound : 0;
            for (long long i = first; i < last; ++i) {
                const int c = base_code(__ldg(seq + i));
                if (c < 0) { run = 0; kmer = 0; continue; }
                kmer = ((kmer << 2) | (uint64_t)c) & kmask;
                run = run < KMER ? run + 1 : run;
                if (run < KMER) continue;
                const uint64_t item = canonical_kmer(kmer);
                const uint64_t h = wang_hash(item);
                hll_offer(regs, p, h);
                if (m_aux) hll_offer(regs_aux, aux_len, h);
                if (m_smh) smh_offer<PermT>(buckets, m_smh, bound, item, perm, undo);
            }
        }
        if (m_smh) {
            // refresh the bound: largest step index held by any bucket (empty buckets count as m-1)
            __syncthreads();
            if (threadIdx.x < 32) {
                uint32_t mx = 0;
                for (uint32_t i = threadIdx.x; i < m_smh; i += 32) {
                    const uint32_t jb = (uint32_t)min((unsigned long long)(m_smh - 1), buckets[i] >> 32);
                    mx = max(mx, jb);
                }
                for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if (threadIdx.x == 0) s_bound = mx;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m_hll; i += SK_THREADS) out_hll[(size_t)g * m_hll + i] = (uint8_t)regs[i];
    for (uint32_t i = threadIdx.x; i < m_aux; i += SK_THREADS) out_aux_hll[(size_t)g * m_aux + i] = (uint8_t)regs_aux[i];
    for (uint32_t i = threadIdx.x; i < m_smh; i += SK_THREADS) out_smh[(size_t)g * m_smh + i] = buckets[i];
}

struct Buf {
    void* p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
};

thread_local char g_sk_err[512];

}  // namespace

extern "C" {

const char* selb200_sketch_last_error(void) { return g_sk_err; }

// Creates the CUDA context of `device` ahead of the first real call; a host driver runs it on a
// side thread while it is still reading files.
int selb200_warmup(int device) {
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) {
        cudaGetLastError();
        return SELB200_ECUDA;
    }
    return SELB200_OK;
}

int selb200_smh_size(int m_arg) {           // SizePow2Policy (sketch/include/sketch/policy.h:14-19)
    if (m_arg < 1) return 0;
    int lg = 0;
    while ((1 << (lg + 1)) <= m_arg) ++lg;
    if (m_arg & (m_arg - 1)) ++lg;
    return 1 << lg;
}

int selb200_sketch_host(int device, int64_t n_genomes, const uint8_t* seq, const int64_t* offsets, int p,
                        int aux_kind, int aux_len, uint8_t* out_hll, void* out_aux) {
#define SK_FAIL(code, ...) do { snprintf(g_sk_err, sizeof g_sk_err, __VA_ARGS__); return code; } while (0)
#define SK_CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) SK_FAIL(SELB200_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); } while (0)
    if (n_genomes < 0 || (n_genomes && (!seq || !offsets || !out_hll))) SK_FAIL(SELB200_EINVAL, "null argument");
    if (p < 4 || p > 14) SK_FAIL(SELB200_EINVAL, "HLL precision %d unsupported by the builder (4..14)", p);
    int m_smh = 0;
    if (aux_kind == SELB200_AUX_SMH) {
        m_smh = selb200_smh_size(aux_len);
        if (m_smh < 1 || m_smh > MAX_SMH) SK_FAIL(SELB200_EINVAL, "SuperMinHash size %d unsupported (1..%d)", aux_len, MAX_SMH);
    } else if (aux_kind == SELB200_AUX_HLL) {
        if (aux_len < 4 || aux_len > 13) SK_FAIL(SELB200_EINVAL, "aux HLL precision %d unsupported (4..13)", aux_len);
    } else if (aux_kind != SELB200_AUX_NONE) {
        SK_FAIL(SELB200_EINVAL, "unknown aux kind %d", aux_kind);
    }
    if (aux_kind != SELB200_AUX_NONE && n_genomes && !out_aux) SK_FAIL(SELB200_EINVAL, "null aux output");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        SK_FAIL(SELB200_ECUDA, "no CUDA device available; the sketch builder has no CPU path");
    }
    SK_CK(cudaSetDevice(device));
    if (n_genomes == 0) return SELB200_OK;
    const size_t total = (size_t)offsets[n_genomes];
    const size_t m_hll = (size_t)1 << p;
    const size_t m_aux = aux_kind == SELB200_AUX_HLL ? (size_t)1 << aux_len : 0;
    Buf d_seq, d_off, d_hll, d_aux, d_perm;
    SK_CK(d_seq.alloc(total));
    SK_CK(d_off.alloc((size_t)(n_genomes + 1) * 8));
    SK_CK(d_hll.alloc((size_t)n_genomes * m_hll));
    SK_CK(d_aux.alloc(aux_kind == SELB200_AUX_HLL ? (size_t)n_genomes * m_aux : (size_t)n_genomes * m_smh * 8));
    SK_CK(cudaMemcpy(d_seq.p, seq, total, cudaMemcpyHostToDevice));
    SK_CK(cudaMemcpy(d_off.p, offsets, (size_t)(n_genomes + 1) * 8, cudaMemcpyHostToDevice));
    const size_t smem = (m_hll + m_aux) * 4 + (size_t)m_smh * 8;
    // genomes are processed in waves so that the permutation scratch stays bounded
    const int wave = 592;   // 4 CTAs per SM on 148 SMs
    const bool wide = m_smh > 256;
    if (m_smh) SK_CK(d_perm.alloc((size_t)wave * SK_THREADS * 2 * m_smh * (wide ? 2 : 1)));
    if (wide) SK_CK(cudaFuncSetAttribute(k_sketch_build<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SK_CK(cudaFuncSetAttribute(k_sketch_build<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t g0 = 0; g0 < n_genomes; g0 += wave) {
        const int cnt = (int)std::min<int64_t>(wave, n_genomes - g0);
        uint8_t* o_hll = (uint8_t*)d_hll.p + (size_t)g0 * m_hll;
        uint8_t* o_auxh = aux_kind == SELB200_AUX_HLL ? (uint8_t*)d_aux.p + (size_t)g0 * m_aux : nullptr;
        unsigned long long* o_smh = aux_kind == SELB200_AUX_SMH ? (unsigned long long*)d_aux.p + (size_t)g0 * m_smh : nullptr;
        if (wide)
            k_sketch_build<uint16_t><<<cnt, SK_THREADS, smem>>>((const uint8_t*)d_seq.p, (const long long*)d_off.p + g0, p,
                                                               aux_kind, aux_kind == SELB200_AUX_SMH ? m_smh : aux_len,
                                                               o_hll, o_auxh, o_smh, (uint16_t*)d_perm.p);
        else
            k_sketch_build<uint8_t><<<cnt, SK_THREADS, smem>>>((const uint8_t*)d_seq.p, (const long long*)d_off.p + g0, p,
                                                              aux_kind, aux_kind == SELB200_AUX_SMH ? m_smh : aux_len,
                                                              o_hll, o_auxh, o_smh, (uint8_t*)d_perm.p);
        SK_CK(cudaGetLastError());
    }
    SK_CK(cudaMemcpy(out_hll, d_hll.p, (size_t)n_genomes * m_hll, cudaMemcpyDeviceToHost));
    if (aux_kind == SELB200_AUX_HLL) SK_CK(cudaMemcpy(out_aux, d_aux.p, (size_t)n_genomes * m_aux, cudaMemcpyDeviceToHost));
    if (aux_kind == SELB200_AUX_SMH) SK_CK(cudaMemcpy(out_aux, d_aux.p, (size_t)n_genomes * m_smh * 8, cudaMemcpyDeviceToHost));
    return SELB200_OK;
#undef SK_FAIL
#undef SK_CK
}

}  // extern "C"
