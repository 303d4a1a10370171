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

#include "kernels/sketch_kernels.inl"

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
