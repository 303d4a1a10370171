// synth.cuh — "synth-v1" synthetic sketch generator, integer-only so that the host
// loop and the device kernel produce identical bytes.
//
// Model: genomes come in clusters.  A member's k-mer set is core(cluster) ∪ private(member),
// disjoint, so
//   * its HLL registers are byte-max(core registers, private registers) — exactly the
//     sketch of a disjoint union (sketch/include/sketch/hll.h:886-894 keeps the max rank
//     per register), each part drawn from the exact register law
//     P(reg <= k) = exp(-(n/m)·2^-k), handed in as 64 thresholds T[k] = floor(2^64·P(reg<=k));
//   * its SuperMinHash buckets are element-wise min(core, private)
//     (sketch/include/sketch/bbmh.h:656-660 keeps the minimum per bucket), each part a
//     uniform draw below R ≈ 2^33/(n/m) (mean of the minimum of n/m uniform 32-bit values).
// Streams are counter-based (two rounds of the splitmix64 finaliser), keyed by
// (seed, tag, cluster-or-genome id, register index).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SELB_SYN_HD __host__ __device__ __forceinline__
#else
#define SELB_SYN_HD inline
#endif

namespace selb {

SELB_SYN_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

SELB_SYN_HD uint64_t synth_u64(uint64_t seed, uint32_t tag, uint32_t part, uint64_t id, uint64_t j) {
    uint64_t x = mix64(seed + 0x9E3779B97F4A7C15ull * (((uint64_t)tag << 8) | part));
    x = mix64(x ^ (id * 0xD6E8FEB86659FD93ull));
    return mix64(x + j * 0x9E3779B97F4A7C15ull);
}

// number of k in [0,64) with u >= T[k]  (T non-decreasing): smallest register value r
// with u < T[r]; T is padded with UINT64_MAX so the result never exceeds q+1.
SELB_SYN_HD uint32_t synth_reg(const uint64_t* T, uint64_t u) {
    uint32_t lo = 0, hi = 64;          // first index with u < T[idx], or 64
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (u >= T[mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

SELB_SYN_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

SELB_SYN_HD uint8_t synth_hll_reg(uint64_t seed, uint32_t tag, int64_t genome, int32_t cluster, uint64_t j,
                                  const uint64_t* thr_core, const uint64_t* thr_priv, uint32_t vmax) {
    const uint32_t rc = synth_reg(thr_core + (size_t)cluster * 64, synth_u64(seed, tag, 1, (uint64_t)cluster, j));
    const uint32_t rp = synth_reg(thr_priv + (size_t)genome * 64, synth_u64(seed, tag, 2, (uint64_t)genome, j));
    uint32_t r = rc > rp ? rc : rp;
    if (r > vmax) r = vmax;
    return (uint8_t)r;
}

SELB_SYN_HD uint64_t synth_smh_bucket(uint64_t seed, uint32_t tag, int64_t genome, int32_t cluster, uint64_t j,
                                      const uint64_t* range_core, const uint64_t* range_priv) {
    // range 0 = that part of the set is empty: its bucket is "never filled" (bbmh.h:566)
    const uint64_t rc = range_core[cluster], rp = range_priv[genome];
    const uint64_t vc = rc ? mulhi64(synth_u64(seed, tag, 3, (uint64_t)cluster, j), rc) : ~0ull;
    const uint64_t vp = rp ? mulhi64(synth_u64(seed, tag, 4, (uint64_t)genome, j), rp) : ~0ull;
    return vc < vp ? vc : vp;
}

}  // namespace selb
