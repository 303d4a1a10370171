// helpers.inl — device helpers shared by the selection kernels (part of selb200.cu, inside its anonymous namespace)

#ifndef SELB_EMUL   // byte-histogram helpers (inline PTX) and warp_claim: not used / replaced on the CPU emulator (tests/emul)
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

#endif   // SELB_EMUL

__device__ __forceinline__ void hist_inc_max16(const uint4& x, const uint4& y, uint32_t bias, uint32_t tb) {
    uint32_t w;
    w = max4_lt128(x.x, y.x) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
    w = max4_lt128(x.y, y.y) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
    w = max4_lt128(x.z, y.z) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
    w = max4_lt128(x.w, y.w) + bias; hist_inc2<0, 1>(w, tb); hist_inc2<2, 3>(w, tb);
}

#ifndef SELB_EMUL
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

#endif   // SELB_EMUL

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
