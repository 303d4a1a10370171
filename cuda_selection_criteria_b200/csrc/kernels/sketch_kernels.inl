// sketch_kernels.inl — device code of the sketch builder (part of sketch_build.cu, inside its anonymous namespace;
// also compiled as host code by tests/emul/emul_sketch.cpp).  See sketch_build.cu for the algorithm and references.
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
#ifndef SELB_EMUL   // the emulator's dynamic shared memory is a global array of this name
    extern __shared__ __align__(16) uint8_t smem_raw[];
#endif
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
