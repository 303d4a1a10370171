// union_planes.inl — K5, bit-plane form: LOP3 max / decode / carry-save counting, TMA staging (part of selb200.cu)
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
// one warp per CTA, 16 CTAs per SM.
// Pairs whose value range does not fit a 32-value window (a register of 32 or more next to a smallest register below 8:
// one genome in 1 700 of the bench workload has such a register, about 0.1 % of the pairs) go to a "wide" list and
// through the byte kernel.
// layout: genome g at planes + g * 6 * m/8 bytes; chunk c (PL_CHUNK_REGS registers, or m if smaller) holds its
//         6 planes back to back: [chunk][plane][chunk_regs/32 words], bit r of word w = register 32w+r
// ============================================================================
// chunk / ring geometry, measured at n=100k (511 521 pairs), registers x stages x CTAs/SM:
//   2048 x 4 x 16: 1.94 ms   4096 x 3 x 12: 2.02   8192 x 2 x 9: 1.86   8192 x 2 x 8: 1.97   8192 x 3 x 6: 2.38
//   2048 x 2 x 16: 1.93      2048 x 3 x 16: 1.92   4096 x 2 x 16: 1.78  <- kept: 16 warps per SM and half
//   the pipeline-control steps of the 2048-register chunks
#ifndef PL_CHUNK_REGS_V
#define PL_CHUNK_REGS_V 4096
#endif
constexpr int PL_CHUNK_REGS = PL_CHUNK_REGS_V;
constexpr int PL_NQ = PL_CHUNK_REGS / 64;   // uint2 per plane of a full chunk

__global__ void __launch_bounds__(256)
k_planes_from_bytes(const uint8_t* __restrict__ regs, long long rows, size_t m, int chunk_regs,
                    uint32_t* __restrict__ planes, uint32_t* __restrict__ gtop = nullptr) {
    // gtop (optional, zeroed by the caller, m >= 4096): gtop[8 g + j] = largest register of the j-th eighth of genome g —
    // the union kernel stops counting a 2048-register step at the group its eighths reach (see k_pair_hist_planes)
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
        if (gtop) {
            uint32_t mx = 0;
#pragma unroll
            for (int sh = 0; sh < 32; sh += 8)
                mx = max(max(mx, (v.x >> sh) & 0xffu), max(max((v.y >> sh) & 0xffu, (v.z >> sh) & 0xffu), (v.w >> sh) & 0xffu));
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) atomicMax(gtop + g * 8 + (long long)bg * 8 / blk_per_genome, mx);
        }
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

#ifndef SELB_EMUL   // tests/emul/cuda_emul.h supplies host versions of these five when the .inl runs on the CPU
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
#endif   // SELB_EMUL

// Experiment switches of the subset form (A/B builds, `make variants`):
//   PL_PACK_C2   the 32 carry counters of a lane in 16 registers (two 16-bit halves: a lane sees at most m/32 registers
//                of a pair, so a carry count stays below 2^15 up to p = 20): frees 16 registers
//   PL_UNROLL    the steps of a full chunk straight-line instead of a loop
#ifndef PL_PACK_C2
#define PL_PACK_C2 0
#endif
#ifndef PL_UNROLL
#define PL_UNROLL 1      // measured 1.546 against 1.566 ms (n = 100k, 511 521 pairs); packing the carry counters: no change
#endif
//   PL_DIRECT    (subset form) 4-bit mask over the four subset masks of a group: a mask with its bit set is counted with two
//                POPC and one three-input add instead of a carry-save step (2 LOP3 + POPC + add).  The carry-save form
//                loads the ALU pipe (LOP3, 16 lanes / clk / scheduler), POPC the XU pipe (4 lanes): with one mask of four
//                counted directly the two pipes carry about the same time per step, and the mask's state register goes
//   PL_BFLY16    epilogue butterfly on 16 words holding two 16-bit totals each (16 shuffles instead of 31); totals of
//                sixteen lanes stay below 2^16 while m <= 2^16, larger sketches take the 32-word butterfly
// measured (n = 100k, 511 521 pairs, same box): neither 1.533 ms, PL_DIRECT=1 1.490, PL_BFLY16 1.531, both 1.475 (kept);
// PL_DIRECT=9 (two masks of four counted directly) 1.570: the XU pipe becomes the limit
#ifndef PL_DIRECT
#define PL_DIRECT 1
#endif
#ifndef PL_BFLY16
#define PL_BFLY16 1
#endif
constexpr int PL_NC2 = PL_PACK_C2 ? 16 : 32;
__device__ __forceinline__ void c2_add(uint32_t (&C2)[PL_NC2], int v, uint32_t cnt) {
#if PL_PACK_C2
    if (v < 16) C2[v] += cnt;
    else C2[v - 16] = cnt * 65536u + C2[v - 16];
#else
    C2[v] += cnt;
#endif
}
__device__ __forceinline__ uint32_t c2_get(const uint32_t (&C2)[PL_NC2], int v) {
#if PL_PACK_C2
    return v < 16 ? (C2[v] & 0xffffu) : (C2[v - 16] >> 16);
#else
    return C2[v];
#endif
}

// One chunk (<= PL_CHUNK_REGS registers) against the running carry-save state.  Per step, lane q holds
// two consecutive words of every plane (LDS.64).  Written stage by stage over the 8 values of a group so
// that eight independent dependency chains are in flight (LOP3 latency 4 at one issue per 2 clocks).
template <int G0, int NQ>   // NQ > 0: uint2 per plane known at compile time (full 2048-register chunks)
__device__ __forceinline__ void plane_chunk(const uint2* __restrict__ sA, const uint2* __restrict__ sB, int nq_rt,
                                            int lane, uint32_t gmask, uint32_t (&S)[32], uint32_t (&C2)[PL_NC2]) {
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
                _Pragma("unroll") for (int j = 0; j < 8; ++j) c2_add(C2, T * 8 + j, (uint32_t)__popc(kk[j])); \
            }                                                                                             \
        }
        SELB_PLANE_GROUP(0, false)
        SELB_PLANE_GROUP(1, false)
        SELB_PLANE_GROUP(2, false)
        SELB_PLANE_GROUP(3, false)
#undef SELB_PLANE_GROUP
    }
}

// The same chunk step in SUBSET form (SELB200_UNION=subsets; written without GPU budget, CPU-checked on the warp
// emulator, to be measured in round 2).  The one-hot form above spends, per 32 registers, 8 decode masks + per
// 8-value group 1 selector + 8 ANDs before the 8 carry-save inputs.  Here the window is cut into eight groups of
// FOUR values (selected by planes 5..2) and a group counts the four SUBSET masks of the two low planes
//   c[{}] = sel,  c[{0}] = sel & M0,  c[{1}] = sel & M1,  c[{0,1}] = sel & M0 & M1        (one LOP3 each)
// i.e. #registers of the group whose low bits CONTAIN the subset; the epilogue turns the four totals back into the
// four bins with two subtract steps (Moebius inversion, exact in integers).  Every counted mask now costs one LOP3
// to form and one to absorb, the decode masks are gone, and groups of four follow the pair's value range more
// closely (a range 6..25 counts 24 bins instead of 32): per 32 registers and a range of 20 values
// 12 (max) + 4 + 6*4 (form) + 24 (absorb) = 64 LOP3 and 12 POPC instead of 88 and 16.
// gmask: bit t = the group of values 8*G0 + 4t .. 8*G0 + 4t + 3 lies inside the pair's range (warp-uniform).
// tops: three bits per eighth of the sketch, the last group of four (relative to the window) that the eighth's registers of
// either genome reach; step s of this chunk lies in eighth (step0 + s) >> tsh and counts the groups of gmask_all up to there
template <int G0, int NQ>
__device__ __forceinline__ void plane_chunk_subsets(const uint2* __restrict__ sA, const uint2* __restrict__ sB, int nq_rt,
                                                    int lane, uint32_t gmask_all, uint32_t (&S)[32], uint32_t (&C2)[PL_NC2],
                                                    uint32_t tops = 0x00ffffffu, int step0 = 0, int tsh = 0) {
    const int nq = NQ > 0 ? NQ : nq_rt;
    constexpr int NP = (G0 == 0) ? 5 : 6;
#if PL_UNROLL
#pragma unroll (NQ > 0 ? NQ / 32 : 1)
#else
#pragma unroll 1
#endif
    for (int q = lane; q < nq; q += 32) {
        const uint32_t gmask = gmask_all & ((2u << ((tops >> (3 * ((step0 + (q >> 5)) >> tsh))) & 7u)) - 1u);
        uint32_t M[2][6];
        {
            uint2 a[NP], b[NP];
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) { a[pl] = sA[pl * nq + q]; b[pl] = sB[pl * nq + q]; }
            uint32_t lt0 = 0u, lt1 = 0u;
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
                lt0 = lop3<0x8E>(a[pl].x, b[pl].x, lt0);
                lt1 = lop3<0x8E>(a[pl].y, b[pl].y, lt1);
            }
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {
                M[0][pl] = lop3<0xCA>(lt0, b[pl].x, a[pl].x);
                M[1][pl] = lop3<0xCA>(lt1, b[pl].y, a[pl].y);
            }
            if (NP == 5) { M[0][5] = 0u; M[1][5] = 0u; }
        }
        // T8: 8-value group of the window (selector from planes 5..3), HALF: its lower / upper four values (plane 2)
#define SELB_SUBSET_GROUP(T8, HALF)                                                                       \
        if (gmask & (1u << (2 * T8 + HALF))) {                                                            \
            constexpr int c0 = 4 * (2 * T8 + HALF);                                                       \
            const uint32_t e0 = HALF ? lop3<0xC0>(H0, M[0][2], 0u) : lop3<0x30>(H0, M[0][2], 0u);         \
            const uint32_t e1 = HALF ? lop3<0xC0>(H1, M[1][2], 0u) : lop3<0x30>(H1, M[1][2], 0u);         \
            uint32_t m0[4], m1[4], kk[4];                                                                 \
            m0[0] = e0; m0[1] = lop3<0xC0>(e0, M[0][0], 0u); m0[2] = lop3<0xC0>(e0, M[0][1], 0u);         \
            m0[3] = lop3<0x80>(e0, M[0][0], M[0][1]);                                                     \
            m1[0] = e1; m1[1] = lop3<0xC0>(e1, M[1][0], 0u); m1[2] = lop3<0xC0>(e1, M[1][1], 0u);         \
            m1[3] = lop3<0x80>(e1, M[1][0], M[1][1]);                                                     \
            _Pragma("unroll") for (int j = 0; j < 4; ++j)                                                 \
                if (!((PL_DIRECT >> j) & 1)) kk[j] = lop3<0xE8>(S[c0 + j], m0[j], m1[j]);                 \
            _Pragma("unroll") for (int j = 0; j < 4; ++j)                                                 \
                if (!((PL_DIRECT >> j) & 1)) S[c0 + j] = lop3<0x96>(S[c0 + j], m0[j], m1[j]);             \
            _Pragma("unroll") for (int j = 0; j < 4; ++j)                                                 \
                c2_add(C2, c0 + j, ((PL_DIRECT >> j) & 1) ? (uint32_t)(__popc(m0[j]) + __popc(m1[j]))     \
                                                          : (uint32_t)__popc(kk[j]));                     \
        }
#define SELB_SUBSET_GROUP8(T8)                                                                            \
        if (gmask & (3u << (2 * T8))) {                                                                   \
            const uint32_t H0 = lop3<(1 << (G0 + T8))>(M[0][5], M[0][4], M[0][3]);                        \
            const uint32_t H1 = lop3<(1 << (G0 + T8))>(M[1][5], M[1][4], M[1][3]);                        \
            SELB_SUBSET_GROUP(T8, 0)                                                                      \
            SELB_SUBSET_GROUP(T8, 1)                                                                      \
        }
        SELB_SUBSET_GROUP8(0)
        SELB_SUBSET_GROUP8(1)
        SELB_SUBSET_GROUP8(2)
        SELB_SUBSET_GROUP8(3)
#undef SELB_SUBSET_GROUP8
#undef SELB_SUBSET_GROUP
    }
}

// pair list -> histogram rows, bit-plane form.  grange[g] = min | max<<8 of genome g's registers.
// One warp per CTA.  Work comes in batches of 32 consecutive pairs claimed from a device counter (dynamic
// balance, no tail): each lane fetches the descriptor of one pair of the batch (rows through `order`, value
// window from grange), so the dependent global loads are paid once per 32 pairs and the warp then reads
// descriptors with shuffles.  The pairs' planes flow chunk by chunk (4096 registers = 2 x 3 KiB) through a
// ring of PL_STAGES shared-memory stages: lane 0 keeps PL_STAGES-1 bulk copies (TMA) in flight ahead of
// the chunk being counted, across pair and batch boundaries.
#ifndef PL_STAGES_V
#define PL_STAGES_V 2
#endif
constexpr int PL_STAGES = PL_STAGES_V;
#ifndef PL_MIN_CTAS
#define PL_MIN_CTAS 16
#endif

// The counting form is a property of the epilogue type, so that the default instantiation keeps its name and its code:
// k_pair_hist_planes<Epi> counts one-hot (plane_chunk), k_pair_hist_planes<EpiSubsets<Epi>> in subset form on groups of
// four values (plane_chunk_subsets, SELB200_UNION=subsets).  Everything but the counting step, the granularity of the
// range mask and two subtract steps in the epilogue is shared.
template <class E> struct EpiSubsets : E {};
template <class E> struct union_form { static constexpr int value = 0; };
template <class E> struct union_form<EpiSubsets<E>> { static constexpr int value = 1; };

template <class Epi>
__global__ void __launch_bounds__(32, PL_MIN_CTAS)
k_pair_hist_planes(const uint32_t* __restrict__ planes, size_t m, int chunk_regs, const uint16_t* __restrict__ grange,
                   SrcPairs src, Epi epi, uint32_t* __restrict__ wide_list, unsigned long long* __restrict__ wide_count,
                   unsigned long long* __restrict__ batch_counter, uint32_t* __restrict__ wide_flag = nullptr, uint32_t epoch = 0u,
                   const uint32_t* __restrict__ gtop = nullptr) {
    // gtop (optional, subset form, m >= 16384 so that an eighth is one or more whole 2048-register steps): the per-eighth
    // maxima written by k_planes_from_bytes.  High values are rare — 2048 registers seldom reach the top of the range of all
    // 2^p — so a step counts the groups up to the larger of the two genomes' maxima over its eighth instead of up to the
    // pair's maximum (C4: 5.4 -> 4.8 groups of four per step); every register of the step lies at or below that group
    // wide_flag (optional): wide_flag[pair] = epoch for every pair handed to the wide list, so that the estimate kernel of
    // the plane pairs can leave those rows to the byte kernel + estimate that run beside it on another stream
    constexpr int FORM = union_form<Epi>::value;
#ifndef SELB_EMUL   // the emulator's dynamic shared memory is a global array of this name
    extern __shared__ __align__(128) uint8_t pl_smem[];
#endif
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
    // Guided claims: a batch is 1/(PL_GUIDE x warps) of the pairs still unclaimed, between 4 and 32 — whole batches of 32
    // while there is plenty of work (the descriptor loads are paid once per batch), small ones towards the end, so that
    // the warps finish within a few pairs of one another (a pair is 6 us of a warp's time, a batch of 32 a seventh of the
    // whole kernel at n = 100k: with fixed batches the last ones left most warps idle for half a batch on average).
#ifndef PL_GUIDE
#define PL_GUIDE 2
#endif
#ifndef PL_BATCH_MAX
#define PL_BATCH_MAX 32      // measured: cap 16 1.365 ms, cap 8 1.383 against 1.356 (plane + byte kernel); PL_GUIDE 3: 1.358
#endif

    // two descriptor sets (batch k lives in set k&1); per lane: one pair of the batch
    uint32_t d_rx0 = 0, d_ry0 = 0, d_ix0 = 0, d_iy0 = 0, d_gm0 = 0, d_rx1 = 0, d_ry1 = 0, d_ix1 = 0, d_iy1 = 0, d_gm1 = 0;
    uint32_t d_tp0 = 0x00ffffffu, d_tp1 = 0x00ffffffu;      // per-eighth top groups of the lane's pair (gtop)
    int tsh = 0;                                            // 2048-register steps per eighth = 1 << tsh
    while (((size_t)16384 << tsh) < m) ++tsh;
    uint32_t mask0 = 0, mask1 = 0;          // lanes of the set holding a pair to do (warp-uniform)
    long long base0 = 0, base1 = 0;         // first pair index of the batch
    bool end0 = false, end1 = false;        // the batch starts past the end of the list: nothing follows
    int filled = -1;

    auto fill = [&](int k) {
        long long start = 0;
        int bsz = 0;
        if (lane == 0) {
            const long long left = npairs - (long long)*reinterpret_cast<volatile unsigned long long*>(batch_counter);
            bsz = (int)max(4ll, min((long long)PL_BATCH_MAX, left / ((long long)PL_GUIDE * gridDim.x)));
            start = (long long)atomicAdd(batch_counter, (unsigned long long)bsz);
        }
        start = __shfl_sync(FULL, start, 0);
        bsz = __shfl_sync(FULL, bsz, 0);
        const long long pi = start + lane;
        bool ok = lane < bsz && pi < npairs;
        uint2 id = make_uint2(0u, 0u), rw = id;
        uint32_t gm = 0, tp = 0x00ffffffu;
        if (ok) {
            rw = src.rows(pi, id);
            const uint32_t ra = grange[rw.x], rb = grange[rw.y];
            const int lo = max((int)(ra & 0xff), (int)(rb & 0xff)), hi = max((int)(ra >> 8), (int)(rb >> 8));
            const int g0 = min(lo >> 3, 4);
            if ((hi >> 3) > g0 + 3) {        // value range wider than the window: the byte kernel does this pair
                wide_list[atomicAdd(wide_count, 1ull)] = (uint32_t)pi;
                if (wide_flag) wide_flag[pi] = epoch;
                ok = false;
            } else {
                uint32_t gmask = 0;
                if (FORM == 0) {
                    for (int t = 0; t < 4; ++t)
                        if ((g0 + t) >= (lo >> 3) && (g0 + t) <= (hi >> 3)) gmask |= 1u << t;
                } else {                     // groups of four values: 8*g0 + 4t .. 8*g0 + 4t + 3
                    for (int t = 0; t < 8; ++t)
                        if ((2 * g0 + t) >= (lo >> 2) && (2 * g0 + t) <= (hi >> 2)) gmask |= 1u << t;
                }
                gm = (uint32_t)g0 | (gmask << 8);
                if (FORM == 1 && gtop) {
                    const uint4* ta = reinterpret_cast<const uint4*>(gtop + (size_t)rw.x * 8);
                    const uint4* tb = reinterpret_cast<const uint4*>(gtop + (size_t)rw.y * 8);
                    const uint4 a0 = ta[0], a1 = ta[1], b0 = tb[0], b1 = tb[1];
                    const uint32_t top[8] = {max(a0.x, b0.x), max(a0.y, b0.y), max(a0.z, b0.z), max(a0.w, b0.w),
                                             max(a1.x, b1.x), max(a1.y, b1.y), max(a1.z, b1.z), max(a1.w, b1.w)};
                    tp = 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j)      // lo <= top[j] <= hi: the group index lies in 0 .. 7
                        tp |= (uint32_t)min(max((int)(top[j] >> 2) - 2 * g0, 0), 7) << (3 * j);
                }
            }
        }
        const uint32_t msk = __ballot_sync(FULL, ok);
        if (k & 1) { d_rx1 = rw.x; d_ry1 = rw.y; d_ix1 = id.x; d_iy1 = id.y; d_gm1 = gm; d_tp1 = tp; mask1 = msk; base1 = start; end1 = start >= npairs; }
        else       { d_rx0 = rw.x; d_ry0 = rw.y; d_ix0 = id.x; d_iy0 = id.y; d_gm0 = gm; d_tp0 = tp; mask0 = msk; base0 = start; end0 = start >= npairs; }
        filled = k;
    };

    struct Cur {               // position in the warp's sequence of (pair, chunk) items; warp-uniform
        int k;                 // batch number
        uint32_t mask;         // pairs of the batch not started yet
        bool valid, done;
        int ch;
        uint32_t rx, ry, ix, iy, gm, tp;
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
                if (FORM == 1 && is_cons) c.tp = __shfl_sync(FULL, odd ? d_tp1 : d_tp0, j);
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
    cons.rx = cons.ry = cons.ix = cons.iy = cons.gm = 0; cons.tp = 0x00ffffffu; cons.pi = 0;
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
    uint32_t S[32], C2[PL_NC2];
#pragma unroll
    for (int v = 0; v < 32; ++v) S[v] = 0;
#pragma unroll
    for (int v = 0; v < PL_NC2; ++v) C2[v] = 0;
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
        if (FORM == 1) {
            if (nq == PL_NQ) {
                switch (g0) {
                    case 0: plane_chunk_subsets<0, PL_NQ>(pa, pb, nq, lane, gmask, S, C2, cons.tp, cons.ch * (PL_NQ / 32), tsh); break;
                    case 1: plane_chunk_subsets<1, PL_NQ>(pa, pb, nq, lane, gmask, S, C2, cons.tp, cons.ch * (PL_NQ / 32), tsh); break;
                    case 2: plane_chunk_subsets<2, PL_NQ>(pa, pb, nq, lane, gmask, S, C2, cons.tp, cons.ch * (PL_NQ / 32), tsh); break;
                    case 3: plane_chunk_subsets<3, PL_NQ>(pa, pb, nq, lane, gmask, S, C2, cons.tp, cons.ch * (PL_NQ / 32), tsh); break;
                    default: plane_chunk_subsets<4, PL_NQ>(pa, pb, nq, lane, gmask, S, C2, cons.tp, cons.ch * (PL_NQ / 32), tsh); break;
                }
            } else {
                switch (g0) {
                    case 0: plane_chunk_subsets<0, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                    case 1: plane_chunk_subsets<1, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                    case 2: plane_chunk_subsets<2, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                    case 3: plane_chunk_subsets<3, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                    default: plane_chunk_subsets<4, 0>(pa, pb, nq, lane, gmask, S, C2); break;
                }
            }
        } else if (nq == PL_NQ) {
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
            for (int v = 0; v < 32; ++v) {
                // a directly counted subset mask (PL_DIRECT, subset form) has no carry-save state: C2 holds its count
                if (FORM == 1 && ((PL_DIRECT >> (v & 3)) & 1)) x[v] = c2_get(C2, v);
                else x[v] = 2u * c2_get(C2, v) + (uint32_t)__popc(S[v]);
                S[v] = 0;
            }
#pragma unroll
            for (int v = 0; v < PL_NC2; ++v) C2[v] = 0;
            if (PL_BFLY16 && m <= 65536) {
                // two totals per word (value v in the low half, v + 16 in the high half: a lane sees m/32 registers, sixteen
                // lanes m/2 <= 2^15), four butterfly stages over the lane's low bits, then the halves part over bit 4
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] += x[i + 16] << 16;
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) {
                    const bool upper = (lane & o) != 0;
#pragma unroll
                    for (int i = 0; i < o; ++i) {
                        const uint32_t send = upper ? x[i] : x[i + o];
                        const uint32_t keep = upper ? x[i + o] : x[i];
                        x[i] = keep + __shfl_xor_sync(FULL, send, o);
                    }
                }
                const uint32_t other = __shfl_xor_sync(FULL, x[0], 16);
                x[0] = (lane & 16) ? (x[0] >> 16) + (other >> 16) : (x[0] & 0xffffu) + (other & 0xffffu);
            } else {
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
            }
            if (FORM == 1) {
                // lane 4t + s holds #registers of group t whose two low bits contain subset s: two subtract steps
                // (s without a bit -= s with the bit) leave the count of low bits == s, i.e. the bin 8*g0 + lane
                const uint32_t y0 = __shfl_xor_sync(FULL, x[0], 1);
                if (!(lane & 1)) x[0] -= y0;
                const uint32_t y1 = __shfl_xor_sync(FULL, x[0], 2);
                if (!(lane & 2)) x[0] -= y1;
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
    static constexpr uint32_t kMask = 0xffffffffu;
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
