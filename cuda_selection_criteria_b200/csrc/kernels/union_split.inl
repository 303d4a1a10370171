// union_split.inl — K5, split form: dense 16-value window on bit planes + exact sparse lists for the rest (part of selb200.cu)
// ============================================================================
// The bit-plane kernel (union_planes.inl) decodes a window of 32 values (4 groups of 8) for every
// register, although the registers of a genome concentrate in ~10 consecutive values: with n/m of a
// few hundred, fewer than 2 % of the registers reach min+16.  This form counts only the 16 values
// [base, base+16), base = 8*(min>>3), with logic instructions, and treats every register >= base+16
// ("high") exactly through a per-genome sorted list:
//
//   record of genome g (built once at load, k_split_build), chunks + 1 slots of 5*chunk_regs/8 bytes each:
//     chunk c : 5 planes [plane][chunk_regs/32 words]: planes 0..3 = bits of (v - base) where v < base+16
//               (0 where high), plane 4 = flag "v >= base+16"
//     last    : u16 ghist[64] (the genome's own histogram, saturated), u16 offs[R+1] (first list entry of each
//               of R equal position ranges, offs[R] = len), then the high list: one u32 (position << 6 | value)
//               per high register, ascending position
//   gmeta[g] = base | len << 8   (len = 0xFFFF: the list does not fit its slot)
//
//   pair (a, b) with equal bases:
//     dense  : e = flag_a | flag_b; max over 4 planes (8 LOP3 / 32 regs), one-hot decode of the low 3 bits (8),
//              2 groups x (8 AND + carry-save adder 2 x 8 / 2 words)        = 51 LOP3 / 32 registers (was 86)
//     high   : count[v] = ghist_a[v] + ghist_b[v] - #{positions high in both with min(va, vb) = v}:
//              every entry of the shorter list is looked up in the longer one — its position range holds
//              one or two candidates (offs) — and a hit decrements bin min(va, vb)
//   pairs with different bases or an overflowing list go to the wide list (byte kernel), like the pairs whose
//   range does not fit the window of the plane kernel.
// The staging ring, the batched descriptors and the producer/consumer walk are those of k_pair_hist_planes;
// a pair is chunks + 1 ring items, the last one carrying both lists.
//
// Measured (B200, n=100k, 511 521 pairs, profiles/r01_ncu_summary.md): bit-identical to the other forms on the
// whole -m gpu suite, but NOT faster: 1.91 ms with a binary-search lookup, 1.87 ms with the offset table,
// 1.83 ms at 20 CTAs/SM (-DSPLIT_MIN_CTAS=20) against 1.79 ms of the plane kernel, although the inner loop
// issues 150 instead of 255 instructions per 64 registers.  ncu: the dense loop is 48 % of the instructions and
// only 27 % of the stall samples; per-item pipeline control (117 instructions x 5 items per pair instead of 4:
// the elected TMA issue, descriptor shuffles, barrier waits) and the list phase take the rest.  The union pass
// is therefore bound by its control path, not by the counting, in BOTH forms; this one stays behind
// SELB200_UNION=split until that path moves to a producer warp.
// ============================================================================
constexpr uint32_t SPLIT_LEN_OVERFLOW = 0xFFFFu;
constexpr int SPLIT_GHIST_BYTES = 128;

__host__ __device__ __forceinline__ uint32_t split_chunk_bytes(int chunk_regs) { return 5u * ((uint32_t)chunk_regs >> 3); }
// position ranges of the offset table: 128 for m >= 16384, never fewer than 8
__host__ __device__ __forceinline__ uint32_t split_ranges(size_t m) {
    const size_t r = m >> 7;
    return r > 128 ? 128u : (r < 8 ? 8u : (uint32_t)r);
}
__host__ __device__ __forceinline__ uint32_t split_hdr_bytes(size_t m) {
    return SPLIT_GHIST_BYTES + ((2u * (split_ranges(m) + 1u) + 15u) & ~15u);
}
__host__ __device__ __forceinline__ uint32_t split_list_cap(size_t m, int chunk_regs) {
    return (split_chunk_bytes(chunk_regs) - split_hdr_bytes(m)) / 4u;
}

// one warp per genome: bytes -> 5 relative planes + high list + saturated histogram copy
__global__ void __launch_bounds__(256)
k_split_build(const uint8_t* __restrict__ regs, long long rows, size_t m, int chunk_regs,
              const uint16_t* __restrict__ grange, const uint32_t* __restrict__ hist, uint8_t* __restrict__ rec,
              uint32_t* __restrict__ gmeta) {
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int blk_per_genome = (int)(m >> 9), blk_per_chunk = chunk_regs >> 9;
    const int nchunks = (int)(m / (size_t)chunk_regs);
    const uint32_t cb = split_chunk_bytes(chunk_regs);
    const uint32_t cap = split_list_cap(m, chunk_regs);
    const uint32_t hdr = split_hdr_bytes(m), nr = split_ranges(m);
    const uint32_t per_range = (uint32_t)(m / nr);    // registers per position range (a power of two)
    const int cw = chunk_regs >> 5;                   // words per plane of a chunk
    const size_t rec_bytes = (size_t)(nchunks + 1) * cb;
    for (long long g = warp0; g < rows; g += nwarps) {
        const uint32_t base = (uint32_t)((grange[g] & 0xff) >> 3) << 3;
        const uint32_t base4 = base * 0x01010101u;
        uint8_t* grec = rec + (size_t)g * rec_bytes;
        uint16_t* ghist = reinterpret_cast<uint16_t*>(grec + (size_t)nchunks * cb);
        uint16_t* offs = ghist + 64;
        uint32_t* list = reinterpret_cast<uint32_t*>(grec + (size_t)nchunks * cb + hdr);
        ghist[lane] = (uint16_t)min(hist[g * 64 + lane], 0xFFFFu);
        ghist[lane + 32] = (uint16_t)min(hist[g * 64 + 32 + lane], 0xFFFFu);
        uint32_t len = 0;                             // warp-uniform: high registers seen so far
        for (int bg = 0; bg < blk_per_genome; ++bg) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(regs + (size_t)g * m + (size_t)bg * 512) + lane);
            uint32_t w[4] = {v.x - base4, v.y - base4, v.z - base4, v.w - base4};   // v >= base: no borrows
            uint32_t hb[4], cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                hb[k] = ((w[k] >> 4) | (w[k] >> 5)) & 0x01010101u;      // relative value >= 16 (values <= 63)
                cnt += (uint32_t)__popc(hb[k]);
                w[k] &= ~(hb[k] * 0xFFu);                               // value planes are zero where high
            }
            const int chunk = bg / blk_per_chunk, bc = bg - chunk * blk_per_chunk;
            uint32_t* dst = reinterpret_cast<uint32_t*>(grec + (size_t)chunk * cb) + (size_t)bc * 16;
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                uint32_t h = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t bits = b < 4 ? ((w[k] >> b) & 0x01010101u) : hb[k];
                    h |= ((bits * 0x10204080u) >> 28) << (4 * k);
                }
                const uint32_t lo = __shfl_sync(FULL, h, 2 * (lane & 15));
                const uint32_t hi = __shfl_sync(FULL, h, 2 * (lane & 15) + 1);
                if (lane < 16) dst[(size_t)b * cw + lane] = lo | (hi << 16);
            }
            if (__any_sync(FULL, cnt != 0u)) {
                uint32_t incl = cnt;                  // inclusive prefix over the lanes: list order = position order
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                uint32_t off = len + incl - cnt;
                if (cnt) {
                    const uint32_t raw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if ((hb[k] >> (8 * j)) & 1u) {
                                const uint32_t pos = (uint32_t)bg * 512u + (uint32_t)lane * 16u + (uint32_t)(4 * k + j);
                                if (off < cap) list[off] = (pos << 6) | ((raw[k] >> (8 * j)) & 63u);
                                ++off;
                            }
                }
                len += __shfl_sync(FULL, incl, 31);
            }
        }
        if (lane == 0) gmeta[g] = base | ((len > cap ? SPLIT_LEN_OVERFLOW : len) << 8);
        if (len <= cap) {                             // offs[r] = entries below position r * per_range
            __syncwarp();
            for (uint32_t r = (uint32_t)lane; r <= nr; r += 32) {
                const uint32_t key = (r * per_range) << 6;
                uint32_t lo = 0, hi = len;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (list[mid] < key) lo = mid + 1; else hi = mid;
                }
                offs[r] = (uint16_t)lo;
            }
        }
    }
}

// One chunk of the dense window against the running carry-save state (16 values).
template <int NQ>   // NQ > 0: uint2 per plane known at compile time
__device__ __forceinline__ void split_chunk(const uint2* __restrict__ sA, const uint2* __restrict__ sB, int nq_rt, int lane,
                                            uint32_t (&S)[16], uint32_t (&C2)[16]) {
    const int nq = NQ > 0 ? NQ : nq_rt;
#pragma unroll 1
    for (int q = lane; q < nq; q += 32) {
        uint32_t M[2][4], H[2][2];
        {
            uint2 a[5], b[5];
#pragma unroll
            for (int pl = 0; pl < 5; ++pl) { a[pl] = sA[pl * nq + q]; b[pl] = sB[pl * nq + q]; }
            const uint32_t e0 = a[4].x | b[4].x, e1 = a[4].y | b[4].y;    // a high register on either side
            uint32_t lt0 = 0u, lt1 = 0u;
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) {      // borrow of a - b, plane by plane: ends as the mask a < b
                lt0 = lop3<0x8E>(a[pl].x, b[pl].x, lt0);
                lt1 = lop3<0x8E>(a[pl].y, b[pl].y, lt1);
            }
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) {      // max = a < b ? b : a
                M[0][pl] = lop3<0xCA>(lt0, b[pl].x, a[pl].x);
                M[1][pl] = lop3<0xCA>(lt1, b[pl].y, a[pl].y);
            }
            H[0][0] = lop3<0x03>(M[0][3], e0, e0);     // ~(M3 | e): values 0..7, not high
            H[1][0] = lop3<0x03>(M[1][3], e1, e1);
            H[0][1] = lop3<0x30>(M[0][3], e0, e0);     // M3 & ~e: values 8..15, not high
            H[1][1] = lop3<0x30>(M[1][3], e1, e1);
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
#pragma unroll
        for (int T = 0; T < 2; ++T) {
            uint32_t m0[8], m1[8], kk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { m0[j] = H[0][T] & L[0][j]; m1[j] = H[1][T] & L[1][j]; }
#pragma unroll
            for (int j = 0; j < 8; ++j) kk[j] = lop3<0xE8>(S[T * 8 + j], m0[j], m1[j]);
#pragma unroll
            for (int j = 0; j < 8; ++j) S[T * 8 + j] = lop3<0x96>(S[T * 8 + j], m0[j], m1[j]);
#pragma unroll
            for (int j = 0; j < 8; ++j) C2[T * 8 + j] += __popc(kk[j]);
        }
    }
}

#ifndef SPLIT_MIN_CTAS
#define SPLIT_MIN_CTAS 16
#endif

template <class Epi>
__global__ void __launch_bounds__(32, SPLIT_MIN_CTAS)
k_pair_hist_split(const uint8_t* __restrict__ rec, size_t m, int chunk_regs, const uint32_t* __restrict__ gmeta,
                  SrcPairs src, Epi epi, uint32_t* __restrict__ wide_list, unsigned long long* __restrict__ wide_count,
                  unsigned long long* __restrict__ batch_counter) {
#ifndef SELB_EMUL   // the emulator's dynamic shared memory is a global array of this name
    extern __shared__ __align__(128) uint8_t pl_smem[];
#endif
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int nchunks = (int)(m / (size_t)chunk_regs);
    const int nitems = nchunks + 1;
    const uint32_t chunk_bytes = split_chunk_bytes(chunk_regs);
    const int nq = chunk_regs >> 6;
    const uint32_t hdr = split_hdr_bytes(m);
    const int rshift = 31 - __clz((int)(m / split_ranges(m)));      // position -> range of the offset table
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(pl_smem);
    const uint32_t bar0 = smem0 + PL_STAGES * 2 * chunk_bytes;
    uint32_t* hsub = reinterpret_cast<uint32_t*>(pl_smem + (size_t)PL_STAGES * 2 * chunk_bytes + 8 * PL_STAGES);
    if (lane == 0)
        for (int st = 0; st < PL_STAGES; ++st) mbar_init(bar0 + 8 * st, 1);
    hsub[lane] = 0u;
    hsub[lane + 32] = 0u;
    __syncwarp();
    const size_t genome_bytes = (size_t)nitems * chunk_bytes;
    const long long npairs = src.count();
    int bsz = 32;
    while (bsz > 4 && npairs < (long long)bsz * gridDim.x * 4) bsz >>= 1;

    // two descriptor sets (batch k lives in set k&1); per lane: one pair of the batch
    uint32_t d_rx0 = 0, d_ry0 = 0, d_ix0 = 0, d_iy0 = 0, d_gm0 = 0, d_rx1 = 0, d_ry1 = 0, d_ix1 = 0, d_iy1 = 0, d_gm1 = 0;
    uint32_t mask0 = 0, mask1 = 0;
    long long base0 = 0, base1 = 0;
    bool end0 = false, end1 = false;
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
            const uint32_t ma = gmeta[rw.x], mb = gmeta[rw.y];
            const uint32_t la = ma >> 8, lb = mb >> 8;
            if ((ma & 0xffu) != (mb & 0xffu) || la == SPLIT_LEN_OVERFLOW || lb == SPLIT_LEN_OVERFLOW) {
                wide_list[atomicAdd(wide_count, 1ull)] = (uint32_t)pi;      // the byte kernel does this pair
                ok = false;
            } else {
                gm = (ma & 0xffu) | (la << 8) | (lb << 20);                 // base, list lengths (<= 608 each)
            }
        }
        const uint32_t msk = __ballot_sync(FULL, ok);
        if (k & 1) { d_rx1 = rw.x; d_ry1 = rw.y; d_ix1 = id.x; d_iy1 = id.y; d_gm1 = gm; mask1 = msk; base1 = bidx * bsz; end1 = bidx * bsz >= npairs; }
        else       { d_rx0 = rw.x; d_ry0 = rw.y; d_ix0 = id.x; d_iy0 = id.y; d_gm0 = gm; mask0 = msk; base0 = bidx * bsz; end0 = bidx * bsz >= npairs; }
        filled = k;
    };

    struct Cur {               // position in the warp's sequence of (pair, item) steps; warp-uniform
        int k;
        uint32_t mask;
        bool valid, done;
        int ch;
        uint32_t rx, ry, ix, iy, gm;
        long long pi;
        const uint8_t* ga;
        const uint8_t* gb;
    };
    Cur cons, prod;
    auto next_pair = [&](Cur& c, bool is_cons) {
        c.valid = false;
        for (;;) {
            // see k_pair_hist_planes: a producer parked on a batch the consumer has left rejoins the consumer
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
                    c.ga = rec + (size_t)c.rx * genome_bytes;
                    c.gb = rec + (size_t)c.ry * genome_bytes;
                }
                c.ch = 0;
                c.valid = true;
                return;
            }
            if ((c.k & 1) ? end1 : end0) { c.done = true; return; }
            if (c.k + 1 > filled) return;
            ++c.k;
            c.mask = (c.k & 1) ? mask1 : mask0;
            if (is_cons && !(((filled & 1) ? end1 : end0))) fill(c.k + 1);
        }
    };
    auto issue = [&](const Cur& c, uint32_t n_issued) {      // lane 0: two bulk copies into the next stage
        const uint32_t st = n_issued % PL_STAGES;
        const uint32_t dst = smem0 + st * 2 * chunk_bytes, bar = bar0 + 8 * st;
        const uint8_t* ga = c.ga + (uint32_t)c.ch * chunk_bytes;
        const uint8_t* gb = c.gb + (uint32_t)c.ch * chunk_bytes;
        uint32_t na = chunk_bytes, nb = chunk_bytes;
        if (c.ch == nchunks) {                               // histogram copy + the list, rounded up to 16 bytes
            na = hdr + ((((c.gm >> 8) & 0xfffu) * 4u + 15u) & ~15u);
            nb = hdr + (((c.gm >> 20) * 4u + 15u) & ~15u);
        }
        mbar_expect_tx(bar, na + nb);
        tma_bulk_g2s(dst, ga, na, bar);
        tma_bulk_g2s(dst + chunk_bytes, gb, nb, bar);
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
        if (++prod.ch >= nitems) next_pair(prod, false);
    }
    uint32_t S[16], C2[16];
#pragma unroll
    for (int v = 0; v < 16; ++v) { S[v] = 0; C2[v] = 0; }
    while (cons.valid) {
        __syncwarp();                          // every lane has finished reading the stage about to be refilled
        if (!prod.valid && !prod.done) next_pair(prod, false);
        if (prod.valid) {
            if (lane == 0) issue(prod, n_issued);
            ++n_issued;
            if (++prod.ch >= nitems) next_pair(prod, false);
        }
        if (n_done == n_issued) {              // cannot happen: the consumer never overtakes the producer
            if (lane == 0) atomicExch(batch_counter + 1, 0xBB00000000000000ull | (n_issued & 0xffffu));
            return;
        }
        const uint32_t st = n_done % PL_STAGES;
        mbar_wait(bar0 + 8 * st, (n_done / PL_STAGES) & 1u);
        ++n_done;
        const uint8_t* stA = pl_smem + (size_t)st * 2 * chunk_bytes;
        const uint8_t* stB = stA + chunk_bytes;
        if (cons.ch < nchunks) {
            const uint2* pa = reinterpret_cast<const uint2*>(stA);
            const uint2* pb = reinterpret_cast<const uint2*>(stB);
            if (nq == PL_NQ) split_chunk<PL_NQ>(pa, pb, nq, lane, S, C2);
            else split_chunk<0>(pa, pb, nq, lane, S, C2);
        } else {
            // ---- high registers: positions present in both lists count once, in bin max(va, vb) ----
            const int base = (int)(cons.gm & 0xffu);
            const int la = (int)((cons.gm >> 8) & 0xfffu), lb = (int)(cons.gm >> 20);
            const uint16_t* gha = reinterpret_cast<const uint16_t*>(stA);
            const uint16_t* ghb = reinterpret_cast<const uint16_t*>(stB);
            const uint32_t* la_p = reinterpret_cast<const uint32_t*>(stA + hdr);
            const uint32_t* lb_p = reinterpret_cast<const uint32_t*>(stB + hdr);
            const bool a_short = la <= lb;
            const uint32_t* X = a_short ? la_p : lb_p;
            const uint32_t* Y = a_short ? lb_p : la_p;
            const uint16_t* offY = (a_short ? ghb : gha) + 64;
            const int lx = a_short ? la : lb;
            for (int i = lane; i < lx; i += 32) {
                const uint32_t ex = X[i], pos = ex >> 6, r = pos >> rshift;
                const uint32_t j1 = offY[r + 1];
                for (uint32_t j = offY[r]; j < j1; ++j) {         // the candidates: Y's entries of the same range
                    const uint32_t ey = Y[j];
                    if ((ey >> 6) == pos) { atomicAdd(&hsub[min(ex & 63u, ey & 63u)], 1u); break; }
                }
            }
            __syncwarp();
            // ---- dense totals: per-lane counts, transposing butterfly; value d ends in lanes 2d and 2d+1 ----
            uint32_t x[16];
#pragma unroll
            for (int v = 0; v < 16; ++v) { x[v] = 2u * C2[v] + (uint32_t)__popc(S[v]); S[v] = 0; C2[v] = 0; }
#pragma unroll
            for (int o = 16; o >= 2; o >>= 1) {
                const bool upper = (lane & o) != 0;
                const int h = o >> 1;
#pragma unroll
                for (int i = 0; i < h; ++i) {
                    const uint32_t send = upper ? x[i] : x[i + h];
                    const uint32_t keep = upper ? x[i + h] : x[i];
                    x[i] = keep + __shfl_xor_sync(FULL, send, o);
                }
            }
            const uint32_t dense = x[0] + __shfl_xor_sync(FULL, x[0], 1);
            const int b0 = lane, b1 = lane + 32;                  // the two bins of this lane
            const uint32_t t0 = __shfl_sync(FULL, dense, (2 * (b0 - base)) & 31);
            const uint32_t t1 = __shfl_sync(FULL, dense, (2 * (b1 - base)) & 31);
            const uint32_t h0 = (uint32_t)gha[b0] + (uint32_t)ghb[b0] - hsub[b0];
            const uint32_t h1 = (uint32_t)gha[b1] + (uint32_t)ghb[b1] - hsub[b1];
            const uint32_t c0 = b0 < base ? 0u : (b0 < base + 16 ? t0 : h0);
            const uint32_t c1 = b1 < base ? 0u : (b1 < base + 16 ? t1 : h1);
            hsub[b0] = 0u;
            hsub[b1] = 0u;
            uint32_t tot = c0 + c1;                               // every register lands in exactly one bin
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            if (tot != (uint32_t)m && lane == 0)
                atomicExch(batch_counter + 1, 0xBC00000000000000ull | ((unsigned long long)(cons.pi & 0xffffffll) << 24) | (tot & 0xffffffu));
            epi(src.slot(cons.pi), make_uint2(cons.ix, cons.iy), c0, c1, (uint32_t)lane);
        }
        if (++cons.ch >= nitems) next_pair(cons, true);
    }
}
