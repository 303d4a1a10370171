// filter_hll.inl — K4'/K4'': hll_a / hll_an tile filters, byte form and bit-plane form (part of selb200.cu)
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
#ifndef SELB_EMUL   // the emulator's dynamic shared memory is a global array of this name
    extern __shared__ __align__(1024) uint32_t hist_dyn[];   // 2 x [nbins][64 threads]
#endif
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
//                                   (subset form: the quad layout of k_aux_planes_quad in the same buffer)
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

// The same planes in QUAD layout, read by the subset form of the filter (k_tile_filter_hll_planes<2>, <3>): the twelve
// plane words of a word PAIR (planes 0..5 of word 2wp, then of word 2wp+1) of one genome sit in three uint4,
//   auxQ[(wp*3 + q)*npad + g]   q = 0: planes 0..3 of word 2wp   q = 1: planes 4,5 of 2wp and 0,1 of 2wp+1   q = 2: planes 2..5 of 2wp+1
// so a filter step is three 128-bit loads per genome (coalesced over the 32 columns of a warp step) instead of twelve
// 32-bit loads with their own index arithmetic.
__global__ void __launch_bounds__(256)
k_aux_planes_quad(const uint8_t* __restrict__ aux, const int32_t* __restrict__ order, long long n, long long npad,
                  int p_aux, uint32_t* __restrict__ auxQ) {
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
            const int slot = (w & 1) * 6 + b;            // 0..11 inside the word pair
            auxQ[(((size_t)(w >> 1) * 3 + (slot >> 2)) * npad + g) * 4 + (slot & 3)] = m;
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
// FORM 0: one-hot masks on groups of eight values (gmask: 4 bits); FORM 1: subset masks on groups of four values
// (gmask: 8 bits) followed by the in-register Moebius step — the counting step of plane_chunk_subsets in
// kernels/union_planes.inl, where the scheme is described (SELB200_HLLFILTER=subsets; CPU-checked, not yet measured)
template <int G0, int FORM>
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
            if (FORM == 1) {
                // quad layout (k_aux_planes_quad): rowp / colp point at the genome's first uint4; three 128-bit loads
                // per genome and step, 32-bit uint4 offsets (the launcher checks that the matrix stays below 2^32)
                const uint4* rq = reinterpret_cast<const uint4*>(rowp);
                const uint4* cq = reinterpret_cast<const uint4*>(colp);
                const uint32_t np32 = (uint32_t)npad, o = (uint32_t)(w >> 1) * 3u * np32;
                const uint4 r0 = __ldg(rq + o), r1 = __ldg(rq + (o + np32)), r2 = __ldg(rq + (o + 2u * np32));
                const uint4 c0 = __ldg(cq + o), c1 = __ldg(cq + (o + np32)), c2 = __ldg(cq + (o + 2u * np32));
                a[0][0] = r0.x; a[0][1] = r0.y; a[0][2] = r0.z; a[0][3] = r0.w; a[0][4] = r1.x; a[0][5] = r1.y;
                a[1][0] = r1.z; a[1][1] = r1.w; a[1][2] = r2.x; a[1][3] = r2.y; a[1][4] = r2.z; a[1][5] = r2.w;
                b[0][0] = c0.x; b[0][1] = c0.y; b[0][2] = c0.z; b[0][3] = c0.w; b[0][4] = c1.x; b[0][5] = c1.y;
                b[1][0] = c1.z; b[1][1] = c1.w; b[1][2] = c2.x; b[1][3] = c2.y; b[1][4] = c2.z; b[1][5] = c2.w;
            } else {
#pragma unroll
            for (int pl = 0; pl < 6; ++pl) {
                const size_t o0 = ((size_t)pl * nw + w) * (size_t)npad, o1 = o0 + (size_t)npad;
                a[0][pl] = __ldg(rowp + o0); a[1][pl] = __ldg(rowp + o1);
                b[0][pl] = __ldg(colp + o0); b[1][pl] = __ldg(colp + o1);
            }
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
        if (FORM == 1) {
#define SELB_AUX_SUBSET_GROUP(T8, HALF)                                                                   \
            if (gmask & (1u << (2 * T8 + HALF))) {                                                        \
                constexpr int c0 = 4 * (2 * T8 + HALF);                                                   \
                const uint32_t e0 = HALF ? lop3<0xC0>(H0, M[0][2], 0u) : lop3<0x30>(H0, M[0][2], 0u);     \
                const uint32_t e1 = HALF ? lop3<0xC0>(H1, M[1][2], 0u) : lop3<0x30>(H1, M[1][2], 0u);     \
                uint32_t m0[4], m1[4], kk[4];                                                             \
                m0[0] = e0; m0[1] = lop3<0xC0>(e0, M[0][0], 0u); m0[2] = lop3<0xC0>(e0, M[0][1], 0u);     \
                m0[3] = lop3<0x80>(e0, M[0][0], M[0][1]);                                                 \
                m1[0] = e1; m1[1] = lop3<0xC0>(e1, M[1][0], 0u); m1[2] = lop3<0xC0>(e1, M[1][1], 0u);     \
                m1[3] = lop3<0x80>(e1, M[1][0], M[1][1]);                                                 \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) kk[j] = lop3<0xE8>(S[c0 + j], m0[j], m1[j]); \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) S[c0 + j] = lop3<0x96>(S[c0 + j], m0[j], m1[j]); \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) C2[c0 + j] += __popc(kk[j]);                \
            }
#define SELB_AUX_SUBSET_GROUP8(T8)                                                                        \
            if (gmask & (3u << (2 * T8))) {                                                               \
                const uint32_t H0 = lop3<(1 << (G0 + T8))>(M[0][5], M[0][4], M[0][3]);                    \
                const uint32_t H1 = lop3<(1 << (G0 + T8))>(M[1][5], M[1][4], M[1][3]);                    \
                SELB_AUX_SUBSET_GROUP(T8, 0)                                                              \
                SELB_AUX_SUBSET_GROUP(T8, 1)                                                              \
            }
            SELB_AUX_SUBSET_GROUP8(0)
            SELB_AUX_SUBSET_GROUP8(1)
            SELB_AUX_SUBSET_GROUP8(2)
            SELB_AUX_SUBSET_GROUP8(3)
#undef SELB_AUX_SUBSET_GROUP8
#undef SELB_AUX_SUBSET_GROUP
            continue;
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
    if (FORM == 1) {
        // x[4t + s] = #registers of group t whose two low bits contain subset s -> the four bins of the group
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            x[4 * t + 0] -= x[4 * t + 1];
            x[4 * t + 2] -= x[4 * t + 3];
            x[4 * t + 0] -= x[4 * t + 2];
            x[4 * t + 1] -= x[4 * t + 3];
        }
    }
}

template <int G0, int FORM>
__device__ __forceinline__ void aux_plane_hist(const uint32_t* __restrict__ rowp, const uint32_t* __restrict__ colp,
                                               long long npad, int nw, uint32_t gmask, uint32_t* __restrict__ hcol,
                                               int nbins) {
    uint32_t x[32];
    aux_plane_pairs<G0, FORM>(rowp, colp, npad, nw, gmask, x);
    // the thread's histogram column: zeros outside the window, the counts inside
    for (int b = 0; b < 8 * G0; ++b) hcol[b * 64] = 0u;
#pragma unroll
    for (int v = 0; v < 32; ++v)
        if (8 * G0 + v < nbins) hcol[(8 * G0 + v) * 64] = x[v];
    for (int b = 8 * G0 + 32; b < nbins; ++b) hcol[b * 64] = 0u;
}

// AN bit 0: 0 = hll_a, 1 = hll_an; AN bit 1: counting form (0 = one-hot, 1 = subsets).  One integer keeps the names of
// the two GPU-validated instantiations <0> and <1> (and with them the SASS identity check of tools/sass_diff.py).
template <int AN>
__global__ void __launch_bounds__(64, HLLP_MIN_CTAS)
k_tile_filter_hll_planes(const uint32_t* __restrict__ auxP, const uint16_t* __restrict__ agrange,
                         const uint32_t* __restrict__ auxT, long long npad, int p_aux, TileWalk tw,
                         const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
                         const unsigned long long* __restrict__ e, double tau, float zs, int order_n,
                         uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count,
                         unsigned long long pair_cap, unsigned long long* __restrict__ unit_counter) {
#ifndef SELB_EMUL
    extern __shared__ __align__(1024) uint32_t hist_dyn[];   // [nbins][64 threads]
#endif
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
                constexpr int FORM = (AN >> 1) & 1;
                uint32_t gmask = 0;
                if (FORM == 0) {
                    for (int tt = 0; tt < 4; ++tt)
                        if ((g0 + tt) >= (vlo >> 3) && (g0 + tt) <= (vhi >> 3)) gmask |= 1u << tt;
                } else {
                    for (int tt = 0; tt < 8; ++tt)
                        if ((2 * g0 + tt) >= (vlo >> 2) && (2 * g0 + tt) <= (vhi >> 2)) gmask |= 1u << tt;
                }
                const uint32_t* rowp = FORM ? auxP + 4 * (size_t)i : auxP + i;      // FORM 1: quad layout, uint4 per genome
                const uint32_t* colp = FORM ? auxP + 4 * (size_t)kc : auxP + kc;
                switch (g0) {
                    case 0: aux_plane_hist<0, FORM>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    case 1: aux_plane_hist<1, FORM>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    case 2: aux_plane_hist<2, FORM>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    case 3: aux_plane_hist<3, FORM>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
                    default: aux_plane_hist<4, FORM>(rowp, colp, npad, nw, gmask, hcol, nbins); break;
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
                const StopHll stop{tau, e[i], e[k], zs, order_n, AN & 1};
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
