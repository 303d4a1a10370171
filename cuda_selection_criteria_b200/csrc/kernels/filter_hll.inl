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
// K4'' : hll_a / hll_an on BIT PLANES of the auxiliary sketches (p_aux >= 6), in two passes.
//
//   pass A  k_tile_filter_hll_bound   every pair of the CB band (tile walk, thread per pair, lane = column, row word =
//           broadcast): the twelve smallest values of the union's registers are counted in registers with the logic of
//           k_pair_hist_planes (LOP3 borrow-chain max, subset masks on groups of four values, carry-save counting), every
//           other register is charged the next value, and the sums are
//           turned, in fp32 and without touching shared memory, into a LOWER BOUND of the union estimate
//           (selb::hll_surely_fails, estimators.cuh).  Both criteria are non-increasing in the estimate, so a pair that
//           fails at the bound fails; everything else — about one pair in a few hundred — goes to the candidate list.
//           The bound is also evaluated after 3/8, 1/2 and 3/4 of the registers, with the unread part
//           replaced by the smaller of the two genomes' own sums over it (atail, built at load): max(a,b) >= a, so the
//           union's harmonic sum and zero count over any set of positions are at most either genome's.  A warp step
//           whose 32 pairs all fail there stops reading; strangers of similar size fail at the first checkpoint of a
//           1024-register sketch, so most steps do.
//   pass B  k_hll_verify              the candidates, thread per candidate: the same counting into the thread's
//           shared-memory histogram column, then the Ertl MLE and the criterion exactly as the reference evaluates them
//           (include/criteria_sketch.hpp:52-64, sketch/include/sketch/hll.h:628-688).  Decisions are the reference's:
//           pass A only ever discards pairs pass B would discard.
//   k_tile_filter_hll_planes (SELB200_HLLFILTER=onepass) is the single-pass form both replace: exact MLE for every CB pair.
//
//   auxQ[(wp*3 + q)*npad + g] (uint4) : the twelve plane words of word pair wp of the g-th sorted genome (k_aux_planes_quad)
//   agrange[g]                        : min | max<<8 of that genome's auxiliary registers
//   atail[g]                          : harmonic sum and zero count of its registers behind the two checkpoints
// The 32 pairs of a warp step share one 32-value window (their genomes sit within the CB band of each other, so their
// register ranges coincide); a step whose pairs do not fit one window goes to pass B, which takes the byte path
// (shared-memory counters over auxT) for such a pair.
// ============================================================================
#ifndef HLLP_MIN_CTAS
#define HLLP_MIN_CTAS 8
#endif
#ifndef HLLB_MIN_CTAS
#define HLLB_MIN_CTAS 12      // pass A: 64 threads x 12 CTAs = 24 warps per SM at <= 80 registers (measured: 10 CTAs 8.34 ms,
                              // 12 CTAs 7.79, 16 CTAs with 32 B spilled 7.67; loading a step ahead: 8.3 — C5, p_aux = 10)
#endif


// The planes in QUAD layout: the twelve plane words of a word PAIR (planes 0..5 of word 2wp, then of word 2wp+1) of one
// genome sit in three uint4,
//   auxQ[(wp*3 + q)*npad + g]   q = 0: planes 0..3 of word 2wp   q = 1: planes 4,5 of 2wp and 0,1 of 2wp+1   q = 2: planes 2..5 of 2wp+1
// so a filter step is three 128-bit loads per genome (coalesced over the 32 columns of a warp step).
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

// checkpoints of pass A, in word pairs (steps of 64 registers): after 1/2 and 3/4 of the sketch, and for sketches of 2048
// registers and more also after 3/8 (ascending; a checkpoint that does not fall strictly between its neighbours is
// dropped).  Strangers of similar size are decided at the first one for tau = 0.9, pairs with some overlap at the later
// ones.  Measured on C5 (filter ms, executed fraction of a full read): p_aux = 12: 3/8 first 21.2 (0.425), 1/2 first 25.7
// (0.541), 1/4 first 26.9 (0.539); p_aux = 10: 8.07 (0.504) against 8.00 (0.545) — a step stops only when all its 32
// pairs are decided, and the bound of a short sketch is too loose for that after 3/8; p_aux = 8 (4 steps): a checkpoint
// after the first step costs more than it stops (5.0 against 4.1).
constexpr int HLLB_NCP = 3;
struct HllCheckpoints { int cp[HLLB_NCP]; int n, nwp; };
__host__ __device__ __forceinline__ HllCheckpoints hll_checkpoints(int p_aux) {
    const int nwp = (1 << p_aux) >> 6;
    const int want[HLLB_NCP] = {nwp >= 32 ? (3 * nwp) / 8 : 0, nwp / 2, (3 * nwp) / 4};
    HllCheckpoints c{{0, 0, 0}, 0, nwp};
    int last = 0;
    for (int q = 0; q < HLLB_NCP; ++q)
        if (want[q] > last && want[q] < nwp) { c.cp[c.n++] = want[q]; last = want[q]; }
    return c;
}
// per genome: sum of 2^-r over the non-empty registers / count of empty ones, behind each checkpoint
struct AuxTail { float z[HLLB_NCP], c[HLLB_NCP]; };

// one warp per genome: smallest / largest register value, and the tail sums pass A bounds the unread part with
// (rounded UP: they stand in for an upper bound)
__global__ void __launch_bounds__(256)
k_aux_range(const uint8_t* __restrict__ aux, const int32_t* __restrict__ order, long long n, int p_aux,
            uint16_t* __restrict__ agrange, AuxTail* __restrict__ atail) {
    const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= n) return;
    const uint8_t* row = aux + ((size_t)order[g] << p_aux);
    const HllCheckpoints cp = hll_checkpoints(p_aux);
    int vmin = 255, vmax = 0;
    double z[HLLB_NCP] = {0., 0., 0.};
    uint32_t cz[HLLB_NCP] = {0, 0, 0};
    for (int j = lane; j < (1 << p_aux); j += 32) {
        const int v = row[j];
        vmin = min(vmin, v); vmax = max(vmax, v);
        const double t = v ? ldexp(1.0, -v) : 0.;
#pragma unroll
        for (int q = 0; q < HLLB_NCP; ++q)
            if (q < cp.n && j >= cp.cp[q] * 64) { z[q] += t; cz[q] += v == 0; }
    }
    for (int o = 16; o; o >>= 1) {
        vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
#pragma unroll
        for (int q = 0; q < HLLB_NCP; ++q) {
            z[q] += __shfl_xor_sync(0xffffffffu, z[q], o);
            cz[q] += __shfl_xor_sync(0xffffffffu, cz[q], o);
        }
    }
    if (lane == 0) {
        agrange[g] = (uint16_t)(min(vmin, vmax) | (vmax << 8));
        if (atail) {
            AuxTail t;
#pragma unroll
            for (int q = 0; q < HLLB_NCP; ++q) { t.z[q] = (float)z[q] * 1.000001f; t.c[q] = (float)cz[q]; }   // rounded UP: they stand in for an upper bound
            atail[g] = t;
        }
    }
}

// word pairs [wp0, wp1) of one (row, column) pair for window G0 into the per-value carry-save state S / C2:
// subset masks on groups of four values (the counting step of plane_chunk_subsets in kernels/union_planes.inl, where the
// scheme is described); gmask: bit t = group t (values 8*G0 + 4t .. +3) can occur
template <int G0>
__device__ __forceinline__ void aux_subset_accum(const uint4* __restrict__ rq, const uint4* __restrict__ cq, uint32_t np32,
                                                 int wp0, int wp1, uint32_t gmask, uint32_t (&S)[32], uint32_t (&C2)[32]) {
#pragma unroll 1
    for (int wp = wp0; wp < wp1; ++wp) {
        uint32_t M[2][6];
        {
            // three 128-bit loads per genome and step, 32-bit uint4 offsets (the launcher checks that the matrix stays below 2^32)
            const uint32_t o = (uint32_t)wp * 3u * np32;
            const uint4 r0 = __ldg(rq + o), r1 = __ldg(rq + (o + np32)), r2 = __ldg(rq + (o + 2u * np32));
            const uint4 c0 = __ldg(cq + o), c1 = __ldg(cq + (o + np32)), c2 = __ldg(cq + (o + 2u * np32));
            const uint32_t a[2][6] = {{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y}, {r1.z, r1.w, r2.x, r2.y, r2.z, r2.w}};
            const uint32_t b[2][6] = {{c0.x, c0.y, c0.z, c0.w, c1.x, c1.y}, {c1.z, c1.w, c2.x, c2.y, c2.z, c2.w}};
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
#define SELB_AUX_SUBSET_GROUP(T8, HALF)                                                                   \
        if (gmask & (1u << (2 * T8 + HALF))) {                                                            \
            constexpr int c0 = 4 * (2 * T8 + HALF);                                                       \
            const uint32_t e0 = HALF ? lop3<0xC0>(H0, M[0][2], 0u) : lop3<0x30>(H0, M[0][2], 0u);         \
            const uint32_t e1 = HALF ? lop3<0xC0>(H1, M[1][2], 0u) : lop3<0x30>(H1, M[1][2], 0u);         \
            uint32_t m0[4], m1[4], kk[4];                                                                 \
            m0[0] = e0; m0[1] = lop3<0xC0>(e0, M[0][0], 0u); m0[2] = lop3<0xC0>(e0, M[0][1], 0u);         \
            m0[3] = lop3<0x80>(e0, M[0][0], M[0][1]);                                                     \
            m1[0] = e1; m1[1] = lop3<0xC0>(e1, M[1][0], 0u); m1[2] = lop3<0xC0>(e1, M[1][1], 0u);         \
            m1[3] = lop3<0x80>(e1, M[1][0], M[1][1]);                                                     \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) kk[j] = lop3<0xE8>(S[c0 + j], m0[j], m1[j]);    \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) S[c0 + j] = lop3<0x96>(S[c0 + j], m0[j], m1[j]); \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) C2[c0 + j] += __popc(kk[j]);                    \
        }
#define SELB_AUX_SUBSET_GROUP8(T8)                                                                        \
        if (gmask & (3u << (2 * T8))) {                                                                   \
            const uint32_t H0 = lop3<(1 << (G0 + T8))>(M[0][5], M[0][4], M[0][3]);                        \
            const uint32_t H1 = lop3<(1 << (G0 + T8))>(M[1][5], M[1][4], M[1][3]);                        \
            SELB_AUX_SUBSET_GROUP(T8, 0)                                                                  \
            SELB_AUX_SUBSET_GROUP(T8, 1)                                                                  \
        }
        SELB_AUX_SUBSET_GROUP8(0)
        SELB_AUX_SUBSET_GROUP8(1)
        SELB_AUX_SUBSET_GROUP8(2)
        SELB_AUX_SUBSET_GROUP8(3)
#undef SELB_AUX_SUBSET_GROUP8
#undef SELB_AUX_SUBSET_GROUP
    }
}

// the four bins of group t from the state: totals 2*C2 + popc(S) count the registers of the group whose two low bits
// CONTAIN subset s; two subtract steps (Moebius inversion, exact in integers) turn them into the four bins
__device__ __forceinline__ void aux_subset_bins(const uint32_t (&S)[32], const uint32_t (&C2)[32], int t, uint32_t (&x)[4]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) x[s] = 2u * C2[4 * t + s] + (uint32_t)__popc(S[4 * t + s]);
    x[0] -= x[1];
    x[2] -= x[3];
    x[0] -= x[2];
    x[1] -= x[3];
}

__device__ __forceinline__ uint32_t aux_gmask(int g0, int vlo, int vhi) {
    uint32_t gmask = 0;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt)
        if ((2 * g0 + tt) >= (vlo >> 2) && (2 * g0 + tt) <= (vhi >> 2)) gmask |= 1u << tt;
    return gmask;
}

template <int G0>
__device__ __forceinline__ void aux_plane_hist(const uint32_t* __restrict__ rowp, const uint32_t* __restrict__ colp,
                                               long long npad, int nw, uint32_t gmask, uint32_t* __restrict__ hcol,
                                               int nbins, bool write = true) {
    uint32_t S[32], C2[32];
#pragma unroll
    for (int v = 0; v < 32; ++v) { S[v] = 0; C2[v] = 0; }
    aux_subset_accum<G0>(reinterpret_cast<const uint4*>(rowp), reinterpret_cast<const uint4*>(colp), (uint32_t)npad, 0, nw >> 1,
                         gmask, S, C2);
    if (!write) return;
    // the thread's histogram column: zeros outside the window, the counts inside
    for (int b = 0; b < 8 * G0; ++b) hcol[b * 64] = 0u;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        uint32_t x[4];
        aux_subset_bins(S, C2, t, x);
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (8 * G0 + 4 * t + s < nbins) hcol[(8 * G0 + 4 * t + s) * 64] = x[s];
    }
    for (int b = 8 * G0 + 32; b < nbins; ++b) hcol[b * 64] = 0u;
}

__device__ __forceinline__ void aux_plane_hist_g(int g0, const uint32_t* __restrict__ rowp, const uint32_t* __restrict__ colp,
                                                 long long npad, int nw, uint32_t gmask, uint32_t* __restrict__ hcol, int nbins,
                                                 bool write = true) {
    switch (g0) {
        case 0: aux_plane_hist<0>(rowp, colp, npad, nw, gmask, hcol, nbins, write); break;
        case 1: aux_plane_hist<1>(rowp, colp, npad, nw, gmask, hcol, nbins, write); break;
        case 2: aux_plane_hist<2>(rowp, colp, npad, nw, gmask, hcol, nbins, write); break;
        case 3: aux_plane_hist<3>(rowp, colp, npad, nw, gmask, hcol, nbins, write); break;
        default: aux_plane_hist<4>(rowp, colp, npad, nw, gmask, hcol, nbins, write); break;
    }
}

// byte path of one pair into the thread's histogram column (register ranges too far apart for one window)
__device__ __forceinline__ void aux_byte_hist(const uint32_t* __restrict__ auxT, long long npad, int words, int i, int kc,
                                              uint32_t* __restrict__ hcol, int nbins, uint32_t bias0, uint32_t tb) {
    for (int b = 0; b < nbins; ++b) hcol[b * 64] = 0u;
    const uint32_t* colp = auxT + kc;
    const uint32_t* row0 = auxT + i;
    for (int j = 0; j < words; ++j) {
        const uint32_t m0 = max4_lt128(__ldg(row0 + (size_t)j * npad), __ldg(colp + (size_t)j * npad)) + bias0;
        hist_inc2<0, 1>(m0, tb);
        hist_inc2<2, 3>(m0, tb);
    }
}

// ---- single pass (SELB200_HLLFILTER=onepass): exact MLE + criterion for every pair of the band -------------------
template <int AN>
__global__ void __launch_bounds__(64, HLLP_MIN_CTAS)
k_tile_filter_hll_planes(const uint32_t* __restrict__ auxQ, const uint16_t* __restrict__ agrange,
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
            if ((vhi >> 3) <= g0 + 3)
                aux_plane_hist_g(g0, auxQ + 4 * (size_t)i, auxQ + 4 * (size_t)kc, npad, nw, aux_gmask(g0, vlo, vhi), hcol, nbins);
            else
                aux_byte_hist(auxT, npad, words, i, kc, hcol, nbins, bias0, tb);
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

// ---- pass A: the bound ------------------------------------------------------------------------------------------------
// The bound needs an UPPER bound of the union's harmonic sum, not its histogram: only the twelve smallest values the step's
// pairs can hold (three groups of four, from the group of vlo up) are counted; every other register read so far is
// larger than those and enters the sum with 2^-vcap, vcap = the first value not counted.  Register values fall off
// geometrically above the smallest one (P(value >= k) ~ load 2^-k), so those registers are a few percent of the sketch
// and each is charged at most 2^-9 of an average term: the sum grows by ~1e-3 of itself, far inside the bound's slack.
// Three groups instead of the five or six a full window holds: 76 instead of ~130 LOP3 per step, 24 state registers
// instead of 64.
constexpr int HLLB_NG = 3, HLLB_NV = 4 * HLLB_NG;

// word pairs [wp0, wp1) of one (row, column) pair: groups T0 .. T0+2 (of four values, counted from value 8*G0) into the
// carry-save state; gmask as in aux_subset_accum (bit t = group t can occur)
template <int G0, int T0>
__device__ __forceinline__ void aux_bound_accum(const uint4* __restrict__ rq, const uint4* __restrict__ cq, uint32_t np32,
                                                int wp0, int wp1, uint32_t gmask, uint32_t (&S)[HLLB_NV], uint32_t (&C2)[HLLB_NV]) {
#pragma unroll 1
    for (int wp = wp0; wp < wp1; ++wp) {
        uint32_t M[2][6];
        {
            const uint32_t o = (uint32_t)wp * 3u * np32;
            const uint4 r0 = __ldg(rq + o), r1 = __ldg(rq + (o + np32)), r2 = __ldg(rq + (o + 2u * np32));
            const uint4 c0 = __ldg(cq + o), c1 = __ldg(cq + (o + np32)), c2 = __ldg(cq + (o + 2u * np32));
            const uint32_t a[2][6] = {{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y}, {r1.z, r1.w, r2.x, r2.y, r2.z, r2.w}};
            const uint32_t b[2][6] = {{c0.x, c0.y, c0.z, c0.w, c1.x, c1.y}, {c1.z, c1.w, c2.x, c2.y, c2.z, c2.w}};
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
        // groups T0 .. T0+2 lie in the eight-value blocks 0 and 1 of the window for T0 = 0 and for T0 = 1
        const uint32_t H[2][2] = {{lop3<(1 << G0)>(M[0][5], M[0][4], M[0][3]), lop3<(1 << G0)>(M[1][5], M[1][4], M[1][3])},
                                  {lop3<(1 << (G0 + 1))>(M[0][5], M[0][4], M[0][3]), lop3<(1 << (G0 + 1))>(M[1][5], M[1][4], M[1][3])}};
#pragma unroll
        for (int u = 0; u < HLLB_NG; ++u) {
            const int t = T0 + u, t8 = t >> 1, half = t & 1;
            if (gmask & (1u << t)) {
                const int c0 = 4 * u;
                const uint32_t e0 = half ? lop3<0xC0>(H[t8][0], M[0][2], 0u) : lop3<0x30>(H[t8][0], M[0][2], 0u);
                const uint32_t e1 = half ? lop3<0xC0>(H[t8][1], M[1][2], 0u) : lop3<0x30>(H[t8][1], M[1][2], 0u);
                uint32_t m0[4], m1[4], kk[4];
                m0[0] = e0; m0[1] = lop3<0xC0>(e0, M[0][0], 0u); m0[2] = lop3<0xC0>(e0, M[0][1], 0u);
                m0[3] = lop3<0x80>(e0, M[0][0], M[0][1]);
                m1[0] = e1; m1[1] = lop3<0xC0>(e1, M[1][0], 0u); m1[2] = lop3<0xC0>(e1, M[1][1], 0u);
                m1[3] = lop3<0x80>(e1, M[1][0], M[1][1]);
#pragma unroll
                for (int j = 0; j < 4; ++j) kk[j] = lop3<0xE8>(S[c0 + j], m0[j], m1[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) S[c0 + j] = lop3<0x96>(S[c0 + j], m0[j], m1[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) C2[c0 + j] += __popc(kk[j]);
            }
        }
    }
}

// upper bounds of the harmonic sum over the non-empty registers and of the empty count among the nread registers read
// so far, in fp32 (Horner from the top counted value down; every count is an exact small integer, the bound's margins
// absorb the roundings)
template <int G0, int T0>
__device__ __forceinline__ void aux_bound_sums(const uint32_t (&S)[HLLB_NV], const uint32_t (&C2)[HLLB_NV], uint32_t nread,
                                               float& z, float& c0) {
    constexpr int v0 = 8 * G0 + 4 * T0;            // first counted value
    constexpr int vfirst = v0 == 0 ? 1 : v0;       // first value of the harmonic sum (0 = empty register)
    constexpr int vcap = v0 + HLLB_NV;             // every register not counted holds at least this
    float acc = 0.f;
    uint32_t cnt = 0;
    c0 = 0.f;
#pragma unroll
    for (int u = HLLB_NG - 1; u >= 0; --u) {
        uint32_t x[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) x[s] = 2u * C2[4 * u + s] + (uint32_t)__popc(S[4 * u + s]);
        x[0] -= x[1]; x[2] -= x[3]; x[0] -= x[2]; x[1] -= x[3];      // Moebius inversion, as in aux_subset_bins
#pragma unroll
        for (int s = 3; s >= 0; --s) {
            cnt += x[s];
            if (v0 + 4 * u + s == 0) c0 = (float)x[s];
            else acc = fmaf(acc, 0.5f, (float)x[s]);
        }
    }
    // acc = sum x[v] 2^-(v - vfirst)
    z = acc * (1.f / (float)(1ull << vfirst)) + (float)(nread - cnt) * (1.f / (float)(1ull << vcap));
}

template <int G0, int T0>
__device__ __forceinline__ void aux_bound_segment(const uint4* __restrict__ rq, const uint4* __restrict__ cq, uint32_t np32,
                                                  int wp0, int wp1, uint32_t gmask, uint32_t (&S)[HLLB_NV],
                                                  uint32_t (&C2)[HLLB_NV], float& z, float& c0) {
    aux_bound_accum<G0, T0>(rq, cq, np32, wp0, wp1, gmask, S, C2);
    aux_bound_sums<G0, T0>(S, C2, 64u * (uint32_t)wp1, z, c0);
}

template <int AN>
__global__ void __launch_bounds__(64, HLLB_MIN_CTAS)
k_tile_filter_hll_bound(const uint32_t* __restrict__ auxQ, const uint16_t* __restrict__ agrange,
                        const AuxTail* __restrict__ atail, long long npad, int p_aux, TileWalk tw,
                        const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
                        const unsigned long long* __restrict__ e, float tau, float zs, int order_n,
                        uint2* __restrict__ cand, unsigned long long* __restrict__ cand_count,
                        unsigned long long cand_cap, unsigned long long* __restrict__ unit_counter) {
    __shared__ int s_unit;
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
    uint32_t steps = 0;                         // word-pair steps this warp executed (statistics: unit_counter[1] = M_STEPS)
    const HllCheckpoints cp = hll_checkpoints(p_aux);
    const float m_f = (float)(1 << p_aux);
    const int q = 64 - p_aux;
    const uint32_t np32 = (uint32_t)npad;
    const int uend = tw.count() * 4;
    for (;;) {
        __syncthreads();
        if (t == 0) s_unit = tw.j0 * 4 + (int)atomicAdd(unit_counter, 1ull);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= uend) break;
        const int2 rc = tw.tile(unit >> 2);
        const int r0 = rc.x * TILE + (unit & 3) * 32, c0 = rc.y * TILE;
        for (int item = (int)w; item < 128; item += 2) {
            const int i = r0 + (item >> 2);
            const int k = c0 + (item & 3) * 32 + (int)lane;
            if (i >= n) continue;
            const bool v = k < n && k >= lo[i] && k <= hi[i];
            if (!__any_sync(0xffffffffu, v)) continue;
            const int kc = (int)min((long long)k, npad - 1), kn = min(kc, n - 1);
            const uint32_t ra = agrange[i], rb = agrange[kn];
            int vlo = v ? max((int)(ra & 0xff), (int)(rb & 0xff)) : 255;
            int vhi = v ? max((int)(ra >> 8), (int)(rb >> 8)) : 0;
            for (int o = 16; o; o >>= 1) {
                vlo = min(vlo, __shfl_xor_sync(0xffffffffu, vlo, o));
                vhi = max(vhi, __shfl_xor_sync(0xffffffffu, vhi, o));
            }
            // every union register of the step's pairs holds at least vlo: the counted groups start at vlo's.  The bound
            // needs no register at q+1 (then the estimate's starting point is its first branch, hll.h:657) and vlo below 40
            // (the templates cover windows 0..4 x the first two groups); zs < 0 would make the criteria non-monotone.
            // Otherwise the step's pairs all go to pass B
            const int g0 = min(vlo >> 3, 4), t0 = (vlo >> 2) - 2 * g0;
            bool alive = v;
            if (t0 <= 1 && vhi <= q && zs >= 0.f) {
                const uint32_t gmask = aux_gmask(g0, vlo, vhi);
                const uint4* rq = reinterpret_cast<const uint4*>(auxQ) + i;
                const uint4* cq = reinterpret_cast<const uint4*>(auxQ) + kc;
                const float e1 = (float)e[i], e2 = (float)e[kn];
                const AuxTail ti = atail[i], tk = atail[kn];
                uint32_t S[HLLB_NV], C2[HLLB_NV];
#pragma unroll
                for (int x = 0; x < HLLB_NV; ++x) { S[x] = 0; C2[x] = 0; }
                int wp = 0;
#pragma unroll 1
                for (int seg = 0; seg <= cp.n; ++seg) {
                    const int wend = seg >= cp.n ? cp.nwp : (seg == 0 ? cp.cp[0] : seg == 1 ? cp.cp[1] : cp.cp[2]);
                    float z, zeros;
                    switch (2 * g0 + t0) {
                        case 0: aux_bound_segment<0, 0>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 1: aux_bound_segment<0, 1>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 2: aux_bound_segment<1, 0>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 3: aux_bound_segment<1, 1>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 4: aux_bound_segment<2, 0>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 5: aux_bound_segment<2, 1>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 6: aux_bound_segment<3, 0>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 7: aux_bound_segment<3, 1>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        case 8: aux_bound_segment<4, 0>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                        default: aux_bound_segment<4, 1>(rq, cq, np32, wp, wend, gmask, S, C2, z, zeros); break;
                    }
                    steps += (uint32_t)(wend - wp);
                    wp = wend;
                    if (seg < cp.n) {              // the unread part: at most the smaller of the two genomes' own sums over it
                        z += seg == 0 ? fminf(ti.z[0], tk.z[0]) : seg == 1 ? fminf(ti.z[1], tk.z[1]) : fminf(ti.z[2], tk.z[2]);
                        zeros += seg == 0 ? fminf(ti.c[0], tk.c[0]) : seg == 1 ? fminf(ti.c[1], tk.c[1]) : fminf(ti.c[2], tk.c[2]);
                    }
                    alive = alive && !selb::hll_surely_fails(AN, tau, zs, order_n, m_f, e1, e2, z, zeros);
                    if (!__any_sync(0xffffffffu, alive)) break;
                }
            }
            if (alive) {
                const unsigned long long slot = warp_claim(cand_count);
                if (slot < cand_cap) cand[slot] = make_uint2((uint32_t)i, (uint32_t)k);
            }
        }
    }
    if (lane == 0 && steps) atomicAdd(unit_counter + 1, (unsigned long long)steps);
}

// ---- pass B: exact decision for the candidates -------------------------------------------------------------------
template <int AN>
__global__ void __launch_bounds__(64, HLLP_MIN_CTAS)
k_hll_verify(const uint32_t* __restrict__ auxQ, const uint16_t* __restrict__ agrange, const uint32_t* __restrict__ auxT,
             long long npad, int p_aux, const uint2* __restrict__ cand, const unsigned long long* __restrict__ cand_count,
             unsigned long long cand_cap, int n, const unsigned long long* __restrict__ e, double tau, float zs, int order_n,
             uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count, unsigned long long pair_cap) {
#ifndef SELB_EMUL
    extern __shared__ __align__(1024) uint32_t hist_dyn[];   // [nbins][64 threads]
#endif
    const int nbins = 64 - p_aux + 2;
    const uint32_t t = threadIdx.x, lane = t & 31, tb = t * 4;
    const int nw = (1 << p_aux) >> 5;
    const int words = (1 << p_aux) >> 2;
    uint32_t* hcol = hist_dyn + t;
    const uint32_t bias0 = hist_bias(hist_dyn);
    const long long ncand = (long long)min(*cand_count, cand_cap);
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + t) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp0 * 32; base < ncand; base += nwarps * 32) {
        const bool valid = base + lane < ncand;
        const uint2 pr = valid ? cand[base + lane] : make_uint2(0u, 0u);
        const int i = (int)pr.x, k = (int)pr.y;
        const uint32_t ra = agrange[i], rb = agrange[k];
        const int vlo = max((int)(ra & 0xff), (int)(rb & 0xff)), vhi = max((int)(ra >> 8), (int)(rb >> 8));
        const int g0 = min(vlo >> 3, 4);
        const bool fits = (vhi >> 3) <= g0 + 3;
        // candidates come from anywhere in the band: group the warp's lanes by window, one counting pass per window
        // (every lane runs it on its own pair; only the group's lanes keep the result)
        uint32_t todo = __ballot_sync(0xffffffffu, valid && fits);
        while (todo) {
            const int g = __shfl_sync(0xffffffffu, g0, __ffs((int)todo) - 1);
            const bool mine = valid && fits && g0 == g;
            uint32_t gmask = mine ? aux_gmask(g0, vlo, vhi) : 0u;
            for (int o = 16; o; o >>= 1) gmask |= __shfl_xor_sync(0xffffffffu, gmask, o);     // a superset counts empty groups: harmless
            aux_plane_hist_g(g, auxQ + 4 * (size_t)i, auxQ + 4 * (size_t)k, npad, nw, gmask, hcol, nbins, mine);
            todo &= ~__ballot_sync(0xffffffffu, mine);
        }
        if (valid && !fits) aux_byte_hist(auxT, npad, words, i, k, hcol, nbins, bias0, tb);
        bool pass = false;
        if (valid) {
            bool stopped = false;
            const StopHll stop{tau, e[i], e[k], zs, order_n, AN & 1};
            const double tu = selb::ertl_mle(hcol, p_aux, 64, stop, &stopped);
            pass = !stopped && stop.crit(tu);
        }
        if (pass) {
            const unsigned long long slot = warp_claim(pair_count);
            if (slot < pair_cap) pairs[slot] = pr;
        }
    }
}
