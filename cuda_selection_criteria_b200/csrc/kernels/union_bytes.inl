// union_bytes.inl — K5, byte form: warp-per-pair register max + histogram with shared-memory counters (part of selb200.cu)
// ============================================================================
// K5: warp-per-pair register max + histogram (primary HLL, m >= 512)
//   reference: sketch/include/sketch/hll.h:1188-1206 (union_size: _mm_max_epu8 + 64-bin counts)
//   CTA = 2 warps, each warp owns its own pairs; NB bins x 64 threads x 4 B static smem.
//   Src  : where pair (row a, row b) number pi comes from, and how many there are
//   Epi  : what happens to the finished histogram (lane L holds bins L and L+32)
// ============================================================================
// kMask: ANDed onto every loaded register word.  The selection path validates its registers at load time (k_max_byte,
// SrcSelf) and passes all-ones; the link-level shims histogram caller-owned, unvalidated bytes and keep six bits, so
// that no byte can address a counter outside the 64-bin array.
struct SrcPairs {            // pair list of the selection path: sorted positions, mapped through `order`
    static constexpr uint32_t kMask = 0xffffffffu;
    const uint2* pairs;
    const int32_t* order;    // nullptr: entries are row indices already
    long long n;             // count, or the capacity when n_dev is given
    const unsigned long long* n_dev;   // optional: the count lives in device memory (no host sync)
    __device__ __forceinline__ long long count() const {
        return n_dev ? (long long)min((unsigned long long)n, *n_dev) : n;
    }
    __device__ __forceinline__ uint2 rows(long long pi, uint2& id) const {
        id = pairs[pi];
        return order ? make_uint2((uint32_t)order[id.x], (uint32_t)order[id.y]) : id;
    }
    __device__ __forceinline__ long long slot(long long pi) const { return pi; }   // histogram row of pair pi
};

struct SrcSelf {             // rows g0..g0+n-1 against themselves: per-genome histograms (max(a,a) = a)
    static constexpr uint32_t kMask = 0xffffffffu;
    long long g0, n;
    const uint32_t* max_seen;   // written by k_max_byte earlier on the stream: a register above
    uint32_t max_ok;            // 64-p+1 would index past the histogram, so nothing is processed
    __device__ __forceinline__ long long count() const { return *max_seen > max_ok ? 0 : n; }
    __device__ __forceinline__ uint2 rows(long long pi, uint2& id) const {
        id = make_uint2((uint32_t)(g0 + pi), (uint32_t)(g0 + pi));
        return id;
    }
    __device__ __forceinline__ long long slot(long long pi) const { return pi; }
};

struct EpiWriteHist {        // histogram rows for k_estimate_emit
    uint32_t* out;
    __device__ __forceinline__ void operator()(long long pi, uint2, uint32_t s0, uint32_t s1, uint32_t lane) const {
        out[pi * 64 + lane] = s0;
        out[pi * 64 + 32 + lane] = s1;
    }
};

template <int NB, class Src, class Epi>
__global__ void __launch_bounds__(64)
k_pair_hist(const uint8_t* __restrict__ regs, size_t row_stride, size_t m, Src src, Epi epi) {
    __shared__ __align__(1024) uint32_t hist[NB * 64];
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5, tb = t * 4;
    if ((long long)blockIdx.x * 2 >= src.count()) return;      // nothing for this CTA (the wide list is usually empty)
    const uint32_t bias = hist_bias(hist);
#pragma unroll 4
    for (int b = 0; b < NB; ++b) hist[b * 64 + t] = 0;
    __syncwarp();
    const int nchunk = (int)(m >> 9);   // 512 B per warp-wide 128-bit load
    const int ngroups = nchunk >> 2;    // software pipeline works on groups of 4 chunks
    const long long nw = (long long)gridDim.x * 2;
    const long long npairs = src.count();
    uint32_t prev0 = 0, prev1 = 0;
    auto ld = [](const uint4* q) {
        uint4 v = __ldg(q);
        if (Src::kMask != 0xffffffffu) { v.x &= Src::kMask; v.y &= Src::kMask; v.z &= Src::kMask; v.w &= Src::kMask; }
        return v;
    };
    for (long long pi = (long long)blockIdx.x * 2 + w; pi < npairs; pi += nw) {
        uint2 id;
        const uint2 rw = src.rows(pi, id);
        const uint4* a = reinterpret_cast<const uint4*>(regs + (size_t)rw.x * row_stride) + lane;
        const uint4* b = reinterpret_cast<const uint4*>(regs + (size_t)rw.y * row_stride) + lane;
        if (ngroups) {
            // two chunks being histogrammed while the next two are in flight (no register rotation)
            uint4 ax0 = ld(a), ay0 = ld(b), ax1 = ld(a + 32), ay1 = ld(b + 32);
            for (int g = 0; g < ngroups; ++g) {
                const uint4 bx0 = ld(a + 64), by0 = ld(b + 64), bx1 = ld(a + 96), by1 = ld(b + 96);
                hist_inc_max16(ax0, ay0, bias, tb);
                hist_inc_max16(ax1, ay1, bias, tb);
                a += 128; b += 128;
                if (g + 1 < ngroups) { ax0 = ld(a); ay0 = ld(b); ax1 = ld(a + 32); ay1 = ld(b + 32); }
                hist_inc_max16(bx0, by0, bias, tb);
                hist_inc_max16(bx1, by1, bias, tb);
            }
        }
        for (int c = ngroups * 4; c < nchunk; ++c, a += 32, b += 32) hist_inc_max16(ld(a), ld(b), bias, tb);
        __syncwarp();
        // transposed, conflict-free column sums: lane L totals bins L and L+32.  Counters are
        // never cleared: they run cumulatively (mod 2^32) and the pair's histogram is the
        // difference to the previous totals, which saves the 64 clearing stores per pair.
        uint32_t s0 = 0, s1 = 0;
        const uint32_t cb = w * 32;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const uint32_t col = cb + ((lane + r) & 31);
            s0 += hist[lane * 64 + col];
            if (lane + 32 < NB) s1 += hist[(lane + 32) * 64 + col];
        }
        __syncwarp();
        epi(src.slot(pi), id, s0 - prev0, s1 - prev1, lane);
        prev0 = s0;
        prev1 = s1;
    }
}
