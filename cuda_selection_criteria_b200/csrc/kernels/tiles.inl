// tiles.inl — K2: CB band per row and the band's tile list, built on the device (part of selb200.cu)
// ============================================================================
// K2: CB band per sorted row
//   reference: src/selection.cpp:278-283 — skip e2==0, break at the first CB failure.
//   Sorted ascending + correctly-rounded fp64 division => the passing set of row i is the
//   contiguous range [lo(i), hi(i)], lo = max(i+1, first index with e>0).
// ============================================================================
__device__ __forceinline__ void cb_bounds_row(int i, const unsigned long long* __restrict__ e, int n, int zeros, double tau,
                                              int32_t* __restrict__ lo, int32_t* __restrict__ hi) {
    const unsigned long long e1 = e[i];
    const int l = max(i + 1, zeros);
    int a = l, b = n;   // first k in [l,n) failing CB
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (selb::crit_cb(tau, e1, e[mid])) a = mid + 1; else b = mid;
    }
    lo[i] = l;
    hi[i] = a - 1;
}
__global__ void k_cb_bounds(const unsigned long long* __restrict__ e, int n, int zeros, double tau,
                            int32_t* __restrict__ lo, int32_t* __restrict__ hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cb_bounds_row(i, e, n, zeros, tau, lo, hi);
}

// ============================================================================
// tile list of the CB band, built on the device (no host round trip):
//   k_rowblock_span : per 128-row block, the column-block span of its band and its pair count
//   cub exclusive scan over the spans -> first tile index of every row block
//   k_tile_table    : (row block, column block) of every tile, so that a filter CTA finds its
//                     tile with one 8-byte load
// meta[] (unsigned long long, device): [0] candidates [1] pairs [2] out [3] near of the current
// range, [4] pairs inside the CB band, [5] tiles of the band, [6] gather: pushed flag
// ============================================================================
enum { M_CAND = 0, M_PAIRS = 1, M_OUT = 2, M_NEAR = 3, M_PAIRS_CB = 4, M_TILES = 5, M_PUSHED = 6, M_WIDE = 7, M_BATCH = 8, M_KERR = 9, M_UNIT = 10, M_STEPS = 11, M_ITEMS = 12, M_ITEMS_MAX = 13, M_SURV = 14, M_WORDS = 16 };

// one warp per row block (rb in [0, nrb]; rb == nrb writes the scan sentinel)
__device__ __forceinline__ void rowblock_span_warp(int rb, int lane, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                                                   int n, int nrb, int32_t* __restrict__ nt, int32_t* __restrict__ cb0,
                                                   unsigned long long* __restrict__ rb_pairs, unsigned long long* __restrict__ meta) {
    if (rb == nrb) { if (lane == 0) nt[nrb] = 0; return; }   // scan sentinel: prefix[nrb] = total
    int cmin = INT32_MAX, cmax = -1;
    unsigned long long cnt = 0;
    for (int i = rb * TILE + lane; i < min(n, (rb + 1) * TILE); i += 32) {
        const int l = lo[i], h = hi[i];
        if (h < l) continue;
        cnt += (unsigned long long)(h - l + 1);
        cmin = min(cmin, l);
        cmax = max(cmax, h);
    }
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    if (lane == 0) {
        nt[rb] = cnt ? cmax / TILE - cmin / TILE + 1 : 0;
        cb0[rb] = cnt ? cmin / TILE : 0;
        rb_pairs[rb] = cnt;
        if (cnt) atomicAdd(meta + M_PAIRS_CB, cnt);
    }
}
__global__ void __launch_bounds__(128)
k_rowblock_span(const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n, int nrb,
                int32_t* __restrict__ nt, int32_t* __restrict__ cb0, unsigned long long* __restrict__ rb_pairs,
                unsigned long long* __restrict__ meta) {
    const int rb = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rb > nrb) return;
    rowblock_span_warp(rb, lane, lo, hi, n, nrb, nt, cb0, rb_pairs, meta);
}

__device__ __forceinline__ void tile_table_warp(int rb, int lane, const int32_t* __restrict__ tile_prefix, const int32_t* __restrict__ cb0,
                                                int nrb, long long tile_cap, int2* __restrict__ tile_rc, unsigned long long* __restrict__ meta) {
    const int a = tile_prefix[rb], cnt = tile_prefix[rb + 1] - a, c0 = cb0[rb];
    for (int t = lane; t < cnt; t += 32)
        if (a + t < tile_cap) tile_rc[a + t] = make_int2(rb, c0 + t);
    if (rb == 0 && lane == 0) meta[M_TILES] = (unsigned long long)tile_prefix[nrb];
}
__global__ void __launch_bounds__(128)
k_tile_table(const int32_t* __restrict__ tile_prefix, const int32_t* __restrict__ cb0, int nrb, long long tile_cap,
             int2* __restrict__ tile_rc, unsigned long long* __restrict__ meta) {
    const int rb = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rb >= nrb) return;
    tile_table_warp(rb, lane, tile_prefix, cb0, nrb, tile_cap, tile_rc, meta);
}

#ifndef SELB_EMUL
// The four steps above as ONE cooperative launch (grid-wide barriers between them): at n = 100k each of them runs for
// 3 - 7 us, so four launches and their gaps are most of the 27 us they took.  The emulator (and a device without
// cooperative launch) runs the separate kernels; both call the same device functions.
__global__ void __launch_bounds__(256)
k_bounds_fused(const unsigned long long* __restrict__ e, int n, int zeros, double tau, int32_t* __restrict__ lo,
               int32_t* __restrict__ hi, int nrb, int32_t* __restrict__ nt, int32_t* __restrict__ tile_prefix,
               int32_t* __restrict__ cb0, unsigned long long* __restrict__ rb_pairs, long long tile_cap,
               int2* __restrict__ tile_rc, unsigned long long* __restrict__ meta) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = tid >> 5, nwarps = nth >> 5;
    for (int i = tid; i < n; i += nth) cb_bounds_row(i, e, n, zeros, tau, lo, hi);
    grid.sync();
    for (int rb = warp; rb <= nrb; rb += nwarps) rowblock_span_warp(rb, lane, lo, hi, n, nrb, nt, cb0, rb_pairs, meta);
    grid.sync();
    if (blockIdx.x == 0) {              // exclusive prefix sums of nt[0 .. nrb] (a few hundred to a few thousand values): one CTA
        typedef cub::BlockScan<int, 256> Scan;
        __shared__ typename Scan::TempStorage tmp;
        __shared__ int carry;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int base = 0; base <= nrb; base += 256) {
            const int idx = base + (int)threadIdx.x;
            const int v = idx <= nrb ? nt[idx] : 0;
            int ex, tot;
            Scan(tmp).ExclusiveSum(v, ex, tot);
            const int c0 = carry;
            if (idx <= nrb) tile_prefix[idx] = c0 + ex;
            __syncthreads();
            if (threadIdx.x == 0) carry = c0 + tot;
            __syncthreads();
        }
    }
    grid.sync();
    for (int rb = warp; rb < nrb; rb += nwarps) tile_table_warp(rb, lane, tile_prefix, cb0, nrb, tile_cap, tile_rc, meta);
}
#endif

// tiles owned by one shard: tile = shard + j * n_shards for j in [0, count)
struct TileWalk {
    const int2* tile_rc;
    const unsigned long long* meta;
    long long tile_cap;
    int shard, n_shards;
    int j0, j1;          // this launch covers j in [j0, j1) (clipped to the shard's tile count)
    __device__ __forceinline__ int count() const {
        const long long total = (long long)min((unsigned long long)tile_cap, meta[M_TILES]);
        const long long mine = total > shard ? (total - shard + n_shards - 1) / n_shards : 0;
        return (int)min((long long)j1, mine);
    }
    __device__ __forceinline__ int2 tile(int j) const { return __ldg(tile_rc + shard + (long long)j * n_shards); }
};
