// estimate_sort.inl — K6 union estimate + emit, K7 print order of sparse outputs (part of selb200.cu)
// ============================================================================
// K6: union estimate -> Jaccard -> tau test -> emit
//   reference: hll.h:1206 (calculate_estimate(counts, ERTL_MLE...)), selection.cpp:286-288
// ============================================================================
// J is non-increasing in t: once it is below tau (and outside the near-tau window) at the MLE's lower bound the pair can
// neither be emitted nor listed as near
struct StopJ {
    double tau, slack;
    unsigned long long e1, e2;
    __device__ __forceinline__ bool operator()(double t_lb) const { return selb::jaccard(e1, e2, t_lb) < tau - slack; }
};
// the same test at the estimator's starting point only: stops there either way and says which way
struct StopJFirst {
    StopJ j;
    bool* fails;
    __device__ __forceinline__ bool operator()(double t_lb) const { *fails = j(t_lb); return true; }
};

// the row's non-empty bins from sixteen independent 128-bit loads (the estimator's own scans for the first and the last
// one are a chain of dependent loads: ~35 round trips before the arithmetic starts); the row is in L1 afterwards
__device__ __forceinline__ unsigned long long hist_row_nonzero(const uint32_t* __restrict__ row64) {
    const uint4* row = reinterpret_cast<const uint4*>(row64);
    unsigned long long nz = 0ull;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint4 v = __ldg(row + j);
        const unsigned long long b = (v.x ? 1ull : 0ull) | (v.y ? 2ull : 0ull) | (v.z ? 4ull : 0ull) | (v.w ? 8ull : 0ull);
        nz |= b << (4 * j);
    }
    return nz;
}

// K6 in two steps, for pair lists that were NOT filtered by an auxiliary sketch (criterion cb: C2's 4.7 M pairs, five
// thousand of which pass).  One passing lane makes its whole warp walk the fp64 secant iterations, so
// k_estimate_screen first settles the pairs that fail J >= tau at the estimator's starting point and lists the rest;
// k_estimate_emit then runs the full estimator on the list, all lanes busy: 0.93 -> 0.54 ms on C2.  After smh_a / hll_a /
// hll_an most pairs of the list are similar enough to survive the screen (it starts 7 % below the estimate), and the
// single kernel is faster: 0.105 against 0.141 ms on C4.
__global__ void __launch_bounds__(128)
k_estimate_screen(const uint32_t* __restrict__ hist, const uint2* __restrict__ pairs,
                  const unsigned long long* __restrict__ npairs_dev, unsigned long long npairs_cap,
                  const unsigned long long* __restrict__ e, int p, double tau,
                  uint32_t* __restrict__ surv, unsigned long long* __restrict__ surv_count,
                  const uint32_t* __restrict__ wide_flag = nullptr, uint32_t epoch = 0u) {
    const long long npairs = (long long)min(*npairs_dev, npairs_cap);
    for (long long pi = blockIdx.x * (long long)blockDim.x + threadIdx.x; pi < npairs;
         pi += (long long)gridDim.x * blockDim.x) {
        if (wide_flag && wide_flag[pi] == epoch) continue;         // a wide pair: its row is being written on the other stream
        const uint2 pr = pairs[pi];
        const unsigned long long nz = hist_row_nonzero(hist + pi * 64);
        const unsigned long long e1 = e[pr.x], e2 = e[pr.y];
        if (nz == 0ull) continue;                                  // no histogram (cannot happen for a counted pair)
        bool fails = false, stopped = false;
        selb::ertl_mle_range(hist + pi * 64, p, 1, __ffsll((long long)nz) - 1, 63 - __clzll((long long)nz),
                             StopJFirst{StopJ{tau, 1e-6 * fabs(tau), e1, e2}, &fails}, &stopped);
        // stopped && fails: settled.  Anything else (not settled at the starting point, or the estimator returned without
        // iterating: saturated or converged at once) goes through the full estimator
        if (stopped && fails) continue;
        const unsigned long long slot = warp_claim(surv_count);
        surv[slot] = (uint32_t)pi;                                 // the list is as long as the pair list: cannot overflow
    }
}

__global__ void __launch_bounds__(128)
k_estimate_emit(const uint32_t* __restrict__ hist, const uint2* __restrict__ pairs,
                const uint32_t* __restrict__ surv, const unsigned long long* __restrict__ count_dev, unsigned long long count_cap,
                const unsigned long long* __restrict__ e, int p, double tau,
                uint64_t* __restrict__ out_keys, double* __restrict__ out_j,
                unsigned long long* __restrict__ out_count, unsigned long long out_cap,
                uint64_t* __restrict__ near_keys, double* __restrict__ near_j,
                unsigned long long* __restrict__ near_count, unsigned long long near_cap,
                const uint32_t* __restrict__ wide_flag = nullptr, uint32_t epoch = 0u) {
    // surv == nullptr: every pair of the list (count_dev = its length); else the listed pairs (what k_estimate_screen
    // kept, or the wide list).  wide_flag: pairs to leave out (they are estimated from the wide list)
    const long long nsurv = (long long)min(*count_dev, count_cap);
    for (long long si = blockIdx.x * (long long)blockDim.x + threadIdx.x; si < nsurv;
         si += (long long)gridDim.x * blockDim.x) {
        const long long pi = surv ? (long long)surv[si] : si;
        if (wide_flag && wide_flag[pi] == epoch) continue;
        const uint2 pr = pairs[pi];
        const unsigned long long nz = hist_row_nonzero(hist + pi * 64);
        const unsigned long long e1 = e[pr.x], e2 = e[pr.y];
        bool stopped = false;
        const double t = selb::ertl_mle_range(hist + pi * 64, p, 1, __ffsll((long long)nz) - 1, 63 - __clzll((long long)nz),
                                              StopJ{tau, 1e-6 * fabs(tau), e1, e2}, &stopped);
        if (stopped) continue;
        const double jac = selb::jaccard(e1, e2, t);
        const uint64_t key = ((uint64_t)pr.x << 32) | pr.y;
        if (jac >= tau) {
            const unsigned long long slot = warp_claim(out_count);
            if (slot < out_cap) { out_keys[slot] = key; out_j[slot] = jac; }
        }
        if (fabs(jac - tau) <= 1e-6 * fabs(tau)) {
            const unsigned long long slot = warp_claim(near_count);
            if (slot < near_cap) { near_keys[slot] = key; near_j[slot] = jac; }
        }
    }
}

// ============================================================================
// K7: (i,k) print order of the reference (selection.cpp:297-300) for SPARSE outputs: bucket by row
// (count -> scan -> scatter), then every element finds its place inside its row by counting the
// smaller columns.  Four small launches instead of the ~9 of a 49-bit radix sort; rows hold a
// handful of pairs (cluster mates), so the quadratic in-row step is a few loads per element.
// ============================================================================
__global__ void __launch_bounds__(256)
k_rowsort_count(const uint64_t* __restrict__ keys, long long cnt, int32_t* __restrict__ rowcnt) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e < cnt) atomicAdd(rowcnt + (keys[e] >> 32), 1);
}

__global__ void __launch_bounds__(256)
k_rowsort_scatter(const uint64_t* __restrict__ keys, const double* __restrict__ jac, long long cnt,
                  int32_t* __restrict__ rowcnt, const int32_t* __restrict__ rowoff,
                  uint64_t* __restrict__ tkeys, double* __restrict__ tj) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= cnt) return;
    const uint64_t key = keys[e];
    const uint32_t i = (uint32_t)(key >> 32);
    const int pos = rowoff[i] + atomicSub(rowcnt + i, 1) - 1;    // counts back down to zero
    tkeys[pos] = key;
    tj[pos] = jac[e];
}

__global__ void __launch_bounds__(256)
k_rowsort_rank(const uint64_t* __restrict__ tkeys, const double* __restrict__ tj, long long cnt,
               const int32_t* __restrict__ rowoff, uint64_t* __restrict__ out_keys, double* __restrict__ out_j) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= cnt) return;
    const uint64_t key = tkeys[e];
    const uint32_t i = (uint32_t)(key >> 32);
    const int a = rowoff[i], b = rowoff[i + 1];
    int r = 0;
    for (int t = a; t < b; ++t) r += tkeys[t] < key;      // keys are unique
    out_keys[a + r] = key;
    out_j[a + r] = tj[e];
}

#ifndef SELB_EMUL
// The print-order sort as ONE cooperative launch: zero the row counters, count, exclusive prefix sums over the n + 1 rows
// (every CTA sums and scans a contiguous segment; the CTAs' totals meet in `blocksum`), scatter, rank.  Same steps as the
// four kernels + memset + library scan above (which stay for the emulator and for devices without cooperative launch):
// seven stream operations of 2 - 6 us each become one.
__global__ void __launch_bounds__(256)
k_rowsort_fused(const uint64_t* __restrict__ keys, const double* __restrict__ jac, long long cnt, int n,
                int32_t* __restrict__ rowcnt, int32_t* __restrict__ rowoff, int32_t* __restrict__ blocksum,
                uint64_t* __restrict__ tkeys, double* __restrict__ tj, uint64_t* __restrict__ out_keys, double* __restrict__ out_j) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    typedef cub::BlockScan<int, 256> Scan;
    typedef cub::BlockReduce<int, 256> Reduce;
    __shared__ union { typename Scan::TempStorage scan; typename Reduce::TempStorage red; } tmp;
    __shared__ int carry;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i <= n; i += nth) rowcnt[i] = 0;
    grid.sync();
    for (long long e = tid; e < cnt; e += nth) atomicAdd(rowcnt + (keys[e] >> 32), 1);
    grid.sync();
    // segment of this CTA in rowcnt[0 .. n]
    const int seg = (n + 1 + (int)gridDim.x - 1) / (int)gridDim.x;
    const int a = min(n + 1, (int)blockIdx.x * seg), b = min(n + 1, a + seg);
    {
        int part = 0;
        for (int i = a + (int)threadIdx.x; i < b; i += 256) part += rowcnt[i];
        const int tot = Reduce(tmp.red).Sum(part);
        if (threadIdx.x == 0) blocksum[blockIdx.x] = tot;
    }
    grid.sync();
    {
        int part = 0;
        for (int q = (int)threadIdx.x; q < (int)blockIdx.x; q += 256) part += blocksum[q];
        const int before = Reduce(tmp.red).Sum(part);
        if (threadIdx.x == 0) carry = before;
        __syncthreads();
        for (int base = a; base < b; base += 256) {
            const int idx = base + (int)threadIdx.x;
            const int v = idx < b ? rowcnt[idx] : 0;
            int ex, tot;
            Scan(tmp.scan).ExclusiveSum(v, ex, tot);
            const int c0 = carry;
            if (idx < b) rowoff[idx] = c0 + ex;
            __syncthreads();
            if (threadIdx.x == 0) carry = c0 + tot;
            __syncthreads();
        }
    }
    grid.sync();
    for (long long e = tid; e < cnt; e += nth) {
        const uint64_t key = keys[e];
        const uint32_t i = (uint32_t)(key >> 32);
        const int pos = rowoff[i] + atomicSub(rowcnt + i, 1) - 1;    // counts back down to zero
        tkeys[pos] = key;
        tj[pos] = jac[e];
    }
    grid.sync();
    for (long long e = tid; e < cnt; e += nth) {
        const uint64_t key = tkeys[e];
        const uint32_t i = (uint32_t)(key >> 32);
        const int ra = rowoff[i], rb = rowoff[i + 1];
        int r = 0;
        for (int t = ra; t < rb; ++t) r += tkeys[t] < key;      // keys are unique
        out_keys[ra + r] = key;
        out_j[ra + r] = tj[e];
    }
}
#endif
