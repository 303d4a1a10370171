// shims.inl — link-level drop-in for the reference's kernel launchers (part of selb200.cu).
//
// Exports, with C++ linkage and the exact signatures of src/selection_kernels_wrapper.hpp:11-45,
//   void launch_kernel_smh  (...)   _Z17launch_kernel_smhPKhPKmPKdPK4int2idiiiiP6ResultPii
//   void launch_kernel_CBsmh(...)   _Z19launch_kernel_CBsmhPKhPKmPKdPK4int2idiiiiP6ResultPii
// so that the reference's selection_cuda.o / time_smh_cuda.o link against libselb200.so
// unchanged.  Contract kept (src/selection_kernels.cu:119-177): the CALLER owns every device
// buffer; out_count is zeroed; work is queued asynchronously on the default stream; nothing is
// returned.  Semantics kept (src/selection_kernels.cu:13-117, include/criteria_sketch_cuda.cuh):
// per listed pair, smh_a band test on aux rows of stride m_aux, then the Flajolet-ORIGINAL union
// estimate of the first 2^14 registers of main rows of stride m_hll, Jaccard on the UNTRUNCATED
// cards, |J| >= tau, Result{i,k,(float)J} appended in atomic order.  (Both reference kernels are
// the same program; "CBsmh" applies no CB test — SURVEY.md §2.3.)  This is the reference GPU
// path's arithmetic, NOT the CPU oracle's (which is Ertl-MLE): it exists for link compatibility
// and as the "existing GPU kernel" speed bar.
//
// Implementation: thread-per-pair smh_a compaction, then the same warp-per-pair histogram kernel
// as the selection path with a fused epilogue (the estimator needs only the zero count and
// sum_r c[r]*2^-r, both exact warp reductions).

struct Result {   // src/selection_kernels_wrapper.hpp:6-9
    int x, y;
    float sim;
};

namespace {

__global__ void k_shim_smh_compact(const uint64_t* __restrict__ aux, const int2* __restrict__ pairs, int total_pairs,
                                   int m_aux, int n_rows, int n_bands, uint32_t* __restrict__ surv,
                                   unsigned long long* __restrict__ surv_count) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_pairs) return;
    const int2 pr = pairs[idx];
    const uint64_t* v1 = aux + (long long)pr.x * m_aux;     // criteria_sketch_cuda.cuh:16-28
    const uint64_t* v2 = aux + (long long)pr.y * m_aux;
    bool hit = false;
    for (int b = 0; b < n_bands && !hit; ++b) {
        bool eq = true;
        for (int r = 0; r < n_rows; ++r)
            if (v1[b * n_rows + r] != v2[b * n_rows + r]) { eq = false; break; }
        hit = eq;
    }
    if (hit) surv[warp_claim(surv_count)] = (uint32_t)idx;
}

struct SrcShim {             // survivors -> caller's int2 pair list (row indices)
    // caller-owned, unvalidated register bytes: six bits are kept and the histogram has all 64 bins, so no byte can
    // address shared memory outside the counters (the reference kernel feeds any byte to ldexp; a byte above 63 is
    // not an HLL register of any precision, and for those inputs the shim's estimate differs from the reference's)
    static constexpr uint32_t kMask = 0x3f3f3f3fu;
    const int2* pairs;
    const uint32_t* surv;
    const unsigned long long* surv_count;
    __device__ __forceinline__ long long count() const { return (long long)*surv_count; }
    __device__ __forceinline__ uint2 rows(long long pi, uint2& id) const {
        const int2 pr = pairs[surv[pi]];
        id = make_uint2((uint32_t)pr.x, (uint32_t)pr.y);
        return id;
    }
    __device__ __forceinline__ long long slot(long long pi) const { return pi; }
};

struct EpiFlajolet {         // criteria_sketch_cuda.cuh:30-65 + selection_kernels.cu:41-59
    const double* cards;
    double tau;
    Result* out;
    int* out_count;
    __device__ __forceinline__ void operator()(long long, uint2 id, uint32_t s0, uint32_t s1, uint32_t lane) const {
        // lane holds counts of register values `lane` and `lane+32`
        double sum = 0.0;
        if (lane > 0) sum += ldexp((double)s0, -(int)lane);
        sum += ldexp((double)s1, -(int)(lane + 32));
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const uint32_t zeros = __shfl_sync(0xffffffffu, s0, 0);
        if (lane != 0) return;
        const int m = 1 << 14;
        const double alpha = 0.7213 / (1 + 1.079 / m);
        double raw = alpha * m * m / (zeros + sum);
        if (raw < 2.5 * m && zeros) raw = m * log(double(m) / zeros);
        // criteria_sketch_cuda.cuh:61-63 negates the UNSIGNED 1ULL<<32 (= 2^64 - 2^32) before the conversion to double:
        // the "corrected" estimate comes out negative and the pair is dropped below — reproduced, not repaired
        else if (raw > (1ULL << 32) / 30.0) raw = (double)(-(1ULL << 32)) * log1p(-raw / (1ULL << 32));
        if (raw == 0.0) return;
        if (!isfinite(raw) || raw < 0.0) return;
        double jac = (cards[id.x] + cards[id.y] - raw) / raw;
        if (jac < 0.0 && jac != 0.0) jac = -jac;
        if (!isfinite(jac)) return;
        if (jac < tau) return;
        const int slot = atomicAdd(out_count, 1);
        out[slot] = Result{(int)id.x, (int)id.y, (float)jac};
    }
};

struct ShimScratch {
    uint32_t* surv = nullptr;
    unsigned long long* count = nullptr;
    size_t cap = 0;
    int sm_count = 0;
};
thread_local ShimScratch g_shim;

void shim_launch(const uint8_t* main_sketches, const uint64_t* aux, const double* cards, const int2* pairs,
                 int total_pairs, double tau, int m_aux, int m_hll, int n_rows, int n_bands, Result* out,
                 int* out_count, int blockSize) {
    cudaMemsetAsync(out_count, 0, sizeof(int), 0);
    if (total_pairs <= 0) return;
    ShimScratch& sc = g_shim;
    if (!sc.count) {
        cudaMalloc(&sc.count, 8);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sc.sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    if (sc.cap < (size_t)total_pairs) {
        if (sc.surv) cudaFree(sc.surv);
        cudaMalloc(&sc.surv, (size_t)total_pairs * 4);
        sc.cap = (size_t)total_pairs;
    }
    cudaMemsetAsync(sc.count, 0, 8, 0);
    if (blockSize < 32 || blockSize > 1024) blockSize = 256;
    const int grid = (total_pairs + blockSize - 1) / blockSize;
    k_shim_smh_compact<<<grid, blockSize, 0, 0>>>(aux, pairs, total_pairs, m_aux, n_rows, n_bands, sc.surv, sc.count);
    SrcShim src{pairs, sc.surv, sc.count};
    EpiFlajolet epi{cards, tau, out, out_count};
    // hll_union_card hard-codes p = 14 (criteria_sketch_cuda.cuh:33); rows are m_hll apart
    launch_pair_hist_t((cudaStream_t)0, sc.sm_count, main_sketches, (size_t)m_hll, (size_t)1 << 14, 14,
                       (int64_t)total_pairs, src, epi);
}

}  // namespace

void launch_kernel_smh(const uint8_t* main_sketches, const uint64_t* aux_sketches, const double* cards,
                       const int2* pairs, int total_pairs, double tau, int m_aux, int m_hll, int n_rows,
                       int n_bands, Result* out, int* out_count, int blockSize) {
    shim_launch(main_sketches, aux_sketches, cards, pairs, total_pairs, tau, m_aux, m_hll, n_rows, n_bands, out,
                out_count, blockSize);
}

void launch_kernel_CBsmh(const uint8_t* main_sketches, const uint64_t* aux_sketches, const double* cards,
                         const int2* pairs, int total_pairs, double tau, int m_aux, int m_hll, int n_rows,
                         int n_bands, Result* out, int* out_count, int blockSize) {
    shim_launch(main_sketches, aux_sketches, cards, pairs, total_pairs, tau, m_aux, m_hll, n_rows, n_bands, out,
                out_count, blockSize);
}
