// synth.cu — "synth-v1" synthetic sketch generator (bench + tests), device kernel and the
// identical host loop.  See synth.cuh for the model.  Not part of the selection path.
#include "../../include/selb200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "synth.cuh"

namespace {

__global__ void k_synth_hll(long long n, int p, const int32_t* __restrict__ cluster,
                            const uint64_t* __restrict__ thr_core, const uint64_t* __restrict__ thr_priv,
                            uint64_t seed, uint32_t tag, uint8_t* __restrict__ out) {
    const size_t m = (size_t)1 << p;
    const uint32_t vmax = (uint32_t)(64 - p + 1);
    const size_t total = (size_t)n * (m / 4);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const long long g = (long long)(idx / (m / 4));
        const uint64_t j0 = (uint64_t)(idx % (m / 4)) * 4;
        const int32_t cl = cluster[g];
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            w |= (uint32_t)selb::synth_hll_reg(seed, tag, g, cl, j0 + b, thr_core, thr_priv, vmax) << (8 * b);
        reinterpret_cast<uint32_t*>(out)[idx] = w;
    }
}

__global__ void k_synth_smh(long long n, int m, const int32_t* __restrict__ cluster,
                            const uint64_t* __restrict__ range_core, const uint64_t* __restrict__ range_priv,
                            uint64_t seed, uint32_t tag, uint64_t* __restrict__ out) {
    const size_t total = (size_t)n * m;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const long long g = (long long)(idx / m);
        const uint64_t j = idx % m;
        out[idx] = selb::synth_smh_bucket(seed, tag, g, cluster[g], j, range_core, range_priv);
    }
}

struct Tmp {
    void* p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
    cudaError_t put(const void* src, size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    }
};

}  // namespace

extern "C" {

int selb200_synth_hll(int on_device, int device, int64_t n, int p, const int32_t* cluster, int64_t n_clusters,
                      const uint64_t* thr_core, const uint64_t* thr_priv, uint64_t seed, uint32_t tag,
                      uint8_t* out) {
    if (n < 0 || p < 4 || p > 20 || !cluster || !thr_core || !thr_priv || !out) return SELB200_EINVAL;
    const size_t m = (size_t)1 << p;
    const uint32_t vmax = (uint32_t)(64 - p + 1);
    if (!on_device) {
#pragma omp parallel for schedule(static)
        for (long long g = 0; g < (long long)n; ++g)
            for (size_t j = 0; j < m; ++j)
                out[(size_t)g * m + j] = selb::synth_hll_reg(seed, tag, g, cluster[g], j, thr_core, thr_priv, vmax);
        return SELB200_OK;
    }
    if (cudaSetDevice(device) != cudaSuccess) return SELB200_ECUDA;
    Tmp dc, dtc, dtp;
    if (dc.put(cluster, (size_t)n * 4) != cudaSuccess || dtc.put(thr_core, (size_t)n_clusters * 64 * 8) != cudaSuccess ||
        dtp.put(thr_priv, (size_t)n * 64 * 8) != cudaSuccess)
        return SELB200_ECUDA;
    if (n) {
        k_synth_hll<<<148 * 16, 256>>>(n, p, (const int32_t*)dc.p, (const uint64_t*)dtc.p, (const uint64_t*)dtp.p, seed,
                                       tag, out);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return SELB200_ECUDA;
    }
    return SELB200_OK;
}

int selb200_synth_smh(int on_device, int device, int64_t n, int m, const int32_t* cluster, int64_t n_clusters,
                      const uint64_t* range_core, const uint64_t* range_priv, uint64_t seed, uint32_t tag,
                      uint64_t* out) {
    if (n < 0 || m < 1 || !cluster || !range_core || !range_priv || !out) return SELB200_EINVAL;
    if (!on_device) {
#pragma omp parallel for schedule(static)
        for (long long g = 0; g < (long long)n; ++g)
            for (int j = 0; j < m; ++j)
                out[(size_t)g * m + j] = selb::synth_smh_bucket(seed, tag, g, cluster[g], (uint64_t)j, range_core, range_priv);
        return SELB200_OK;
    }
    if (cudaSetDevice(device) != cudaSuccess) return SELB200_ECUDA;
    Tmp dc, drc, drp;
    if (dc.put(cluster, (size_t)n * 4) != cudaSuccess || drc.put(range_core, (size_t)n_clusters * 8) != cudaSuccess ||
        drp.put(range_priv, (size_t)n * 8) != cudaSuccess)
        return SELB200_ECUDA;
    if (n) {
        k_synth_smh<<<148 * 8, 256>>>(n, m, (const int32_t*)dc.p, (const uint64_t*)drc.p, (const uint64_t*)drp.p, seed,
                                      tag, out);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return SELB200_ECUDA;
    }
    return SELB200_OK;
}

}  // extern "C"
