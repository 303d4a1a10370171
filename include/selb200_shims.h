/*
 * selb200_shims.h — link-level boundary: the reference's kernel launchers, re-implemented.
 *
 * Declarations identical to src/selection_kernels_wrapper.hpp:6-45 of the reference (C++ linkage,
 * mangled _Z17launch_kernel_smhPKhPKmPKdPK4int2idiiiiP6ResultPii and
 * _Z19launch_kernel_CBsmhPKhPKmPKdPK4int2idiiiiP6ResultPii).  libselb200.so exports both, so
 * src/selection_cuda.cpp and experiments/src/time_smh_cuda.cpp link against it unchanged
 * (INTEGRATION.md §2).  Caller owns all device buffers; out_count is zeroed; the work is queued
 * asynchronously on the default stream (src/selection_kernels.cu:119-177).
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

struct Result {
    int x, y;
    float sim;
};

void launch_kernel_smh(const uint8_t* main_sketches, const uint64_t* aux_sketches, const double* cards,
                       const int2* pairs, int total_pairs, double tau, int m_aux, int m_hll, int n_rows,
                       int n_bands, Result* out, int* out_count, int blockSize);

void launch_kernel_CBsmh(const uint8_t* main_sketches, const uint64_t* aux_sketches, const double* cards,
                         const int2* pairs, int total_pairs, double tau, int m_aux, int m_hll, int n_rows,
                         int n_bands, Result* out, int* out_count, int blockSize);
