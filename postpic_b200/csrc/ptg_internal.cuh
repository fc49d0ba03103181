// ptg_internal.cuh -- declarations shared by the translation units of libptg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include "ptg_stencil.cuh"

namespace ptg {

// ptg_deposit_global.cu
cudaError_t deposit_global(int ndim, int shape, const DepositParams &P, bool rows, int sm_count,
                           cudaStream_t st);

// ptg_util.cu
cudaError_t minmax_f64(int64_t n, const double *x, int64_t stride, double *out2, int sm_count,
                       cudaStream_t st);
cudaError_t scale_f64(double *grid, int64_t count, double factor, int sm_count, cudaStream_t st);
cudaError_t selftest_div6(int64_t n, const double *in, double *out, cudaStream_t st);

}  // namespace ptg
