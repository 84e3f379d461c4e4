//
// cuda-b200 kernel strategy: B200-native (sm_100a) fp64 CSR SpMV behind the spmv-acc strategy API.
//
// Drop this directory into the reference tree as src/acc/cuda-b200 and select it with
// -DKERNEL_STRATEGY=CUDA_B200 (see INTEGRATION.md for the six registration steps of README.md:73-115).
// The launcher follows the convention of the other strategies, e.g. adaptive_sparse_spmv
// (src/acc/hip-adaptive/adaptive.h:10-11): same argument list, void return, returns without synchronising.
//
#ifndef SPMV_ACC_CUDA_B200_SPMV_H
#define SPMV_ACC_CUDA_B200_SPMV_H

#include "api/handle.h"
#include "api/types.h"

/**
 * y = alpha * A * x + beta * y on the current device (null stream).
 * @param h_csr_desc accepted for signature compatibility and never dereferenced: it aliases device memory when the
 *        call comes through the deprecated sparse_spmv (src/acc/api/spmv_imp.cpp:14-17).
 * @throws std::runtime_error when the CUDA library reports an error (the benchmark harness skips a strategy that
 *         throws: benchmark/csr_spmv.hpp:52-62).
 */
void cuda_b200_sparse_spmv(int trans, const double alpha, const double beta, const csr_desc<int, double> h_csr_desc,
                           const csr_desc<int, double> d_csr_desc, const double *x, double *y);

/**
 * Same computation with the analyse / kernel / destroy phases timed separately (microseconds, CUDA events) into the
 * handle, like csr_adaptive_plus_sparse_spmv<true> (src/acc/hip-csr-adaptive-plus/csr_adaptive_plus_spmv.cpp:92-129).
 */
void cuda_b200_sparse_spmv_profile(SpMVAccHanele *handle, int trans, const double alpha, const double beta,
                                   const csr_desc<int, double> h_csr_desc, const csr_desc<int, double> d_csr_desc,
                                   const double *x, double *y);

/** Drops the cached analysis of every matrix. The cache re-validates a fingerprint of the row pointers on every call
 *  (include/spmv_b200.h); this is for callers that rewrite row pointers in place between calls. */
void cuda_b200_invalidate_plans();

#endif // SPMV_ACC_CUDA_B200_SPMV_H
