// Context baseline for the bench report: cusparseSpMV on the same device buffers (not part of the product path).
// Mirrors the reference's comparator benchmark/benchmark_cusparse.hpp:27-67 (generic API, CSR, 32-bit indices,
// CUDA_R_64F), with the handle / descriptor / buffer creation hoisted out of the timed call.
#include <cuda_runtime.h>
#include <cusparse.h>
#include <stdint.h>

#define CTX_API extern "C" __attribute__((visibility("default")))

struct ctx_cusparse {
  cusparseHandle_t handle = nullptr;
  cusparseSpMatDescr_t mat = nullptr;
  cusparseDnVecDescr_t vx = nullptr, vy = nullptr;
  void *buffer = nullptr;
  size_t buffer_bytes = 0;
  cusparseSpMVAlg_t alg = CUSPARSE_SPMV_ALG_DEFAULT;
  int m = 0, n = 0;
  const double *x_bound = nullptr;
  double *y_bound = nullptr;
};

CTX_API int spmv_b200_ctx_cusparse_destroy(ctx_cusparse *c) {
  if (!c) return 0;
  if (c->vx) cusparseDestroyDnVec(c->vx);
  if (c->vy) cusparseDestroyDnVec(c->vy);
  if (c->mat) cusparseDestroySpMat(c->mat);
  if (c->handle) cusparseDestroy(c->handle);
  if (c->buffer) cudaFree(c->buffer);
  delete c;
  return 0;
}

// alg: 0 = CUSPARSE_SPMV_ALG_DEFAULT, 1 = CUSPARSE_SPMV_CSR_ALG1, 2 = CUSPARSE_SPMV_CSR_ALG2
CTX_API int spmv_b200_ctx_cusparse_create(ctx_cusparse **out, int m, int n, long long nnz, const int *d_rowptr,
                                          const int *d_col, const double *d_val, const double *d_x, double *d_y,
                                          int alg) {
  ctx_cusparse *c = new ctx_cusparse();
  c->m = m;
  c->n = n;
  c->alg = alg == 1 ? CUSPARSE_SPMV_CSR_ALG1 : (alg == 2 ? CUSPARSE_SPMV_CSR_ALG2 : CUSPARSE_SPMV_ALG_DEFAULT);
  const double one = 1.0;
  int st = 0;
  if ((st = cusparseCreate(&c->handle)) ||
      (st = cusparseCreateCsr(&c->mat, m, n, nnz, (void *)d_rowptr, (void *)d_col, (void *)d_val, CUSPARSE_INDEX_32I,
                              CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F)) ||
      (st = cusparseCreateDnVec(&c->vx, n, (void *)d_x, CUDA_R_64F)) ||
      (st = cusparseCreateDnVec(&c->vy, m, (void *)d_y, CUDA_R_64F)) ||
      (st = cusparseSpMV_bufferSize(c->handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, c->mat, c->vx, &one, c->vy,
                                    CUDA_R_64F, c->alg, &c->buffer_bytes))) {
    spmv_b200_ctx_cusparse_destroy(c);
    return 100 + st;
  }
  if (cudaMalloc(&c->buffer, c->buffer_bytes ? c->buffer_bytes : 16) != cudaSuccess) {
    spmv_b200_ctx_cusparse_destroy(c);
    return 2;
  }
#if CUSPARSE_VERSION >= 12100
  st = cusparseSpMV_preprocess(c->handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, c->mat, c->vx, &one, c->vy,
                               CUDA_R_64F, c->alg, c->buffer);
  if (st) {
    spmv_b200_ctx_cusparse_destroy(c);
    return 200 + st;
  }
#endif
  c->x_bound = d_x;
  c->y_bound = d_y;
  *out = c;
  return 0;
}

CTX_API int spmv_b200_ctx_cusparse_spmv(ctx_cusparse *c, double alpha, double beta, void *stream) {
  cusparseSetStream(c->handle, static_cast<cudaStream_t>(stream));
  return (int)cusparseSpMV(c->handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, c->mat, c->vx, &beta, c->vy,
                           CUDA_R_64F, c->alg, c->buffer);
}
