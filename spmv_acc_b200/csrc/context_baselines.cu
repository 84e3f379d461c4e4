// Context baseline for the bench report: cusparseSpMV on the same device buffers (not part of the product path).
// Mirrors the reference's comparator benchmark/benchmark_cusparse.hpp:27-67 (generic API, CSR, 32-bit indices,
// CUDA_R_64F), with the handle / descriptor / buffer creation hoisted out of the timed call.
#include <cuda_runtime.h>
#include <cusparse.h>
#include <cub/device/device_spmv.cuh>
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

// ---------------------------------------------------------------------------------------------------------------
// Gather bound (context, not product): streams colindex (and value) exactly like SpMV and gathers x[colindex[k]],
// but with NO row structure: every thread keeps private sums, nothing is reduced per row, one store per thread.
// Its time is a lower bound for any CSR kernel that gathers x element-wise for the same column stream: it pays the
// same HBM streams, the same L1TEX wavefronts (one per distinct 128-byte line per warp gather) and the same L2 / DRAM
// sector traffic for x, and nothing else.
//   mode bit 0: also stream `value` and multiply;  bits 1-3: flavour of the gather load
//   (0 ld.global.nc, 1 L1::no_allocate, 2 ld.global.cg (L2 only), 3 L1::evict_first, 4 L1::evict_last)
// `smem_bytes` of dynamic shared memory are requested per CTA only to shrink L1 (the unified array is 256 KB per SM):
// it measures how the gather rate depends on the L1 capacity left next to staged tiles.
template <int FL> __device__ __forceinline__ double gb_load(const double *p) {
  double g;
  if (FL == 1)
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(g) : "l"(p));
  else if (FL == 2)
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(g) : "l"(p));
  else if (FL == 3)
    asm volatile("ld.global.nc.L1::evict_first.f64 %0, [%1];" : "=d"(g) : "l"(p));
  else if (FL == 4)
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(g) : "l"(p));
  else
    g = __ldg(p);
  return g;
}

template <int U, bool VAL, int FL>
__global__ void __launch_bounds__(256) k_gather_bound(const int *__restrict__ col, const double *__restrict__ val,
                                                      const double *__restrict__ x, long long nnz,
                                                      double *__restrict__ out) {
  const long long chunk = 256LL * U;
  double acc = 0.0;
  for (long long base = blockIdx.x * chunk; base < nnz; base += (long long)gridDim.x * chunk) {
    int c[U];
    double v[U], g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long k = base + u * 256 + threadIdx.x;
      c[u] = k < nnz ? __ldcs(col + k) : -1;
      if (VAL)
        v[u] = k < nnz ? __ldcs(val + k) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      g[u] = c[u] >= 0 ? gb_load<FL>(x + c[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u)
      acc += VAL ? v[u] * g[u] : g[u];
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

typedef void (*GatherKernel)(const int *, const double *, const double *, long long, double *);

// out must hold grid*256 doubles; returns the grid size when out == nullptr
CTX_API int spmv_b200_ctx_gather_bound(long long nnz, const int *d_col, const double *d_val, const double *d_x,
                                       double *d_out, int mode, int ctas_per_sm, int smem_bytes, void *stream) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * (ctas_per_sm > 0 ? ctas_per_sm : 8);
  if (!d_out)
    return grid;
  static const GatherKernel table[2][5] = {
      {k_gather_bound<8, false, 0>, k_gather_bound<8, false, 1>, k_gather_bound<8, false, 2>,
       k_gather_bound<8, false, 3>, k_gather_bound<8, false, 4>},
      {k_gather_bound<8, true, 0>, k_gather_bound<8, true, 1>, k_gather_bound<8, true, 2>, k_gather_bound<8, true, 3>,
       k_gather_bound<8, true, 4>}};
  const int fl = (mode >> 1) & 7;
  if (fl > 4)
    return -2;
  GatherKernel k = table[mode & 1][fl];
  if (smem_bytes > 0) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
      return -3;
  }
  k<<<grid, 256, smem_bytes > 0 ? smem_bytes : 0, static_cast<cudaStream_t>(stream)>>>(d_col, d_val, d_x, nnz, d_out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------
// cub::DeviceSpmv::CsrMV (merge-based, y = A*x), the second comparator of the reference harness
// (benchmark/cub/spmv.cu:30-37: size query, cudaMalloc of the buffer, one call). Context only.
// ---------------------------------------------------------------------------------------------------------------
struct ctx_cub {
  void *buffer = nullptr;
  size_t bytes = 0;
};

CTX_API int spmv_b200_ctx_cub_destroy(ctx_cub *c) {
  if (!c) return 0;
  if (c->buffer) cudaFree(c->buffer);
  delete c;
  return 0;
}

CTX_API int spmv_b200_ctx_cub_create(ctx_cub **out, int m, int n, int nnz, const int *d_rowptr, const int *d_col,
                                     const double *d_val, const double *d_x, double *d_y) {
  ctx_cub *c = new ctx_cub();
  if (cub::DeviceSpmv::CsrMV<double>(nullptr, c->bytes, d_val, d_rowptr, d_col, d_x, d_y, m, n, nnz) != cudaSuccess ||
      cudaMalloc(&c->buffer, c->bytes ? c->bytes : 16) != cudaSuccess) {
    spmv_b200_ctx_cub_destroy(c);
    return 1;
  }
  *out = c;
  return 0;
}

CTX_API int spmv_b200_ctx_cub_spmv(ctx_cub *c, int m, int n, int nnz, const int *d_rowptr, const int *d_col,
                                   const double *d_val, const double *d_x, double *d_y, void *stream) {
  return (int)cub::DeviceSpmv::CsrMV<double>(c->buffer, c->bytes, d_val, d_rowptr, d_col, d_x, d_y, m, n, nnz,
                                             static_cast<cudaStream_t>(stream));
}
