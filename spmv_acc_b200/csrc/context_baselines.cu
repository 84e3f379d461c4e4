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
// Gather bound with SM affinity (context, not product): does it help to let one group of SMs gather only from one half
// of x? Data read by every SM is cached in both halves of the L2 (one per die), so a table larger than ~63 MB misses;
// if the SMs of one die read only one half of the table, an L2 half holds its home lines plus the far copies of that
// half only. Two column streams (indices into the lower / upper half of x); every CTA reads its SM id, picks its
// group by `map`, and takes chunks of its group's stream from a ticket counter until the stream is empty.
//   map 0: smid & 1          map 1: (smid >> 1) & 1 (TPC parity)     map 2: smid >= #SMs / 2
//   map 3: ((smid >> 1) % 8) < 4      map 4: (smid / 18) & 1 ... guesses of how SM ids map to the two dies
//   map 9: blockIdx.x & 1 (control: groups unrelated to the SM)
__device__ __forceinline__ unsigned smid() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}

__global__ void __launch_bounds__(256) k_gather_affine(long long nnz_half, const int *__restrict__ col_lo,
                                                       const int *__restrict__ col_hi, const double *__restrict__ x,
                                                       double *__restrict__ out, unsigned long long *tickets, int map,
                                                       int nsm) {
  __shared__ long long s_base;
  const unsigned sm = smid();
  int group;
  switch (map) {
  case 0: group = sm & 1; break;
  case 1: group = (sm >> 1) & 1; break;
  case 2: group = sm >= (unsigned)nsm / 2; break;
  case 3: group = ((sm >> 1) % 8) < 4; break;
  case 4: group = (sm / 18) & 1; break;
  default: group = blockIdx.x & 1; break;
  }
  const int *__restrict__ col = group ? col_hi : col_lo;
  constexpr long long kChunk = 256 * 8;
  double sum = 0.0;
  for (;;) {
    if (threadIdx.x == 0)
      s_base = (long long)atomicAdd(tickets + group, (unsigned long long)kChunk);
    __syncthreads();
    const long long base = s_base;
    __syncthreads();
    if (base >= nnz_half)
      break;
    int c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long k = base + threadIdx.x + 256 * j;
      c[j] = k < nnz_half ? __ldg(col + k) : -1;
    }
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = c[j] >= 0 ? __ldg(x + c[j]) : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sum += v[j];
  }
  out[(size_t)blockIdx.x * 256 + threadIdx.x] = sum;
}

// d_tickets: two zeroed 64-bit counters; out: grid * 256 doubles; returns the grid size when d_out == nullptr
CTX_API int spmv_b200_ctx_gather_affine(long long nnz_half, const int *d_col_lo, const int *d_col_hi, const double *d_x,
                                        double *d_out, unsigned long long *d_tickets, int map, int ctas_per_sm,
                                        void *stream) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * (ctas_per_sm > 0 ? ctas_per_sm : 8);
  if (!d_out)
    return grid;
  if (cudaMemsetAsync(d_tickets, 0, 2 * sizeof(unsigned long long), static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return -2;
  k_gather_affine<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(nnz_half, d_col_lo, d_col_hi, d_x, d_out,
                                                                       d_tickets, map, sms);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------
// Gather bound through the TMA unit (context, not product): the same column stream, but x is fetched with
// cp.async.bulk.tensor ... tile::gather4 instead of LSU loads. x is described as a 2-D tensor [n/2][2] of fp64 (16-byte
// rows, the smallest box TMA accepts); one instruction names four rows (col >> 1) and lands their 4 x 16 bytes in
// shared memory, from where the wanted half (col & 1) is read. The LSU gather rate is capped by L1TEX at one 128-byte
// line per clock and SM (profiles/r1_gather_bound.jsonl: 257 G gathers/s); this measures whether the TMA path, which
// does not go through L1TEX, gets past that. Two stages of 1024 elements per CTA; every thread issues one gather4
// per stage. Returns the kernel's result in out[] (one partial sum per thread) so that it can be checked against the
// LSU kernel.
#include <cuda.h>

template <int STAGES>
__global__ void __launch_bounds__(256) k_gather4_bound(const __grid_constant__ CUtensorMap tmap,
                                                       const int *__restrict__ col, long long nnz,
                                                       double *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char g4_smem[]; // STAGES x 256 slots of 128 bytes (64 used)
  __shared__ __align__(8) unsigned long long bar[STAGES];
  const int tid = threadIdx.x;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bar[0]);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long chunk = 1024, stride = (long long)gridDim.x * chunk;
  double acc = 0.0;
  int4 cpend[STAGES];
  long long base = (long long)blockIdx.x * chunk;
  auto issue = [&](int s, long long b) {
    int4 c = make_int4(0, 0, 0, 0);
    const long long k = b + 4 * tid;
    if (k + 3 < nnz)
      c = __ldcs(reinterpret_cast<const int4 *>(col + k));
    else
      for (int j = 0; j < 4; ++j)
        if (k + j < nnz)
          (&c.x)[j] = __ldcs(col + k + j);
    cpend[s] = c;
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(256u * 64u)
                   : "memory");
    __syncwarp();
    const unsigned dst = (unsigned)__cvta_generic_to_shared(g4_smem + ((size_t)s * 256 + tid) * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
                 "%3, %4, %5, %6}], [%7];" ::"r"(dst),
                 "l"(&tmap), "r"(0), "r"(c.x >> 1), "r"(c.y >> 1), "r"(c.z >> 1), "r"(c.w >> 1), "r"(bar0 + 8u * s)
                 : "memory");
  };
  // (the lane-0 arrive of warp 0 and the copies of the other warps may come in any order: the only pending arrival
  // is that one arrive, so the phase cannot complete before the expected bytes are known)
  int filled = 0;
  for (; filled < STAGES && base + (long long)filled * stride < nnz; ++filled)
    issue(filled, base + (long long)filled * stride);
  for (int it = 0; base < nnz; ++it, base += stride) {
    const int s = it % STAGES;
    const unsigned parity = (unsigned)((it / STAGES) & 1);
    unsigned done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(bar0 + 8u * s), "r"(parity)
                   : "memory");
    const int4 c = cpend[s];
    const double *slot = reinterpret_cast<const double *>(g4_smem + ((size_t)s * 256 + tid) * 128);
    const long long k = base + 4 * tid;
    double v = 0.0;
    if (k < nnz) v += slot[0 + (c.x & 1)];
    if (k + 1 < nnz) v += slot[2 + (c.y & 1)];
    if (k + 2 < nnz) v += slot[4 + (c.z & 1)];
    if (k + 3 < nnz) v += slot[6 + (c.w & 1)];
    acc += v;
    __syncthreads(); // every thread has read its slot of this stage before it is refilled
    const long long nb = base + (long long)STAGES * stride;
    if (nb < nnz)
      issue(s, nb);
  }
  out[blockIdx.x * 256 + tid] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// out must hold grid*256 doubles; returns the grid size when d_out == nullptr; box_rows = second box dimension of the
// tensor map (1 or 4: the two readings of the gather4 box rule; the caller checks which one the driver accepts and
// which one gives the right sums)
CTX_API int spmv_b200_ctx_gather4_bound(long long nnz, const int *d_col, const double *d_x, long long n, double *d_out,
                                        int ctas_per_sm, int box_rows, void *stream) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * (ctas_per_sm > 0 ? ctas_per_sm : 3);
  if (!d_out)
    return grid;
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return -10;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {2, (cuuint64_t)(n / 2)};
  const cuuint64_t strides[1] = {16};
  const cuuint32_t box[2] = {2, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(d_x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return -100 - (int)r;
  constexpr int STAGES = 2;
  const size_t smem = (size_t)STAGES * 256 * 128;
  if (cudaFuncSetAttribute(k_gather4_bound<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return -3;
  k_gather4_bound<STAGES><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(tmap, d_col, nnz, d_out);
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
