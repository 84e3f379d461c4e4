// Streaming SpMV kernels for sm_100a (parts 2-4 of the hot path).
//
// Every CTA streams one nnz-balanced row block ("tile", see analysis.cu): the contiguous `value` and `colindex`
// ranges of the tile are brought into shared memory with two TMA bulk copies (cp.async.bulk, completion on an
// mbarrier, L2 evict-first because they are read exactly once), the row-pointer segment is staged next to them,
// x is gathered through the read-only path, and the alpha/beta epilogue is fused into the y store.
//
//   k_spmv_rows<TMA, false>  SHORT tiles  : one thread per row (rows of <= short_max nnz). Lanes of a warp walk
//                                           consecutive rows, so for banded/stencil matrices the x gathers of a
//                                           warp fall into 2-3 cache lines.
//   k_spmv_rows<TMA, true>   MEDIUM tiles : 2^k lanes per row, warp-shuffle segmented sum.
//   k_spmv_mixed<TMA>        MIXED tiles  : rows of any length plus fragments of rows that are split across tiles:
//                                           products are formed in place in shared memory (one nnz per thread and
//                                           step), short rows are summed by one thread, long rows by one warp, row
//                                           fragments by the whole CTA into a per-tile partial.
//   k_fixup                  second pass  : one warp per split row adds the row's partials in a fixed order and
//                                           applies the epilogue. No atomics anywhere: results are bitwise
//                                           reproducible from run to run.
//
// Reference analogues (behaviour, not code): thread-row (src/acc/hip-thread-row/thread_row.inl:18-98), vector-row
// (src/acc/hip-vector-row/vector_row_native.hpp:83-121), line-enhance (src/acc/hip-line-enhance/
// line_enhance_spmv_imp.inl:11-95), flat (src/acc/hip-flat/flat_imp_one_pass.hpp:15-77, which uses atomicAdd and
// assumes beta == 1), merge-path reduction + update (benchmark/merge-path/merge_path_reduction.h:80-136,
// merge_path_update.h:8-64). The epilogue y = alpha*sum + beta*y follows cli/verification.cpp:64.
#include "internal.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}

__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                             unsigned long long *bar, unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, "
               "[%3], %4;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

__device__ __forceinline__ double ld_stream_f64(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ int ld_stream_s32(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// tile streaming: global -> shared. Element i of the tile lives at smem index (i - a0), a0 = elem_begin & ~3.
// ---------------------------------------------------------------------------------------------------------------
template <bool TMA>
__device__ __forceinline__ void tile_issue_loads(const SpmvArgs &a, int a0, int e0, int e1, double *sval, int *scol,
                                                 unsigned long long *bar, int tid) {
  const int span = e1 - a0;
  if (TMA) {
    if (tid == 0)
      mbar_init(bar, 1);
    __syncthreads();
    // both copies need 16-byte sizes: round the element count up to a multiple of 4, but never read past the
    // end of the arrays (only the last tile can be short; its <= 3 tail elements are loaded by threads).
    const long long avail = a.nnz - (long long)a0;
    int cnt = (span + 3) & ~3;
    if ((long long)cnt > avail)
      cnt = (int)(avail & ~3LL);
    if (tid == 0) {
      if (cnt > 0) {
        const unsigned long long pol = policy_evict_first();
        mbar_arrive_expect_tx(bar, (uint32_t)cnt * 12u);
        tma_bulk_g2s(sval, a.val + a0, (uint32_t)cnt * 8u, bar, pol);
        tma_bulk_g2s(scol, a.col + a0, (uint32_t)cnt * 4u, bar, pol);
      } else {
        mbar_arrive(bar);
      }
    }
    for (int i = cnt + tid; i < span; i += kThreads) {
      sval[i] = ld_stream_f64(a.val + a0 + i);
      scol[i] = ld_stream_s32(a.col + a0 + i);
    }
  } else {
    for (int i = (e0 - a0) + tid; i < span; i += kThreads) {
      sval[i] = ld_stream_f64(a.val + a0 + i);
      scol[i] = ld_stream_s32(a.col + a0 + i);
    }
  }
}

// sum_{k = k0, k0+stride, ... < e} sval[k] * x[scol[k]], accumulated left to right, gathers issued four at a time
__device__ __forceinline__ double row_dot(const double *__restrict__ sval, const int *__restrict__ scol,
                                          const double *__restrict__ x, int k, const int e, const int stride) {
  double sum = 0.0;
  for (; k + 3 * stride < e; k += 4 * stride) {
    const int c0 = scol[k], c1 = scol[k + stride], c2 = scol[k + 2 * stride], c3 = scol[k + 3 * stride];
    const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
    const double v0 = sval[k], v1 = sval[k + stride], v2 = sval[k + 2 * stride], v3 = sval[k + 3 * stride];
    sum = fma(v0, x0, sum);
    sum = fma(v1, x1, sum);
    sum = fma(v2, x2, sum);
    sum = fma(v3, x3, sum);
  }
  for (; k < e; k += stride)
    sum = fma(sval[k], __ldg(x + scol[k]), sum);
  return sum;
}

__device__ __forceinline__ void store_y(const SpmvArgs &a, int row, double sum) {
  // cli/verification.cpp:64  y[i] = alpha * y0 + beta * y[i]
  const double yv = a.read_y ? a.y[row] : 0.0;
  a.y[row] = a.alpha * sum + a.beta * yv;
}

// ---------------------------------------------------------------------------------------------------------------
// SHORT (VEC = false) and MEDIUM (VEC = true) tiles: every owned row lies completely inside the tile
// ---------------------------------------------------------------------------------------------------------------
template <bool TMA, bool VEC>
__global__ void __launch_bounds__(kThreads) k_spmv_rows(const SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  double *sval = reinterpret_cast<double *>(smem_raw);
  int *scol = reinterpret_cast<int *>(sval + a.cap);
  int *srow = scol + a.cap;

  const int tid = threadIdx.x;
  const int t = a.list ? a.list[blockIdx.x] : (int)blockIdx.x;
  const int r0 = a.tile_row[t], r1 = a.tile_row[t + 1];
  const int e0 = a.tile_elem[t], e1 = a.tile_elem[t + 1];
  const int a0 = e0 & ~3;
  const int nrows = r1 - r0;

  tile_issue_loads<TMA>(a, a0, e0, e1, sval, scol, &bar, tid);

  int lv = 0; // log2(lanes per row)
  if (VEC) {
    const int avg = (e1 - e0) / (nrows > 0 ? nrows : 1);
    const int want = (avg + a.vec_div - 1) / a.vec_div;
    while ((1 << lv) < want && lv < 5)
      ++lv;
  }
  const int V = 1 << lv;
  const int G = kThreads >> lv; // rows per pass
  const int g = tid >> lv;
  const int l = tid & (V - 1);

  for (int cb = 0;; cb += kRowChunk) {
    int nr = nrows - cb;
    if (nr > kRowChunk)
      nr = kRowChunk;
    if (cb > 0)
      __syncthreads(); // srow is about to be overwritten
    for (int i = tid; i <= nr; i += kThreads)
      srow[i] = __ldg(a.rowptr + r0 + cb + i);
    __syncthreads();
    if (TMA && cb == 0)
      mbar_wait(&bar, 0);

    for (int rb = 0; rb < nr; rb += G) {
      const int r = rb + g;
      const bool act = r < nr;
      double sum = 0.0;
      if (act) {
        const int s = srow[r] - a0, e = srow[r + 1] - a0;
        sum = row_dot(sval, scol, a.x, s + l, e, V);
      }
      if (VEC) {
        for (int off = V >> 1; off > 0; off >>= 1)
          sum += __shfl_down_sync(0xffffffffu, sum, off, V);
      }
      if (act && l == 0)
        store_y(a, r0 + cb + r, sum);
    }
    if (cb + kRowChunk >= nrows)
      break;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// MIXED tiles
// ---------------------------------------------------------------------------------------------------------------
// deterministic CTA-wide sum of sval[lo, hi): strided per-thread partials, xor-shuffle tree, warp partials in order
__device__ __forceinline__ double block_sum(const double *__restrict__ sval, int lo, int hi, double *swarp, int tid) {
  double s = 0.0;
  for (int k = lo + tid; k < hi; k += kThreads)
    s += sval[k];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((tid & 31) == 0)
    swarp[tid >> 5] = s;
  __syncthreads();
  double total = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w)
      total += swarp[w];
  }
  __syncthreads();
  return total; // valid in thread 0
}

template <bool TMA>
__global__ void __launch_bounds__(kThreads) k_spmv_mixed(const SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ double swarp[kThreads / 32];
  __shared__ int nlong;
  double *sval = reinterpret_cast<double *>(smem_raw);
  int *scol = reinterpret_cast<int *>(sval + a.cap);
  int *srow = scol + a.cap;
  int *slong = srow + (kRowChunk + 1); // rows longer than kSerialMax in the current chunk

  const int tid = threadIdx.x;
  const int t = a.list ? a.list[blockIdx.x] : (int)blockIdx.x;
  const int r0 = a.tile_row[t], r1 = a.tile_row[t + 1];
  const int e0 = a.tile_elem[t], e1 = a.tile_elem[t + 1];
  const int a0 = e0 & ~3;
  const bool split_begin = a.tile_split[t] != 0;
  const bool split_end = a.tile_split[t + 1] != 0;
  // the last owned row continues in the next tile: its part in this tile is the tail fragment
  const bool has_tail = split_end && (r1 > r0);
  const int nrows = (r1 - r0) - (has_tail ? 1 : 0);

  tile_issue_loads<TMA>(a, a0, e0, e1, sval, scol, &bar, tid);
  const int first_row_start = __ldg(a.rowptr + r0);                   // r0 <= m
  const int tail_start = has_tail ? __ldg(a.rowptr + r1 - 1) : e1;    // first element of the tail fragment
  __syncthreads();
  if (TMA)
    mbar_wait(&bar, 0);

  // products in place: one nnz per thread and step, coalesced smem access
  for (int i = (e0 - a0) + tid; i < e1 - a0; i += kThreads)
    sval[i] *= __ldg(a.x + scol[i]);
  __syncthreads();

  if (split_begin) { // leading elements belong to a row that started in an earlier tile
    const int hend = first_row_start < e1 ? first_row_start : e1;
    const double s = block_sum(sval, e0 - a0, hend - a0, swarp, tid);
    if (tid == 0)
      a.partials[2 * (size_t)t] = s;
  }
  if (has_tail) {
    const double s = block_sum(sval, tail_start - a0, e1 - a0, swarp, tid);
    if (tid == 0)
      a.partials[2 * (size_t)t + 1] = s;
  }

  const int warp = tid >> 5, lane = tid & 31;
  for (int cb = 0; cb < nrows; cb += kRowChunk) {
    int nr = nrows - cb;
    if (nr > kRowChunk)
      nr = kRowChunk;
    __syncthreads(); // srow / slong / nlong reuse
    for (int i = tid; i <= nr; i += kThreads)
      srow[i] = __ldg(a.rowptr + r0 + cb + i);
    if (tid == 0)
      nlong = 0;
    __syncthreads();
    // pass 1: one thread per row; long rows are queued
    for (int r = tid; r < nr; r += kThreads) {
      const int s = srow[r] - a0, e = srow[r + 1] - a0;
      if (e - s <= kSerialMax) {
        double sum = 0.0;
        for (int k = s; k < e; ++k)
          sum += sval[k];
        store_y(a, r0 + cb + r, sum);
      } else {
        slong[atomicAdd(&nlong, 1)] = r;
      }
    }
    __syncthreads();
    // pass 2: one warp per queued row (the value of a row does not depend on its queue position)
    const int nl = nlong;
    for (int i = warp; i < nl; i += kThreads / 32) {
      const int r = slong[i];
      const int s = srow[r] - a0, e = srow[r + 1] - a0;
      double sum = 0.0;
      for (int k = s + lane; k < e; k += 32)
        sum += sval[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
      if (lane == 0)
        store_y(a, r0 + cb + r, sum);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// second pass: rows split across tiles
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fixup(const FixupArgs f) {
  const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= f.nsplit)
    return;
  const int row = f.split_row[w];
  const int t0 = f.split_t0[w], t1 = f.split_t1[w];
  const int nfrag = t1 - t0 + 1; // tail of t0, then the heads of t0+1 .. t1
  double sum = 0.0;
  for (int j = lane; j < nfrag; j += 32)
    sum += (j == 0) ? f.partials[2 * (size_t)t0 + 1] : f.partials[2 * (size_t)(t0 + j)];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) {
    const double yv = f.read_y ? f.y[row] : 0.0;
    f.y[row] = f.alpha * sum + f.beta * yv;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static size_t smem_for(const spmv_b200_plan *p, bool mixed) {
  size_t b = (size_t)p->cap * 12 + sizeof(int) * (kRowChunk + 1);
  if (mixed)
    b += sizeof(int) * ((size_t)p->cap / (kSerialMax + 1) + 8);
  return (b + 15) & ~(size_t)15;
}

template <typename K> static int set_smem(K kernel, size_t bytes) {
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SPMV_B200_OK;
}

int kernels_configure(spmv_b200_plan *p) {
  p->cap = p->T + p->medium_max + 8;
  p->smem_bytes = smem_for(p, true);
  int dev = 0, max_optin = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (p->smem_bytes > (size_t)max_optin) {
    set_error("tile_nnz too large for the shared memory of this device");
    return SPMV_B200_ERR_ARG;
  }
  p->device = dev;
  int rc;
  const size_t sr = smem_for(p, false), sm = smem_for(p, true);
  if ((rc = set_smem(k_spmv_rows<true, false>, sr)) || (rc = set_smem(k_spmv_rows<true, true>, sr)) ||
      (rc = set_smem(k_spmv_rows<false, false>, sr)) || (rc = set_smem(k_spmv_rows<false, true>, sr)) ||
      (rc = set_smem(k_spmv_mixed<true>, sm)) || (rc = set_smem(k_spmv_mixed<false>, sm)))
    return rc;
  return SPMV_B200_OK;
}

int kernels_launch(const spmv_b200_plan *p, double alpha, double beta, const double *x, double *y,
                   cudaStream_t stream) {
  if (p->m == 0)
    return SPMV_B200_OK;
  SpmvArgs a;
  a.rowptr = p->rowptr;
  a.col = p->col;
  a.val = p->val;
  a.x = x;
  a.y = y;
  a.alpha = alpha;
  a.beta = beta;
  a.tile_row = p->tile_row;
  a.tile_elem = p->tile_elem;
  a.tile_split = p->tile_split;
  a.list = nullptr;
  a.partials = p->partials;
  a.nnz = p->elem_end; // absolute index one past the last element (rowptr may be a view: rowptr[0] != 0)
  a.cap = p->cap;
  a.vec_div = p->vec_div;
  a.read_y = (beta == 0.0 && (p->flags & SPMV_B200_FLAG_BETA0_SKIP_Y)) ? 0 : 1;

  const size_t sr = smem_for(p, false), sm = smem_for(p, true);
  const bool tma = p->uses_tma;
  if (p->count[SPMV_B200_KIND_SHORT] > 0) {
    a.list = p->list[SPMV_B200_KIND_SHORT];
    if (tma)
      k_spmv_rows<true, false><<<p->count[SPMV_B200_KIND_SHORT], kThreads, sr, stream>>>(a);
    else
      k_spmv_rows<false, false><<<p->count[SPMV_B200_KIND_SHORT], kThreads, sr, stream>>>(a);
  }
  if (p->count[SPMV_B200_KIND_MEDIUM] > 0) {
    a.list = p->list[SPMV_B200_KIND_MEDIUM];
    if (tma)
      k_spmv_rows<true, true><<<p->count[SPMV_B200_KIND_MEDIUM], kThreads, sr, stream>>>(a);
    else
      k_spmv_rows<false, true><<<p->count[SPMV_B200_KIND_MEDIUM], kThreads, sr, stream>>>(a);
  }
  if (p->count[SPMV_B200_KIND_MIXED] > 0) {
    a.list = p->list[SPMV_B200_KIND_MIXED];
    if (tma)
      k_spmv_mixed<true><<<p->count[SPMV_B200_KIND_MIXED], kThreads, sm, stream>>>(a);
    else
      k_spmv_mixed<false><<<p->count[SPMV_B200_KIND_MIXED], kThreads, sm, stream>>>(a);
  }
  if (p->nsplit > 0) {
    FixupArgs f;
    f.split_row = p->split_rows;
    f.split_t0 = p->split_rows + p->nsplit;
    f.split_t1 = p->split_rows + 2 * (size_t)p->nsplit;
    f.partials = p->partials;
    f.y = y;
    f.alpha = alpha;
    f.beta = beta;
    f.nsplit = p->nsplit;
    f.read_y = a.read_y;
    const int warps_per_cta = kThreads / 32;
    k_fixup<<<(p->nsplit + warps_per_cta - 1) / warps_per_cta, kThreads, 0, stream>>>(f);
  }
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

} // namespace b200
