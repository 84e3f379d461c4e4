// One-time row analysis of a CSR matrix (part 1 of the hot path): bins rows by nnz, cuts the nnz range into
// nnz-balanced row blocks ("tiles"), decides which long rows are split across tiles, and tags every tile with the
// kernel kind that will stream it. Every output is an integer array and is reproduced bit-for-bit by the CPU
// restatement in oracle/analysis_port.c.
//
// Reference analogues (what this replaces, not how): the serial host analysis of csr-adaptive-plus
// (src/acc/hip-csr-adaptive-plus/csr_adaptive_plus_analyze.cpp:12-98), the break-point pre-pass of flat
// (src/acc/hip-flat/flat_imp.inl:107-152), the merge-path partition (benchmark/merge-path/merge_path_partition.h:7-17)
// and the 4-sample selector of adaptive (src/acc/hip-adaptive/adaptive.cpp:16-67).
//
// Specification (T = tile_nnz, L = medium_max, S = short_max, base = rowptr[0], end = rowptr[m]):
//   Tiles balance nnz AND rows (merge-path coordinate): the items of row r are its non-zeros followed by one
//   end-of-row marker, so the first item of row r sits at position f(r) = (rowptr[r] - base) + r of the merged list and
//   element x of row r at position (x - base) + r. Tile t covers positions [t*T, (t+1)*T): at most T non-zeros and at
//   most T rows, however many rows are empty.
//   ntiles      = max(1, ceil((end - base + m) / T)) if m > 0 else 0
//   tile_row[t] = 0 if t == 0; m if t == ntiles; else min{ r in [0,m] : f(r) >= t*T }      (rows are owned by the tile
//                 that holds their first item)
//   cut(t)      = base + t*T - (tile_row[t] - 1)   = number of non-zeros in front of position t*T, as an index
//   a row "straddles" boundary t (0 < t < ntiles) when cut(t) < rowptr[tile_row[t]]; it is row tile_row[t]-1.
//   tile_split[t] = 1 iff the straddling row is longer than L (its partial sums are combined by the fix-up kernel);
//                   a straddling row of length <= L is streamed entirely by the tile that owns it.
//   tile_elem[t]  = base if t == 0; end if t == ntiles; cut(t) if tile_split[t]; else rowptr[tile_row[t]]
//   tile_part[t]  = max{ r in [0,m] : rowptr[r] <= min(base + t*T, end) }   (the reference's merge-path partition S[t],
//                   benchmark/merge-path/merge_path_partition.h:7-17; exported for the cross-check only)
//   row r is owned by tile min(floor(f(r)/T), ntiles-1); tile_maxlen[t] = max nnz of the rows it owns.
//   tile_kind[t]  = MIXED if tile_split[t] or tile_split[t+1] or tile_maxlen[t] > L or the tile owns more than 512 rows
//                   or (tile_maxlen[t] > S and tile_maxlen[t] * rows_t > 2 * nnz_t);
//                   SHORT if tile_maxlen[t] <= S; else MEDIUM.
//   bin(r) = SHORT if nnz_r <= S; MEDIUM if nnz_r <= L; LONG if nnz_r <= T; else VERYLONG.
//   split row r (first split boundary t, i.e. tile_row[t-1] <= r = tile_row[t]-1): fragments live in tiles
//   t-1 .. t1 with t1 = min(floor(((rowptr[r+1] - 1 - base) + r) / T), ntiles-1).
#include <algorithm>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "internal.cuh"

namespace b200 {

__device__ __forceinline__ int lower_bound_rowptr(const int *__restrict__ rowptr, int m, long long target) {
  // first index in [0, m] with rowptr[idx] >= target (rowptr[m] >= target is guaranteed by the caller)
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if ((long long)__ldg(rowptr + mid) < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// first index in [0, m] with f(idx) = rowptr[idx] - base + idx >= target (f is strictly increasing, f(m) >= target)
__device__ __forceinline__ int lower_bound_merge(const int *__restrict__ rowptr, int m, long long base,
                                                 long long target) {
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if ((long long)__ldg(rowptr + mid) - base + mid < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int last_le_rowptr(const int *__restrict__ rowptr, int m, long long target) {
  // largest index in [0, m] with rowptr[idx] <= target (rowptr[0] <= target is guaranteed by the caller)
  int lo = 0, hi = m + 1;
  while (lo < hi - 1) {
    const int mid = lo + ((hi - lo) >> 1);
    if ((long long)__ldg(rowptr + mid) <= target)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
    k_tile_bounds(const int *__restrict__ rowptr, int m, int ntiles, int T, int medium_max, int *__restrict__ tile_row,
                  int *__restrict__ tile_elem, unsigned char *__restrict__ tile_split, int *__restrict__ tile_part,
                  int *__restrict__ tile_aux) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles)
    return;
  const long long base = __ldg(rowptr);
  const long long end = __ldg(rowptr + m);
  const long long pos = (long long)t * T; // position in the merged (non-zeros + row ends) list
  long long target = base + pos;          // nnz-only target of the reference's partition (cross-check array)
  if (target > end)
    target = end;
  int row, elem, aux = 0;
  unsigned char split = 0;
  if (t == 0) {
    row = 0;
    elem = (int)base;
  } else if (t == ntiles) {
    row = m;
    elem = (int)end;
  } else {
    row = lower_bound_merge(rowptr, m, base, pos); // >= 1 because f(0) = 0 < pos
    const int row_start = __ldg(rowptr + row);
    const long long cut = base + pos - (row - 1); // non-zeros in front of the boundary (index into value/colindex)
    elem = row_start;
    if (cut < (long long)row_start) { // row-1 straddles this boundary
      const int len = row_start - __ldg(rowptr + row - 1);
      if (len > medium_max) {
        split = 1;
        elem = (int)cut;
        aux = row_start; // end of the split row
      }
    }
  }
  tile_row[t] = row;
  tile_elem[t] = elem;
  tile_split[t] = split;
  tile_aux[t] = aux;
  tile_part[t] = last_le_rowptr(rowptr, m, target);
}

__global__ void __launch_bounds__(256)
    k_row_stats(const int *__restrict__ rowptr, int m, int ntiles, int T, int short_max, int medium_max,
                int *__restrict__ tile_maxlen, unsigned long long *__restrict__ hist /* [8] rows[4] | nnz[4] */) {
  __shared__ unsigned long long sh[8];
  if (threadIdx.x < 8)
    sh[threadIdx.x] = 0ull;
  __syncthreads();
  const long long base = __ldg(rowptr);
  // per-thread counters, then one warp reduction and one shared atomic per warp and counter (integer sums are
  // order independent, so the result does not depend on the schedule)
  unsigned long long cnt[4] = {0, 0, 0, 0}, sum[4] = {0, 0, 0, 0};
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (long long)gridDim.x * blockDim.x) {
    const int s = __ldg(rowptr + r);
    const int len = __ldg(rowptr + r + 1) - s;
    const int bin = len <= short_max ? 0 : (len <= medium_max ? 1 : (len <= T ? 2 : 3));
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      cnt[b] += (bin == b) ? 1ull : 0ull;
      sum[b] += (bin == b) ? (unsigned long long)len : 0ull;
    }
    long long t = ((long long)s - base + r) / T;
    if (t > ntiles - 1)
      t = ntiles - 1;
    // most rows of a tile are no longer than the value already stored: test before paying for the atomic
    if (len > 0 && len > tile_maxlen[t])
      atomicMax(tile_maxlen + t, len);
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      cnt[b] += __shfl_xor_sync(0xffffffffu, cnt[b], off);
      sum[b] += __shfl_xor_sync(0xffffffffu, sum[b], off);
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (cnt[b])
        atomicAdd(&sh[b], cnt[b]);
      if (sum[b])
        atomicAdd(&sh[4 + b], sum[b]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 8 && sh[threadIdx.x] != 0ull)
    atomicAdd(hist + threadIdx.x, sh[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
    k_tile_kind(int ntiles, int short_max, int medium_max, const int *__restrict__ tile_maxlen,
                const unsigned char *__restrict__ tile_split, const int *__restrict__ tile_row,
                const int *__restrict__ tile_elem, unsigned char *__restrict__ tile_kind) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles)
    return;
  const int ml = tile_maxlen[t];
  unsigned char k;
  // tiles with more than two passes' worth of rows (mostly empty / one-element rows) also go to the MIXED kernel,
  // whose cost grows with the non-zeros and not with the number of rows
  // ... and so do tiles whose longest row is more than twice the tile's average row (a lane group that owns
  // such a row would keep the whole CTA waiting)
  const long long rows = tile_row[t + 1] - tile_row[t];
  const long long elems = tile_elem[t + 1] - tile_elem[t];
  const bool skewed = ml > short_max && (long long)ml * rows > 2 * elems;
  if (tile_split[t] || tile_split[t + 1] || ml > medium_max || rows > kSparseTileRows || skewed)
    k = SPMV_B200_KIND_MIXED;
  else if (ml <= short_max)
    k = SPMV_B200_KIND_SHORT;
  else
    k = SPMV_B200_KIND_MEDIUM;
  tile_kind[t] = k;
}

__global__ void __launch_bounds__(256)
    k_tile_desc(const int *__restrict__ rowptr, int ntiles, const int *__restrict__ tile_row,
                const int *__restrict__ tile_elem, const unsigned char *__restrict__ tile_split,
                TileDesc *__restrict__ desc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles)
    return;
  TileDesc d;
  d.r0 = tile_row[t];
  d.r1 = tile_row[t + 1];
  d.e0 = tile_elem[t];
  d.e1 = tile_elem[t + 1];
  d.flags = (tile_split[t] ? 1 : 0) | (tile_split[t + 1] ? 2 : 0);
  const int first_row_start = __ldg(rowptr + d.r0); // r0 <= m
  d.head_end = first_row_start < d.e1 ? first_row_start : d.e1;
  d.tail_start = ((d.flags & 2) && d.r1 > d.r0) ? __ldg(rowptr + d.r1 - 1) : d.e1;
  d.tile = t;
  desc[t] = d;
}

// ---- direct form (one warp per row block, no shared memory): row-start bit flags, non-empty row list ----
// bits[k >> 5] bit (k & 31) is set iff element k (absolute index into value / colindex) is the first element of a row;
// nzflag[r] = 1 iff row r is non-empty (nzflag[m] = 0 so that the exclusive scan yields the total at index m).
__global__ void __launch_bounds__(256) k_row_start_bits(const int *__restrict__ rowptr, int m,
                                                         unsigned int *__restrict__ bits, int *__restrict__ nzflag) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= m; r += (long long)gridDim.x * blockDim.x) {
    int f = 0;
    if (r < m) {
      const int s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
      if (e > s) {
        f = 1;
        atomicOr(bits + (s >> 5), 1u << (s & 31)); // integer OR: order independent
      }
    }
    nzflag[r] = f;
  }
}

__global__ void __launch_bounds__(256) k_nz_rows(const int *__restrict__ nzflag, const int *__restrict__ nzprefix,
                                                  int m, int *__restrict__ nz_rows) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (long long)gridDim.x * blockDim.x)
    if (nzflag[r])
      nz_rows[nzprefix[r]] = (int)r;
}

__global__ void __launch_bounds__(256) k_desc_direct(int ntiles, const TileDesc *__restrict__ all,
                                                      const int *__restrict__ nzprefix,
                                                      TileDesc *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles)
    return;
  TileDesc d = all[t];
  d.head_end = nzprefix[d.r0]; // number of non-empty rows in front of the tile (r0 <= m)
  out[t] = d;
}

__global__ void __launch_bounds__(256) k_gather_desc(const int *__restrict__ list, int n,
                                                      const TileDesc *__restrict__ all, TileDesc *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = all[list[i]];
}

__global__ void __launch_bounds__(256) k_row_bins(const int *__restrict__ rowptr, int m, int T, int short_max,
                                                   int medium_max, unsigned char *__restrict__ out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (long long)gridDim.x * blockDim.x) {
    const int len = __ldg(rowptr + r + 1) - __ldg(rowptr + r);
    out[r] = (unsigned char)(len <= short_max ? 0 : (len <= medium_max ? 1 : (len <= T ? 2 : 3)));
  }
}

__global__ void __launch_bounds__(256) k_shard_bounds(const int *__restrict__ rowptr, int m, int nshards,
                                                       int *__restrict__ bounds) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > nshards)
    return;
  const long long base = __ldg(rowptr);
  const long long total = (long long)__ldg(rowptr + m) - base;
  if (g == 0)
    bounds[g] = 0;
  else if (g == nshards)
    bounds[g] = m;
  else
    bounds[g] = lower_bound_rowptr(rowptr, m, base + (total * g) / nshards);
}

__global__ void __launch_bounds__(256)
    k_col_block_bitmap(const int *__restrict__ col, long long nnz, int block_shift, unsigned int *__restrict__ bitmap) {
  // bitmap is one 32-bit word per block (0/1); integer OR is order independent, hence deterministic.
  int last = -1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
    const int b = __ldg(col + i) >> block_shift;
    if (b != last) {
      if (bitmap[b] == 0u)
        bitmap[b] = 1u;
      last = b;
    }
  }
}

// Gather-coalescing statistic: the row kernels map the lanes of a warp to consecutive rows, so the x gather of a
// warp touches as many 128-byte lines as there are distinct (colindex >> 4) among 32 consecutive rows at the same
// position inside the row. Sampled on kGatherSamples groups of 32 rows (middle element of every non-empty row):
// stat[0] = active lanes, stat[1] = distinct lines. Stencil / banded matrices give ~0.1 lines per lane, random
// column patterns ~1.0. Integer counts of a fixed sample: deterministic, restated in oracle/analysis_port.c.
constexpr int kGatherSamples = 4096;
__global__ void __launch_bounds__(256)
    k_gather_stat(const int *__restrict__ rowptr, const int *__restrict__ col, int m, int medium_max,
                  unsigned long long *stat) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= kGatherSamples)
    return;
  const long long span = m > 32 ? (long long)(m - 32) : 0;
  const long long r = (span * warp) / kGatherSamples + lane;
  int line = -1, len = 0;
  if (r < m) {
    const int s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    len = e - s;
    if (e > s)
      line = __ldg(col + s + ((e - s) >> 1)) >> 4;
  }
  const unsigned active = __ballot_sync(0xffffffffu, line >= 0);
  const unsigned same = __match_any_sync(0xffffffffu, line);
  const bool leader = (line >= 0) && ((__ffs(same) - 1) == lane);
  const unsigned leaders = __ballot_sync(0xffffffffu, leader);
  // stat[2] = non-zeros of the sampled rows, stat[3] = those in rows longer than medium_max (row-length skew)
  unsigned long long nz = (unsigned long long)len, nzl = len > medium_max ? (unsigned long long)len : 0ull;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    nz += __shfl_xor_sync(0xffffffffu, nz, off);
    nzl += __shfl_xor_sync(0xffffffffu, nzl, off);
  }
  if (lane == 0) {
    atomicAdd(stat, (unsigned long long)__popc(active));
    atomicAdd(stat + 1, (unsigned long long)__popc(leaders));
    atomicAdd(stat + 2, nz);
    atomicAdd(stat + 3, nzl);
  }
}

static inline int grid_for(long long n, int threads, int cap) {
  long long g = (n + threads - 1) / threads;
  if (g < 1)
    g = 1;
  if (g > cap)
    g = cap;
  return (int)g;
}

int analysis_prepare(spmv_b200_plan *p, cudaStream_t stream) {
  // base / end of the nnz range (rowptr may be a view into a larger matrix: rowptr[0] need not be 0)
  if (p->m == 0) {
    p->nnz = 0;
    p->elem_end = 0;
    return SPMV_B200_OK;
  }
  int h_be[2];
  B200_CUDA(cudaMemcpyAsync(&h_be[0], p->rowptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaMemcpyAsync(&h_be[1], p->rowptr + p->m, sizeof(int), cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaStreamSynchronize(stream));
  const long long total = (long long)h_be[1] - (long long)h_be[0];
  if (total < 0) {
    set_error("rowptr is not monotone: rowptr[m] < rowptr[0]");
    return SPMV_B200_ERR_ARG;
  }
  if (p->nnz >= 0 && p->nnz != total) {
    set_error("nnz argument does not match rowptr[m] - rowptr[0]");
    return SPMV_B200_ERR_ARG;
  }
  p->nnz = total;
  p->elem_base = h_be[0];
  p->elem_end = h_be[1];
  p->gather_active = p->gather_lines = p->sample_nnz = p->sample_nnz_long = 0;
  if (total > 0 && p->col) {
    unsigned long long h_stat[4];
    DeviceScratch stat_buf;
    B200_CUDA(stat_buf.alloc(sizeof(h_stat)));
    unsigned long long *d_stat = stat_buf.as<unsigned long long>();
    B200_CUDA(cudaMemsetAsync(d_stat, 0, sizeof(h_stat), stream));
    k_gather_stat<<<kGatherSamples * 32 / 256, 256, 0, stream>>>(p->rowptr, p->col, p->m, p->medium_max, d_stat);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(h_stat, d_stat, sizeof(h_stat), cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    p->gather_active = (long long)h_stat[0];
    p->gather_lines = (long long)h_stat[1];
    p->sample_nnz = (long long)h_stat[2];
    p->sample_nnz_long = (long long)h_stat[3];
  }
  return SPMV_B200_OK;
}

int analysis_run(spmv_b200_plan *p, cudaStream_t stream) {
  const int m = p->m;
  if (m == 0) {
    p->ntiles = 0;
    return SPMV_B200_OK;
  }
  const long long total = p->nnz;
  const int h_be[2] = {(int)p->elem_base, (int)p->elem_end};
  const long long nt = std::max<long long>(1, (total + m + p->T - 1) / p->T);
  if (nt > 0x7ffffff0LL) {
    set_error("too many tiles");
    return SPMV_B200_ERR_ARG;
  }
  const int ntiles = (int)nt;
  p->ntiles = ntiles;

  DeviceScratch aux_buf, hist_buf; // freed on every return path
  size_t ws = 0;
  B200_CUDA(cudaMalloc(&p->tile_row, sizeof(int) * (ntiles + 1)));
  B200_CUDA(cudaMalloc(&p->tile_elem, sizeof(int) * (ntiles + 1)));
  B200_CUDA(cudaMalloc(&p->tile_part, sizeof(int) * (ntiles + 1)));
  B200_CUDA(cudaMalloc(&p->tile_split, ntiles + 1));
  B200_CUDA(cudaMalloc(&p->tile_maxlen, sizeof(int) * ntiles));
  B200_CUDA(cudaMalloc(&p->tile_kind, ntiles));
  B200_CUDA(aux_buf.alloc(sizeof(int) * (ntiles + 1)));
  B200_CUDA(hist_buf.alloc(sizeof(unsigned long long) * 8));
  int *tile_aux = aux_buf.as<int>();
  unsigned long long *d_hist = hist_buf.as<unsigned long long>();
  ws += sizeof(int) * (size_t)(ntiles + 1) * 3 + (ntiles + 1) + sizeof(int) * (size_t)ntiles + ntiles;
  B200_CUDA(cudaMemsetAsync(p->tile_maxlen, 0, sizeof(int) * ntiles, stream));
  B200_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 8, stream));

  k_tile_bounds<<<grid_for(ntiles + 1, 256, 1 << 30), 256, 0, stream>>>(p->rowptr, m, ntiles, p->T, p->medium_max,
                                                                         p->tile_row, p->tile_elem, p->tile_split,
                                                                         p->tile_part, tile_aux);
  k_row_stats<<<grid_for(m, 256, 148 * 16), 256, 0, stream>>>(p->rowptr, m, ntiles, p->T, p->short_max, p->medium_max,
                                                               p->tile_maxlen, d_hist);
  k_tile_kind<<<grid_for(ntiles, 256, 1 << 30), 256, 0, stream>>>(ntiles, p->short_max, p->medium_max, p->tile_maxlen,
                                                                   p->tile_split, p->tile_row, p->tile_elem,
                                                                   p->tile_kind);
  B200_CUDA(cudaMalloc(&p->desc_all, sizeof(TileDesc) * (size_t)ntiles));
  ws += sizeof(TileDesc) * (size_t)ntiles;
  k_tile_desc<<<grid_for(ntiles, 256, 1 << 30), 256, 0, stream>>>(p->rowptr, ntiles, p->tile_row, p->tile_elem,
                                                                   p->tile_split, p->desc_all);
  B200_CUDA(cudaGetLastError());
  if (p->direct) {
    const size_t words = (size_t)((p->elem_end + 31) / 32) + 16; // slack: a warp reads whole 128-element windows
    DeviceScratch nzflag_buf, nzprefix_buf, scan_buf;
    size_t scan_bytes = 0;
    B200_CUDA(cudaMalloc(&p->row_start_bits, sizeof(unsigned int) * words));
    B200_CUDA(cudaMemsetAsync(p->row_start_bits, 0, sizeof(unsigned int) * words, stream));
    B200_CUDA(nzflag_buf.alloc(sizeof(int) * ((size_t)m + 1)));
    B200_CUDA(nzprefix_buf.alloc(sizeof(int) * ((size_t)m + 1)));
    int *nzflag = nzflag_buf.as<int>(), *nzprefix = nzprefix_buf.as<int>();
    k_row_start_bits<<<grid_for((long long)m + 1, 256, 148 * 16), 256, 0, stream>>>(p->rowptr, m, p->row_start_bits,
                                                                                    nzflag);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, nzflag, nzprefix, m + 1, stream));
    B200_CUDA(scan_buf.alloc(scan_bytes));
    B200_CUDA(cub::DeviceScan::ExclusiveSum(scan_buf.p, scan_bytes, nzflag, nzprefix, m + 1, stream));
    int h_nz = 0;
    B200_CUDA(cudaMemcpyAsync(&h_nz, nzprefix + m, sizeof(int), cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    p->n_nz_rows = h_nz;
    // padded by one tile's worth of rows: the direct kernel requests (prefetch, no fault semantics relied upon) the
    // row ids of a block by its row count, which can exceed its number of non-empty rows
    B200_CUDA(cudaMalloc(&p->nz_rows, sizeof(int) * ((size_t)h_nz + 1 + (size_t)p->T + 64)));
    k_nz_rows<<<grid_for(m, 256, 148 * 16), 256, 0, stream>>>(nzflag, nzprefix, m, p->nz_rows);
    B200_CUDA(cudaMalloc(&p->desc_direct, sizeof(TileDesc) * (size_t)ntiles));
    k_desc_direct<<<grid_for(ntiles, 256, 1 << 30), 256, 0, stream>>>(ntiles, p->desc_all, nzprefix, p->desc_direct);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(stream));
    ws += sizeof(unsigned int) * words + sizeof(int) * ((size_t)h_nz + 1) + sizeof(TileDesc) * (size_t)ntiles;
  }

  // finalise on the host: compact per-kind tile lists (ascending tile id) and the split-row table
  std::vector<unsigned char> h_kind(ntiles), h_split(ntiles + 1);
  std::vector<int> h_row(ntiles + 1), h_aux(ntiles + 1);
  unsigned long long h_hist[8];
  B200_CUDA(cudaMemcpyAsync(h_kind.data(), p->tile_kind, ntiles, cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaMemcpyAsync(h_split.data(), p->tile_split, ntiles + 1, cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaMemcpyAsync(h_row.data(), p->tile_row, sizeof(int) * (ntiles + 1), cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaMemcpyAsync(h_aux.data(), tile_aux, sizeof(int) * (ntiles + 1), cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaMemcpyAsync(h_hist, d_hist, sizeof(h_hist), cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaStreamSynchronize(stream));
  for (int b = 0; b < 4; ++b) {
    p->bin_rows[b] = (long long)h_hist[b];
    p->bin_nnz[b] = (long long)h_hist[4 + b];
  }

  std::vector<int> lists[3];
  for (int t = 0; t < ntiles; ++t)
    lists[h_kind[t]].push_back(t);
  p->h_tile_row = h_row;
  for (int k = 0; k < 3; ++k) {
    p->h_list[k] = lists[k];
    p->count[k] = (int)lists[k].size();
    // when every tile has the same kind the kernel indexes tiles by blockIdx directly (no list)
    if (p->count[k] > 0 && p->count[k] < ntiles) {
      B200_CUDA(cudaMalloc(&p->list[k], sizeof(int) * lists[k].size()));
      B200_CUDA(cudaMemcpyAsync(p->list[k], lists[k].data(), sizeof(int) * lists[k].size(), cudaMemcpyHostToDevice,
                                stream));
      B200_CUDA(cudaMalloc(&p->desc[k], sizeof(TileDesc) * lists[k].size()));
      k_gather_desc<<<grid_for((long long)lists[k].size(), 256, 1 << 30), 256, 0, stream>>>(
          p->list[k], (int)lists[k].size(), p->desc_all, p->desc[k]);
      B200_CUDA(cudaGetLastError());
      ws += (sizeof(int) + sizeof(TileDesc)) * lists[k].size();
    } else if (p->count[k] == ntiles) {
      p->desc[k] = p->desc_all; // one kind owns every tile: no compaction needed
    }
  }
  // split rows: boundary t is the first split boundary of its row iff the row is owned by tile t-1
  std::vector<int> srow, st0, st1;
  const long long base = h_be[0];
  for (int t = 1; t < ntiles; ++t) {
    if (!h_split[t])
      continue;
    const int r = h_row[t] - 1;
    if (h_row[t - 1] <= r) { // row r starts in tile t-1
      srow.push_back(r);
      st0.push_back(t - 1);
      long long t1 = ((long long)h_aux[t] - 1 - base + r) / p->T;
      if (t1 > ntiles - 1)
        t1 = ntiles - 1;
      st1.push_back((int)t1);
    }
  }
  p->nsplit = (int)srow.size();
  if (p->nsplit > 0) {
    std::vector<int> packed;
    packed.reserve(3 * srow.size());
    packed.insert(packed.end(), srow.begin(), srow.end());
    packed.insert(packed.end(), st0.begin(), st0.end());
    packed.insert(packed.end(), st1.begin(), st1.end());
    B200_CUDA(cudaMalloc(&p->split_rows, sizeof(int) * packed.size()));
    B200_CUDA(cudaMemcpyAsync(p->split_rows, packed.data(), sizeof(int) * packed.size(), cudaMemcpyHostToDevice,
                              stream));
    B200_CUDA(cudaMalloc(&p->partials, sizeof(double) * 2 * (size_t)ntiles));
    B200_CUDA(cudaMemsetAsync(p->partials, 0, sizeof(double) * 2 * (size_t)ntiles, stream));
    ws += sizeof(int) * packed.size() + sizeof(double) * 2 * (size_t)ntiles;
  }
  B200_CUDA(cudaStreamSynchronize(stream));
  p->workspace_bytes = ws;
  return SPMV_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// staged-x form: per row block, the 128-byte lines of x it references, their runs, and 16-bit local column indices
// ---------------------------------------------------------------------------------------------------------------
// Specification (bit-exact restatement: oracle/analysis_port.c:port_xstage). For tile t with elements [e0, e1):
//   lines(t)   = sorted set of { col[k] >> 4 : e0 <= k < e1 }            (a line = 16 consecutive entries of x)
//   rank(l)    = number of lines of the set smaller than l
//   segments   = maximal runs of consecutive line numbers; segment s = (first line, rank of its first line)
//   lcol[k]    = (rank(col[k] >> 4) << 4) | (col[k] & 15)                (offset into the staged copy, in doubles)
//   the tile qualifies iff it has elements, last line - first line < kXspanLinesMax, |lines| <= kXlinesMax and
//   #segments <= kXsegMax; the plan takes the form iff every tile qualifies.
//   Too many segments (unstructured meshes: many short runs a few lines apart) are repaired by staging the lines in
//   between as well: for g = 1, 2, 4, 8, 16, 32 in turn, runs separated by at most g missing lines are merged; the first
//   g that leaves <= kXsegMax segments is taken, and the tile qualifies if the merged segments hold <= kXlinesMax
//   lines (at most kXrunsCap runs are examined). rank(l) is then the position of l in the merged segments:
//   rank(l) = off[s] + (l - line[s]) for the segment s that contains l (identical to the definition above when nothing
//   was merged).
// One CTA per tile: a bitmap of the line span in shared memory (integer OR), a scan of its popcounts for the ranks.
__global__ void __launch_bounds__(256)
    k_xstage_build(const int *__restrict__ col, const TileDesc *__restrict__ desc, int ntiles, long long lcol_base,
                   unsigned short *__restrict__ lcol, XDesc *__restrict__ xdesc, int *__restrict__ status /* [2]: failed, max lines */) {
  extern __shared__ unsigned int xs_smem[];
  unsigned int *bitmap = xs_smem;                                          // [kXspanLinesMax / 32]
  int *prefix = reinterpret_cast<int *>(xs_smem + kXspanLinesMax / 32);    // lines in front of each word
  __shared__ int s_red[2][8];
  __shared__ int s_scan[9], s_runs[9];
  __shared__ int s_rs[kXrunsCap], s_re[kXrunsCap];                     // run starts / ends (merge path only)
  __shared__ int s_mline[kXsegMax], s_moff[kXsegMax], s_mseg, s_mlines; // merged segment table
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (t >= ntiles)
    return;
  XDesc *xd = xdesc + t;
  __shared__ int s_abort;
  if (tid == 0)
    s_abort = *reinterpret_cast<volatile int *>(status); // another tile has already failed: the plan will not use the form
  __syncthreads();
  if (s_abort != 0)
    return;
  const int e0 = desc[t].e0, e1 = desc[t].e1;
  // span of lines
  int lo = 0x7fffffff, hi = -1;
  for (int k = e0 + tid; k < e1; k += 256) {
    const int l = __ldg(col + k) >> 4;
    lo = l < lo ? l : lo;
    hi = l > hi ? l : hi;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const int l2 = __shfl_xor_sync(0xffffffffu, lo, off), h2 = __shfl_xor_sync(0xffffffffu, hi, off);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if (lane == 0) {
    s_red[0][warp] = lo;
    s_red[1][warp] = hi;
  }
  __syncthreads();
  lo = s_red[0][0];
  hi = s_red[1][0];
  for (int w = 1; w < 8; ++w) {
    lo = s_red[0][w] < lo ? s_red[0][w] : lo;
    hi = s_red[1][w] > hi ? s_red[1][w] : hi;
  }
  bool ok = e1 > e0 && hi >= lo && (hi - lo) < kXspanLinesMax;
  int nwords = ok ? ((hi - lo) >> 5) + 1 : 0;
  if (ok) {
    for (int w = tid; w < nwords; w += 256)
      bitmap[w] = 0u;
    __syncthreads();
    for (int k = e0 + tid; k < e1; k += 256) {
      const int l = (__ldg(col + k) >> 4) - lo;
      atomicOr(bitmap + (l >> 5), 1u << (l & 31));
    }
    __syncthreads();
    // ranks: exclusive scan of the popcounts of the words; runs: set bits whose predecessor bit is clear
    const int per = (nwords + 255) / 256; // consecutive words per thread
    int cnt = 0, runs = 0;
    for (int j = 0; j < per; ++j) {
      const int w = tid * per + j;
      if (w < nwords) {
        const unsigned b = bitmap[w];
        const unsigned prev = w > 0 ? bitmap[w - 1] >> 31 : 0u;
        cnt += __popc(b);
        runs += __popc(b & ~((b << 1) | prev));
      }
    }
    int icnt = cnt, iruns = runs;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int c2 = __shfl_up_sync(0xffffffffu, icnt, off), r2 = __shfl_up_sync(0xffffffffu, iruns, off);
      if (lane >= off) {
        icnt += c2;
        iruns += r2;
      }
    }
    if (lane == 31) {
      s_scan[warp + 1] = icnt;
      s_runs[warp + 1] = iruns;
    }
    __syncthreads();
    if (tid == 0) {
      s_scan[0] = 0;
      s_runs[0] = 0;
      for (int w = 1; w <= 8; ++w) {
        s_scan[w] += s_scan[w - 1];
        s_runs[w] += s_runs[w - 1];
      }
    }
    __syncthreads();
    const int nlines = s_scan[8], nruns = s_runs[8];
    const bool plain = nlines <= kXlinesMax && nruns <= kXsegMax;
    bool merged = false;
    if (!plain && nruns > kXsegMax && nruns <= kXrunsCap && nlines <= kXlinesMax) { // (uniform over the CTA)
      if (tid == 0) { // serial: only tiles of irregular meshes come here
        int nrs = 0, nre = 0;
        unsigned in = 0u;
        for (int w = 0; w < nwords; ++w) {
          const unsigned b = bitmap[w];
          const unsigned next_lsb = (w + 1 < nwords) ? (bitmap[w + 1] & 1u) : 0u;
          unsigned starts = b & ~((b << 1) | in), ends = b & ~((b >> 1) | (next_lsb << 31));
          while (starts) {
            s_rs[nrs++] = 32 * w + __ffs(starts) - 1;
            starts &= starts - 1;
          }
          while (ends) {
            s_re[nre++] = 32 * w + __ffs(ends); // one past the last line of the run
            ends &= ends - 1;
          }
          in = b >> 31;
        }
        int pick = 0;
        for (int g = 1; g <= 32 && !pick; g <<= 1) {
          int cnt = 1;
          for (int i = 1; i < nrs; ++i)
            cnt += (s_rs[i] - s_re[i - 1]) > g;
          if (cnt <= kXsegMax)
            pick = g;
        }
        int nseg = -1, total = 0;
        if (pick) {
          nseg = 0;
          int start = s_rs[0];
          for (int i = 1; i <= nrs; ++i)
            if (i == nrs || (s_rs[i] - s_re[i - 1]) > pick) {
              s_mline[nseg] = start;
              s_moff[nseg] = total;
              total += s_re[i - 1] - start;
              ++nseg;
              if (i < nrs)
                start = s_rs[i];
            }
          if (total > kXlinesMax)
            nseg = -1;
        }
        s_mseg = nseg;
        s_mlines = total;
      }
      __syncthreads();
      merged = s_mseg > 0;
    }
    ok = plain || merged;
    if (merged) {
      const int nseg = s_mseg;
      if (tid < kXsegMax) {
        xd->line[tid] = tid < nseg ? lo + s_mline[tid] : 0;
        xd->off[tid] = tid < nseg ? (unsigned short)s_moff[tid] : (unsigned short)0;
      }
      if (tid < 6)
        xd->pad[tid] = 0;
      if (tid == 0) {
        xd->nseg = nseg;
        xd->nlines = s_mlines;
        atomicMax(status + 1, s_mlines);
      }
      for (int k = e0 + tid; k < e1; k += 256) {
        const int c = __ldg(col + k);
        const int l = (c >> 4) - lo;
        int sgm = 0;
        for (int j = 1; j < nseg; ++j)
          sgm = s_mline[j] <= l ? j : sgm;
        const int rank = s_moff[sgm] + (l - s_mline[sgm]);
        lcol[(long long)k - lcol_base] = (unsigned short)((rank << 4) | (c & 15));
      }
    } else if (ok) {
      int base = s_scan[warp] + icnt - cnt, rbase = s_runs[warp] + iruns - runs; // exclusive prefixes of this thread
      for (int j = 0; j < per; ++j) {
        const int w = tid * per + j;
        if (w < nwords) {
          const unsigned b = bitmap[w];
          const unsigned prev = w > 0 ? bitmap[w - 1] >> 31 : 0u;
          prefix[w] = base;
          unsigned starts = b & ~((b << 1) | prev);
          while (starts) { // segment table: first line and rank of every run
            const int bit = __ffs(starts) - 1;
            starts &= starts - 1;
            xd->line[rbase] = lo + 32 * w + bit;
            xd->off[rbase] = (unsigned short)(base + __popc(b & ((1u << bit) - 1u)));
            ++rbase;
          }
          base += __popc(b);
        }
      }
      if (tid == 0) {
        xd->nseg = nruns;
        xd->nlines = nlines;
        atomicMax(status + 1, nlines);
      }
      for (int s = nruns + tid; s < kXsegMax; s += 256) {
        xd->line[s] = 0;
        xd->off[s] = 0;
      }
      if (tid < 6)
        xd->pad[tid] = 0;
      __syncthreads();
      for (int k = e0 + tid; k < e1; k += 256) {
        const int c = __ldg(col + k);
        const int l = (c >> 4) - lo;
        const int rank = prefix[l >> 5] + __popc(bitmap[l >> 5] & ((1u << (l & 31)) - 1u));
        lcol[(long long)k - lcol_base] = (unsigned short)((rank << 4) | (c & 15));
      }
    }
  }
  if (!ok && tid == 0) {
    xd->nseg = -1;
    atomicExch(status, 1);
  }
}

int analysis_xstage(spmv_b200_plan *p, cudaStream_t stream) {
  p->xstage = false;
  // regular matrices only: one row-kernel kind owns every row block, nothing is split, TMA is usable
  const bool one_kind = p->ntiles > 0 && (p->count[SPMV_B200_KIND_SHORT] == p->ntiles ||
                                           p->count[SPMV_B200_KIND_MEDIUM] == p->ntiles);
  if (!one_kind || p->direct || p->nsplit > 0 || !p->uses_tma || p->irregular || p->nnz == 0 ||
      (p->flags & SPMV_B200_FLAG_NO_XSTAGE))
    return SPMV_B200_OK;
  p->lcol_base = p->elem_base & ~7LL; // keeps lcol + (a0 - base) 16-byte aligned for a0 multiple of 8
  const size_t n16 = (size_t)(p->elem_end - p->lcol_base) + 16;
  DeviceScratch status;
  B200_CUDA(status.alloc(2 * sizeof(int)));
  B200_CUDA(cudaMemsetAsync(status.p, 0, 2 * sizeof(int), stream));
  cudaError_t e = cudaMalloc(&p->lcol, sizeof(unsigned short) * n16);
  if (e == cudaSuccess)
    e = cudaMalloc(&p->xdesc, sizeof(XDesc) * (size_t)p->ntiles);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(p->lcol, 0, sizeof(unsigned short) * n16, stream);
  int h_status[2] = {1, 0};
  if (e == cudaSuccess) {
    const size_t smem = sizeof(unsigned int) * (kXspanLinesMax / 32) * 2;
    k_xstage_build<<<p->ntiles, 256, smem, stream>>>(p->col, p->desc_all, p->ntiles, p->lcol_base, p->lcol, p->xdesc,
                                                     status.as<int>());
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h_status, status.p, sizeof(h_status), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess || h_status[0] != 0) { // some row block does not qualify: the plan keeps the gather kernels
    cudaFree(p->lcol);
    cudaFree(p->xdesc);
    p->lcol = nullptr;
    p->xdesc = nullptr;
    B200_CUDA(e);
    return SPMV_B200_OK;
  }
  p->xstage = true;
  p->xlines = h_status[1];
  p->workspace_bytes += sizeof(unsigned short) * n16 + sizeof(XDesc) * (size_t)p->ntiles;
  return SPMV_B200_OK;
}

// descriptors in a caller-given tile order (the fused halo loop walks boundary row blocks first)
int analysis_gather_descs(const TileDesc *d_all, const int *h_order, int n, TileDesc **d_out, cudaStream_t stream) {
  *d_out = nullptr;
  if (n <= 0)
    return SPMV_B200_OK;
  int *d_order = nullptr;
  TileDesc *out = nullptr;
  B200_CUDA(cudaMalloc(&d_order, sizeof(int) * (size_t)n));
  cudaError_t e = cudaMalloc(&out, sizeof(TileDesc) * (size_t)n);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_order, h_order, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) {
    k_gather_desc<<<grid_for(n, 256, 1 << 30), 256, 0, stream>>>(d_order, n, d_all, out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(stream);
  cudaFree(d_order);
  if (e != cudaSuccess) {
    cudaFree(out);
    B200_CUDA(e);
  }
  *d_out = out;
  return SPMV_B200_OK;
}

// smallest / largest column index referenced by every tile (one warp per tile): tells a sharded caller which row blocks
// read entries of x owned by other GPUs. Integer min / max: order independent.
__global__ void __launch_bounds__(256) k_tile_col_range(const int *__restrict__ col, int ntiles,
                                                         const int *__restrict__ tile_elem, int *__restrict__ cmin,
                                                         int *__restrict__ cmax) {
  const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= ntiles)
    return;
  int lo = 0x7fffffff, hi = -1;
  for (int k = tile_elem[t] + lane; k < tile_elem[t + 1]; k += 32) {
    const int c = __ldg(col + k);
    lo = c < lo ? c : lo;
    hi = c > hi ? c : hi;
  }
  for (int off = 16; off > 0; off >>= 1) {
    const int l2 = __shfl_xor_sync(0xffffffffu, lo, off), h2 = __shfl_xor_sync(0xffffffffu, hi, off);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if (lane == 0) {
    cmin[t] = lo;
    cmax[t] = hi;
  }
}

int analysis_tile_col_range(const spmv_b200_plan *p, int *h_min, int *h_max, cudaStream_t stream) {
  if (p->ntiles == 0)
    return SPMV_B200_OK;
  int *d = nullptr;
  B200_CUDA(cudaMalloc(&d, sizeof(int) * 2 * (size_t)p->ntiles));
  k_tile_col_range<<<grid_for((long long)p->ntiles * 32, 256, 1 << 30), 256, 0, stream>>>(p->col, p->ntiles,
                                                                                           p->tile_elem, d,
                                                                                           d + p->ntiles);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h_min, d, sizeof(int) * (size_t)p->ntiles, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h_max, d + p->ntiles, sizeof(int) * (size_t)p->ntiles, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(stream);
  cudaFree(d);
  B200_CUDA(e);
  return SPMV_B200_OK;
}

int analysis_row_bins(const spmv_b200_plan *p, unsigned char *d_out, cudaStream_t stream) {
  if (p->m == 0)
    return SPMV_B200_OK;
  k_row_bins<<<grid_for(p->m, 256, 148 * 16), 256, 0, stream>>>(p->rowptr, p->m, p->T, p->short_max, p->medium_max,
                                                                 d_out);
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

int shard_bounds_run(int m, long long nnz, const int *d_rowptr, int nshards, int *h_bounds, cudaStream_t stream) {
  (void)nnz;
  if (nshards < 1 || m < 0) {
    set_error("shard_bounds: nshards must be >= 1 and m >= 0");
    return SPMV_B200_ERR_ARG;
  }
  if (m == 0) {
    for (int g = 0; g <= nshards; ++g)
      h_bounds[g] = 0;
    return SPMV_B200_OK;
  }
  int *d_bounds = nullptr;
  B200_CUDA(cudaMalloc(&d_bounds, sizeof(int) * (nshards + 1)));
  k_shard_bounds<<<grid_for(nshards + 1, 256, 1 << 20), 256, 0, stream>>>(d_rowptr, m, nshards, d_bounds);
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(h_bounds, d_bounds, sizeof(int) * (nshards + 1), cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaStreamSynchronize(stream));
  B200_CUDA(cudaFree(d_bounds));
  return SPMV_B200_OK;
}

int col_block_bitmap_run(long long nnz, const int *d_col, int n, int block_shift, unsigned char *h_bitmap,
                         cudaStream_t stream) {
  if (block_shift < 0 || block_shift > 30 || n < 0) {
    set_error("col_block_bitmap: bad block_shift or n");
    return SPMV_B200_ERR_ARG;
  }
  const long long nblocks = ((long long)n + (1LL << block_shift) - 1) >> block_shift;
  if (nblocks == 0)
    return SPMV_B200_OK;
  unsigned int *d_bm = nullptr;
  B200_CUDA(cudaMalloc(&d_bm, sizeof(unsigned int) * nblocks));
  B200_CUDA(cudaMemsetAsync(d_bm, 0, sizeof(unsigned int) * nblocks, stream));
  if (nnz > 0) {
    k_col_block_bitmap<<<grid_for(nnz, 256, 148 * 16), 256, 0, stream>>>(d_col, nnz, block_shift, d_bm);
    B200_CUDA(cudaGetLastError());
  }
  std::vector<unsigned int> h(nblocks);
  B200_CUDA(cudaMemcpyAsync(h.data(), d_bm, sizeof(unsigned int) * nblocks, cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaStreamSynchronize(stream));
  B200_CUDA(cudaFree(d_bm));
  for (long long b = 0; b < nblocks; ++b)
    h_bitmap[b] = h[b] ? 1 : 0;
  return SPMV_B200_OK;
}

} // namespace b200
