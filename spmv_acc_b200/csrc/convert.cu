// COO -> CSR on the device: the step in front of the hot path for MatrixMarket input. Device-side form
// of the reference's host conversion matrix_market::to_csr (cli/sparse_format.h:100-128: sort the entries by
// (row, col), count the rows, prefix-sum; duplicates are kept). Here: one stable radix sort of 64-bit keys
// (row << 32 | col) carrying the entry index, a gather of the values, and one binary search per row for the offsets.
// Equal (row, col) pairs keep their input order (the reference's std::sort leaves that order unspecified).
#include <cstdint>

#include <cub/device/device_radix_sort.cuh>

#include "internal.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) k_coo_keys(long long nnz, const int *__restrict__ row,
                                                  const int *__restrict__ col, int m, int n,
                                                  unsigned long long *__restrict__ keys, int *__restrict__ idx,
                                                  int *__restrict__ bad) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
    const int r = row[i], c = col[i];
    if (r < 0 || r >= m || c < 0 || c >= n)
      *bad = 1; // "index out of bounds in matrix market file" (cli/matrix_market_reader.hpp)
    keys[i] = ((unsigned long long)(unsigned)r << 32) | (unsigned)c;
    idx[i] = (int)i;
  }
}

__global__ void __launch_bounds__(256) k_coo_scatter(long long nnz, const unsigned long long *__restrict__ keys,
                                                     const int *__restrict__ idx, const double *__restrict__ val,
                                                     int *__restrict__ col_out, double *__restrict__ val_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
    col_out[i] = (int)(keys[i] & 0xffffffffull);
    val_out[i] = val[idx[i]];
  }
}

// rowptr[r] = number of entries with row < r = lower_bound(sorted keys, r << 32)
__global__ void __launch_bounds__(256) k_coo_rowptr(int m, long long nnz, const unsigned long long *__restrict__ keys,
                                                    int *__restrict__ rowptr) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > m)
    return;
  const unsigned long long target = (unsigned long long)r << 32;
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    if (keys[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  rowptr[r] = (int)lo;
}

int coo_to_csr_run(int m, int n, long long nnz, const int *d_row, const int *d_col, const double *d_val,
                   int *d_rowptr_out, int *d_col_out, double *d_val_out, cudaStream_t stream) {
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  int *idx = nullptr, *idx2 = nullptr, *bad = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = SPMV_B200_OK, h_bad = 0;
  const size_t cnt = nnz > 0 ? (size_t)nnz : 1;
  auto cleanup = [&]() {
    cudaFree(keys);
    cudaFree(keys2);
    cudaFree(idx);
    cudaFree(idx2);
    cudaFree(bad);
    cudaFree(tmp);
  };
#define COO_CUDA(call)                                                                                                 \
  do {                                                                                                                 \
    cudaError_t _e = (call);                                                                                           \
    if (_e != cudaSuccess) {                                                                                           \
      cleanup();                                                                                                       \
      return cuda_fail(_e, #call, __FILE__, __LINE__);                                                                 \
    }                                                                                                                  \
  } while (0)
  COO_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * cnt));
  COO_CUDA(cudaMalloc(&keys2, sizeof(unsigned long long) * cnt));
  COO_CUDA(cudaMalloc(&idx, sizeof(int) * cnt));
  COO_CUDA(cudaMalloc(&idx2, sizeof(int) * cnt));
  COO_CUDA(cudaMalloc(&bad, sizeof(int)));
  COO_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
  if (nnz > 0) {
    const int grid = (int)std::min<long long>((nnz + 255) / 256, 148 * 16);
    k_coo_keys<<<grid, 256, 0, stream>>>(nnz, d_row, d_col, m, n, keys, idx, bad);
    COO_CUDA(cudaGetLastError());
    COO_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    COO_CUDA(cudaStreamSynchronize(stream));
    if (h_bad) {
      cleanup();
      set_error("coo_to_csr: index out of bounds");
      return SPMV_B200_ERR_ARG;
    }
    int row_bits = 1;
    while (row_bits < 32 && (1ll << row_bits) < (long long)m)
      ++row_bits;
    const int end_bit = 32 + row_bits; // only the bits that can differ are sorted
    COO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, (int)nnz, 0, end_bit, stream));
    COO_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    COO_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, idx2, (int)nnz, 0, end_bit, stream));
    k_coo_scatter<<<grid, 256, 0, stream>>>(nnz, keys2, idx2, d_val, d_col_out, d_val_out);
    COO_CUDA(cudaGetLastError());
  }
  k_coo_rowptr<<<(m + 1 + 255) / 256, 256, 0, stream>>>(m, nnz, keys2, d_rowptr_out);
  COO_CUDA(cudaGetLastError());
  COO_CUDA(cudaStreamSynchronize(stream));
#undef COO_CUDA
  cleanup();
  return rc;
}

} // namespace b200
