// ORACLE — TEST INFRASTRUCTURE ONLY.
// Door onto the reference's merge-path `partition` kernel (benchmark/merge-path/merge_path_partition.h:7-17),
// compiled in place from /root/reference by `make ref-gpu`. Launch shape follows merge_path_spmv.cu:44
// (<<<512, 256>>>, ITEMS_PER_BLOCK = 512 * 4 = 2048). Used by the GPU tests to pin the TILE_PART array of the
// CUDA analysis (tile_nnz = 2048) against the reference's own output.
#include <cuda_runtime.h>

#include "merge_path_partition.h"

extern "C" int ref_gpu_merge_path_partition_2048(const int *d_rowptr, int m, int count, int *d_S) {
  partition<int, 256, 2048><<<512, 256>>>(d_rowptr, m, count, d_S);
  return (int)cudaDeviceSynchronize();
}
