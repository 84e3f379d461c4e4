// Internal declarations shared by the translation units of libspmv_b200.so.
// Nothing here is part of the ABI; the ABI is include/spmv_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "spmv_b200.h"

namespace b200 {

constexpr int kThreads = 256;     // threads per CTA of every streaming kernel
constexpr int kRowChunk = 512;    // row pointers staged in shared memory per pass (rows per pass = kRowChunk)
constexpr int kSerialMax = 8;     // MIXED kernel: rows up to this length are reduced by one thread
constexpr int kGroupMax = 96;     // MIXED kernel: rows up to this length are reduced by 8 lanes, longer ones by a warp
constexpr int kSparseTileRows = 512; // tiles owning more rows than this are streamed by the MIXED kernel
constexpr int kDefaultTile = 0; // 0 = chosen from the average row length (see auto_tile in capi.cu)
constexpr int kDefaultShort = 8;
constexpr int kDefaultMedium = 128;
constexpr int kDefaultVecDiv = 16;

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define B200_CUDA(call)                                                                                               \
  do {                                                                                                                 \
    cudaError_t _e = (call);                                                                                           \
    if (_e != cudaSuccess)                                                                                             \
      return ::b200::cuda_fail(_e, #call, __FILE__, __LINE__);                                                         \
  } while (0)

// scratch device memory / events that must not outlive an early error return (B200_CUDA returns from the function)
struct DeviceScratch {
  void *p = nullptr;
  DeviceScratch() = default;
  DeviceScratch(const DeviceScratch &) = delete;
  DeviceScratch &operator=(const DeviceScratch &) = delete;
  ~DeviceScratch() {
    if (p)
      cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

// Everything a CTA needs to know about its tile, packed into 32 bytes so that the prologue of a CTA is a single
// round trip to memory (two 128-bit loads) before the TMA copies can be issued. Built per tile kind by the analysis.
struct __align__(16) TileDesc {
  int r0, r1;     // owned rows [r0, r1)
  int e0, e1;     // streamed nnz [e0, e1)
  int head_end;   // split leading boundary: the head fragment is [e0, head_end) (= min(rowptr[r0], e1))
  int tail_start; // split trailing boundary: the tail fragment is [tail_start, e1) (= rowptr[r1-1])
  int tile;       // global tile id (index of its partial sums)
  int flags;      // bit 0: leading boundary split, bit 1: trailing boundary split
};

// Staged-x form ("XS") of the row kernels: the entries of x a row block references are brought into shared memory as
// whole 128-byte lines by TMA, and the block's column indices are stored a second time as 16-bit offsets into that
// staged copy (plan array `lcol`, 2 bytes per non-zero instead of 4). XDesc lists the runs of consecutive lines
// ("segments") of one row block: segment s covers lines [line[s], line[s] + off[s+1] - off[s]) of x and lands at line
// off[s] of the staged copy (off[nseg] = nlines). 128 bytes, one per row block.
constexpr int kXsegMax = 16;          // segments per row block (a 27-point stencil has 9, a 5-point stencil 3)
constexpr int kXlinesMax = 256;       // staged lines per row block (32 KB), upper limit
constexpr int kXspanLinesMax = 65536; // lines between the smallest and the largest column of a row block (bitmap size)
constexpr int kXrunsCap = 1024;       // runs of lines examined when too many segments are merged across small gaps
struct __align__(16) XDesc {
  int nseg;   // -1: the row block does not qualify
  int nlines; // staged lines
  int line[kXsegMax];
  unsigned short off[kXsegMax];
  int pad[6];
};
static_assert(sizeof(XDesc) == 128, "XDesc is one 128-byte line");

// Rows whose result is also stored into other GPUs' memory (fused halo push of the iterated multi-GPU loop):
// for row in [row_lo[j], row_hi[j]) the epilogue stores y to dst[j][row] as well (dst is a peer-mapped pointer,
// already offset so that it is indexed by the shard-local row).
constexpr int kMaxPush = SPMV_B200_MAX_PUSH;
struct PushArgs {
  int count;
  unsigned int multicast_mask; // bit j: dst[j] is an NVLink multicast address (stored to with multimem.st)
  int align_rows;              // pushed row blocks deal rows by absolute row index (SPMV_B200_HALO_ALIGN_PUSH)
  int row_lo[kMaxPush], row_hi[kMaxPush];
  double *dst[kMaxPush];
};

// Arguments of the streaming kernels (passed by value).
struct SpmvArgs {
  const int *__restrict__ rowptr;
  const int *__restrict__ col;
  const double *__restrict__ val;
  const double *__restrict__ x;
  double *__restrict__ y;
  double alpha, beta;
  const TileDesc *__restrict__ desc;            // descriptors of the tiles of this kind, in ascending tile order
  double *__restrict__ partials;                // [2*ntiles]: head fragment, tail fragment of each tile
  long long nnz; // absolute index one past the last element of the matrix (bounds the TMA size)
  int ntiles;   // tiles of this kind (persistent kernels walk [0, ntiles) with stride gridDim.x)
  int cap;      // element capacity of the shared-memory tile
  int vec_div;  // MEDIUM: lanes per row = pow2ceil(avg / vec_div)
  const unsigned int *__restrict__ row_start_bits; // direct form only
  const int *__restrict__ nz_rows;                 // direct form only
  int read_y;   // 0: beta == 0 and SPMV_B200_FLAG_BETA0_SKIP_Y
  int gather_na; // 1: x gathers use L1::no_allocate
  int stream_keep; // 1: matrix streams with the normal L2 policy instead of evict-first (small matrices)
  // staged-x form only
  const unsigned short *__restrict__ lcol; // 16-bit offsets into the staged x, indexed like col minus lcol_base
  const XDesc *__restrict__ xdesc;         // by tile id
  long long lcol_base;
  int n;        // entries of x (the last line of x may be short)
  int xcap;     // doubles of shared memory reserved for the staged x
  int m;        // rows of the plan (rowptr has m + 1 entries)
  int ring_stages, ring_stage_bytes; // staged-x ring kernel: shared-memory stages per CTA and bytes per stage
  PushArgs push;
};

// Ordering of the iterations of neighbouring GPUs inside the SpMV kernel of the fused halo loop (capi.cu,
// spmv_b200_halo_loop_*): CTAs [0, n_boundary) stream the boundary row blocks. Each of them waits until every
// neighbour's flag has reached the epoch (the neighbour's rows of the current x have arrived and it no longer reads the
// buffer about to be overwritten) before it gathers x; the last of them to finish raises this rank's flag in the
// neighbours' memory and advances the epoch. state = {epoch, finished boundary CTAs, error}.
struct HaloSync {
  unsigned int *wait[kMaxPush];   // local words written by the neighbours
  unsigned int *signal[kMaxPush]; // words in the neighbours' memory written by this rank
  int n_neigh;
  int n_boundary;
  unsigned int *state;
  unsigned long long timeout_ns; // a wait that lasts longer sets state[2] and stops signalling (0 = wait for ever)
};

struct FixupArgs {
  const int *__restrict__ split_row; // [nsplit]
  const int *__restrict__ split_t0;  // [nsplit] tile that owns the row (tail fragment)
  const int *__restrict__ split_t1;  // [nsplit] last tile holding a fragment of the row
  const double *__restrict__ partials;
  double *__restrict__ y;
  double alpha, beta;
  int nsplit;
  int read_y;
  PushArgs push;
};

} // namespace b200

struct spmv_b200_plan {
  int m = 0, n = 0;
  long long nnz = 0;      // rowptr[m] - rowptr[0]
  long long gather_active = 0, gather_lines = 0; // sampled gather-coalescing statistic (analysis.cu)
  long long sample_nnz = 0, sample_nnz_long = 0;  // non-zeros of the sampled rows / of those longer than medium_max
  long long elem_base = 0; // rowptr[0]
  long long elem_end = 0; // rowptr[m]
  const int *rowptr = nullptr;
  const int *col = nullptr;
  const double *val = nullptr;
  int T = 0, short_max = 0, medium_max = 0, vec_div = 0;
  unsigned flags = 0;
  int device = 0;
  bool uses_tma = false;
  int ntiles = 0;
  int cap = 0;
  size_t smem_bytes = 0;
  size_t persist_bytes = 0, max_window_bytes = 0; // L2 persistence for x (SPMV_B200_FLAG_L2_PERSIST_X)
  int variant_short = 0, variant_medium = 0; // kernel variants (option flag bits 8-11 / 12-15)
  int mixed_threads = 256;                   // CTA size of the MIXED kernel (option flag bits 20-21 override)
  // staged-x form (regular matrices: every row block references few runs of consecutive entries of x)
  bool xstage = false;
  unsigned short *lcol = nullptr;
  b200::XDesc *xdesc = nullptr;
  long long lcol_base = 0;
  int xlines = 0; // largest number of staged lines of any row block
  int stream_keep = 0; // matrix streams with the normal L2 policy (see policy_stream in kernels.cu)
  int comm_sms = 0;    // SMs the persistent form leaves free for a concurrently running collective
  int ring_ctas = 0, ring_stages = 0; // persistent ring form of the staged-x kernels: CTAs per SM, stages per CTA (0: off)
  // device arrays owned by the plan
  int *tile_row = nullptr;
  int *tile_elem = nullptr;
  unsigned char *tile_split = nullptr;
  int *tile_part = nullptr;
  int *tile_maxlen = nullptr;
  unsigned char *tile_kind = nullptr;
  int *list[3] = {nullptr, nullptr, nullptr};
  b200::TileDesc *desc_all = nullptr;                       // [ntiles] descriptors in tile order
  b200::TileDesc *desc[3] = {nullptr, nullptr, nullptr};    // per kind (aliases desc_all when one kind owns all tiles)
  int count[3] = {0, 0, 0};
  // direct (warp-per-tile, no shared memory) form for matrices with irregular gathers (SPMV_B200_FLAG_DIRECT / auto)
  bool direct = false;
  bool irregular = false;                  // gather-coalescing statistic > 0.5 lines per gather (set at plan creation)
  unsigned int *row_start_bits = nullptr;  // bit k (absolute element index) set iff element k is the first of its row
  int *nz_rows = nullptr;                  // ascending ids of the non-empty rows
  int n_nz_rows = 0;
  b200::TileDesc *desc_direct = nullptr;   // [ntiles] like desc_all, with head_end = number of non-empty rows < r0
  int nsplit = 0;
  int *split_rows = nullptr; // [3*nsplit]: row, t0, t1 (struct of arrays: rows | t0 | t1)
  double *partials = nullptr;
  // host copies used to launch a sub-range of tiles (pipelined host-buffer path)
  std::vector<int> h_list[3];   // tile ids per kind, ascending
  std::vector<int> h_tile_row;  // [ntiles+1]
  long long bin_rows[4] = {0, 0, 0, 0};
  long long bin_nnz[4] = {0, 0, 0, 0};
  size_t workspace_bytes = 0;
};

namespace b200 {

// analysis.cu
int analysis_gather_descs(const TileDesc *d_all, const int *h_order, int n, TileDesc **d_out, cudaStream_t stream);
int analysis_prepare(spmv_b200_plan *p, cudaStream_t stream); // reads rowptr[0], rowptr[m]
int analysis_run(spmv_b200_plan *p, cudaStream_t stream);
int analysis_xstage(spmv_b200_plan *p, cudaStream_t stream); // sets p->xstage (and lcol / xdesc / xlines) if it qualifies
int analysis_row_bins(const spmv_b200_plan *p, unsigned char *d_out, cudaStream_t stream);
int analysis_tile_col_range(const spmv_b200_plan *p, int *h_min, int *h_max, cudaStream_t stream);
int shard_bounds_run(int m, long long nnz, const int *d_rowptr, int nshards, int *h_bounds, cudaStream_t stream);
int col_block_bitmap_run(long long nnz, const int *d_col, int n, int block_shift, unsigned char *h_bitmap,
                         cudaStream_t stream);

// convert.cu
int coo_to_csr_run(int m, int n, long long nnz, const int *d_row, const int *d_col, const double *d_val,
                   int *d_rowptr_out, int *d_col_out, double *d_val_out, cudaStream_t stream);

// kernels.cu
int kernels_configure(spmv_b200_plan *p);
int kernels_configure_xs(spmv_b200_plan *p); // after analysis_xstage
void kernels_release(spmv_b200_plan *p);     // device-wide state a plan took (L2 carve-out)
int kernels_launch(const spmv_b200_plan *p, double alpha, double beta, const double *x, double *y,
                   cudaStream_t stream, const PushArgs *push = nullptr);
// tiles [tile_lo, tile_hi) only; the plan must have no split rows (their partial sums cross tile ranges)
int kernels_launch_tiles(const spmv_b200_plan *p, double alpha, double beta, const double *x, double *y, int tile_lo,
                         int tile_hi, cudaStream_t stream, const PushArgs *push = nullptr);
// fused halo loop: can the whole shard run as ONE launch with the flag protocol inside the kernel?
bool kernels_halo_single_launch_ok(const spmv_b200_plan *p);
// one iteration as one launch: `desc` lists every tile of the plan, boundary row blocks first (y = A*x, beta = 0)
int kernels_launch_halo(const spmv_b200_plan *p, const TileDesc *desc, const double *x, double *y,
                        const PushArgs *push, const HaloSync &sync, cudaStream_t stream);
// the same protocol as separate one-CTA kernels (plans the single-launch kernel does not cover)
int kernels_halo_wait(const HaloSync &sync, cudaStream_t stream);
int kernels_halo_signal(const HaloSync &sync, cudaStream_t stream);

} // namespace b200
