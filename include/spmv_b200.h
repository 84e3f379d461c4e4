/*
 * spmv_b200.h — C ABI of the B200-native fp64 CSR SpMV engine (y = alpha*A*x + beta*y).
 *
 * This is the drop-in boundary for the `cuda-b200` kernel strategy of hpcde/spmv-acc.
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns 0 on success
 * and a non-zero status otherwise (the text is available from spmv_b200_last_error()).
 * The C++ strategy launcher (src/acc/cuda-b200/cuda_b200_spmv.cpp) turns a non-zero status
 * into std::runtime_error, which the reference harness catches (benchmark/csr_spmv.hpp:52-62).
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   sparse_csr_spmv(trans, alpha, beta, h_csr_desc, d_csr_desc, dx, dy)   src/acc/api/spmv.h:20-21
 *        -> spmv_b200_csr_spmv            (stateless, plan cache behind it)
 *   sparse_spmv(trans, alpha, beta, m, n, rowptr, colindex, value, x, y)   src/acc/api/spmv.h:27-28
 *        -> spmv_b200_sparse_spmv         (same 10 arguments, nnz read on the device)
 *   csr_adaptive_plus analyze / kernel / destroy phases                    src/acc/hip-csr-adaptive-plus/csr_adaptive_plus_spmv.cpp:16-73
 *        -> spmv_b200_plan_create / spmv_b200_execute / spmv_b200_plan_destroy
 *   host_spmv caller pattern of the CLI (H2D y0, spmv, D2H y)              cli/main.cpp:99-118
 *        -> spmv_b200_hostmat_create[_device] / spmv_b200_hostmat_spmv / spmv_b200_hostmat_destroy
 *
 * All device pointers are owned by the caller and must stay valid for the life of a plan.
 * Indices are 0-based int32, values fp64, rows may be empty, columns need not be sorted.
 */
#ifndef SPMV_B200_H
#define SPMV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPMV_B200_ABI_VERSION 2

/* status codes */
#define SPMV_B200_OK 0
#define SPMV_B200_ERR_ARG 1
#define SPMV_B200_ERR_CUDA 2
#define SPMV_B200_ERR_UNSUPPORTED 3
#define SPMV_B200_ERR_TIMEOUT 4 /* fused halo loop: a neighbour GPU did not signal in time */

/* option flags */
#define SPMV_B200_FLAG_NO_TMA 1u       /* stream tiles with plain loads instead of cp.async.bulk            */
#define SPMV_B200_FLAG_BETA0_SKIP_Y 2u /* beta==0: do not read y (cuSPARSE semantics); default reads y so   */
                                       /* that NaN/Inf in y propagate exactly like cli/verification.cpp:64  */
#define SPMV_B200_FLAG_GATHER_NO_L1 8u /* gather x with L1::no_allocate                                             */
#define SPMV_B200_FLAG_NO_XSTAGE 0x40000u /* never use the staged-x form (x segments of a row block in shared memory  */
                                          /* + 16-bit local column indices; automatic for stencil / banded matrices) */
#define SPMV_B200_FLAG_DIRECT 0x40u    /* force the direct form: one warp per row block, value / colindex              */
                                       /* streamed straight into registers, no shared memory (all of the unified      */
                                       /* L1/shared array stays L1 for the x gathers); automatic for irregular gathers */
#define SPMV_B200_FLAG_NO_DIRECT 0x80u /* never use the direct form                                                   */
#define SPMV_B200_FLAG_L2_PERSIST_X 4u /* mark x as persisting in L2 for the SpMV launches (irregular gathers)      */

/* row bins (by nnz per row) and tile kinds (which per-bin kernel streams a row block) */
enum { SPMV_B200_BIN_SHORT = 0, SPMV_B200_BIN_MEDIUM = 1, SPMV_B200_BIN_LONG = 2, SPMV_B200_BIN_VERYLONG = 3 };
enum { SPMV_B200_KIND_SHORT = 0, SPMV_B200_KIND_MEDIUM = 1, SPMV_B200_KIND_MIXED = 2 };

typedef struct spmv_b200_options {
  int32_t tile_nnz;   /* items (non-zeros + rows) per row block, multiple of 256 in [256,16384]; 0 = automatic */
  int32_t short_max;  /* rows with nnz <= short_max are SHORT; 0 = default (8)                       */
  int32_t medium_max; /* rows with nnz <= medium_max are MEDIUM, multiple of 4, <= tile_nnz; 0 = 128 */
  int32_t vec_div;    /* MEDIUM kernel: lanes per row = pow2ceil(avg_nnz / vec_div); 0 = default (16) */
  uint32_t flags;     /* SPMV_B200_FLAG_* ; bits 8-11 / 12-15 select SHORT / MEDIUM kernel variants (tuning) */
} spmv_b200_options;

typedef struct spmv_b200_plan_info {
  int32_t m, n;
  int64_t nnz;
  int32_t tile_nnz, short_max, medium_max, vec_div;
  uint32_t flags;
  int32_t uses_tma;          /* 1 if tiles are streamed with cp.async.bulk                      */
  int32_t ntiles;            /* number of nnz-balanced row blocks                                */
  int32_t tiles_per_kind[3]; /* SHORT / MEDIUM / MIXED                                           */
  int32_t nsplit_rows;       /* rows whose partial sums are combined by the fix-up kernel       */
  int32_t launches_per_execute;
  int32_t direct;            /* 1 if the direct (warp-per-row-block, no shared memory) form is used */
  int64_t bin_rows[4];       /* rows per bin: short / medium / long / very long                 */
  int64_t bin_nnz[4];        /* nnz per bin                                                      */
  int64_t gather_active;     /* sampled gathers (lanes) of the gather-coalescing statistic      */
  int64_t gather_lines;      /* distinct 128-byte lines of x those gathers touch                */
  int64_t smem_bytes;        /* dynamic shared memory per CTA of the streaming kernels          */
  int64_t workspace_bytes;   /* device memory owned by the plan                                 */
  int32_t xstage;            /* 1 if the staged-x form is used (x segments in shared memory, 16-bit local indices) */
  int32_t xstage_lines;      /* largest number of 128-byte lines of x any row block stages      */
  int32_t ring_ctas;         /* staged-x form as a persistent ring: CTAs per SM (0: one row block per CTA) */
  int32_t ring_stages;       /* ... and shared-memory stages per CTA                            */
} spmv_b200_plan_info;

/* arrays that spmv_b200_plan_export can copy to the host (for bit-exact analysis checks) */
enum {
  SPMV_B200_EXPORT_TILE_ROW = 0,   /* int32 [ntiles+1]  first row owned by each tile (lower_bound of t*T in rowptr) */
  SPMV_B200_EXPORT_TILE_ELEM = 1,  /* int32 [ntiles+1]  first nnz streamed by each tile                              */
  SPMV_B200_EXPORT_TILE_SPLIT = 2, /* uint8 [ntiles+1]  1 if a long row is split at the tile's leading boundary      */
  SPMV_B200_EXPORT_TILE_KIND = 3,  /* uint8 [ntiles]    SPMV_B200_KIND_*                                             */
  SPMV_B200_EXPORT_TILE_PART = 4,  /* int32 [ntiles+1]  largest r with rowptr[r] <= t*T (merge-path partition)       */
  SPMV_B200_EXPORT_ROW_BIN = 5,    /* uint8 [m]         SPMV_B200_BIN_* per row                                      */
  SPMV_B200_EXPORT_SPLIT_ROWS = 6, /* int32 [3*nsplit]  (row, first tile, last tile) per split row                   */
  SPMV_B200_EXPORT_TILE_MAXLEN = 7, /* int32 [ntiles]    longest row owned by each tile                               */
  /* direct form only (empty otherwise): */
  SPMV_B200_EXPORT_ROW_START_BITS = 8, /* uint32 [ceil(rowptr[m]/32)] bit k set iff element k is the first of its row */
  SPMV_B200_EXPORT_NZ_ROWS = 9,        /* int32 [non-empty rows]      their ids, ascending                           */
  SPMV_B200_EXPORT_TILE_NZBASE = 10,   /* int32 [ntiles]              number of non-empty rows in front of each tile  */
  /* staged-x form only (empty otherwise): */
  SPMV_B200_EXPORT_LCOL = 11,  /* uint16 [nnz]      (rank of the 128-byte line of x among the lines the element's row block
                                                     references) * 16 + (colindex & 15)                                   */
  SPMV_B200_EXPORT_XDESC = 12  /* int32 [32*ntiles] per row block: nseg, nlines, first line of each of <= 16 runs of
                                                     consecutive lines, then their ranks as uint16 pairs, padding           */
};

/* Fused halo push: rows [row_lo[j], row_hi[j]) of y are also stored to dst[j][row] (device pointers, typically
 * peer memory of another GPU mapped through CUDA IPC, offset so that they are indexed by the local row).
 * Bit j of multicast_mask says that dst[j] is an NVLink multicast address (cuMulticast* / NVSwitch: one store is
 * replicated by the switch into the memory of every GPU bound to the multicast object); the kernels then use
 * multimem.st for it, and one destination replaces one per peer. */
#define SPMV_B200_MAX_PUSH 8
typedef struct spmv_b200_push {
  int32_t count;
  uint32_t multicast_mask;
  int32_t row_lo[SPMV_B200_MAX_PUSH];
  int32_t row_hi[SPMV_B200_MAX_PUSH];
  double *dst[SPMV_B200_MAX_PUSH];
} spmv_b200_push;

typedef struct spmv_b200_plan spmv_b200_plan;
typedef struct spmv_b200_hostmat spmv_b200_hostmat;

int spmv_b200_abi_version(void);
const char *spmv_b200_last_error(void);

/* ---- plan lifecycle: analyze once, execute many times (stream is a cudaStream_t, may be NULL) ---- */
int spmv_b200_plan_create(spmv_b200_plan **out, int32_t m, int32_t n, int64_t nnz, const int32_t *d_rowptr,
                          const int32_t *d_colidx, const double *d_val, const spmv_b200_options *opt, void *stream);
int spmv_b200_execute(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y, void *stream);
/* tiles [tile_lo, tile_hi) only (row blocks in ascending row order, see SPMV_B200_EXPORT_TILE_ROW); lets a caller
 * overlap the rows whose results other GPUs wait for with the rest. Refused for plans with split rows. */
int spmv_b200_execute_tiles(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y,
                            int32_t tile_lo, int32_t tile_hi, void *stream);
/* execute + store the rows listed in `push` into other GPUs' memory from the same kernels (no separate exchange) */
int spmv_b200_execute_push(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y,
                           const spmv_b200_push *push, void *stream);
/* the same for tiles [tile_lo, tile_hi) only: lets the rows other GPUs wait for be multiplied and pushed first */
int spmv_b200_execute_tiles_push(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y,
                                 int32_t tile_lo, int32_t tile_hi, const spmv_b200_push *push, void *stream);
/* smallest / largest column index referenced by each tile (h_min / h_max: int32 [ntiles]; INT32_MAX / -1 if empty):
 * a sharded caller uses it to find the row blocks that read entries of x owned by other GPUs */
int spmv_b200_plan_tile_col_range(spmv_b200_plan *plan, int32_t *h_min, int32_t *h_max, void *stream);
/* Multi-GPU callers that run a collective (NCCL) next to spmv_b200_execute_tiles: leave `sms` SMs free. Plans in the
 * persistent form fill every SM with CTAs that stay until the launch ends, so a communication kernel launched beside
 * them would otherwise wait for the whole SpMV. 0 (default) = use every SM. Applies to execute / execute_tiles. */
int spmv_b200_plan_set_comm_sms(spmv_b200_plan *plan, int32_t sms);
int spmv_b200_enable_peer_access(int32_t peer_device);

/* The iterated loop x <- A*x of one rank (one row shard) with the halo exchange fused into the SpMV kernels. Iteration k
 * multiplies buf[k%2] into rows [row_lo,row_hi) of buf[(k+1)%2]; rows other GPUs reference are stored straight into
 * their buffers as well (push[(k+1)%2], NVLink stores). Iterations of neighbouring GPUs are ordered by 32-bit flags in
 * each other's memory, no collective and no host synchronisation: the boundary row blocks of iteration k wait until
 * every neighbour's flag is >= k, and the last of them to finish raises this rank's flag in the neighbours' memory to
 * k+1; the interior row blocks need neither. With a plan of one row-kernel kind the whole iteration is ONE launch
 * (flag wait and flag store inside the SpMV kernel, boundary row blocks scheduled first), and runs of iterations are
 * replayed from a CUDA graph built at creation (20 iterations per graph launch, SPMV_B200_HALO_GRAPH_CHUNK in the
 * environment; the epoch the flags are compared with lives in device memory, so one graph serves every run); other plans fall back to wait kernel / boundary launches / flag kernel / interior
 * launches. n_boundary == 0: the whole shard is one launch that waits first and signals last.
 * A wait that lasts longer than SPMV_B200_FLAG_TIMEOUT_MS (environment, default 10000, 0 = for ever) marks the loop as
 * failed, stops signalling (so the failure reaches the neighbours instead of stale rows) and makes
 * spmv_b200_halo_loop_sync return SPMV_B200_ERR_TIMEOUT. */
#define SPMV_B200_MAX_RANGES 16
#define SPMV_B200_HALO_NO_GRAPH 1u     /* enqueue every iteration separately */
#define SPMV_B200_HALO_MULTI_LAUNCH 2u /* never use the single-launch kernel */
#define SPMV_B200_HALO_ALIGN_PUSH 4u   /* pushed row blocks deal their rows to the lane groups by absolute row index, so
                                          that every peer store is a full aligned 128-byte line (costs such a block a
                                          second pass now and then: worth it when the link, not the SpMV, bounds the
                                          iteration, i.e. all-gather pushes on 4 or more GPUs) */
typedef struct spmv_b200_halo_loop_desc {
  spmv_b200_plan *plan;
  double *buf[2];
  int32_t row_lo, row_hi;
  int32_t n_neigh;
  uint32_t *wait_flags[SPMV_B200_MAX_PUSH];   /* local words written by the neighbours             */
  uint32_t *signal_flags[SPMV_B200_MAX_PUSH]; /* words in the neighbours' memory written by this rank */
  spmv_b200_push push[2];                     /* by parity of the destination buffer                  */
  int32_t n_boundary, n_interior;
  int32_t boundary[2 * SPMV_B200_MAX_RANGES]; /* (tile_lo, tile_hi) pairs                             */
  int32_t interior[2 * SPMV_B200_MAX_RANGES];
  uint32_t flags;                             /* SPMV_B200_HALO_*                                     */
} spmv_b200_halo_loop_desc;
typedef struct spmv_b200_halo_loop_info {
  int64_t iterations_enqueued;
  int32_t single_launch;          /* 1: one kernel launch per iteration                       */
  int32_t uses_graph;             /* iterations replayed per CUDA graph launch (0: no graph) */
  int32_t launches_per_iteration;
  int32_t boundary_row_blocks;
} spmv_b200_halo_loop_info;
typedef struct spmv_b200_halo_loop spmv_b200_halo_loop;
int spmv_b200_halo_loop_create(spmv_b200_halo_loop **out, const spmv_b200_halo_loop_desc *desc);
int spmv_b200_halo_loop_run(spmv_b200_halo_loop *loop, int32_t iterations, void *stream); /* asynchronous */
int spmv_b200_halo_loop_sync(spmv_b200_halo_loop *loop, void *stream); /* waits; reports flag time-outs */
int spmv_b200_halo_loop_get_info(const spmv_b200_halo_loop *loop, spmv_b200_halo_loop_info *info);
int spmv_b200_halo_loop_destroy(spmv_b200_halo_loop *loop);
/* exchange buffers other GPUs (other processes) store into: allocated with cudaMalloc on the current device and
 * exported as a 64-byte CUDA IPC handle; the consumer opens the handle with ITS device current, which is what maps
 * the memory for its kernels (cudaIpcMemLazyEnablePeerAccess). peer_free / peer_close release them. */
#define SPMV_B200_IPC_HANDLE_BYTES 64
int spmv_b200_peer_alloc(void **d_ptr, int64_t bytes, uint8_t handle_out[SPMV_B200_IPC_HANDLE_BYTES]);
int spmv_b200_peer_open(const uint8_t handle[SPMV_B200_IPC_HANDLE_BYTES], void **d_ptr);
int spmv_b200_peer_close(void *d_ptr);
int spmv_b200_peer_free(void *d_ptr);
int spmv_b200_plan_destroy(spmv_b200_plan *plan);
int spmv_b200_plan_get_info(const spmv_b200_plan *plan, spmv_b200_plan_info *info);
/* copies one analysis array to host memory; returns the number of bytes through *bytes_out. dst may be NULL to query */
int spmv_b200_plan_export(spmv_b200_plan *plan, int32_t what, void *h_dst, int64_t capacity_bytes, int64_t *bytes_out);

/* ---- stateless entry points with the reference's argument lists (device pointers, null stream) ---- */
int spmv_b200_csr_spmv(int32_t trans, double alpha, double beta, int32_t m, int32_t n, int32_t nnz,
                       const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_val, const double *d_x,
                       double *d_y);
int spmv_b200_sparse_spmv(int32_t trans, double alpha, double beta, int32_t m, int32_t n, const int32_t *d_rowptr,
                          const int32_t *d_colidx, const double *d_val, const double *d_x, double *d_y);
/* The analysis behind these two calls is cached per (pointers, shape, device). Every hit re-reads a fingerprint of the
 * row pointers on the device (rowptr[0], rowptr[m], 2048 evenly spaced entries) and analyses again when it differs, so
 * an allocator that hands the same addresses to another matrix cannot resurrect a stale plan. A matrix whose row
 * pointers are rewritten in place without changing any sampled entry must still be announced with
 * spmv_b200_cache_invalidate(). SPMV_B200_CACHE_TRUST=1 in the environment skips the check. */
int spmv_b200_cache_invalidate(void); /* drop every cached plan */
int spmv_b200_cache_size(void);
int64_t spmv_b200_cache_revalidations(void); /* hits whose fingerprint did not match (the matrix was analysed again) */

/* ---- host-buffer path: matrix uploaded and analysed once, vectors copied every call ---- */
int spmv_b200_hostmat_create(spmv_b200_hostmat **out, int32_t m, int32_t n, int64_t nnz, const int32_t *h_rowptr,
                             const int32_t *h_colidx, const double *h_val, const spmv_b200_options *opt);
/* the same for a matrix that already lives in device memory (the CLI's create_device_data, cli/utils.hpp:94-116, done by
 * the caller): the arrays are not copied and must outlive the handle; nnz < 0 = read it from rowptr on the device */
int spmv_b200_hostmat_create_device(spmv_b200_hostmat **out, int32_t m, int32_t n, int64_t nnz,
                                    const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_val,
                                    const spmv_b200_options *opt);
/* y0 = h_y is copied in unless beta == 0 and the plan was created with SPMV_B200_FLAG_BETA0_SKIP_Y */
int spmv_b200_hostmat_spmv(spmv_b200_hostmat *hm, double alpha, double beta, const double *h_x, double *h_y);
/* columns [col_lo, col_hi) the matrix references: the only part of h_x that hostmat_spmv reads and copies */
int spmv_b200_hostmat_x_range(const spmv_b200_hostmat *hm, int32_t *col_lo, int32_t *col_hi);
int spmv_b200_hostmat_destroy(spmv_b200_hostmat *hm);
/* one-shot: upload, analyse, multiply, download, free */
int spmv_b200_host_spmv(double alpha, double beta, int32_t m, int32_t n, int64_t nnz, const int32_t *h_rowptr,
                        const int32_t *h_colidx, const double *h_val, const double *h_x, double *h_y);

/* ---- COO -> CSR on the device (0-based int32 row / col, fp64 values, any order, duplicates kept): entries sorted by
 * (row, col), equal pairs in input order. Device-side form of matrix_market::to_csr, cli/sparse_format.h:100-128.
 * Outputs are caller-allocated device arrays: rowptr[m+1], col[nnz], val[nnz]. ---- */
int spmv_b200_coo_to_csr(int32_t m, int32_t n, int64_t nnz, const int32_t *d_row, const int32_t *d_col,
                         const double *d_val, int32_t *d_rowptr_out, int32_t *d_col_out, double *d_val_out, void *stream);

/* ---- row sharding for multi-GPU runs: bounds[g] = lower_bound(rowptr, g*nnz/nshards), bounds[nshards] = m ---- */
int spmv_b200_shard_bounds(int32_t m, int64_t nnz, const int32_t *d_rowptr, int32_t nshards, int32_t *h_bounds,
                           void *stream);
/* blocks of x (block = 2^block_shift entries) referenced by the matrix: h_bitmap[b] = 1 if any colidx lies in block b */
int spmv_b200_col_block_bitmap(int64_t nnz, const int32_t *d_colidx, int32_t n, int32_t block_shift,
                               uint8_t *h_bitmap, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPMV_B200_H */
