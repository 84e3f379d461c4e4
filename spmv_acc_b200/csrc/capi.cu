// C ABI of libspmv_b200.so (declared in include/spmv_b200.h).
//
// Plan lifecycle mirrors the analyze / kernel / destroy phases of the reference's csr-adaptive-plus strategy
// (src/acc/hip-csr-adaptive-plus/csr_adaptive_plus_spmv.cpp:16-73); the stateless entry points mirror
// sparse_csr_spmv (src/acc/api/spmv.h:20-21) and the deprecated sparse_spmv (src/acc/api/spmv_imp.cpp:10-18).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <new>
#include <vector>

#include "internal.cuh"

namespace b200 {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  std::snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  g_last_error = buf;
  return SPMV_B200_ERR_CUDA;
}

// Tile size when the caller does not fix it: one pass of the 256 threads over the rows of a tile. A tile of T items
// (non-zeros + rows) holds T/(avg+1) rows; the row kernels give each row V = pow2ceil(avg / vec_div) lanes, so T = avg * 256 / V makes every
// lane group own exactly one row (fewest round trips to memory per CTA). Clamped to [1024, 4096], multiple of 256.
static bool irregular_gathers(const spmv_b200_plan *p) {
  // more than half a cache line per gathered element (sampled): x gathers do not coalesce across rows
  return p->gather_active > 0 && 2 * p->gather_lines > p->gather_active;
}

// Direct form (one warp per 256-item row block, no shared memory) when the x gathers do not coalesce: then the kernel is
// bound by the number of gathers in flight, every one of which holds a 128-byte line of L1, and shared memory for
// staged tiles is taken from the same 256 KB array (profiles/: gather rate against the shared-memory carve-out).
static bool auto_direct(const spmv_b200_plan *p) {
  // Measured (profiles/r1_sweep_direct_*.jsonl): power-law matrices (irregular gathers AND a quarter or more of the
  // sampled non-zeros in rows longer than medium_max) run 8 % faster in the direct form with 2048-item row blocks than
  // in the tiled MIXED kernel (C4: 1.54 ms against 1.67 ms); matrices with uniform short rows and random columns
  // (C3) are faster in the tiled MEDIUM kernel (1.84 ms against 2.2 ms) and stay there.
  return irregular_gathers(p) && 4 * p->sample_nnz_long > p->sample_nnz;
}

static int auto_tile(const spmv_b200_plan *p) {
  if (p->m <= 0 || p->nnz <= 0)
    return 2048;
  // Irregular gathers are bound by the number of L1 misses in flight, which scales with the part of the unified
  // L1/shared array left to L1: small tiles keep the shared-memory carve-out at about half of it (ncu: profiles/).
  // Measured (profiles/): uniform row lengths with random columns (x misses L2) are fastest at 1024; power-law
  // matrices, whose long rows and hot columns go together, at 2048.
  if (irregular_gathers(p))
    return (4 * p->sample_nnz_long > p->sample_nnz) ? 2048 : 1024;
  const double avg = (double)p->nnz / (double)p->m;
  int want = (int)((avg + p->vec_div - 1) / p->vec_div);
  int V = 1;
  while (V < want && V < 32)
    V <<= 1;
  const double t = (avg + 1.0) * (kThreads / V); // tiles count non-zeros and rows: avg + 1 items per row
  // rounded down (one more row than lane groups would cost a second pass), with 5% slack so that an average just
  // below a whole number of nnz per row (matrix boundary effects, e.g. 4.999 for the 5-point stencil) still counts
  int T = (int)(t / 256.0 + 0.05) * 256;
  if (T < 1024)
    T = 1024;
  if (T > 4096)
    T = 4096;
  return T;
}

static int free_plan_arrays(spmv_b200_plan *p) {
  for (int k = 0; k < 3; ++k)
    if (p->desc[k] == p->desc_all)
      p->desc[k] = nullptr; // alias, freed once below
  void *ptrs[] = {p->tile_row, p->tile_elem, p->tile_split, p->tile_part,  p->tile_maxlen, p->tile_kind, p->list[0],
                  p->list[1],  p->list[2],   p->split_rows, p->partials,   p->desc_all,    p->desc[0],   p->desc[1],
                  p->desc[2],  p->row_start_bits, p->nz_rows, p->desc_direct};
  int rc = SPMV_B200_OK;
  for (void *q : ptrs)
    if (q && cudaFree(q) != cudaSuccess)
      rc = SPMV_B200_ERR_CUDA;
  return rc;
}

static void reset_plan_arrays(spmv_b200_plan *p) {
  p->tile_row = p->tile_elem = p->tile_part = p->tile_maxlen = nullptr;
  p->tile_split = p->tile_kind = nullptr;
  for (int k = 0; k < 3; ++k) {
    p->list[k] = nullptr;
    p->desc[k] = nullptr;
    p->count[k] = 0;
    p->h_list[k].clear();
  }
  p->desc_all = p->desc_direct = nullptr;
  p->split_rows = nullptr;
  p->partials = nullptr;
  p->row_start_bits = nullptr;
  p->nz_rows = nullptr;
  p->nsplit = 0;
  p->n_nz_rows = 0;
  p->h_tile_row.clear();
}

} // namespace b200

using namespace b200;

extern "C" {

int spmv_b200_abi_version(void) { return SPMV_B200_ABI_VERSION; }

const char *spmv_b200_last_error(void) { return g_last_error.c_str(); }

int spmv_b200_plan_create(spmv_b200_plan **out, int32_t m, int32_t n, int64_t nnz, const int32_t *d_rowptr,
                          const int32_t *d_colidx, const double *d_val, const spmv_b200_options *opt, void *stream) {
  if (!out) {
    set_error("plan_create: out is NULL");
    return SPMV_B200_ERR_ARG;
  }
  *out = nullptr;
  if (m < 0 || n < 0 || nnz > 0x7fffffffLL) {
    set_error("plan_create: m, n must be >= 0 and nnz < 2^31 (int32 indices)");
    return SPMV_B200_ERR_ARG;
  }
  if (m > 0 && !d_rowptr) {
    set_error("plan_create: rowptr is NULL");
    return SPMV_B200_ERR_ARG;
  }
  spmv_b200_plan *p = new (std::nothrow) spmv_b200_plan();
  if (!p) {
    set_error("plan_create: out of host memory");
    return SPMV_B200_ERR_ARG;
  }
  p->m = m;
  p->n = n;
  p->nnz = nnz; // < 0: take it from rowptr
  p->rowptr = d_rowptr;
  p->col = d_colidx;
  p->val = d_val;
  p->T = (opt && opt->tile_nnz) ? opt->tile_nnz : 2048; // provisional; the automatic choice needs nnz (below)
  p->short_max = (opt && opt->short_max) ? opt->short_max : kDefaultShort;
  p->medium_max = (opt && opt->medium_max) ? opt->medium_max : kDefaultMedium;
  p->vec_div = (opt && opt->vec_div) ? opt->vec_div : kDefaultVecDiv;
  p->flags = opt ? opt->flags : 0u;
  if (p->T < 256 || p->T > 16384 || (p->T % 256) != 0 || p->medium_max < 4 || (p->medium_max % 4) != 0 ||
      p->medium_max > p->T || p->short_max < 1 || p->short_max > p->medium_max || p->vec_div < 1) {
    set_error("plan_create: invalid options (tile_nnz multiple of 256 in [256,16384]; medium_max multiple of 4 in "
              "[4,tile_nnz]; 1 <= short_max <= medium_max; vec_div >= 1)");
    delete p;
    return SPMV_B200_ERR_ARG;
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(d_val) & 15u) == 0) &&
                       ((reinterpret_cast<uintptr_t>(d_colidx) & 15u) == 0);
  p->uses_tma = aligned && !(p->flags & SPMV_B200_FLAG_NO_TMA);
  int rc = analysis_prepare(p, static_cast<cudaStream_t>(stream));
  if (rc == SPMV_B200_OK && !(opt && opt->vec_div) && irregular_gathers(p))
    p->vec_div = 8;
  if (rc == SPMV_B200_OK)
    p->irregular = irregular_gathers(p);
  // direct form: forced by flag; automatic choice below (auto_direct)
  if (rc == SPMV_B200_OK)
    p->direct = !(p->flags & SPMV_B200_FLAG_NO_DIRECT) && p->nnz > 0 &&
                ((p->flags & SPMV_B200_FLAG_DIRECT) || auto_direct(p));
  if (rc == SPMV_B200_OK && !(opt && opt->tile_nnz))
    p->T = p->direct ? 2048 : auto_tile(p);
  if (p->T < p->medium_max)
    p->T = (p->medium_max + 255) / 256 * 256;
  if (rc == SPMV_B200_OK && !((p->flags >> 8) & 0xf) && p->m > 0 && (double)p->nnz / p->m > 6.0)
    p->flags |= 1u << 8; // SHORT rows of 7+ nnz: gather 8 per round instead of 6 (variant table in kernels.cu)
  if (rc == SPMV_B200_OK)
    rc = kernels_configure(p);
  if (rc == SPMV_B200_OK)
    rc = analysis_run(p, static_cast<cudaStream_t>(stream));
  // Automatic tile size, second look: a tile that owns more rows than the row kernels have lane groups costs the CTA a
  // second pass over its rows. T was derived from the average row; if rows a little shorter than the average (domain
  // boundaries of a stencil) push more than 15 % of the tiles over the limit, one step down is faster (measured on an
  // interior z-slab of the 27-point stencil, 1/8 of 384^3: 21.8 % of the tiles over, 0.447 ms with T = 3584, 0.407 ms
  // with T = 3328).
  if (rc == SPMV_B200_OK && !(opt && opt->tile_nnz) && !p->direct && p->T > 1024 && p->ntiles > 0 &&
      p->count[SPMV_B200_KIND_MIXED] == 0) {
    const double avg = (double)p->nnz / (double)p->m;
    const int want = (int)((avg + p->vec_div - 1) / p->vec_div);
    int V = 1;
    while (V < want && V < 32)
      V <<= 1;
    const int G = kThreads / V;
    long long over = 0;
    for (int t = 0; t < p->ntiles; ++t)
      over += (p->h_tile_row[t + 1] - p->h_tile_row[t]) > G ? 1 : 0;
    if (20 * over > 3 * (long long)p->ntiles) { // > 15 %: C2 (12.5 % at T = 1536, still the faster choice) stays
      free_plan_arrays(p);
      reset_plan_arrays(p);
      p->T -= 256;
      rc = kernels_configure(p);
      if (rc == SPMV_B200_OK)
        rc = analysis_run(p, static_cast<cudaStream_t>(stream));
    }
  }
  if (rc != SPMV_B200_OK) {
    free_plan_arrays(p);
    delete p;
    return rc;
  }
  if (p->nnz > 0 && (!d_colidx || !d_val)) {
    set_error("plan_create: colidx / val is NULL but the matrix has non-zeros");
    free_plan_arrays(p);
    delete p;
    return SPMV_B200_ERR_ARG;
  }
  *out = p;
  return SPMV_B200_OK;
}

int spmv_b200_execute(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y, void *stream) {
  if (!plan) {
    set_error("execute: plan is NULL");
    return SPMV_B200_ERR_ARG;
  }
  if ((plan->m > 0 && !d_y) || (plan->nnz > 0 && !d_x)) {
    set_error("execute: x or y is NULL");
    return SPMV_B200_ERR_ARG;
  }
  return kernels_launch(plan, alpha, beta, d_x, d_y, static_cast<cudaStream_t>(stream));
}

int spmv_b200_execute_tiles(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y,
                            int32_t tile_lo, int32_t tile_hi, void *stream) {
  if (!plan) {
    set_error("execute_tiles: plan is NULL");
    return SPMV_B200_ERR_ARG;
  }
  if (tile_lo < 0 || tile_hi > plan->ntiles || tile_lo > tile_hi) {
    set_error("execute_tiles: tile range out of bounds");
    return SPMV_B200_ERR_ARG;
  }
  if ((plan->m > 0 && !d_y) || (plan->nnz > 0 && !d_x)) {
    set_error("execute_tiles: x or y is NULL");
    return SPMV_B200_ERR_ARG;
  }
  if (tile_lo == tile_hi)
    return SPMV_B200_OK;
  return kernels_launch_tiles(plan, alpha, beta, d_x, d_y, tile_lo, tile_hi, static_cast<cudaStream_t>(stream));
}

static int convert_push(const spmv_b200_plan *plan, const spmv_b200_push *push, PushArgs *pa, const char *who) {
  if (push->count < 0 || push->count > SPMV_B200_MAX_PUSH) {
    set_error(std::string(who) + ": push.count out of range");
    return SPMV_B200_ERR_ARG;
  }
  pa->count = push->count;
  for (int j = 0; j < kMaxPush; ++j) {
    pa->row_lo[j] = j < push->count ? push->row_lo[j] : 0;
    pa->row_hi[j] = j < push->count ? push->row_hi[j] : 0;
    pa->dst[j] = j < push->count ? push->dst[j] : nullptr;
    if (j < push->count &&
        (!pa->dst[j] || pa->row_lo[j] < 0 || pa->row_hi[j] > plan->m || pa->row_lo[j] > pa->row_hi[j])) {
      set_error(std::string(who) + ": bad push range");
      return SPMV_B200_ERR_ARG;
    }
  }
  return SPMV_B200_OK;
}

int spmv_b200_execute_push(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y,
                           const spmv_b200_push *push, void *stream) {
  if (!plan || !push) {
    set_error("execute_push: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  if ((plan->m > 0 && !d_y) || (plan->nnz > 0 && !d_x)) {
    set_error("execute_push: x or y is NULL");
    return SPMV_B200_ERR_ARG;
  }
  PushArgs pa;
  if (int rc = convert_push(plan, push, &pa, "execute_push"))
    return rc;
  return kernels_launch(plan, alpha, beta, d_x, d_y, static_cast<cudaStream_t>(stream), &pa);
}

int spmv_b200_execute_tiles_push(spmv_b200_plan *plan, double alpha, double beta, const double *d_x, double *d_y,
                                 int32_t tile_lo, int32_t tile_hi, const spmv_b200_push *push, void *stream) {
  if (!plan || !push) {
    set_error("execute_tiles_push: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  if (tile_lo < 0 || tile_hi > plan->ntiles || tile_lo > tile_hi) {
    set_error("execute_tiles_push: tile range out of bounds");
    return SPMV_B200_ERR_ARG;
  }
  if ((plan->m > 0 && !d_y) || (plan->nnz > 0 && !d_x)) {
    set_error("execute_tiles_push: x or y is NULL");
    return SPMV_B200_ERR_ARG;
  }
  PushArgs pa;
  if (int rc = convert_push(plan, push, &pa, "execute_tiles_push"))
    return rc;
  if (tile_lo == tile_hi)
    return SPMV_B200_OK;
  return kernels_launch_tiles(plan, alpha, beta, d_x, d_y, tile_lo, tile_hi, static_cast<cudaStream_t>(stream), &pa);
}

int spmv_b200_plan_tile_col_range(spmv_b200_plan *plan, int32_t *h_min, int32_t *h_max, void *stream) {
  if (!plan || !h_min || !h_max) {
    set_error("plan_tile_col_range: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  return analysis_tile_col_range(plan, h_min, h_max, static_cast<cudaStream_t>(stream));
}

// stream memory operations of the driver API, fetched at run time (no link-time dependency on libcuda)
namespace {
typedef int (*StreamMemOp32)(void *stream, unsigned long long addr, unsigned int value, unsigned int flags);
StreamMemOp32 driver_fn(const char *name) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<StreamMemOp32>(fn);
}
} // namespace

// The flag may live in another GPU's memory (IPC mapping): a one-thread kernel with a system-scope fence is the
// portable way to publish it after the stores of the preceding kernels in the stream.
__global__ void k_write_flag(uint32_t *flag, uint32_t value) {
  __threadfence_system();
  *reinterpret_cast<volatile uint32_t *>(flag) = value;
}

int spmv_b200_stream_write_flag(void *stream, uint32_t *d_flag, uint32_t value) {
  if (!d_flag) {
    set_error("stream_write_flag: flag is NULL");
    return SPMV_B200_ERR_ARG;
  }
  k_write_flag<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(d_flag, value);
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

struct FlagList {
  uint32_t *p[SPMV_B200_MAX_PUSH];
  int n;
};
__global__ void k_write_flags(FlagList f, uint32_t value) {
  __threadfence_system();
  if ((int)threadIdx.x < f.n)
    *reinterpret_cast<volatile uint32_t *>(f.p[threadIdx.x]) = value;
}

int spmv_b200_stream_write_flags(void *stream, uint32_t *const *d_flags, int32_t count, uint32_t value) {
  if (count < 0 || count > SPMV_B200_MAX_PUSH || (count > 0 && !d_flags)) {
    set_error("stream_write_flags: bad argument");
    return SPMV_B200_ERR_ARG;
  }
  if (count == 0)
    return SPMV_B200_OK;
  FlagList f;
  f.n = count;
  for (int i = 0; i < SPMV_B200_MAX_PUSH; ++i)
    f.p[i] = i < count ? d_flags[i] : nullptr;
  k_write_flags<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, value);
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

// Waiting inside a one-thread kernel instead of a stream memory operation: the flags are local memory written by the
// neighbours over NVLink. The spin is bounded (about 10 s) so that a protocol error cannot hang the device.
__global__ void k_wait_flags(FlagList f, uint32_t value) {
  if ((int)threadIdx.x < f.n) {
    const volatile uint32_t *p = reinterpret_cast<volatile uint32_t *>(f.p[threadIdx.x]);
    const long long t0 = clock64();
    while (*p < value && clock64() - t0 < 20000000000LL) {
    }
  }
  __threadfence_system();
}

int spmv_b200_stream_wait_flags(void *stream, uint32_t *const *d_flags, int32_t count, uint32_t value) {
  if (count < 0 || count > SPMV_B200_MAX_PUSH || (count > 0 && !d_flags)) {
    set_error("stream_wait_flags: bad argument");
    return SPMV_B200_ERR_ARG;
  }
  if (count == 0)
    return SPMV_B200_OK;
  static const bool use_memop = [] {
    const char *e = getenv("SPMV_B200_FLAG_WAIT");
    return e && std::string(e) == "memop";
  }();
  if (use_memop) {
    for (int i = 0; i < count; ++i)
      if (int rc = spmv_b200_stream_wait_flag(stream, d_flags[i], value))
        return rc;
    return SPMV_B200_OK;
  }
  FlagList f;
  f.n = count;
  for (int i = 0; i < SPMV_B200_MAX_PUSH; ++i)
    f.p[i] = i < count ? d_flags[i] : nullptr;
  k_wait_flags<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, value);
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

int spmv_b200_stream_wait_flag(void *stream, uint32_t *d_flag, uint32_t value) {
  static StreamMemOp32 fn = driver_fn("cuStreamWaitValue32");
  if (!fn) {
    set_error("cuStreamWaitValue32 is not available");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  const int rc = fn(stream, (unsigned long long)(uintptr_t)d_flag, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
  if (rc != 0) {
    set_error("cuStreamWaitValue32 failed with CUresult " + std::to_string(rc));
    return SPMV_B200_ERR_CUDA;
  }
  return SPMV_B200_OK;
}

int spmv_b200_halo_loop_run(const spmv_b200_halo_loop_desc *d, int32_t first_iteration, int32_t iterations,
                            void *stream) {
  if (!d || !d->plan || !d->buf[0] || !d->buf[1] || first_iteration < 0 || iterations < 0) {
    set_error("halo_loop_run: bad argument");
    return SPMV_B200_ERR_ARG;
  }
  if (d->n_neigh < 0 || d->n_neigh > SPMV_B200_MAX_PUSH || d->n_boundary < 0 || d->n_boundary > SPMV_B200_MAX_RANGES ||
      d->n_interior < 0 || d->n_interior > SPMV_B200_MAX_RANGES || d->row_lo < 0 || d->row_hi < d->row_lo ||
      d->row_hi - d->row_lo != d->plan->m) {
    set_error("halo_loop_run: inconsistent descriptor");
    return SPMV_B200_ERR_ARG;
  }
  for (int32_t i = 0; i < iterations; ++i) {
    const int32_t k = first_iteration + i;
    int rc;
    if (k > 0 && (rc = spmv_b200_stream_wait_flags(stream, d->wait_flags, d->n_neigh, (uint32_t)k)))
      return rc;
    const double *src = d->buf[k & 1];
    double *ys = d->buf[(k + 1) & 1] + d->row_lo;
    const spmv_b200_push *push = &d->push[(k + 1) & 1];
    if (d->n_boundary > 0) {
      for (int r = 0; r < d->n_boundary; ++r)
        if ((rc = spmv_b200_execute_tiles_push(d->plan, 1.0, 0.0, src, ys, d->boundary[2 * r], d->boundary[2 * r + 1],
                                               push, stream)))
          return rc;
      if ((rc = spmv_b200_stream_write_flags(stream, d->signal_flags, d->n_neigh, (uint32_t)(k + 1))))
        return rc;
      for (int r = 0; r < d->n_interior; ++r)
        if ((rc = spmv_b200_execute_tiles(d->plan, 1.0, 0.0, src, ys, d->interior[2 * r], d->interior[2 * r + 1],
                                          stream)))
          return rc;
    } else {
      if ((rc = spmv_b200_execute_push(d->plan, 1.0, 0.0, src, ys, push, stream)))
        return rc;
      if ((rc = spmv_b200_stream_write_flags(stream, d->signal_flags, d->n_neigh, (uint32_t)(k + 1))))
        return rc;
    }
  }
  return SPMV_B200_OK;
}

int spmv_b200_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev)
    return SPMV_B200_OK;
  int can = 0;
  B200_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) {
    set_error("enable_peer_access: the devices have no peer path");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return SPMV_B200_OK;
  }
  B200_CUDA(e);
  return SPMV_B200_OK;
}

int spmv_b200_peer_alloc(void **d_ptr, int64_t bytes, uint8_t handle_out[SPMV_B200_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == SPMV_B200_IPC_HANDLE_BYTES, "IPC handle size");
  if (!d_ptr || !handle_out || bytes <= 0) {
    set_error("peer_alloc: bad argument");
    return SPMV_B200_ERR_ARG;
  }
  void *p = nullptr;
  B200_CUDA(cudaMalloc(&p, static_cast<size_t>(bytes)));
  cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess)
    e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    B200_CUDA(e);
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *d_ptr = p;
  return SPMV_B200_OK;
}

int spmv_b200_peer_open(const uint8_t handle[SPMV_B200_IPC_HANDLE_BYTES], void **d_ptr) {
  if (!d_ptr || !handle) {
    set_error("peer_open: bad argument");
    return SPMV_B200_ERR_ARG;
  }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void *p = nullptr;
  B200_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ptr = p;
  return SPMV_B200_OK;
}

int spmv_b200_peer_close(void *d_ptr) {
  if (d_ptr)
    B200_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return SPMV_B200_OK;
}

int spmv_b200_peer_free(void *d_ptr) {
  if (d_ptr)
    B200_CUDA(cudaFree(d_ptr));
  return SPMV_B200_OK;
}

int spmv_b200_plan_destroy(spmv_b200_plan *plan) {
  if (!plan)
    return SPMV_B200_OK;
  const int rc = free_plan_arrays(plan);
  delete plan;
  if (rc != SPMV_B200_OK)
    set_error("plan_destroy: cudaFree failed");
  return rc;
}

int spmv_b200_plan_get_info(const spmv_b200_plan *p, spmv_b200_plan_info *info) {
  if (!p || !info) {
    set_error("plan_get_info: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  std::memset(info, 0, sizeof(*info));
  info->m = p->m;
  info->n = p->n;
  info->nnz = p->nnz;
  info->tile_nnz = p->T;
  info->short_max = p->short_max;
  info->medium_max = p->medium_max;
  info->vec_div = p->vec_div;
  info->flags = p->flags;
  info->uses_tma = p->uses_tma ? 1 : 0;
  info->ntiles = p->ntiles;
  int launches = 0;
  for (int k = 0; k < 3; ++k) {
    info->tiles_per_kind[k] = p->count[k];
    launches += p->count[k] > 0 ? 1 : 0;
  }
  launches += p->nsplit > 0 ? 1 : 0;
  info->nsplit_rows = p->nsplit;
  if (p->direct)
    launches = (p->ntiles > 0 ? 1 : 0) + (p->nsplit > 0 ? 1 : 0);
  info->launches_per_execute = launches;
  info->direct = p->direct ? 1 : 0;
  for (int b = 0; b < 4; ++b) {
    info->bin_rows[b] = p->bin_rows[b];
    info->bin_nnz[b] = p->bin_nnz[b];
  }
  info->gather_active = p->gather_active;
  info->gather_lines = p->gather_lines;
  info->smem_bytes = (int64_t)p->smem_bytes;
  info->workspace_bytes = (int64_t)p->workspace_bytes;
  return SPMV_B200_OK;
}

int spmv_b200_plan_export(spmv_b200_plan *p, int32_t what, void *h_dst, int64_t capacity_bytes, int64_t *bytes_out) {
  if (!p) {
    set_error("plan_export: plan is NULL");
    return SPMV_B200_ERR_ARG;
  }
  const void *src = nullptr;
  int64_t bytes = 0;
  const int64_t nt = p->ntiles;
  switch (what) {
  case SPMV_B200_EXPORT_TILE_ROW:
    src = p->tile_row;
    bytes = p->m > 0 ? 4 * (nt + 1) : 0;
    break;
  case SPMV_B200_EXPORT_TILE_ELEM:
    src = p->tile_elem;
    bytes = p->m > 0 ? 4 * (nt + 1) : 0;
    break;
  case SPMV_B200_EXPORT_TILE_SPLIT:
    src = p->tile_split;
    bytes = p->m > 0 ? (nt + 1) : 0;
    break;
  case SPMV_B200_EXPORT_TILE_KIND:
    src = p->tile_kind;
    bytes = nt;
    break;
  case SPMV_B200_EXPORT_TILE_PART:
    src = p->tile_part;
    bytes = p->m > 0 ? 4 * (nt + 1) : 0;
    break;
  case SPMV_B200_EXPORT_TILE_MAXLEN:
    src = p->tile_maxlen;
    bytes = 4 * nt;
    break;
  case SPMV_B200_EXPORT_SPLIT_ROWS:
    src = p->split_rows;
    bytes = 12 * (int64_t)p->nsplit;
    break;
  case SPMV_B200_EXPORT_ROW_BIN:
    bytes = p->m;
    break;
  case SPMV_B200_EXPORT_ROW_START_BITS:
    src = p->row_start_bits;
    bytes = p->direct ? 4 * ((p->elem_end + 31) / 32) : 0;
    break;
  case SPMV_B200_EXPORT_NZ_ROWS:
    src = p->nz_rows;
    bytes = p->direct ? 4 * (int64_t)p->n_nz_rows : 0;
    break;
  case SPMV_B200_EXPORT_TILE_NZBASE:
    bytes = p->direct ? 4 * nt : 0;
    break;
  default:
    set_error("plan_export: unknown array id");
    return SPMV_B200_ERR_ARG;
  }
  if (bytes_out)
    *bytes_out = bytes;
  if (!h_dst || bytes == 0)
    return SPMV_B200_OK;
  if (capacity_bytes < bytes) {
    set_error("plan_export: destination too small");
    return SPMV_B200_ERR_ARG;
  }
  if (what == SPMV_B200_EXPORT_ROW_BIN) {
    unsigned char *d_tmp = nullptr;
    B200_CUDA(cudaMalloc(&d_tmp, (size_t)bytes));
    int rc = analysis_row_bins(p, d_tmp, nullptr);
    if (rc == SPMV_B200_OK && cudaMemcpy(h_dst, d_tmp, (size_t)bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = cuda_fail(cudaGetLastError(), "cudaMemcpy(row bins)", __FILE__, __LINE__);
    cudaFree(d_tmp);
    return rc;
  }
  if (what == SPMV_B200_EXPORT_TILE_NZBASE) { // field head_end of the direct descriptors
    std::vector<TileDesc> h((size_t)nt);
    B200_CUDA(cudaMemcpy(h.data(), p->desc_direct, sizeof(TileDesc) * (size_t)nt, cudaMemcpyDeviceToHost));
    for (int64_t t = 0; t < nt; ++t)
      static_cast<int32_t *>(h_dst)[t] = h[(size_t)t].head_end;
    return SPMV_B200_OK;
  }
  B200_CUDA(cudaMemcpy(h_dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
  return SPMV_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// stateless entry points: a small plan cache keyed on the device pointers and the shape
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct CacheKey {
  const void *rowptr, *col, *val;
  int m, n, device;
  bool operator==(const CacheKey &o) const {
    return rowptr == o.rowptr && col == o.col && val == o.val && m == o.m && n == o.n && device == o.device;
  }
};
struct CacheEntry {
  CacheKey key;
  spmv_b200_plan *plan;
};
std::mutex g_cache_mutex;
std::list<CacheEntry> g_cache; // most recently used first
constexpr size_t kCacheCapacity = 16;

int cached_plan(int m, int n, long long nnz, const int *rowptr, const int *col, const double *val,
                spmv_b200_plan **out) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  const CacheKey key{rowptr, col, val, m, n, dev};
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  for (auto it = g_cache.begin(); it != g_cache.end(); ++it) {
    if (it->key == key && (nnz < 0 || it->plan->nnz == nnz)) {
      g_cache.splice(g_cache.begin(), g_cache, it);
      *out = g_cache.front().plan;
      return SPMV_B200_OK;
    }
  }
  spmv_b200_plan *p = nullptr;
  const int rc = spmv_b200_plan_create(&p, m, n, nnz, rowptr, col, val, nullptr, nullptr);
  if (rc != SPMV_B200_OK)
    return rc;
  g_cache.push_front(CacheEntry{key, p});
  while (g_cache.size() > kCacheCapacity) {
    // plans may still have kernels in flight on the null stream; cudaFree synchronises implicitly
    spmv_b200_plan_destroy(g_cache.back().plan);
    g_cache.pop_back();
  }
  *out = p;
  return SPMV_B200_OK;
}
} // namespace

int spmv_b200_csr_spmv(int32_t trans, double alpha, double beta, int32_t m, int32_t n, int32_t nnz,
                       const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_val, const double *d_x,
                       double *d_y) {
  if (trans != 0) { // src/acc/api/types.h:8 — only operation_none is supported by the reference as well
    set_error("csr_spmv: only operation_none (trans = 0) is supported");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  spmv_b200_plan *p = nullptr;
  const int rc = cached_plan(m, n, nnz, d_rowptr, d_colidx, d_val, &p);
  if (rc != SPMV_B200_OK)
    return rc;
  return spmv_b200_execute(p, alpha, beta, d_x, d_y, nullptr);
}

int spmv_b200_sparse_spmv(int32_t trans, double alpha, double beta, int32_t m, int32_t n, const int32_t *d_rowptr,
                          const int32_t *d_colidx, const double *d_val, const double *d_x, double *d_y) {
  // src/acc/api/spmv_imp.cpp:14 reads rowptr[hm] on the host; here nnz comes from the device inside the analysis
  if (trans != 0) {
    set_error("sparse_spmv: only operation_none (trans = 0) is supported");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  spmv_b200_plan *p = nullptr;
  const int rc = cached_plan(m, n, -1, d_rowptr, d_colidx, d_val, &p);
  if (rc != SPMV_B200_OK)
    return rc;
  return spmv_b200_execute(p, alpha, beta, d_x, d_y, nullptr);
}

int spmv_b200_cache_invalidate(void) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  int rc = SPMV_B200_OK;
  for (auto &e : g_cache)
    if (spmv_b200_plan_destroy(e.plan) != SPMV_B200_OK)
      rc = SPMV_B200_ERR_CUDA;
  g_cache.clear();
  return rc;
}

int spmv_b200_cache_size(void) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  return (int)g_cache.size();
}

// ---------------------------------------------------------------------------------------------------------------
// host-buffer path (the CLI's pattern: matrix uploaded once, y0 copied in and y copied out around every call,
// cli/utils.hpp:94-116 and cli/main.cpp:99-118)
// ---------------------------------------------------------------------------------------------------------------
struct spmv_b200_hostmat {
  int m = 0, n = 0;
  long long nnz = 0;
  int *d_rowptr = nullptr;
  int *d_col = nullptr;
  double *d_val = nullptr;
  double *d_x = nullptr;
  double *d_y = nullptr;
  spmv_b200_plan *plan = nullptr;
  cudaStream_t stream = nullptr;                 // compute (and the non-pipelined path)
  cudaStream_t s_in = nullptr, s_out = nullptr;  // host->device and device->host copy streams (pipelined path)
  cudaEvent_t ev_x = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;       // per chunk: y0 chunk arrived / chunk computed
  // columns the matrix references: only x[col_lo, col_hi) is copied to the device; per tile for the pipelined path
  int col_lo = 0, col_hi = 0;
  std::vector<int> tile_cmax;
};

constexpr int kHostChunksMax = 64;
static int host_chunks() { // row chunks of the pipelined host-buffer path (SPMV_B200_HOST_CHUNKS overrides)
  const char *e = getenv("SPMV_B200_HOST_CHUNKS");
  const int c = e ? atoi(e) : 8;
  return c < 1 ? 1 : (c > kHostChunksMax ? kHostChunksMax : c);
}

int spmv_b200_hostmat_destroy(spmv_b200_hostmat *hm) {
  if (!hm)
    return SPMV_B200_OK;
  if (hm->stream)
    cudaStreamSynchronize(hm->stream);
  spmv_b200_plan_destroy(hm->plan);
  cudaFree(hm->d_rowptr);
  cudaFree(hm->d_col);
  cudaFree(hm->d_val);
  cudaFree(hm->d_x);
  cudaFree(hm->d_y);
  for (cudaEvent_t e : hm->ev_in)
    cudaEventDestroy(e);
  for (cudaEvent_t e : hm->ev_done)
    cudaEventDestroy(e);
  if (hm->ev_x)
    cudaEventDestroy(hm->ev_x);
  for (cudaStream_t st : {hm->stream, hm->s_in, hm->s_out})
    if (st)
      cudaStreamDestroy(st);
  delete hm;
  return SPMV_B200_OK;
}

static int hostmat_build(spmv_b200_hostmat *hm, const int32_t *h_rowptr, const int32_t *h_colidx, const double *h_val,
                         const spmv_b200_options *opt) {
  const size_t m = (size_t)hm->m, n = (size_t)hm->n, nnz = (size_t)hm->nnz;
  B200_CUDA(cudaStreamCreateWithFlags(&hm->stream, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&hm->s_in, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&hm->s_out, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreateWithFlags(&hm->ev_x, cudaEventDisableTiming));
  hm->ev_in.resize(kHostChunksMax);
  hm->ev_done.resize(kHostChunksMax);
  for (int c = 0; c < kHostChunksMax; ++c) {
    B200_CUDA(cudaEventCreateWithFlags(&hm->ev_in[c], cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&hm->ev_done[c], cudaEventDisableTiming));
  }
  B200_CUDA(cudaMalloc(&hm->d_rowptr, sizeof(int) * (m + 1)));
  B200_CUDA(cudaMalloc(&hm->d_col, sizeof(int) * (nnz ? nnz : 1)));
  B200_CUDA(cudaMalloc(&hm->d_val, sizeof(double) * (nnz ? nnz : 1)));
  B200_CUDA(cudaMalloc(&hm->d_x, sizeof(double) * (n ? n : 1)));
  B200_CUDA(cudaMalloc(&hm->d_y, sizeof(double) * (m ? m : 1)));
  B200_CUDA(cudaMemcpyAsync(hm->d_rowptr, h_rowptr, sizeof(int) * (m + 1), cudaMemcpyHostToDevice, hm->stream));
  if (nnz) {
    B200_CUDA(cudaMemcpyAsync(hm->d_col, h_colidx, sizeof(int) * nnz, cudaMemcpyHostToDevice, hm->stream));
    B200_CUDA(cudaMemcpyAsync(hm->d_val, h_val, sizeof(double) * nnz, cudaMemcpyHostToDevice, hm->stream));
  }
  int rc = spmv_b200_plan_create(&hm->plan, hm->m, hm->n, hm->nnz, hm->d_rowptr, hm->d_col, hm->d_val, opt, hm->stream);
  if (rc != SPMV_B200_OK)
    return rc;
  hm->col_lo = 0;
  hm->col_hi = hm->n;
  if (hm->plan->ntiles > 0 && nnz > 0) {
    std::vector<int> cmin((size_t)hm->plan->ntiles);
    hm->tile_cmax.resize((size_t)hm->plan->ntiles);
    if ((rc = analysis_tile_col_range(hm->plan, cmin.data(), hm->tile_cmax.data(), hm->stream)))
      return rc;
    int lo = 0x7fffffff, hi = -1;
    for (int t = 0; t < hm->plan->ntiles; ++t) {
      lo = cmin[(size_t)t] < lo ? cmin[(size_t)t] : lo;
      hi = hm->tile_cmax[(size_t)t] > hi ? hm->tile_cmax[(size_t)t] : hi;
    }
    if (hi >= lo && lo >= 0 && hi < hm->n) {
      hm->col_lo = lo;
      hm->col_hi = hi + 1;
    }
  }
  return SPMV_B200_OK;
}

int spmv_b200_hostmat_x_range(const spmv_b200_hostmat *hm, int32_t *col_lo, int32_t *col_hi) {
  if (!hm || !col_lo || !col_hi) {
    set_error("hostmat_x_range: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  *col_lo = hm->col_lo;
  *col_hi = hm->col_hi;
  return SPMV_B200_OK;
}

int spmv_b200_hostmat_create(spmv_b200_hostmat **out, int32_t m, int32_t n, int64_t nnz, const int32_t *h_rowptr,
                             const int32_t *h_colidx, const double *h_val, const spmv_b200_options *opt) {
  if (!out || m < 0 || n < 0 || nnz < 0 || nnz > 0x7fffffffLL || !h_rowptr || (nnz > 0 && (!h_colidx || !h_val))) {
    set_error("hostmat_create: invalid argument");
    return SPMV_B200_ERR_ARG;
  }
  *out = nullptr;
  spmv_b200_hostmat *hm = new (std::nothrow) spmv_b200_hostmat();
  if (!hm) {
    set_error("hostmat_create: out of host memory");
    return SPMV_B200_ERR_ARG;
  }
  hm->m = m;
  hm->n = n;
  hm->nnz = nnz;
  const int rc = hostmat_build(hm, h_rowptr, h_colidx, h_val, opt);
  if (rc != SPMV_B200_OK) {
    const std::string keep = g_last_error;
    spmv_b200_hostmat_destroy(hm);
    g_last_error = keep;
    return rc;
  }
  *out = hm;
  return SPMV_B200_OK;
}

int spmv_b200_hostmat_spmv(spmv_b200_hostmat *hm, double alpha, double beta, const double *h_x, double *h_y) {
  if (!hm || (hm->n > 0 && !h_x) || (hm->m > 0 && !h_y)) {
    set_error("hostmat_spmv: invalid argument");
    return SPMV_B200_ERR_ARG;
  }
  const spmv_b200_plan *p = hm->plan;
  // Pipelined path (no split rows, enough tiles): x goes up first; then the rows are walked in chunks of tiles, the y0
  // chunk c+1 travels host->device while chunk c is multiplied and the y chunk c-1 travels device->host, so the two
  // PCIe directions are busy at the same time. Rows of a chunk are final once its kernels have run.
  const int kHostChunks = host_chunks();
  if (p->nsplit == 0 && p->ntiles >= 4 * kHostChunks && hm->m > 0) {
    // x travels in pieces as well: a chunk of row blocks needs x up to the largest column it references, so for banded
    // matrices the first kernel starts after 1/chunks of the input has arrived; a matrix whose first rows reference the
    // last columns simply gets the whole referenced range up front.
    int x_sent = hm->col_lo;
    for (int c = 0; c < kHostChunks; ++c) {
      const int tlo = (int)((long long)p->ntiles * c / kHostChunks);
      const int thi = (int)((long long)p->ntiles * (c + 1) / kHostChunks);
      const size_t rlo = (size_t)p->h_tile_row[tlo], rhi = (size_t)p->h_tile_row[thi];
      int need = x_sent;
      if (hm->tile_cmax.empty())
        need = hm->col_hi;
      else
        for (int t = tlo; t < thi; ++t)
          need = hm->tile_cmax[(size_t)t] + 1 > need ? hm->tile_cmax[(size_t)t] + 1 : need;
      if (need > x_sent) {
        B200_CUDA(cudaMemcpyAsync(hm->d_x + x_sent, h_x + x_sent, sizeof(double) * (size_t)(need - x_sent),
                                  cudaMemcpyHostToDevice, hm->s_in));
        x_sent = need;
      }
      if (rhi > rlo)
        B200_CUDA(cudaMemcpyAsync(hm->d_y + rlo, h_y + rlo, sizeof(double) * (rhi - rlo), cudaMemcpyHostToDevice,
                                  hm->s_in));
      B200_CUDA(cudaEventRecord(hm->ev_in[c], hm->s_in));
      B200_CUDA(cudaStreamWaitEvent(hm->stream, hm->ev_in[c], 0));
      const int rc = kernels_launch_tiles(p, alpha, beta, hm->d_x, hm->d_y, tlo, thi, hm->stream);
      if (rc != SPMV_B200_OK)
        return rc;
      B200_CUDA(cudaEventRecord(hm->ev_done[c], hm->stream));
      B200_CUDA(cudaStreamWaitEvent(hm->s_out, hm->ev_done[c], 0));
      if (rhi > rlo)
        B200_CUDA(cudaMemcpyAsync(h_y + rlo, hm->d_y + rlo, sizeof(double) * (rhi - rlo), cudaMemcpyDeviceToHost,
                                  hm->s_out));
    }
    B200_CUDA(cudaStreamSynchronize(hm->s_out));
    B200_CUDA(cudaStreamSynchronize(hm->stream));
    return SPMV_B200_OK;
  }
  if (hm->col_hi > hm->col_lo)
    B200_CUDA(cudaMemcpyAsync(hm->d_x + hm->col_lo, h_x + hm->col_lo, sizeof(double) * (size_t)(hm->col_hi - hm->col_lo),
                              cudaMemcpyHostToDevice, hm->stream));
  if (hm->m > 0)
    B200_CUDA(cudaMemcpyAsync(hm->d_y, h_y, sizeof(double) * (size_t)hm->m, cudaMemcpyHostToDevice, hm->stream));
  const int rc = spmv_b200_execute(hm->plan, alpha, beta, hm->d_x, hm->d_y, hm->stream);
  if (rc != SPMV_B200_OK)
    return rc;
  if (hm->m > 0)
    B200_CUDA(cudaMemcpyAsync(h_y, hm->d_y, sizeof(double) * (size_t)hm->m, cudaMemcpyDeviceToHost, hm->stream));
  B200_CUDA(cudaStreamSynchronize(hm->stream));
  return SPMV_B200_OK;
}

int spmv_b200_host_spmv(double alpha, double beta, int32_t m, int32_t n, int64_t nnz, const int32_t *h_rowptr,
                        const int32_t *h_colidx, const double *h_val, const double *h_x, double *h_y) {
  spmv_b200_hostmat *hm = nullptr;
  int rc = spmv_b200_hostmat_create(&hm, m, n, nnz, h_rowptr, h_colidx, h_val, nullptr);
  if (rc != SPMV_B200_OK)
    return rc;
  rc = spmv_b200_hostmat_spmv(hm, alpha, beta, h_x, h_y);
  const std::string keep = g_last_error;
  spmv_b200_hostmat_destroy(hm);
  g_last_error = keep;
  return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// sharding helpers
// ---------------------------------------------------------------------------------------------------------------
int spmv_b200_shard_bounds(int32_t m, int64_t nnz, const int32_t *d_rowptr, int32_t nshards, int32_t *h_bounds,
                           void *stream) {
  if (!h_bounds || (m > 0 && !d_rowptr)) {
    set_error("shard_bounds: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  return shard_bounds_run(m, nnz, d_rowptr, nshards, h_bounds, static_cast<cudaStream_t>(stream));
}

int spmv_b200_coo_to_csr(int32_t m, int32_t n, int64_t nnz, const int32_t *d_row, const int32_t *d_col,
                         const double *d_val, int32_t *d_rowptr_out, int32_t *d_col_out, double *d_val_out, void *stream) {
  if (m < 0 || n < 0 || nnz < 0 || nnz > 0x7fffffffLL || !d_rowptr_out ||
      (nnz > 0 && (!d_row || !d_col || !d_val || !d_col_out || !d_val_out))) {
    set_error("coo_to_csr: invalid argument");
    return SPMV_B200_ERR_ARG;
  }
  return coo_to_csr_run(m, n, nnz, d_row, d_col, d_val, d_rowptr_out, d_col_out, d_val_out,
                        static_cast<cudaStream_t>(stream));
}

int spmv_b200_col_block_bitmap(int64_t nnz, const int32_t *d_colidx, int32_t n, int32_t block_shift,
                               uint8_t *h_bitmap, void *stream) {
  if (!h_bitmap || (nnz > 0 && !d_colidx)) {
    set_error("col_block_bitmap: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  return col_block_bitmap_run(nnz, d_colidx, n, block_shift, h_bitmap, static_cast<cudaStream_t>(stream));
}

} // extern "C"
