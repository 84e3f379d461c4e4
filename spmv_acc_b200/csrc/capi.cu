// C ABI of libspmv_b200.so (declared in include/spmv_b200.h).
//
// Plan lifecycle mirrors the analyze / kernel / destroy phases of the reference's csr-adaptive-plus strategy
// (src/acc/hip-csr-adaptive-plus/csr_adaptive_plus_spmv.cpp:16-73); the stateless entry points mirror
// sparse_csr_spmv (src/acc/api/spmv.h:20-21) and the deprecated sparse_spmv (src/acc/api/spmv_imp.cpp:10-18).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
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
                  p->desc[2],  p->row_start_bits, p->nz_rows, p->desc_direct, p->lcol, p->xdesc};
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
  p->lcol = nullptr;
  p->xdesc = nullptr;
  p->xstage = false;
  p->xlines = 0;
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
  // Automatic tile size, third look: 1024 is the choice for irregular gathers with uniform rows (C3: MEDIUM row blocks,
  // small tiles leave L1 to the gathers). If the row lengths are skewed as well, most row blocks end up in the MIXED
  // kernel, which is faster with 2048 (measured on log-normal row lengths, profiles/r2_sweep_suitesparse_shapes.jsonl:
  // 0.0554 against 0.0607 ms and 0.2949 against 0.3025 ms).
  if (rc == SPMV_B200_OK && !(opt && opt->tile_nnz) && !p->direct && p->irregular && p->T == 1024 && p->ntiles > 0 &&
      2 * (long long)p->count[SPMV_B200_KIND_MIXED] > (long long)p->ntiles) {
    free_plan_arrays(p);
    reset_plan_arrays(p);
    p->T = 2048;
    rc = kernels_configure(p);
    if (rc == SPMV_B200_OK)
      rc = analysis_run(p, static_cast<cudaStream_t>(stream));
  }
  // staged-x form for regular matrices (x segments of every row block in shared memory, 16-bit local column indices)
  if (rc == SPMV_B200_OK && p->nnz > 0 && d_colidx)
    rc = analysis_xstage(p, static_cast<cudaStream_t>(stream));
  if (rc == SPMV_B200_OK)
    rc = kernels_configure_xs(p);
  if (rc != SPMV_B200_OK) {
    kernels_release(p);
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
  pa->align_rows = 0;
  pa->multicast_mask = push->count >= 32 ? push->multicast_mask : (push->multicast_mask & ((1u << push->count) - 1u));
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

int spmv_b200_plan_set_comm_sms(spmv_b200_plan *plan, int32_t sms) {
  if (!plan || sms < 0) {
    set_error("plan_set_comm_sms: plan is NULL or sms < 0");
    return SPMV_B200_ERR_ARG;
  }
  plan->comm_sms = sms;
  return SPMV_B200_OK;
}

int spmv_b200_plan_tile_col_range(spmv_b200_plan *plan, int32_t *h_min, int32_t *h_max, void *stream) {
  if (!plan || !h_min || !h_max) {
    set_error("plan_tile_col_range: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  return analysis_tile_col_range(plan, h_min, h_max, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------------
// fused halo loop: x <- A*x on one row shard, halo rows pushed into the neighbours' buffers by the SpMV kernels
// ---------------------------------------------------------------------------------------------------------------
struct spmv_b200_halo_loop {
  spmv_b200_halo_loop_desc d;
  PushArgs push[2];
  HaloSync sync;
  unsigned int *state = nullptr; // device: epoch, finished boundary CTAs, error
  TileDesc *desc_order = nullptr; // every tile, boundary row blocks first (single-launch mode)
  bool single_launch = false;
  bool use_graph = true;
  long long k = 0; // iterations enqueued so far
  int chunk = 0;   // iterations per graph launch (even; the graph always starts at an even iteration)
  cudaGraphExec_t graph = nullptr;
};

static int halo_enqueue_iteration(spmv_b200_halo_loop *L, int parity, cudaStream_t stream);
static int halo_build_graph(spmv_b200_halo_loop *L);

static unsigned long long halo_timeout_ns() {
  const char *e = getenv("SPMV_B200_FLAG_TIMEOUT_MS");
  const long long ms = e ? atoll(e) : 10000; // a peer that does not answer for 10 s is taken for dead; 0 = wait for ever
  return ms <= 0 ? 0ull : (unsigned long long)ms * 1000000ull;
}

int spmv_b200_halo_loop_create(spmv_b200_halo_loop **out, const spmv_b200_halo_loop_desc *d) {
  if (!out || !d || !d->plan || !d->buf[0] || !d->buf[1]) {
    set_error("halo_loop_create: bad argument");
    return SPMV_B200_ERR_ARG;
  }
  *out = nullptr;
  const spmv_b200_plan *p = d->plan;
  if (d->n_neigh < 0 || d->n_neigh > SPMV_B200_MAX_PUSH || d->n_boundary < 0 || d->n_boundary > SPMV_B200_MAX_RANGES ||
      d->n_interior < 0 || d->n_interior > SPMV_B200_MAX_RANGES || d->row_lo < 0 || d->row_hi < d->row_lo ||
      d->row_hi - d->row_lo != p->m) {
    set_error("halo_loop_create: inconsistent descriptor");
    return SPMV_B200_ERR_ARG;
  }
  // the ranges must be disjoint, inside [0, ntiles) and (when a boundary is given) cover every tile
  std::vector<int> order;
  std::vector<char> seen((size_t)p->ntiles, 0);
  int nb_tiles = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int n = pass == 0 ? d->n_boundary : d->n_interior;
    const int32_t *r = pass == 0 ? d->boundary : d->interior;
    for (int i = 0; i < n; ++i) {
      if (r[2 * i] < 0 || r[2 * i + 1] > p->ntiles || r[2 * i] > r[2 * i + 1]) {
        set_error("halo_loop_create: tile range out of bounds");
        return SPMV_B200_ERR_ARG;
      }
      for (int t = r[2 * i]; t < r[2 * i + 1]; ++t) {
        if (seen[(size_t)t]) {
          set_error("halo_loop_create: tile ranges overlap");
          return SPMV_B200_ERR_ARG;
        }
        seen[(size_t)t] = 1;
        order.push_back(t);
      }
    }
    if (pass == 0)
      nb_tiles = (int)order.size();
  }
  if (d->n_boundary > 0 && (int)order.size() != p->ntiles) {
    set_error("halo_loop_create: boundary + interior ranges do not cover every row block");
    return SPMV_B200_ERR_ARG;
  }
  if (d->n_boundary > 0 && p->nsplit > 0) {
    set_error("halo_loop_create: a plan with split rows cannot be run boundary-first");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  if (d->n_boundary == 0) { // no split schedule: the whole shard is "boundary" (waits first, signals when all is stored)
    order.resize((size_t)p->ntiles);
    for (int t = 0; t < p->ntiles; ++t)
      order[(size_t)t] = t;
    nb_tiles = p->ntiles;
  }
  spmv_b200_halo_loop *L = new (std::nothrow) spmv_b200_halo_loop();
  if (!L) {
    set_error("halo_loop_create: out of host memory");
    return SPMV_B200_ERR_ARG;
  }
  L->d = *d;
  for (int b = 0; b < 2; ++b) {
    if (int rc = convert_push(p, &d->push[b], &L->push[b], "halo_loop_create")) {
      delete L;
      return rc;
    }
    L->push[b].align_rows = (d->flags & SPMV_B200_HALO_ALIGN_PUSH) ? 1 : 0;
  }
  cudaError_t e = cudaMalloc(&L->state, 4 * sizeof(unsigned int));
  if (e == cudaSuccess)
    e = cudaMemset(L->state, 0, 4 * sizeof(unsigned int));
  if (e != cudaSuccess) {
    cudaFree(L->state);
    delete L;
    B200_CUDA(e);
  }
  for (int j = 0; j < SPMV_B200_MAX_PUSH; ++j) {
    L->sync.wait[j] = j < d->n_neigh ? d->wait_flags[j] : nullptr;
    L->sync.signal[j] = j < d->n_neigh ? d->signal_flags[j] : nullptr;
    if (j < d->n_neigh && (!d->wait_flags[j] || !d->signal_flags[j])) {
      cudaFree(L->state);
      delete L;
      set_error("halo_loop_create: NULL flag pointer");
      return SPMV_B200_ERR_ARG;
    }
  }
  L->sync.n_neigh = d->n_neigh;
  L->sync.n_boundary = d->n_neigh > 0 ? nb_tiles : 0; // a rank without neighbours orders nothing
  L->sync.state = L->state;
  L->sync.timeout_ns = halo_timeout_ns();
  L->use_graph = !(d->flags & SPMV_B200_HALO_NO_GRAPH);
  L->single_launch = !(d->flags & SPMV_B200_HALO_MULTI_LAUNCH) && kernels_halo_single_launch_ok(p);
  if (L->single_launch) {
    if (int rc = analysis_gather_descs(p->desc_all, order.data(), p->ntiles, &L->desc_order, nullptr)) {
      cudaFree(L->state);
      delete L;
      return rc;
    }
  }
  if (L->use_graph) {
    const char *e = getenv("SPMV_B200_HALO_GRAPH_CHUNK");
    int c = e ? atoi(e) : 20;
    L->chunk = c < 2 ? 2 : (c > 1000 ? 1000 : (c & ~1));
    if (int rc = halo_build_graph(L)) {
      const std::string keep = g_last_error;
      spmv_b200_halo_loop_destroy(L);
      g_last_error = keep;
      return rc;
    }
  }
  *out = L;
  return SPMV_B200_OK;
}

// one iteration, enqueued (also under stream capture): parity of the source buffer is a launch parameter, the epoch
// the flags are compared with lives in device memory
static int halo_enqueue_iteration(spmv_b200_halo_loop *L, int parity, cudaStream_t stream) {
  const spmv_b200_halo_loop_desc &d = L->d;
  const double *src = d.buf[parity];
  double *ys = d.buf[parity ^ 1] + d.row_lo;
  const PushArgs *push = &L->push[parity ^ 1];
  if (L->single_launch)
    return kernels_launch_halo(d.plan, L->desc_order, src, ys, push, L->sync, stream);
  int rc;
  if (d.n_neigh > 0 && (rc = kernels_halo_wait(L->sync, stream)))
    return rc;
  if (d.n_boundary > 0) {
    for (int r = 0; r < d.n_boundary; ++r)
      if (d.boundary[2 * r + 1] > d.boundary[2 * r] &&
          (rc = kernels_launch_tiles(d.plan, 1.0, 0.0, src, ys, d.boundary[2 * r], d.boundary[2 * r + 1], stream, push)))
        return rc;
    if (d.n_neigh > 0 && (rc = kernels_halo_signal(L->sync, stream)))
      return rc;
    for (int r = 0; r < d.n_interior; ++r)
      if (d.interior[2 * r + 1] > d.interior[2 * r] &&
          (rc = kernels_launch_tiles(d.plan, 1.0, 0.0, src, ys, d.interior[2 * r], d.interior[2 * r + 1], stream)))
        return rc;
    return SPMV_B200_OK;
  }
  if ((rc = kernels_launch(d.plan, 1.0, 0.0, src, ys, stream, push)))
    return rc;
  return d.n_neigh > 0 ? kernels_halo_signal(L->sync, stream) : SPMV_B200_OK;
}

// `chunk` iterations captured into one executable graph (on a stream of our own: the caller's may be the legacy null
// stream, which cannot be captured). Nothing is executed here.
static int halo_build_graph(spmv_b200_halo_loop *L) {
  cudaStream_t cs = nullptr;
  cudaGraph_t graph = nullptr;
  B200_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  int rc = SPMV_B200_OK;
  if (e == cudaSuccess) {
    for (int i = 0; i < L->chunk && rc == SPMV_B200_OK; ++i)
      rc = halo_enqueue_iteration(L, i & 1, cs);
    e = cudaStreamEndCapture(cs, &graph);
  }
  if (e == cudaSuccess && rc == SPMV_B200_OK)
    e = cudaGraphInstantiate(&L->graph, graph, 0);
  if (graph)
    cudaGraphDestroy(graph);
  cudaStreamDestroy(cs);
  if (rc != SPMV_B200_OK)
    return rc;
  B200_CUDA(e);
  return SPMV_B200_OK;
}

int spmv_b200_halo_loop_run(spmv_b200_halo_loop *L, int32_t iterations, void *stream_) {
  if (!L || iterations < 0) {
    set_error("halo_loop_run: bad argument");
    return SPMV_B200_ERR_ARG;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int left = iterations, rc;
  if (left > 0 && (L->k & 1)) { // graph launches start at an even iteration
    if ((rc = halo_enqueue_iteration(L, 1, stream)))
      return rc;
    ++L->k;
    --left;
  }
  while (L->graph && left >= L->chunk) {
    B200_CUDA(cudaGraphLaunch(L->graph, stream));
    L->k += L->chunk;
    left -= L->chunk;
  }
  for (; left > 0; --left, ++L->k)
    if ((rc = halo_enqueue_iteration(L, (int)(L->k & 1), stream)))
      return rc;
  return SPMV_B200_OK;
}

int spmv_b200_halo_loop_sync(spmv_b200_halo_loop *L, void *stream) {
  if (!L) {
    set_error("halo_loop_sync: loop is NULL");
    return SPMV_B200_ERR_ARG;
  }
  B200_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  unsigned int st[4] = {0, 0, 0, 0};
  B200_CUDA(cudaMemcpy(st, L->state, sizeof(st), cudaMemcpyDeviceToHost));
  if (st[2] != 0u) {
    set_error("halo loop: a neighbour's flag did not reach iteration " + std::to_string(st[2] & 0x7fffffffu) +
              " in time; x is not valid from that iteration on");
    return SPMV_B200_ERR_TIMEOUT;
  }
  return SPMV_B200_OK;
}

int spmv_b200_halo_loop_get_info(const spmv_b200_halo_loop *L, spmv_b200_halo_loop_info *info) {
  if (!L || !info) {
    set_error("halo_loop_get_info: NULL argument");
    return SPMV_B200_ERR_ARG;
  }
  std::memset(info, 0, sizeof(*info));
  info->iterations_enqueued = L->k;
  info->single_launch = L->single_launch ? 1 : 0;
  info->uses_graph = L->graph ? L->chunk : 0;
  info->boundary_row_blocks = L->sync.n_boundary;
  int launches = 1;
  if (!L->single_launch) {
    const spmv_b200_halo_loop_desc &d = L->d;
    spmv_b200_plan_info pi;
    spmv_b200_plan_get_info(d.plan, &pi);
    launches = (d.n_neigh > 0 ? 2 : 1) + (d.n_boundary > 0 ? d.n_boundary + d.n_interior : pi.launches_per_execute);
  }
  info->launches_per_iteration = launches;
  return SPMV_B200_OK;
}

int spmv_b200_halo_loop_destroy(spmv_b200_halo_loop *L) {
  if (!L)
    return SPMV_B200_OK;
  if (L->graph)
    cudaGraphExecDestroy(L->graph);
  cudaFree(L->desc_order);
  cudaFree(L->state);
  delete L;
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
  kernels_release(plan);
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
  info->xstage = p->xstage ? 1 : 0;
  info->xstage_lines = p->xstage ? p->xlines : 0;
  info->ring_ctas = p->xstage ? p->ring_ctas : 0;
  info->ring_stages = p->xstage ? p->ring_stages : 0;
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
  case SPMV_B200_EXPORT_LCOL:
    src = p->xstage ? p->lcol + (p->elem_base - p->lcol_base) : nullptr;
    bytes = p->xstage ? 2 * p->nnz : 0;
    break;
  case SPMV_B200_EXPORT_XDESC:
    src = p->xdesc;
    bytes = p->xstage ? 128 * nt : 0;
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
    DeviceScratch tmp;
    B200_CUDA(tmp.alloc((size_t)bytes));
    if (int rc = analysis_row_bins(p, tmp.as<unsigned char>(), nullptr))
      return rc;
    B200_CUDA(cudaMemcpy(h_dst, tmp.p, (size_t)bytes, cudaMemcpyDeviceToHost));
    return SPMV_B200_OK;
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
// stateless entry points: a small plan cache keyed on the device pointers and the shape, validated on every hit
// ---------------------------------------------------------------------------------------------------------------
// sparse_csr_spmv has no handle (src/acc/api/spmv.h:20-21), so the one-time analysis is kept in a cache. Pointers and
// shape alone do not identify a matrix: a caching allocator hands the same addresses to the next matrix of the same
// size. Everything a plan derives (row blocks, split rows, row-start flags) is a function of the row pointers, so every
// hit re-reads a fingerprint of them on the device -- rowptr[0], rowptr[m] and kFingerSamples evenly spaced entries,
// hashed position-wise -- and a mismatch drops the plan and analyses again. Cost: one single-CTA kernel and a read of
// 8 bytes of mapped host memory per call (SPMV_B200_CACHE_TRUST=1 in the environment skips it for callers that
// guarantee immutable matrices; the plan API never pays it).
namespace {
constexpr int kFingerSamples = 2048;

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_fingerprint(const int *__restrict__ rowptr, int m,
                                                      unsigned long long *__restrict__ out) {
  __shared__ unsigned long long sh[8];
  unsigned long long h = 0ull; // wrapping sum of position-keyed hashes: independent of the order of the additions
  for (int i = threadIdx.x; i <= kFingerSamples; i += blockDim.x) {
    const long long pos = ((long long)m * i) / kFingerSamples; // i == kFingerSamples: rowptr[m]
    h += mix64(((unsigned long long)pos << 32) ^ (unsigned int)__ldg(rowptr + pos));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    h += __shfl_xor_sync(0xffffffffu, h, off);
  if ((threadIdx.x & 31) == 0)
    sh[threadIdx.x >> 5] = h;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0ull;
    for (int w = 0; w < 8; ++w)
      t += sh[w];
    *out = t | 1ull; // never 0: 0 marks "not written yet"
  }
}

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
  unsigned long long fingerprint;
};
std::mutex g_cache_mutex;
std::list<CacheEntry> g_cache; // most recently used first
constexpr size_t kCacheCapacity = 16;
unsigned long long *g_finger_host = nullptr; // mapped pinned word the fingerprint kernel writes
long long g_cache_revalidations = 0;         // hits whose fingerprint did not match (plan analysed again)

bool cache_trusted() {
  static const bool t = [] {
    const char *e = getenv("SPMV_B200_CACHE_TRUST");
    return e && e[0] == '1';
  }();
  return t;
}

// fingerprint of rowptr as it is now in device memory (null stream; returns after the kernel has finished)
int read_fingerprint(const int *rowptr, int m, unsigned long long *out) {
  *out = 1ull;
  if (m <= 0 || !rowptr)
    return SPMV_B200_OK;
  if (!g_finger_host)
    B200_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&g_finger_host), sizeof(unsigned long long), cudaHostAllocMapped));
  unsigned long long *d_out = nullptr;
  B200_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_out), g_finger_host, 0));
  *g_finger_host = 0ull;
  k_fingerprint<<<1, 256, 0, nullptr>>>(rowptr, m, d_out);
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(nullptr));
  *out = *reinterpret_cast<volatile unsigned long long *>(g_finger_host);
  return SPMV_B200_OK;
}

// Looks the matrix up (or analyses it) and enqueues the SpMV, all under the cache lock: another host thread cannot
// evict or invalidate the plan between the lookup and the launches.
int cached_spmv(int m, int n, long long nnz, const int *rowptr, const int *col, const double *val, double alpha,
                double beta, const double *x, double *y) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  const CacheKey key{rowptr, col, val, m, n, dev};
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  unsigned long long fp = 1ull;
  const bool validate = !cache_trusted();
  if (validate)
    if (int rc = read_fingerprint(rowptr, m, &fp))
      return rc;
  for (auto it = g_cache.begin(); it != g_cache.end(); ++it) {
    if (!(it->key == key))
      continue;
    if ((nnz >= 0 && it->plan->nnz != nnz) || (validate && it->fingerprint != fp)) {
      // same addresses, other contents: the old analysis is worthless
      ++g_cache_revalidations;
      spmv_b200_plan_destroy(it->plan);
      g_cache.erase(it);
      break;
    }
    g_cache.splice(g_cache.begin(), g_cache, it);
    return spmv_b200_execute(g_cache.front().plan, alpha, beta, x, y, nullptr);
  }
  spmv_b200_plan *p = nullptr;
  const int rc = spmv_b200_plan_create(&p, m, n, nnz, rowptr, col, val, nullptr, nullptr);
  if (rc != SPMV_B200_OK)
    return rc;
  g_cache.push_front(CacheEntry{key, p, fp});
  while (g_cache.size() > kCacheCapacity) {
    // plans may still have kernels in flight on the null stream; cudaFree synchronises implicitly
    spmv_b200_plan_destroy(g_cache.back().plan);
    g_cache.pop_back();
  }
  return spmv_b200_execute(p, alpha, beta, x, y, nullptr);
}
} // namespace

int spmv_b200_csr_spmv(int32_t trans, double alpha, double beta, int32_t m, int32_t n, int32_t nnz,
                       const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_val, const double *d_x,
                       double *d_y) {
  if (trans != 0) { // src/acc/api/types.h:8 — only operation_none is supported by the reference as well
    set_error("csr_spmv: only operation_none (trans = 0) is supported");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  return cached_spmv(m, n, nnz, d_rowptr, d_colidx, d_val, alpha, beta, d_x, d_y);
}

int spmv_b200_sparse_spmv(int32_t trans, double alpha, double beta, int32_t m, int32_t n, const int32_t *d_rowptr,
                          const int32_t *d_colidx, const double *d_val, const double *d_x, double *d_y) {
  // src/acc/api/spmv_imp.cpp:14 reads rowptr[hm] on the host; here nnz comes from the device inside the analysis
  if (trans != 0) {
    set_error("sparse_spmv: only operation_none (trans = 0) is supported");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  return cached_spmv(m, n, -1, d_rowptr, d_colidx, d_val, alpha, beta, d_x, d_y);
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

int64_t spmv_b200_cache_revalidations(void) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  return (int64_t)g_cache_revalidations;
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
  bool owns_matrix = true; // false: rowptr / col / val are the caller's device arrays (hostmat_create_device)
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
// Row chunks of the pipelined host-buffer path (SPMV_B200_HOST_CHUNKS overrides): pieces of ~28 MB per direction, at
// least 8 and at most 16 of them. Measured on C5 (453 MB each way; both PCIe directions busy at once run 45.8 GB/s each
// on these boxes = 9.9 ms): 4 chunks 13.2 ms, 8: 12.2, 16: 11.8, 32: 11.9 (profiles/r2_e2e_chunks_c5.jsonl); on C2
// (134 MB each way) 8 chunks were best in round 1.
static int host_chunks(size_t bytes_per_direction) {
  const char *e = getenv("SPMV_B200_HOST_CHUNKS");
  int c = e ? atoi(e) : (int)(bytes_per_direction / (28u << 20));
  if (!e)
    c = c < 8 ? 8 : (c > 16 ? 16 : c);
  return c < 1 ? 1 : (c > kHostChunksMax ? kHostChunksMax : c);
}

int spmv_b200_hostmat_destroy(spmv_b200_hostmat *hm) {
  if (!hm)
    return SPMV_B200_OK;
  if (hm->stream)
    cudaStreamSynchronize(hm->stream);
  spmv_b200_plan_destroy(hm->plan);
  if (hm->owns_matrix) {
    cudaFree(hm->d_rowptr);
    cudaFree(hm->d_col);
    cudaFree(hm->d_val);
  }
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
  const size_t m = (size_t)hm->m, n = (size_t)hm->n, nnz = hm->nnz > 0 ? (size_t)hm->nnz : 0;
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
  if (hm->owns_matrix) {
    B200_CUDA(cudaMalloc(&hm->d_rowptr, sizeof(int) * (m + 1)));
    B200_CUDA(cudaMalloc(&hm->d_col, sizeof(int) * (nnz ? nnz : 1)));
    B200_CUDA(cudaMalloc(&hm->d_val, sizeof(double) * (nnz ? nnz : 1)));
  }
  B200_CUDA(cudaMalloc(&hm->d_x, sizeof(double) * (n ? n : 1)));
  B200_CUDA(cudaMalloc(&hm->d_y, sizeof(double) * (m ? m : 1)));
  if (hm->owns_matrix) {
    B200_CUDA(cudaMemcpyAsync(hm->d_rowptr, h_rowptr, sizeof(int) * (m + 1), cudaMemcpyHostToDevice, hm->stream));
    if (nnz) {
      B200_CUDA(cudaMemcpyAsync(hm->d_col, h_colidx, sizeof(int) * nnz, cudaMemcpyHostToDevice, hm->stream));
      B200_CUDA(cudaMemcpyAsync(hm->d_val, h_val, sizeof(double) * nnz, cudaMemcpyHostToDevice, hm->stream));
    }
  }
  int rc = spmv_b200_plan_create(&hm->plan, hm->m, hm->n, hm->nnz, hm->d_rowptr, hm->d_col, hm->d_val, opt, hm->stream);
  if (rc != SPMV_B200_OK)
    return rc;
  hm->col_lo = 0;
  hm->col_hi = hm->n;
  if (hm->plan->ntiles > 0 && hm->plan->nnz > 0) {
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

int spmv_b200_hostmat_create_device(spmv_b200_hostmat **out, int32_t m, int32_t n, int64_t nnz,
                                    const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_val,
                                    const spmv_b200_options *opt) {
  if (!out || m < 0 || n < 0 || nnz > 0x7fffffffLL || (m > 0 && !d_rowptr)) {
    set_error("hostmat_create_device: invalid argument");
    return SPMV_B200_ERR_ARG;
  }
  *out = nullptr;
  spmv_b200_hostmat *hm = new (std::nothrow) spmv_b200_hostmat();
  if (!hm) {
    set_error("hostmat_create_device: out of host memory");
    return SPMV_B200_ERR_ARG;
  }
  hm->m = m;
  hm->n = n;
  hm->nnz = nnz; // < 0: the analysis reads it from rowptr
  hm->owns_matrix = false;
  hm->d_rowptr = const_cast<int *>(d_rowptr);
  hm->d_col = const_cast<int *>(d_colidx);
  hm->d_val = const_cast<double *>(d_val);
  int rc = hostmat_build(hm, nullptr, nullptr, nullptr, opt);
  if (rc == SPMV_B200_OK)
    hm->nnz = hm->plan->nnz;
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
  // beta == 0 with SPMV_B200_FLAG_BETA0_SKIP_Y: the kernels never read y, so y0 does not travel either
  const bool send_y0 = !(beta == 0.0 && (p->flags & SPMV_B200_FLAG_BETA0_SKIP_Y));
  // Pipelined path (no split rows, enough tiles): x goes up first; then the rows are walked in chunks of tiles, the y0
  // chunk c+1 travels host->device while chunk c is multiplied and the y chunk c-1 travels device->host, so the two
  // PCIe directions are busy at the same time. Rows of a chunk are final once its kernels have run.
  const int kHostChunks = host_chunks(sizeof(double) * (size_t)hm->m);
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
      if (rhi > rlo && send_y0)
        B200_CUDA(cudaMemcpyAsync(hm->d_y + rlo, h_y + rlo, sizeof(double) * (rhi - rlo), cudaMemcpyHostToDevice,
                                  hm->s_in));
      B200_CUDA(cudaEventRecord(hm->ev_in[c], hm->s_in));
      B200_CUDA(cudaStreamWaitEvent(hm->stream, hm->ev_in[c], 0));
      const int rc = kernels_launch_tiles(p, alpha, beta, hm->d_x, hm->d_y, tlo, thi, hm->stream);
      if (rc != SPMV_B200_OK) { // copies of earlier chunks are still in flight on the caller's buffers
        cudaStreamSynchronize(hm->s_in);
        cudaStreamSynchronize(hm->stream);
        cudaStreamSynchronize(hm->s_out);
        return rc;
      }
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
  if (hm->m > 0 && send_y0)
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
