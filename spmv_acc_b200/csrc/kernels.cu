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
//   k_spmv_rows<.., XS>      staged-x form of both: the runs of x a row block references are brought into shared memory
//                                           by TMA as well and addressed through 16-bit local column indices (10 bytes
//                                           per non-zero from HBM instead of 12, no gathers through L1); regular
//                                           matrices (stencils, banded) only, decided by the analysis.
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
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

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
// L2 policy of the matrix streams: evict-first when the matrix is much larger than the L2 (every byte is read once per
// SpMV and should not displace x); normal when it is small enough for part of it to survive until the next SpMV
__device__ __forceinline__ unsigned long long policy_stream(int keep) {
  unsigned long long pol;
  if (keep)
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  else
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

__device__ __forceinline__ void tma_bulk_g2s_nohint(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                    unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
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

// x gather through the read-only path; `na` selects L1::no_allocate (the line is not kept in L1)
__device__ __forceinline__ double gather_x(const double *__restrict__ x, int c, int na) {
  if (na == 2) // experiment: no gather at all (wrong results; isolates the cost of the x traffic)
    return 1.0;
  if (na)
    return ld_stream_f64(x + c);
  return __ldg(x + c);
}

// one round trip: the 32-byte descriptor of tile i of this kind as two 128-bit loads
__device__ __forceinline__ TileDesc load_desc(const TileDesc *__restrict__ desc, int i) {
  const int4 *p = reinterpret_cast<const int4 *>(desc + i);
  const int4 lo = __ldg(p), hi = __ldg(p + 1);
  TileDesc d;
  d.r0 = lo.x;
  d.r1 = lo.y;
  d.e0 = lo.z;
  d.e1 = lo.w;
  d.head_end = hi.x;
  d.tail_start = hi.y;
  d.tile = hi.z;
  d.flags = hi.w;
  return d;
}

// ---------------------------------------------------------------------------------------------------------------
// tile streaming: global -> shared. Element i of the tile lives at smem index (i - a0), a0 = elem_begin & ~3.
// ---------------------------------------------------------------------------------------------------------------
template <bool TMA, int NT = kThreads>
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
        const unsigned long long pol = policy_stream(a.stream_keep);
        mbar_arrive_expect_tx(bar, (uint32_t)cnt * 12u);
        tma_bulk_g2s(sval, a.val + a0, (uint32_t)cnt * 8u, bar, pol);
        tma_bulk_g2s(scol, a.col + a0, (uint32_t)cnt * 4u, bar, pol);
      } else {
        mbar_arrive(bar);
      }
    }
    for (int i = cnt + tid; i < span; i += NT) {
      sval[i] = ld_stream_f64(a.val + a0 + i);
      scol[i] = ld_stream_s32(a.col + a0 + i);
    }
  } else {
    for (int i = (e0 - a0) + tid; i < span; i += NT) {
      sval[i] = ld_stream_f64(a.val + a0 + i);
      scol[i] = ld_stream_s32(a.col + a0 + i);
    }
  }
}

// Staged-x form: value, the 16-bit local column indices and the segments of x the tile references, all by TMA
// (a0 = elem_begin & ~7: 16-byte alignment of the 2-byte index stream). Warp 0 issues the copies: lane s brings
// segment s of x (whole 128-byte lines; only the very last line of x can be short), lane 0 the two streams. The
// transaction count is posted by lane 0 in the same operation as the barrier's only arrival, so the phase cannot
// complete before every copy has been accounted for, whatever the order in which the lanes issue.
__device__ __forceinline__ void tile_issue_loads_xs(const SpmvArgs &a, int tile, int a0, int e1, double *sval,
                                                    unsigned short *slcol, double *sx, unsigned long long *bar,
                                                    int tid) {
  const int span = e1 - a0;
  if (tid == 0)
    mbar_init(bar, 1);
  __syncthreads();
  const long long avail = a.nnz - (long long)a0;
  int cnt = (span + 7) & ~7;
  if ((long long)cnt > avail)
    cnt = (int)(avail & ~7LL);
  if (tid < 32) {
    const XDesc *__restrict__ xd = a.xdesc + tile;
    const int nseg = __ldg(&xd->nseg);
    unsigned int xbytes = 0;
    int line = 0, off = 0;
    if (tid < nseg) {
      line = __ldg(&xd->line[tid]);
      off = (int)__ldg(&xd->off[tid]);
      const int end = tid + 1 < nseg ? (int)__ldg(&xd->off[tid + 1]) : __ldg(&xd->nlines);
      long long elems = (long long)(end - off) * 16;
      const long long left = (long long)a.n - (long long)line * 16;
      if (elems > left)
        elems = left;
      xbytes = (unsigned int)(elems & ~1LL) * 8u;
      if (elems & 1) // n is odd and this is the last entry of x: a 16-byte copy would read past the vector
        sx[(long long)off * 16 + elems - 1] = __ldg(a.x + (long long)line * 16 + elems - 1);
    }
    unsigned int total = xbytes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      total += __shfl_xor_sync(0xffffffffu, total, o);
    total += (unsigned int)cnt * 10u;
    if (tid == 0) {
      if (total > 0)
        mbar_arrive_expect_tx(bar, total);
      else
        mbar_arrive(bar);
    }
    __syncwarp();
    if (tid == 0 && cnt > 0) {
      const unsigned long long pol = policy_stream(a.stream_keep);
      tma_bulk_g2s(sval, a.val + a0, (uint32_t)cnt * 8u, bar, pol);
      tma_bulk_g2s(slcol, a.lcol + ((long long)a0 - a.lcol_base), (uint32_t)cnt * 2u, bar, pol);
    }
    if (xbytes > 0) // x is what neighbouring row blocks read again: default L2 policy
      tma_bulk_g2s_nohint(sx + (long long)off * 16, a.x + (long long)line * 16, xbytes, bar);
  }
  for (int i = cnt + tid; i < span; i += kThreads) {
    sval[i] = ld_stream_f64(a.val + a0 + i);
    slcol[i] = __ldg(a.lcol + ((long long)a0 - a.lcol_base) + i);
  }
}

// the y store of every kernel: local store plus, for rows another GPU waits for, a store into that GPU's memory.
// MC: the kernel also understands NVLink multicast destinations (the row kernels and the ring; the MIXED, direct and
// fix-up kernels do not -- their register budgets are tight and multicast pushes are refused for plans that need them).
template <bool MC = false>
__device__ __forceinline__ void emit_y(double *__restrict__ y, const PushArgs &push, int row, double v) {
  y[row] = v;
  if (push.count) {
#pragma unroll 1
    for (int j = 0; j < push.count; ++j)
      if (row >= push.row_lo[j] && row < push.row_hi[j]) {
        if (MC && ((push.multicast_mask >> j) & 1u)) // the switch replicates the store into every GPU's copy
          asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(push.dst[j] + row), "d"(v) : "memory");
        else
          push.dst[j][row] = v;
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SHORT (VEC = false) and MEDIUM (VEC = true) tiles: every owned row lies completely inside the tile
// ---------------------------------------------------------------------------------------------------------------
// One slot = one (value, column) pair of a row. In the staged-x form the slot loads are written as predicated PTX on
// 32-bit shared-space addresses: a slot behind the end of its row keeps its zero-initialised registers and contributes
// fma(0, 0, sum). The C++ form of the same thing -- (k < e) ? sx[lcol[k]] : 0 -- compiles into a branch per slot around
// two dependent loads plus a generic-to-shared address conversion per access (22 instructions per slot in the SASS, 54
// thread instructions per non-zero on the 27-point stencil: with no long-latency load left in the loop the kernel was
// bound by instruction issue, ncu 54 % issue slots busy at 49 % DRAM); here a slot is 8-10 instructions, branch-free.
// staged-x form: x[col] = sx[lcol[k]], everything in shared memory
__device__ __forceinline__ void slot_xs(double &xv, double &vv, bool p, uint32_t a_lcol, uint32_t a_val, uint32_t sx_s) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\t.reg .u16 h;\n\t"
               "setp.ne.u32 p, %2, 0;\n\t"
               "@p ld.shared.u16 h, [%3];\n\t"
               "@p ld.shared.f64 %1, [%4];\n\t"
               "@p cvt.u32.u16 t, h;\n\t"
               "@p shl.b32 t, t, 3;\n\t"
               "@p add.u32 t, t, %5;\n\t"
               "@p ld.shared.f64 %0, [t];\n\t}"
               : "+d"(xv), "+d"(vv)
               : "r"((uint32_t)p), "r"(a_lcol), "r"(a_val), "r"(sx_s));
}
// log2 of the lanes per row of a MEDIUM row block: the smallest lv with 2^lv >= ceil(avg / vec_div), avg = floor(elems /
// rows), at most 5. Written without the two integer divisions (each costs ~20 instructions, paid per row block and warp):
// 2^lv * vec_div >= floor(elems / rows)  <=>  (2^lv * vec_div + 1) * rows > elems.
__device__ __forceinline__ int lanes_log2(int elems, int nrows, int vec_div) {
  const int n = nrows > 0 ? nrows : 1;
  int lv = 0;
  while (lv < 5 && (long long)((vec_div << lv) + 1) * n <= (long long)elems)
    ++lv;
  return lv;
}

// W = x gathers issued back to back per row and round (gather form): all W gathers of a round are in flight before the
// first FMA, and y is fetched before the tile has landed, so that a CTA exposes one round trip to memory per phase
// (tile, gathers) instead of one per batch of four elements. The staged-x form has no long-latency load in the loop
// and walks rounds of 4 slots.
// Rows of one tile whose value / colindex are (being) staged in shared memory; waits for the TMA phase `parity` of `bar`.
template <bool TMA, bool VEC, int W, bool ROT, bool XS>
__device__ __forceinline__ void rows_tile(const SpmvArgs &a, const double *__restrict__ sval,
                                          const int *__restrict__ scol, const unsigned short *__restrict__ slcol,
                                          const double *__restrict__ sx, int *__restrict__ srow,
                                          unsigned long long *bar, uint32_t parity, int r0, int nrows, int a0, int e0,
                                          int e1, int tid) {
  const int lv = VEC ? lanes_log2(e1 - e0, nrows, a.vec_div) : 0; // log2(lanes per row)
  const int V = 1 << lv;
  const int G = kThreads >> lv; // rows per lane-group pass
  const int g = tid >> lv;
  const int l = tid & (V - 1);
  const uint32_t sval_s = XS ? smem_u32(sval) : 0u, sx_s = XS ? smem_u32(sx) : 0u, scol_s = XS ? smem_u32(slcol) : 0u;

  // y of the first pass, requested while the tile is still in flight
  const double ypre = (a.read_y && l == 0 && g < nrows && g < kRowChunk) ? a.y[r0 + g] : 0.0;

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
      mbar_wait(bar, parity);

    for (int rb = 0; rb < nr; rb += G) {
      const int r = rb + g;
      const bool act = r < nr;
      int k = act ? srow[r] - a0 + l : 0;
      const int e = act ? srow[r + 1] - a0 : 0;
      double sum = 0.0;
      const double yv = (cb == 0 && rb == 0) ? ypre : ((a.read_y && act && l == 0) ? a.y[r0 + cb + r] : 0.0);
      if (XS) {
        // rounds of 4 predicated slots, everything from shared memory
        for (; k < e; k += 4 << lv) {
          double xv[4], vv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kk = k + (j << lv);
            xv[j] = 0.0;
            vv[j] = 0.0;
            slot_xs(xv[j], vv[j], kk < e, scol_s + 2u * (uint32_t)kk, sval_s + 8u * (uint32_t)kk, sx_s);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sum = fma(vv[j], xv[j], sum);
        }
      } else {
        // Gather form: bound by the latency of the W gathers in flight, not by instruction issue; the compiler's
        // branchy code keeps all W gathers in flight inside the register budget (a predicated-PTX version of this
        // loop spilled or serialised the gathers at 40 registers).
        // Irregular-gather plans (ROT): lane groups of a warp walk the W slots of a round in rotated order. With rows
        // whose length is a multiple of 16 (32 nnz per row in C3) every group would otherwise read the same
        // shared-memory banks in the same step (ncu: 8-way conflicts, 130 M extra wavefronts on C3). The rotation
        // depends only on the group's position in the tile.
        int rot = ROT ? (g % W) * V : 0; // offset of the first slot this group reads in a round
        while (k < e) {
          double xv[W];
#pragma unroll
          for (int j = 0; j < W; ++j) {
            int kk = k + j * V;
            if (ROT) {
              kk += rot;
              kk -= (kk >= k + W * V) ? W * V : 0; // wrap around inside the round
            }
            xv[j] = (kk < e) ? gather_x(a.x, scol[kk], a.gather_na) : 0.0;
          }
          if (ROT)
            asm volatile("" : "+r"(rot)); // recompute the indices below instead of keeping W of them in registers
#pragma unroll
          for (int j = 0; j < W; ++j) {
            int kk = k + j * V;
            if (ROT) {
              kk += rot;
              kk -= (kk >= k + W * V) ? W * V : 0;
            }
            if (kk < e)
              sum = fma(sval[kk], xv[j], sum);
          }
          k += W * V;
        }
      }
      if (VEC) {
        for (int off = V >> 1; off > 0; off >>= 1)
          sum += __shfl_down_sync(0xffffffffu, sum, off, V);
      }
      if (act && l == 0) // cli/verification.cpp:64  y[i] = alpha * y0 + beta * y[i]
        emit_y<true>(a.y, a.push, r0 + cb + r, a.alpha * sum + a.beta * yv);
    }
    if (cb + kRowChunk >= nrows)
      break;
  }
}

// one CTA = one tile: shared-memory layout, loads, rows
//   gather form: sval[cap] | scol[cap] (int32) | srow[kRowChunk + 1]
//   staged-x   : sval[cap] | sx[xcap] | srow[kRowChunk + 1] | slcol[cap] (uint16)
template <bool TMA, bool VEC, int W, bool ROT, bool XS>
__device__ __forceinline__ void rows_cta(const SpmvArgs &a, unsigned char *smem_raw, unsigned long long *bar) {
  const int tid = threadIdx.x;
  const TileDesc d = load_desc(a.desc, blockIdx.x);
  double *sval = reinterpret_cast<double *>(smem_raw);
  if (XS) {
    double *sx = sval + a.cap;
    int *srow = reinterpret_cast<int *>(sx + a.xcap);
    unsigned short *slcol = reinterpret_cast<unsigned short *>(srow + (kRowChunk + 4));
    const int a0 = d.e0 & ~7;
    tile_issue_loads_xs(a, d.tile, a0, d.e1, sval, slcol, sx, bar, tid);
    rows_tile<true, VEC, W, false, true>(a, sval, nullptr, slcol, sx, srow, bar, 0u, d.r0, d.r1 - d.r0, a0, d.e0, d.e1,
                                         tid);
  } else {
    int *scol = reinterpret_cast<int *>(sval + a.cap);
    int *srow = scol + a.cap;
    const int a0 = d.e0 & ~3;
    tile_issue_loads<TMA>(a, a0, d.e0, d.e1, sval, scol, bar, tid);
    rows_tile<TMA, VEC, W, ROT, false>(a, sval, scol, nullptr, nullptr, srow, bar, 0u, d.r0, d.r1 - d.r0, a0, d.e0, d.e1,
                                       tid);
  }
}

// register budgets: 40 for MEDIUM and for SHORT with 8 gathers per round, 32 for SHORT with <= 6; the staged-x MEDIUM
// kernel is limited to 4-5 resident CTAs by its shared memory and takes 48 (RB = 5) unless tuning bit 24 asks for 40
constexpr int rows_min_blocks(bool vec, int w, bool xs, int rb) { return rb ? rb : ((xs && vec) ? 5 : ((vec || w > 6) ? 6 : 8)); }
template <bool TMA, bool VEC, int W, bool ROT = false, bool XS = false, int RB = 0>
__global__ void __launch_bounds__(kThreads, rows_min_blocks(VEC, W, XS, RB)) k_spmv_rows(const SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  rows_cta<TMA, VEC, W, ROT, XS>(a, smem_raw, &bar);
}

// ---------------------------------------------------------------------------------------------------------------
// Fused halo loop: the same row kernel with the ordering of neighbouring GPUs' iterations inside it
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// one thread: blocks until every neighbour's flag has reached the epoch
__device__ __forceinline__ void halo_wait(const HaloSync &h) {
  const unsigned int k = ld_volatile_u32(h.state);
  if (k == 0u)
    return;
  const unsigned long long t0 = global_timer_ns();
  for (int j = 0; j < h.n_neigh; ++j) {
    while (ld_volatile_u32(h.wait[j]) < k) {
      if (ld_volatile_u32(h.state + 2) != 0u)
        return; // another CTA has already given up
      if (h.timeout_ns != 0ull && global_timer_ns() - t0 > h.timeout_ns) {
        atomicCAS(h.state + 2, 0u, 0x80000000u | k); // sticky: the host reads it in spmv_b200_halo_loop_sync
        return;
      }
      __nanosleep(64);
    }
  }
  __threadfence_system(); // the neighbours' pushed rows are ordered before their flag
}

// one thread, after all stores of the boundary row blocks: publish epoch + 1 to the neighbours (never after a timeout:
// a rank that multiplied a stale halo must not hand its rows on as if they were good; its neighbours then time out too)
__device__ __forceinline__ void halo_signal(const HaloSync &h) {
  const unsigned int k = ld_volatile_u32(h.state); // the epoch only changes here
  __threadfence_system();
  st_volatile_u32(h.state, k + 1u);
  if (ld_volatile_u32(h.state + 2) == 0u)
    for (int j = 0; j < h.n_neigh; ++j)
      st_volatile_u32(h.signal[j], k + 1u);
}

// The flag protocol sits in front of and behind the row kernel proper, where nothing of it is live, so the body keeps
// the register budget of k_spmv_rows (without that the compiler spent 55 registers and cost the kernel a resident
// CTA per SM: 3.63 ms against 3.30 ms per iteration on C5). The wait is satisfied long before it is reached in the
// steady state (the neighbour raised its flag early in ITS previous iteration), so waiting before the TMA copies are
// issued costs the boundary row blocks (a few percent of the CTAs) a few L2 round trips. In the staged-x form the wait
// has to come first anyway: the x segments the TMA copies read include the halo entries.
template <bool VEC, int W, bool XS>
__global__ void __launch_bounds__(kThreads, rows_min_blocks(VEC, W, XS, 0)) k_spmv_rows_halo(const SpmvArgs a, const HaloSync h) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  if ((int)blockIdx.x < h.n_boundary) {
    if (threadIdx.x == 0)
      halo_wait(h);
    __syncthreads();
  }
  rows_cta<true, VEC, W, false, XS>(a, smem_raw, &bar);
  if ((int)blockIdx.x < h.n_boundary) {
    __syncthreads(); // every row of this block has been stored (locally and into the neighbours' buffers)
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(h.state + 1, 1u) == (unsigned int)h.n_boundary - 1u) { // last boundary row block of the iteration
        st_volatile_u32(h.state + 1, 0u);
        halo_signal(h);
      }
    }
  }
}

__global__ void k_halo_wait(const HaloSync h) { halo_wait(h); }
__global__ void k_halo_signal(const HaloSync h) { halo_signal(h); }

// ---------------------------------------------------------------------------------------------------------------
// Staged-x form as a persistent ring: one wave of CTAs walks the row blocks, S stages of shared memory per CTA
// ---------------------------------------------------------------------------------------------------------------
// The one-row-block-per-CTA kernels alternate between waiting for their TMA copies and computing; with the 45-50 KB a
// staged-x row block of the 27-point stencil needs, four CTAs fit on an SM and the bytes in flight dip every time one of
// them computes, retires or starts (ncu: 61 % of the DRAM peak, the barrier the top stall). Here the copies of row block
// i + (S - 1) * grid are issued before row block i is summed, so S - 1 stages per CTA are always in flight, and
// everything a row block needs -- value, 16-bit local column indices, its segments of x, its row pointers -- arrives
// through the same mbarrier: the loop has no global load left except y (when beta != 0) and one CTA barrier per row
// block (the stage may be refilled once every thread has left it).
// Stage layout (offsets multiples of 128 bytes): sval[cap] | sx[xcap] | srow[kRingRowCap] (int32) | slcol[cap] (uint16).
constexpr int kRingStagesMax = 8;
constexpr int kRingRowCap = kSparseTileRows + 8; // SHORT / MEDIUM row blocks own at most kSparseTileRows rows

struct RingStage {
  double *sval, *sx;
  int *srow;
  unsigned short *slcol;
};
__device__ __forceinline__ RingStage ring_stage(unsigned char *base, int cap, int xcap) {
  RingStage st;
  st.sval = reinterpret_cast<double *>(base);
  st.sx = st.sval + ((cap + 15) & ~15);
  st.srow = reinterpret_cast<int *>(st.sx + xcap);
  st.slcol = reinterpret_cast<unsigned short *>(st.srow + ((kRingRowCap + 31) & ~31));
  return st;
}

// What the producer warp needs to issue the copies of one row block: its descriptor and (one segment per lane) its
// XDesc entry. Both are fetched ahead of their use (the descriptor two row blocks ahead, the segment table one), so that
// the two dependent global round trips are not paid between two row blocks.
struct RingFetch {
  TileDesc d;
  int nseg, line, off, end; // this lane's segment: lines [line, line + end - off) of x to line `off` of the stage
};
__device__ __forceinline__ void ring_fetch_segments(const SpmvArgs &a, RingFetch &f, int lane) {
  const XDesc *__restrict__ xd = a.xdesc + f.d.tile;
  f.nseg = __ldg(&xd->nseg);
  // (lanes beyond nseg read entries that exist -- the table has kXsegMax slots -- and ignore them)
  const int sl = lane < kXsegMax ? lane : kXsegMax - 1;
  f.line = __ldg(&xd->line[sl]);
  f.off = (int)__ldg(&xd->off[sl]);
  const int nxt = (int)__ldg(&xd->off[sl + 1 < kXsegMax ? sl + 1 : kXsegMax - 1]);
  const int nlines = __ldg(&xd->nlines); // (unconditional: a load predicated on nseg would wait for nseg first)
  f.end = lane + 1 < f.nseg ? nxt : nlines;
}

// producer warp: all copies of one row block into one stage (see tile_issue_loads_xs for the x segments); srow[0] will
// hold rowptr[r0 & ~3]; the four descriptor fields the consumers need go to sdesc. Everything written here with plain
// stores precedes lane 0's arrive on the stage's barrier, which the consumers wait for.
__device__ __forceinline__ void ring_issue(const SpmvArgs &a, const RingFetch &f, const RingStage &st, int *sdesc,
                                           unsigned long long *bar, int lane) {
  const TileDesc &d = f.d;
  if (lane == 0) {
    sdesc[0] = d.r0;
    sdesc[1] = d.r1;
    sdesc[2] = d.e0;
    sdesc[3] = d.e1;
  }
  const int a0 = d.e0 & ~7;
  const int span = d.e1 - a0;
  const long long avail = a.nnz - (long long)a0;
  int cnt = (span + 7) & ~7;
  if ((long long)cnt > avail)
    cnt = (int)(avail & ~7LL);
  const int rbase = d.r0 & ~3;
  const int rspan = d.r1 + 1 - rbase; // row pointers rbase .. r1
  int rcnt = (rspan + 3) & ~3;
  if (rcnt > a.m + 1 - rbase)
    rcnt = (a.m + 1 - rbase) & ~3;
  unsigned int xbytes = 0;
  if (lane < f.nseg) {
    long long elems = (long long)(f.end - f.off) * 16;
    const long long left = (long long)a.n - (long long)f.line * 16;
    if (elems > left)
      elems = left;
    xbytes = (unsigned int)(elems & ~1LL) * 8u;
    if (elems & 1)
      st.sx[(long long)f.off * 16 + elems - 1] = __ldg(a.x + (long long)f.line * 16 + elems - 1);
  }
  // what 16-byte copies cannot bring (ends of the arrays): plain loads
  for (int i = cnt + lane; i < span; i += 32) {
    st.sval[i] = ld_stream_f64(a.val + a0 + i);
    st.slcol[i] = __ldg(a.lcol + ((long long)a0 - a.lcol_base) + i);
  }
  for (int i = rcnt + lane; i < rspan; i += 32)
    st.srow[i] = __ldg(a.rowptr + rbase + i);
  unsigned int total = xbytes;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    total += __shfl_xor_sync(0xffffffffu, total, o);
  total += (unsigned int)cnt * 10u + (unsigned int)rcnt * 4u;
  __syncwarp();
  if (lane == 0) {
    mbar_arrive_expect_tx(bar, total);
    const unsigned long long pol = policy_stream(a.stream_keep);
    if (cnt > 0) {
      tma_bulk_g2s(st.sval, a.val + a0, (uint32_t)cnt * 8u, bar, pol);
      tma_bulk_g2s(st.slcol, a.lcol + ((long long)a0 - a.lcol_base), (uint32_t)cnt * 2u, bar, pol);
    }
    if (rcnt > 0)
      tma_bulk_g2s_nohint(st.srow, a.rowptr + rbase, (uint32_t)rcnt * 4u, bar);
  }
  __syncwarp();
  if (xbytes > 0)
    tma_bulk_g2s_nohint(st.sx + (long long)f.off * 16, a.x + (long long)f.line * 16, xbytes, bar);
}

// barrier of the kThreads consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }

// kThreads consumer threads (8 warps) and kRingProducers producer warps. Per stage two mbarriers: `full` (the producer's
// arrive plus the bytes of the copies) and `empty` (one arrive per consumer warp once it has read its last element of
// the stage). The producers run up to S row blocks ahead of the consumers; neither side ever waits at a CTA-wide
// barrier. One producer warp needs ~450 dependent instructions per row block (~2700 cycles: ncu showed it never waiting
// for a free stage while the consumers waited 35 % of their time for a full one), so the row blocks of a CTA alternate
// between two of them.
constexpr int kRingProducers = 2;
constexpr int kRingThreads = kThreads + 32 * kRingProducers;
constexpr int kRingConsumerWarps = kThreads / 32;

template <bool VEC, bool HALO>
__global__ void __launch_bounds__(kRingThreads, VEC ? 2 : 3) k_spmv_ring(const SpmvArgs a, const HaloSync h) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kRingStagesMax], empty[kRingStagesMax];
  __shared__ int sdesc[kRingStagesMax][4];
  const int tid = threadIdx.x, lane = tid & 31;
  const int S = a.ring_stages, G = (int)gridDim.x, nt = a.ntiles;
  if ((int)blockIdx.x >= nt)
    return;
  if (tid == 0)
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kRingConsumerWarps);
    }
  __syncthreads();
  if (tid >= kThreads) { // ---- producer warps: warp p issues the CTA's row blocks p, p + P, p + 2P, ... ----
    const int p = (tid - kThreads) >> 5, step = kRingProducers * G;
    RingFetch cur, nxt; // cur: complete; nxt: descriptor only
    int i = (int)blockIdx.x + p * G;
    if (i >= nt)
      return;
    cur.d = load_desc(a.desc, i);
    if (i + step < nt)
      nxt.d = load_desc(a.desc, i + step);
    ring_fetch_segments(a, cur, lane);
    int s = p; // stage of the CTA's k-th row block: k mod S, barrier phase (k / S) & 1
    uint32_t eparity = 1u; // a fresh barrier passes a wait on parity 1: the first S stages are free
    while (s >= S) {
      s -= S;
      eparity ^= 1u;
    }
    bool halo_seen = false; // the neighbours' flags are polled once per producer warp and launch
    for (; i < nt; i += step) {
      mbar_wait(&empty[s], eparity);
      if (HALO && i < h.n_boundary && !halo_seen) { // the x segments of a boundary row block include halo entries
        if (lane == 0)
          halo_wait(h);
        __syncwarp();
        halo_seen = true;
      }
      ring_issue(a, cur, ring_stage(smem_raw + (size_t)s * a.ring_stage_bytes, a.cap, a.xcap), sdesc[s], &full[s], lane);
      if (i + step < nt) {
        cur.d = nxt.d;
        ring_fetch_segments(a, cur, lane); // its descriptor was requested one round ago
        if (i + 2 * step < nt)
          nxt.d = load_desc(a.desc, i + 2 * step);
      }
      s += kRingProducers;
      while (s >= S) {
        s -= S;
        eparity ^= 1u;
      }
    }
    return;
  }
  // ---- consumers ----
  int s = 0;
  uint32_t parity = 0u; // stage and barrier phase of the row block being summed (no division by the run-time S)
  for (int i = (int)blockIdx.x; i < nt; i += G) {
    const RingStage st = ring_stage(smem_raw + (size_t)s * a.ring_stage_bytes, a.cap, a.xcap);
    mbar_wait(&full[s], parity);
    const int r0 = sdesc[s][0], nrows = sdesc[s][1] - r0, e0 = sdesc[s][2], e1 = sdesc[s][3];
    const int a0 = e0 & ~7;
    const int *srow = st.srow + (r0 & 3);
    const int lv = VEC ? lanes_log2(e1 - e0, nrows, a.vec_div) : 0; // lanes per row, as in rows_tile
    const int Gr = kThreads >> lv, g = tid >> lv, l = tid & ((1 << lv) - 1);
    const uint32_t sval_s = smem_u32(st.sval), sx_s = smem_u32(st.sx), scol_s = smem_u32(st.slcol);
    // SPMV_B200_HALO_ALIGN_PUSH: a row block whose rows are pushed to other GPUs deals its rows to the lane groups by
    // absolute row index, so that the 16 (or 8, 32) rows a warp finishes together are one aligned line of the
    // destination: a peer store of a full 128-byte line needs no read-modify-write of partial sectors on the receiving
    // side (all-gather push at 8 GPUs: 0.954 -> 0.771 ms per iteration). It may cost the block a second pass, so the
    // caller asks for it only where the link and not the SpMV bounds the iteration (2 GPUs: 1.35 -> 1.47 ms with it).
    int shift = 0;
    if (a.push.count && a.push.align_rows) {
      bool pushed = false;
      for (int j = 0; j < a.push.count; ++j)
        pushed |= r0 < a.push.row_hi[j] && r0 + nrows > a.push.row_lo[j];
      shift = pushed ? (r0 & 15) : 0;
    }
    for (int rb = -shift; rb < nrows; rb += Gr) {
      const int r = rb + g;
      const bool act = r >= 0 && r < nrows;
      int k = act ? srow[r] - a0 + l : 0;
      const int e = act ? srow[r + 1] - a0 : 0;
      const double yv = (a.read_y && act && l == 0) ? a.y[r0 + r] : 0.0; // in flight during the row sum
      double sum = 0.0;
      for (; k < e; k += 4 << lv) {
        double xv[4], vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = k + (j << lv);
          xv[j] = 0.0;
          vv[j] = 0.0;
          slot_xs(xv[j], vv[j], kk < e, scol_s + 2u * (uint32_t)kk, sval_s + 8u * (uint32_t)kk, sx_s);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sum = fma(vv[j], xv[j], sum);
      }
      if (VEC) {
        for (int off = (1 << lv) >> 1; off > 0; off >>= 1)
          sum += __shfl_down_sync(0xffffffffu, sum, off, 1 << lv);
      }
      if (act && l == 0) // cli/verification.cpp:64  y[i] = alpha * y0 + beta * y[i]
        emit_y<true>(a.y, a.push, r0 + r, a.alpha * sum + a.beta * yv);
    }
    __syncwarp();
    if (lane == 0)
      mbar_arrive(&empty[s]); // this warp has left stage s
    if (HALO && i < h.n_boundary && i + G >= h.n_boundary) { // this CTA's last boundary row block
      // One system-scope fence per CTA and launch, not per row block: it waits for the NVLink stores of the pushed
      // rows to be acknowledged (microseconds), and with one per boundary row block the single-launch loop ran 3 %
      // behind the multi-launch one at two GPUs. The barrier orders every consumer warp's stores of all this CTA's
      // boundary row blocks before thread 0's fence; the counter counts CTAs that own boundary row blocks.
      consumers_sync();
      if (tid == 0) {
        __threadfence_system();
        const unsigned int owners = (unsigned int)(h.n_boundary < G ? h.n_boundary : G);
        if (atomicAdd(h.state + 1, 1u) == owners - 1u) { // last CTA of the iteration to finish its boundary row blocks
          st_volatile_u32(h.state + 1, 0u);
          halo_signal(h);
        }
      }
    }
    if (++s == S) {
      s = 0;
      parity ^= 1u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// MIXED tiles
// ---------------------------------------------------------------------------------------------------------------
// deterministic CTA-wide sum of sval[lo, hi): strided per-thread partials, xor-shuffle tree, warp partials in order
template <int NT>
__device__ __forceinline__ double block_sum(const double *__restrict__ sval, int lo, int hi, double *swarp, int tid) {
  double s = 0.0;
  for (int k = lo + tid; k < hi; k += NT)
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
    for (int w = 0; w < NT / 32; ++w)
      total += swarp[w];
  }
  __syncthreads();
  return total; // valid in thread 0
}

// NT threads per CTA (64 / 128 / 256): small tiles run with small CTAs so that more independent CTAs are resident per
// SM; the phases of one CTA (descriptor, tile, gathers, row sums) are dependent round trips to memory, and the number of
// CTAs in different phases is what hides them.
template <bool TMA, int NT>
__global__ void __launch_bounds__(NT) k_spmv_mixed(const SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ double swarp[NT / 32];
  __shared__ int nlong, nmid;
  double *sval = reinterpret_cast<double *>(smem_raw);
  int *scol = reinterpret_cast<int *>(sval + a.cap);
  int *srow = scol + a.cap;
  const int qcap = a.cap / (kSerialMax + 1) + 8;
  int *smid = srow + (kRowChunk + 1); // rows of kSerialMax+1 .. kGroupMax products in the current chunk
  int *slong = smid + qcap;           // longer rows

  const int tid = threadIdx.x;
  const TileDesc d = load_desc(a.desc, blockIdx.x);
  const int t = d.tile;
  const int r0 = d.r0, r1 = d.r1, e0 = d.e0, e1 = d.e1;
  const int a0 = e0 & ~3;
  const bool split_begin = (d.flags & 1) != 0;
  const bool split_end = (d.flags & 2) != 0;
  // the last owned row continues in the next tile: its part in this tile is the tail fragment
  const bool has_tail = split_end && (r1 > r0);
  const int nrows = (r1 - r0) - (has_tail ? 1 : 0);

  tile_issue_loads<TMA, NT>(a, a0, e0, e1, sval, scol, &bar, tid);
  const int tail_start = d.tail_start; // first element of the tail fragment
  // while the tile is in flight: row pointers of the first chunk of rows and the y values pass 1 will need
  constexpr int kPre = kRowChunk / NT; // pass-1 iterations of a full chunk
  double ypre[kPre];
  {
    const int nr0 = nrows < kRowChunk ? nrows : kRowChunk;
    for (int i = tid; i <= nr0; i += NT)
      srow[i] = __ldg(a.rowptr + r0 + i);
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int r = tid + j * NT;
      ypre[j] = (a.read_y && r < nr0) ? a.y[r0 + r] : 0.0;
    }
  }
  __syncthreads();
  if (TMA)
    mbar_wait(&bar, 0);

  // products in place. All gathers of a batch of 8 strided elements are issued before the first multiply; the values
  // and indices are read into registers first because the compiler cannot prove that sval and scol do not alias.
  {
    const int end = e1 - a0;
    for (int base = (e0 - a0) + tid; base < end; base += 8 * NT) {
      int c[8];
      double xv[8], v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = base + j * NT;
        c[j] = (i < end) ? scol[i] : -1;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        xv[j] = (c[j] >= 0) ? gather_x(a.x, c[j], a.gather_na) : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = base + j * NT;
        v[j] = (i < end) ? sval[i] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = base + j * NT;
        if (i < end)
          sval[i] = v[j] * xv[j];
      }
    }
  }
  __syncthreads();

  if (split_begin) { // leading elements belong to a row that started in an earlier tile
    const double s = block_sum<NT>(sval, e0 - a0, d.head_end - a0, swarp, tid);
    if (tid == 0)
      a.partials[2 * (size_t)t] = s;
  }
  if (has_tail) {
    const double s = block_sum<NT>(sval, tail_start - a0, e1 - a0, swarp, tid);
    if (tid == 0)
      a.partials[2 * (size_t)t + 1] = s;
  }

  const int warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 3, gl = tid & 7; // 8-lane groups for rows of medium length
  for (int cb = 0; cb < nrows; cb += kRowChunk) {
    int nr = nrows - cb;
    if (nr > kRowChunk)
      nr = kRowChunk;
    __syncthreads(); // srow / queues reuse
    if (cb > 0) {
      for (int i = tid; i <= nr; i += NT)
        srow[i] = __ldg(a.rowptr + r0 + cb + i);
    }
    if (tid == 0) {
      nmid = 0;
      nlong = 0;
    }
    __syncthreads();
    // pass 1: one thread per row sums rows of <= kSerialMax products; longer rows are queued by length class
    // (the value of a row does not depend on its queue position, only on its own fixed summation order)
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int r = tid + j * NT;
      if (r >= nr)
        break;
      const int s = srow[r] - a0, e = srow[r + 1] - a0;
      const int len = e - s;
      if (len <= kSerialMax) {
        const double yv = cb == 0 ? ypre[j] : (a.read_y ? a.y[r0 + cb + r] : 0.0);
        double sum = 0.0;
        for (int k = s; k < e; ++k)
          sum += sval[k];
        emit_y(a.y, a.push, r0 + cb + r, a.alpha * sum + a.beta * yv);
      } else if (len <= kGroupMax) {
        smid[atomicAdd(&nmid, 1)] = r;
      } else {
        slong[atomicAdd(&nlong, 1)] = r;
      }
    }
    __syncthreads();
    // pass 2a: eight lanes per queued row of medium length
    const int nm = nmid;
    for (int ib = 0; ib < nm; ib += NT / 8) {
      const int i = ib + grp;
      const bool act = i < nm;
      const int r = act ? smid[i] : 0;
      const int s = srow[r] - a0, e = act ? srow[r + 1] - a0 : s;
      const double yv = (act && gl == 0 && a.read_y) ? a.y[r0 + cb + r] : 0.0;
      double sum = 0.0;
      for (int k = s + gl; k < e; k += 8)
        sum += sval[k];
      sum += __shfl_down_sync(0xffffffffu, sum, 4, 8);
      sum += __shfl_down_sync(0xffffffffu, sum, 2, 8);
      sum += __shfl_down_sync(0xffffffffu, sum, 1, 8);
      if (act && gl == 0)
        emit_y(a.y, a.push, r0 + cb + r, a.alpha * sum + a.beta * yv);
    }
    // pass 2b: one warp per queued long row
    const int nl = nlong;
    for (int i = warp; i < nl; i += NT / 32) {
      const int r = slong[i];
      const int s = srow[r] - a0, e = srow[r + 1] - a0;
      const double yv = (lane == 0 && a.read_y) ? a.y[r0 + cb + r] : 0.0;
      double sum = 0.0;
      for (int k = s + lane; k < e; k += 32)
        sum += sval[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
      if (lane == 0)
        emit_y(a.y, a.push, r0 + cb + r, a.alpha * sum + a.beta * yv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Direct form: one warp per row block, no shared memory (matrices whose x gathers do not coalesce across rows)
// ---------------------------------------------------------------------------------------------------------------
// A gather in flight holds a 128-byte line of L1, and L1 is what the shared-memory carve-out leaves of the 256 KB
// unified array: with tiles staged in shared memory the gather rate of an SM drops by 2-4x (profiles/, gather bound
// against carve-out). Here value / colindex go straight into registers and rows are delimited by the row-start bit
// flags of the analysis instead of row pointers.
//
// Element ownership is lane-contiguous: a warp walks its block in 128-element windows aligned to 32 elements, and
// gather instruction j of a window serves elements 32j + lane. Consecutive lanes then gather for consecutive non-zeros,
// which in a sorted row are neighbouring columns, so one instruction touches every 128-byte line of x once (with four
// consecutive elements per lane the four gather instructions of a window each touched the lines of the whole window:
// 456 M sectors for 268 M non-zeros on the R-MAT matrix). The 32 row-start flags of sub-window j are exactly one word of
// the flag array, known to every lane, so the segmented sums need no flag shuffles:
//   * every lane keeps an open sum `acc` of its own products since the last row start (no warp reduction per window:
//     a window inside one long row costs four adds per lane);
//   * a sub-window with row starts closes the open row (one butterfly over acc + the products in front of the first
//     start), finishes the rows that begin and end inside it with a segmented scan whose predicate is a bit test on the
//     flag word, and leaves the products behind the last start in acc;
//   * the row a finished sum belongs to is nz_rows[nzbase + ordinal of its row start]; the open sums at the two ends of
//     a block are the fragments of rows split across blocks (partials -> k_fixup), or the block's last row.
// No barrier, no atomics; the order of every addition is fixed by the block geometry (bitwise reproducible).
// Analogue of the reference's flat / merge-path kernels (src/acc/hip-flat/flat_imp_one_pass.hpp:15-77,
// benchmark/merge-path/merge_path_reduction.h:80-136) without atomicAdd and without the per-element row search.
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// loads of the direct kernel as volatile asm: issued in program order (see the element loop)
__device__ __forceinline__ int ld_cs_s32(const int *p) {
  int v;
  asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_cs_f64(const double *p) {
  double v;
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_nc_f64(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_nc_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void finish_row(const SpmvArgs &a, int row, double sum) {
  const double yv = a.read_y ? a.y[row] : 0.0; // cli/verification.cpp:64: y is read even when beta == 0
  emit_y(a.y, a.push, row, a.alpha * sum + a.beta * yv);
}

// one 32-bit field of a tile descriptor, fetched again where it is needed (the asm keeps the compiler from merging the
// load with the one at the top of the kernel and carrying the value in a register through the element loop)
__device__ __forceinline__ int desc_field(const TileDesc *__restrict__ desc, int i, int field) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(reinterpret_cast<const int *>(desc + i) + field));
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// MINB = resident CTAs per SM the register allocation aims for (8: 32 registers, 64 warps; 6: 40 registers, 48 warps)
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_spmv_warp(const SpmvArgs a) {
  const int lane = threadIdx.x & 31;
  const int ti = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (ti >= a.ntiles)
    return;
  // Only e0 / e1 / nzbase stay in registers across the element loop; the other descriptor fields are needed once per
  // block (its first row start, its end) and are fetched again there (an L1 hit) instead of occupying four registers.
  int e0, e1, nzi; // nzi: index into nz_rows of the next row start (starts at the number of non-empty rows in front of
                   // this block = field head_end of the direct descriptors)
  {
    const TileDesc d = load_desc(a.desc, ti);
    e0 = d.e0, e1 = d.e1, nzi = d.head_end;
    // Everything a finished row needs later (its y, its id, the row pointers of the empty-row pass) is requested now,
    // without a destination register, so that those dependent accesses hit L1 instead of adding round trips to the
    // chain descriptor -> value/colindex/flags -> x gathers. One 128-byte line per lane and array (blocks own <= T rows).
    const int nrows = d.r1 - d.r0;
    for (int i = lane * 16; i < nrows; i += 32 * 16)
      prefetch_l1(a.y + d.r0 + i);
    for (int i = lane * 32; i <= nrows; i += 32 * 32) {
      prefetch_l1(a.rowptr + d.r0 + i);
      prefetch_l1(a.nz_rows + nzi + i);
    }
  }

  double acc = 0.0;     // this lane's products since the last row start
  bool started = false; // a row start has been met
  // flag words of the four sub-windows of a window: lane l fetches word (l & 3). Requested one window ahead, so that the
  // load never sits behind the wait for the gathers in the instruction stream (the array is padded by a window).
  unsigned wnext = ld_nc_u32(a.row_start_bits + ((e0 & ~31) >> 5) + (lane & 3));
  for (int rb = e0 & ~31; rb < e1; rb += 128) {
    // Issue order: flag word of the next window, colindex, x gathers, value. The column registers are dead once
    // the gathers are issued, so the value loads reuse them: 16 data registers at the peak instead of 20. Elements of
    // the neighbouring blocks (first / last window) are predicated off: (unsigned)(i - e0) < span <=> e0 <= i < e1.
    const unsigned span = (unsigned)(e1 - e0);
    const int i0 = rb + lane;
    unsigned wq = wnext;
    wnext = ld_nc_u32(a.row_start_bits + (rb >> 5) + 4 + (lane & 3));
    double p[4], v[4];
    {
      int c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        c[j] = (unsigned)(i0 + 32 * j - e0) < span ? ld_cs_s32(a.col + i0 + 32 * j) : -1;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        p[j] = c[j] >= 0 ? ld_nc_f64(a.x + c[j]) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      v[j] = (unsigned)(i0 + 32 * j - e0) < span ? ld_cs_f64(a.val + i0 + 32 * j) : 0.0;
    // row starts of the neighbouring blocks do not count; a sub-window that has row starts broadcasts its word by shuffle
    {
      const int lo = e0 - (rb + 32 * (lane & 3)), hi = e1 - (rb + 32 * (lane & 3));
      unsigned keep = hi <= 0 ? 0u : (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u));
      if (lo > 0)
        keep &= lo >= 32 ? 0u : ~((1u << lo) - 1u);
      wq &= keep;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      p[j] *= v[j];
    if (!__any_sync(0xffffffffu, wq != 0u)) { // the whole window lies inside one row (long rows): four adds per lane
      acc += (p[0] + p[1]) + (p[2] + p[3]);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned wj = __shfl_sync(0xffffffffu, wq, j);
      if (wj == 0u) { // warp-uniform
        acc += p[j];
        continue;
      }
      // the open row ends in front of the first start of this sub-window
      const int first = __ffs(wj) - 1;
      const double open = warp_sum(acc + (lane < first ? p[j] : 0.0));
      if (lane == 0) {
        if (started)
          finish_row(a, __ldg(a.nz_rows + nzi - 1), open);
        else if (desc_field(a.desc, ti, 7) & 1) // in front of the block's first row start: fragment of a split row
          a.partials[2 * (size_t)desc_field(a.desc, ti, 6)] = open;
      }
      const int cnt = __popc(wj);
      int last = first;
      if (cnt > 1) {
        // rows that begin and end inside the sub-window: segmented inclusive scan over the lanes; lane L takes the
        // value of lane L - off iff no row starts in (L - off, L], which is a bit test on the flag word
        double s = p[j];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, s, off);
          if (lane >= off && ((wj >> (lane - off + 1)) & ((1u << off) - 1u)) == 0u)
            s += o;
        }
        last = 31 - __clz(wj);
        // a segment ends at lane L iff a row starts at L + 1; its ordinal is that of the last start at or below L
        if (lane >= first && lane < last && ((wj >> (lane + 1)) & 1u))
          finish_row(a, __ldg(a.nz_rows + nzi + __popc(wj & ((2u << lane) - 1u)) - 1), s);
      }
      acc = lane >= last ? p[j] : 0.0; // products behind the last start stay open
      nzi += cnt;
      started = true;
    }
  }
  // rows without elements only need the epilogue
  {
    const int r1 = desc_field(a.desc, ti, 1);
    for (int r = desc_field(a.desc, ti, 0) + lane; r < r1; r += 32)
      if (__ldg(a.rowptr + r) == __ldg(a.rowptr + r + 1))
        finish_row(a, r, 0.0);
  }
  const double open = warp_sum(acc); // the open sum at the end of the block
  if (lane == 0) {
    const int t = desc_field(a.desc, ti, 6), flags = desc_field(a.desc, ti, 7);
    if (!started) {
      if (flags & 1)
        a.partials[2 * (size_t)t] = open; // the whole block lies inside one row
    } else if (flags & 2) {
      a.partials[2 * (size_t)t + 1] = open;
    } else {
      finish_row(a, __ldg(a.nz_rows + nzi - 1), open);
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
    emit_y(f.y, f.push, row, f.alpha * sum + f.beta * yv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int cap_for(const spmv_b200_plan *p, int kind) {
  // a tile streams at most T + (longest row that may hang over its trailing boundary) - 1 elements, plus alignment slack
  return p->T + (kind == SPMV_B200_KIND_SHORT ? ((p->short_max + 3) & ~3) : p->medium_max) + 8;
}

static size_t smem_for(const spmv_b200_plan *p, int kind) {
  const int cap = cap_for(p, kind);
  size_t b = (size_t)cap * 12 + sizeof(int) * (kRowChunk + 1);
  if (kind == SPMV_B200_KIND_MIXED) // two row queues
    b += 2 * sizeof(int) * ((size_t)cap / (kSerialMax + 1) + 8);
  return (b + 15) & ~(size_t)15;
}

// staged-x form: the streams start at an element index that is a multiple of 8 (16-byte alignment of the 2-byte
// indices), hence 8 more elements of slack; sx holds the plan's largest set of lines
static int xs_cap_for(const spmv_b200_plan *p, int kind) { return cap_for(p, kind) + 8; }
static int xs_xcap(const spmv_b200_plan *p) { return 16 * ((p->xlines + 7) & ~7); }
static size_t xs_smem_for(const spmv_b200_plan *p, int kind) {
  const size_t b = (size_t)xs_cap_for(p, kind) * 10 + sizeof(double) * (size_t)xs_xcap(p) + sizeof(int) * (kRowChunk + 4);
  return (b + 15) & ~(size_t)15;
}

typedef void (*RowsKernel)(const SpmvArgs);
typedef void (*HaloKernel)(const SpmvArgs, const HaloSync);

template <typename K> static int set_smem(K kernel, size_t bytes);

// ---- staged-x ring (persistent CTAs): geometry ----
static size_t ring_stage_bytes(const spmv_b200_plan *p, int kind) {
  const size_t cap = (size_t)xs_cap_for(p, kind);
  size_t b = ((cap + 15) & ~(size_t)15) * 8 + (size_t)xs_xcap(p) * 8 + (size_t)((kRingRowCap + 31) & ~31) * 4 + cap * 2;
  return (b + 127) & ~(size_t)127;
}
static HaloKernel ring_kernel(int kind, bool halo) {
  if (kind == SPMV_B200_KIND_SHORT)
    return halo ? k_spmv_ring<false, true> : k_spmv_ring<false, false>;
  return halo ? k_spmv_ring<true, true> : k_spmv_ring<true, false>;
}
// CTAs per SM and stages per CTA for the plan's tile size: two CTAs (three for SHORT row blocks, whose kernel fits 64
// registers) with as many stages as fit, at least two; SPMV_B200_RING_CTAS / SPMV_B200_RING_STAGES override
static bool ring_geometry(const spmv_b200_plan *p, int kind, int *ctas, int *stages) {
  int max_optin = 0;
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device) != cudaSuccess)
    return false;
  const size_t per_sm = 228 * 1024, stage = ring_stage_bytes(p, kind);
  const char *ec = getenv("SPMV_B200_RING_CTAS"), *es = getenv("SPMV_B200_RING_STAGES");
  const int cdef = kind == SPMV_B200_KIND_SHORT ? 3 : 2;
  for (int c = (ec && atoi(ec) > 0) ? atoi(ec) : cdef; c >= 1; --c) {
    const size_t budget = per_sm / (size_t)c - 2048; // 1 KB per CTA is reserved by the system, some static shared memory
    int s = (int)(budget / stage);
    if ((size_t)s * stage > (size_t)max_optin)
      s = (int)((size_t)max_optin / stage);
    if (es && atoi(es) > 0 && atoi(es) < s)
      s = atoi(es);
    if (s > kRingStagesMax)
      s = kRingStagesMax;
    if (s >= 2) {
      int occ = 0; // registers may allow fewer CTAs than the shared memory does (the query needs the raised limit)
      if (set_smem(ring_kernel(kind, true), (size_t)s * stage) != SPMV_B200_OK ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel(kind, true), kRingThreads, (size_t)s * stage) !=
              cudaSuccess ||
          occ < c) {
        cudaGetLastError();
        continue;
      }
      *ctas = c;
      *stages = s;
      return true;
    }
  }
  return false;
}
struct RowsVariant {
  RowsKernel tma, plain; // tiles by TMA / by plain loads (misaligned value / colindex)
  RowsKernel xs;         // staged-x form (nullptr: not built for this variant)
  RowsKernel xs40;       // the same with a 40-register budget (tuning bit 24; nullptr: same as xs)
  HaloKernel halo, halo_xs;
  const char *name;
};
// variant tables (index = option bits, 0 = default); every entry is a separate instantiation of the row kernels
static const RowsVariant kShortVariants[] = {
    {k_spmv_rows<true, false, 6>, k_spmv_rows<false, false, 6>, k_spmv_rows<true, false, 6, false, true>, nullptr,
     k_spmv_rows_halo<false, 6, false>, k_spmv_rows_halo<false, 6, true>, "W6"},
    {k_spmv_rows<true, false, 8>, k_spmv_rows<false, false, 8>, k_spmv_rows<true, false, 8, false, true>, nullptr,
     k_spmv_rows_halo<false, 8, false>, k_spmv_rows_halo<false, 8, true>, "W8"},
    {k_spmv_rows<true, false, 4>, k_spmv_rows<false, false, 4>, nullptr, nullptr, nullptr, nullptr, "W4"},
};
static const RowsVariant kMediumVariants[] = {
    {k_spmv_rows<true, true, 8>, k_spmv_rows<false, true, 8>, k_spmv_rows<true, true, 8, false, true>,
     k_spmv_rows<true, true, 8, false, true, 6>, k_spmv_rows_halo<true, 8, false>, k_spmv_rows_halo<true, 8, true>, "W8"},
    {k_spmv_rows<true, true, 4>, k_spmv_rows<false, true, 4>, k_spmv_rows<true, true, 4, false, true>,
     k_spmv_rows<true, true, 4, false, true, 6>, nullptr, nullptr, "W4"},
};
// MEDIUM kernels with the slot rotation (plans with irregular gathers); same order as kMediumVariants
static const RowsVariant kMediumVariantsRot[] = {
    {k_spmv_rows<true, true, 8, true>, k_spmv_rows<false, true, 8, true>, nullptr, nullptr, nullptr, nullptr, "W8r"},
    {k_spmv_rows<true, true, 4, true>, k_spmv_rows<false, true, 4, true>, nullptr, nullptr, nullptr, nullptr, "W4r"},
};
constexpr int kNumShortVariants = sizeof(kShortVariants) / sizeof(kShortVariants[0]);
constexpr int kNumMediumVariants = sizeof(kMediumVariants) / sizeof(kMediumVariants[0]);

static RowsKernel mixed_kernel(bool tma, int threads) {
  switch (threads) {
  case 64: return tma ? k_spmv_mixed<true, 64> : k_spmv_mixed<false, 64>;
  case 128: return tma ? k_spmv_mixed<true, 128> : k_spmv_mixed<false, 128>;
  default: return tma ? k_spmv_mixed<true, 256> : k_spmv_mixed<false, 256>;
  }
}

// slot rotation of the MEDIUM kernel: only where gathers do not coalesce across rows anyway (C3: 1.836 ms against
// 1.882 ms); on a stencil it breaks the coalescing of neighbouring rows (C5s: 1.08 ms against 0.97 ms). Tuning bit 26
// switches it off.
static bool rotate_slots(const spmv_b200_plan *p) { return p->irregular && !((p->flags >> 26) & 1u); }

static const RowsVariant &rows_variant(const spmv_b200_plan *p, int kind) {
  if (kind == SPMV_B200_KIND_SHORT)
    return kShortVariants[p->variant_short];
  return rotate_slots(p) ? kMediumVariantsRot[p->variant_medium] : kMediumVariants[p->variant_medium];
}

// The dynamic shared-memory limit is an attribute of the kernel function, not of a plan: several plans with different
// tile sizes may be alive at once, so the limit is only ever raised (per device).
template <typename K> static int set_smem(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> limit;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = limit[std::make_pair(reinterpret_cast<const void *>(kernel), dev)];
    if (bytes <= cur)
      bytes = cur;
    else
      cur = bytes;
  }
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // development knob: SPMV_B200_CARVEOUT=<percent of the unified L1/shared array given to shared memory>
  if (const char *env = getenv("SPMV_B200_CARVEOUT")) {
    const int pct = atoi(env);
    if (pct >= 0 && pct <= 100)
      B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  }
  return SPMV_B200_OK;
}

int kernels_configure(spmv_b200_plan *p) {
  p->cap = cap_for(p, SPMV_B200_KIND_MIXED);
  p->smem_bytes = smem_for(p, SPMV_B200_KIND_MIXED);
  int dev = 0, max_optin = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // every kernel that may be launched must fit the opt-in shared-memory limit of the device: say so here, in words,
  // instead of letting cudaFuncSetAttribute fail with an opaque error
  for (int k = 0; k < 3; ++k)
    if (smem_for(p, k) > (size_t)max_optin) {
      set_error("tile_nnz too large for the shared memory of this device");
      return SPMV_B200_ERR_ARG;
    }
  p->device = dev;
  {
    // matrices whose streams fit the L2 with room to spare are kept there between SpMVs (normal policy instead of
    // evict-first): 67 MB stand-in for largebasis 15.2 -> 13.2 us; at 345 MB (boneS10) the same switch costs 19 %
    // (profiles/r2_sweep_small_matrices_l2_policy.jsonl). Tuning bit 27 inverts the choice.
    int l2 = 0;
    B200_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    const bool fits = (double)p->nnz * 12.0 <= 0.6 * (double)l2;
    p->stream_keep = (int)(fits != (((p->flags >> 27) & 1u) != 0));
  }
  if (p->flags & SPMV_B200_FLAG_L2_PERSIST_X) {
    int persist_max = 0, window_max = 0;
    B200_CUDA(cudaDeviceGetAttribute(&persist_max, cudaDevAttrMaxPersistingL2CacheSize, dev));
    B200_CUDA(cudaDeviceGetAttribute(&window_max, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    if (persist_max > 0) // device-wide carve-out of L2 for persisting lines (released again by kernels_release)
      B200_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_max));
    p->persist_bytes = (size_t)persist_max;
    p->max_window_bytes = (size_t)window_max;
  }
  // CTA size of the MIXED kernel: 8 items per thread (one batch of gathers), option bits 20-21 override (tuning)
  p->mixed_threads = p->T <= 512 ? 64 : (p->T <= 1024 ? 128 : 256);
  switch ((p->flags >> 20) & 3u) {
  case 1: p->mixed_threads = 64; break;
  case 2: p->mixed_threads = 128; break;
  case 3: p->mixed_threads = 256; break;
  default: break;
  }
  p->variant_short = (int)((p->flags >> 8) & 0xf);
  p->variant_medium = (int)((p->flags >> 12) & 0xf);
  if (p->variant_short >= kNumShortVariants || p->variant_medium >= kNumMediumVariants) {
    set_error("unknown kernel variant in the option flags");
    return SPMV_B200_ERR_ARG;
  }
  int rc;
  const RowsVariant &vs = kShortVariants[p->variant_short], &vm = kMediumVariants[p->variant_medium];
  const RowsVariant &vr = kMediumVariantsRot[p->variant_medium];
  if ((rc = set_smem(vs.tma, smem_for(p, SPMV_B200_KIND_SHORT))) ||
      (rc = set_smem(vs.plain, smem_for(p, SPMV_B200_KIND_SHORT))) ||
      (rc = set_smem(vm.tma, smem_for(p, SPMV_B200_KIND_MEDIUM))) ||
      (rc = set_smem(vm.plain, smem_for(p, SPMV_B200_KIND_MEDIUM))) ||
      (rc = set_smem(vr.tma, smem_for(p, SPMV_B200_KIND_MEDIUM))) ||
      (rc = set_smem(vr.plain, smem_for(p, SPMV_B200_KIND_MEDIUM))) ||
      (rc = set_smem(mixed_kernel(true, p->mixed_threads), smem_for(p, SPMV_B200_KIND_MIXED))) ||
      (rc = set_smem(mixed_kernel(false, p->mixed_threads), smem_for(p, SPMV_B200_KIND_MIXED))))
    return rc;
  if (vs.halo && (rc = set_smem(vs.halo, smem_for(p, SPMV_B200_KIND_SHORT))))
    return rc;
  if (vm.halo && (rc = set_smem(vm.halo, smem_for(p, SPMV_B200_KIND_MEDIUM))))
    return rc;
  return SPMV_B200_OK;
}

// second half, after the analysis has decided on the staged-x form (analysis_xstage): drops the form again if its
// kernel variant was not built or its tile does not fit the shared memory of the device
int kernels_configure_xs(spmv_b200_plan *p) {
  if (!p->xstage)
    return SPMV_B200_OK;
  const int kind = p->count[SPMV_B200_KIND_SHORT] == p->ntiles ? SPMV_B200_KIND_SHORT : SPMV_B200_KIND_MEDIUM;
  const RowsVariant &v = rows_variant(p, kind);
  int max_optin = 0;
  B200_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device));
  if (!v.xs || xs_smem_for(p, kind) > (size_t)max_optin) {
    p->xstage = false;
    return SPMV_B200_OK;
  }
  int rc;
  if ((rc = set_smem(v.xs, xs_smem_for(p, kind))))
    return rc;
  if (v.xs40 && (rc = set_smem(v.xs40, xs_smem_for(p, kind))))
    return rc;
  if (v.halo_xs && (rc = set_smem(v.halo_xs, xs_smem_for(p, kind))))
    return rc;
  p->smem_bytes = xs_smem_for(p, kind);
  // the persistent ring form of the same kernels (tuning bit 25 switches it off)
  p->ring_ctas = p->ring_stages = 0;
  if (!((p->flags >> 25) & 1u) && ring_geometry(p, kind, &p->ring_ctas, &p->ring_stages)) {
    const size_t bytes = (size_t)p->ring_stages * ring_stage_bytes(p, kind);
    if ((rc = set_smem(ring_kernel(kind, false), bytes)) || (rc = set_smem(ring_kernel(kind, true), bytes)))
      return rc;
    p->smem_bytes = bytes;
  }
  return SPMV_B200_OK;
}

// gives the device-wide L2 carve-out back (a plan with SPMV_B200_FLAG_L2_PERSIST_X took it): other plans of the process
// would otherwise run with a fraction of the L2 (measured: C3 1.83 ms -> 2.34 ms after such a plan had existed)
void kernels_release(spmv_b200_plan *p) {
  if ((p->flags & SPMV_B200_FLAG_L2_PERSIST_X) && p->persist_bytes > 0) {
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    p->persist_bytes = 0;
  }
}

// Launch with an optional L2 access-policy window that marks x as persisting (gathers of x are the only re-used
// data of an SpMV; value / colindex are streamed with evict-first). The window is a per-launch attribute, so the
// caller's stream state is not modified.
static cudaError_t launch_spmv(RowsKernel k, int grid, size_t smem, cudaStream_t stream, const SpmvArgs &a,
                               const spmv_b200_plan *p, int threads = kThreads) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if ((p->flags & SPMV_B200_FLAG_L2_PERSIST_X) && p->persist_bytes > 0 && p->n > 0) {
    size_t bytes = sizeof(double) * (size_t)p->n;
    if (bytes > p->max_window_bytes)
      bytes = p->max_window_bytes;
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = const_cast<double *>(a.x);
    attr[0].val.accessPolicyWindow.num_bytes = bytes;
    attr[0].val.accessPolicyWindow.hitRatio =
        bytes <= p->persist_bytes ? 1.0f : (float)((double)p->persist_bytes / (double)bytes);
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, k, a);
}

// index range [lo, hi) of the tiles of kind k whose tile id lies in [tile_lo, tile_hi)
static void kind_range(const spmv_b200_plan *p, int k, int tile_lo, int tile_hi, int *lo, int *hi) {
  const std::vector<int> &l = p->h_list[k];
  *lo = (int)(std::lower_bound(l.begin(), l.end(), tile_lo) - l.begin());
  *hi = (int)(std::lower_bound(l.begin(), l.end(), tile_hi) - l.begin());
}

static void fill_args(const spmv_b200_plan *p, double alpha, double beta, const double *x, double *y,
                      const PushArgs *push, SpmvArgs *a) {
  *a = SpmvArgs{};
  a->rowptr = p->rowptr;
  a->col = p->col;
  a->val = p->val;
  a->x = x;
  a->y = y;
  a->alpha = alpha;
  a->beta = beta;
  a->partials = p->partials;
  a->nnz = p->elem_end; // absolute index one past the last element (rowptr may be a view: rowptr[0] != 0)
  a->vec_div = p->vec_div;
  a->gather_na = (p->flags & (1u << 16)) ? 2 : ((p->flags & SPMV_B200_FLAG_GATHER_NO_L1) ? 1 : 0);
  a->stream_keep = p->stream_keep;
  a->read_y = (beta == 0.0 && (p->flags & SPMV_B200_FLAG_BETA0_SKIP_Y)) ? 0 : 1;
  a->row_start_bits = p->row_start_bits;
  a->nz_rows = p->nz_rows;
  a->lcol = p->lcol;
  a->xdesc = p->xdesc;
  a->lcol_base = p->lcol_base;
  a->n = p->n;
  a->m = p->m;
  if (push)
    a->push = *push;
  else
    a->push.count = 0;
}

// the staged-x form copies whole segments of x with TMA: x must be 16-byte aligned (cudaMalloc gives 256)
static bool use_xs(const spmv_b200_plan *p, const double *x) {
  return p->xstage && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
}
// the ring form also brings the row pointers by TMA: a row-pointer window that starts at an odd multiple of 4 bytes
// (a shard given as rowptr[lo : hi + 1] of its parent) keeps the one-row-block-per-CTA kernels
static bool use_ring(const spmv_b200_plan *p, const double *x) {
  return use_xs(p, x) && p->ring_stages >= 2 && (reinterpret_cast<uintptr_t>(p->rowptr) & 15u) == 0;
}

static int launch_range(const spmv_b200_plan *p, double alpha, double beta, const double *x, double *y, int tile_lo,
                        int tile_hi, bool with_fixup, cudaStream_t stream, const PushArgs *push = nullptr) {
  if (p->m == 0)
    return SPMV_B200_OK;
  if (push && push->count > 0 && push->multicast_mask != 0 &&
      (p->direct || p->nsplit > 0 || p->count[SPMV_B200_KIND_MIXED] > 0)) {
    set_error("multicast push destinations need a plan whose row blocks are all SHORT or MEDIUM (no direct form, no "
              "split rows)");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  SpmvArgs a;
  fill_args(p, alpha, beta, x, y, push, &a);
  const bool tma = p->uses_tma;
  const bool whole = tile_lo <= 0 && tile_hi >= p->ntiles;
  if (p->direct) {
    const int lo = tile_lo < 0 ? 0 : tile_lo, hi = tile_hi > p->ntiles ? p->ntiles : tile_hi;
    if (hi > lo) {
      a.desc = p->desc_direct + lo;
      a.ntiles = hi - lo;
      a.cap = 0;
      const int wpc = kThreads / 32;
      // tuning bit 23: register allocation for 6 resident CTAs per SM instead of 8
      RowsKernel kw = ((p->flags >> 23) & 1u) ? k_spmv_warp<6> : k_spmv_warp<8>;
      B200_CUDA(launch_spmv(kw, (a.ntiles + wpc - 1) / wpc, 0, stream, a, p));
    }
  }
  for (int k = 0; k < 3 && !p->direct; ++k) {
    if (p->count[k] == 0)
      continue;
    int lo = 0, hi = p->count[k];
    if (!whole)
      kind_range(p, k, tile_lo, tile_hi, &lo, &hi);
    if (hi <= lo)
      continue;
    a.desc = p->desc[k] + lo;
    a.cap = cap_for(p, k);
    a.ntiles = hi - lo;
    if (k == SPMV_B200_KIND_MIXED) {
      B200_CUDA(launch_spmv(mixed_kernel(tma, p->mixed_threads), a.ntiles, smem_for(p, k), stream, a, p,
                            p->mixed_threads));
      continue;
    }
    const RowsVariant &v = rows_variant(p, k);
    if (use_ring(p, x)) {
      a.cap = xs_cap_for(p, k);
      a.xcap = xs_xcap(p);
      a.ring_stages = p->ring_stages;
      a.ring_stage_bytes = (int)ring_stage_bytes(p, k);
      int sms = 0;
      B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device));
      if (p->comm_sms > 0) // room for a collective running beside this launch (spmv_b200_plan_set_comm_sms)
        sms = sms - p->comm_sms > 8 ? sms - p->comm_sms : 8;
      const int grid = a.ntiles < sms * p->ring_ctas ? a.ntiles : sms * p->ring_ctas;
      HaloSync none = {};
      ring_kernel(k, false)<<<grid, kRingThreads, (size_t)p->ring_stages * ring_stage_bytes(p, k), stream>>>(a, none);
    } else if (use_xs(p, x)) {
      a.cap = xs_cap_for(p, k);
      a.xcap = xs_xcap(p);
      B200_CUDA(launch_spmv((((p->flags >> 24) & 1u) && v.xs40) ? v.xs40 : v.xs, a.ntiles, xs_smem_for(p, k), stream, a, p));
    } else {
      B200_CUDA(launch_spmv(tma ? v.tma : v.plain, a.ntiles, smem_for(p, k), stream, a, p));
    }
  }
  if (with_fixup && p->nsplit > 0) {
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
    f.push = a.push;
    const int warps_per_cta = kThreads / 32;
    k_fixup<<<(p->nsplit + warps_per_cta - 1) / warps_per_cta, kThreads, 0, stream>>>(f);
  }
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

// ---- fused halo loop ----
// the plan's kind if the whole shard can run as one launch of a row kernel with the flag protocol inside, else -1
static int halo_kind(const spmv_b200_plan *p) {
  if (!p->uses_tma || p->direct || p->nsplit > 0 || p->ntiles == 0 || rotate_slots(p))
    return -1;
  for (int kind : {SPMV_B200_KIND_SHORT, SPMV_B200_KIND_MEDIUM})
    if (p->count[kind] == p->ntiles && rows_variant(p, kind).halo && (!p->xstage || rows_variant(p, kind).halo_xs))
      return kind;
  return -1;
}

bool kernels_halo_single_launch_ok(const spmv_b200_plan *p) { return halo_kind(p) >= 0; }

int kernels_launch_halo(const spmv_b200_plan *p, const TileDesc *desc, const double *x, double *y,
                        const PushArgs *push, const HaloSync &sync, cudaStream_t stream) {
  const int kind = halo_kind(p);
  if (kind < 0) {
    set_error("kernels_launch_halo: plan not covered by the single-launch kernel");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  const RowsVariant &v = rows_variant(p, kind);
  SpmvArgs a;
  fill_args(p, 1.0, 0.0, x, y, push, &a);
  a.read_y = 0; // y is a slice of the next x: never read (SPMV_B200_FLAG_BETA0_SKIP_Y semantics)
  a.desc = desc;
  a.ntiles = p->ntiles;
  if (use_ring(p, x)) {
    a.cap = xs_cap_for(p, kind);
    a.xcap = xs_xcap(p);
    a.ring_stages = p->ring_stages;
    a.ring_stage_bytes = (int)ring_stage_bytes(p, kind);
    int sms = 0;
    B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device));
    const int grid = p->ntiles < sms * p->ring_ctas ? p->ntiles : sms * p->ring_ctas;
    ring_kernel(kind, true)<<<grid, kRingThreads, (size_t)p->ring_stages * ring_stage_bytes(p, kind), stream>>>(a, sync);
  } else if (use_xs(p, x)) {
    a.cap = xs_cap_for(p, kind);
    a.xcap = xs_xcap(p);
    v.halo_xs<<<p->ntiles, kThreads, xs_smem_for(p, kind), stream>>>(a, sync);
  } else {
    a.cap = cap_for(p, kind);
    v.halo<<<p->ntiles, kThreads, smem_for(p, kind), stream>>>(a, sync);
  }
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

int kernels_halo_wait(const HaloSync &sync, cudaStream_t stream) {
  k_halo_wait<<<1, 1, 0, stream>>>(sync);
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

int kernels_halo_signal(const HaloSync &sync, cudaStream_t stream) {
  k_halo_signal<<<1, 1, 0, stream>>>(sync);
  B200_CUDA(cudaGetLastError());
  return SPMV_B200_OK;
}

int kernels_launch(const spmv_b200_plan *p, double alpha, double beta, const double *x, double *y,
                   cudaStream_t stream, const PushArgs *push) {
  return launch_range(p, alpha, beta, x, y, 0, p->ntiles, true, stream, push);
}

int kernels_launch_tiles(const spmv_b200_plan *p, double alpha, double beta, const double *x, double *y, int tile_lo,
                         int tile_hi, cudaStream_t stream, const PushArgs *push) {
  if (p->nsplit > 0) {
    set_error("kernels_launch_tiles: the plan has split rows");
    return SPMV_B200_ERR_UNSUPPORTED;
  }
  return launch_range(p, alpha, beta, x, y, tile_lo, tile_hi, false, stream, push);
}

} // namespace b200
