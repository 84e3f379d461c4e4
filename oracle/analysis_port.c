/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see host_spmv_port.c).
 *
 * Plain-C, serial restatement of the row analysis of spmv_acc_b200/csrc/analysis.cu. Every array it produces must be
 * bit-identical to what the CUDA analysis exports through spmv_b200_plan_export.
 *
 * Reference anchors:
 *   port_merge_path_partition  follows benchmark/merge-path/merge_path_partition.h:7-17 together with
 *       largest_less_equal_binary_search (benchmark/merge-path/merge_path_utils.h:32-44): S[t] = largest r in
 *       [0, m] with rowptr[r] <= t * ITEMS. The CUDA analysis exports the same array as TILE_PART; the GPU tests
 *       additionally run the reference's own `partition` kernel (oracle/_ref/libref_gpu.so) against it.
 *   port_flat_break_points_v2   follows src/acc/hip-flat/flat_imp.inl:134-152: break_points[j] = row containing
 *       element j * stride (only entries hit by a non-empty row are written).
 *   port_adaptive_plus_blocks   is NOT restated here: the reference's csr_adaptive_plus_analyze_imp is compiled in
 *       place into oracle/_ref/libref_oracle.so and used as a cross-check of the nnz-balance property only.
 *
 * Specification of port_analysis: identical to the comment block at the top of analysis.cu (tiles balance non-zeros
 * and rows: row r starts at merged position f(r) = rowptr[r] - base + r).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int lower_bound_rowptr(const int *rowptr, int m, long long target) {
  int lo = 0, hi = m;
  while (lo < hi) {
    int mid = lo + ((hi - lo) >> 1);
    if ((long long)rowptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* first index in [0, m] with f(idx) = rowptr[idx] - base + idx >= target */
static int lower_bound_merge(const int *rowptr, int m, long long base, long long target) {
  int lo = 0, hi = m;
  while (lo < hi) {
    int mid = lo + ((hi - lo) >> 1);
    if ((long long)rowptr[mid] - base + mid < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* merge_path_utils.h:32-44 with left = 0, right = m + 1 (merge_path_partition.h:14) */
static int largest_less_equal(const int *rowptr, int left, int right, long long target) {
  while (left < right - 1) {
    int mid = (left + right) >> 1;
    if ((long long)rowptr[mid] <= target)
      left = mid;
    else
      right = mid;
  }
  return left;
}

void port_merge_path_partition(const int *rowptr, int m, int count, int items_per_block, int *S) {
  for (int idx = 0; idx < count; idx++) {
    S[idx] = largest_less_equal(rowptr, 0, m + 1, (long long)idx * items_per_block);
  }
}

void port_flat_break_points_v2(const int *rowptr, int m, int stride, int *break_points) {
  for (int i = 0; i < m; i++) {
    int p1 = rowptr[i] / stride;
    if (rowptr[i] % stride != 0)
      p1++;
    const int p2 = (rowptr[i + 1] - 1) / stride;
    for (int j = p1; j <= p2; j++)
      break_points[j] = i;
  }
}

int port_analysis_ntiles(const int *rowptr, int m, int T) {
  if (m == 0)
    return 0;
  long long total = (long long)rowptr[m] - (long long)rowptr[0];
  long long nt = (total + m + T - 1) / T;
  return (int)(nt < 1 ? 1 : nt);
}

/* Fills tile_row/tile_elem/tile_part [ntiles+1], tile_split [ntiles+1], tile_maxlen/tile_kind [ntiles],
 * row_bin [m], bin_rows/bin_nnz [4]. Returns the number of split rows written to split_rows (3 ints per row:
 * rows | first tiles | last tiles as three consecutive arrays of length cap_split; pass NULL to only count). */
int port_analysis(const int *rowptr, int m, int T, int short_max, int medium_max, int *tile_row, int *tile_elem,
                  unsigned char *tile_split, int *tile_part, int *tile_maxlen, unsigned char *tile_kind,
                  unsigned char *row_bin, long long *bin_rows, long long *bin_nnz, int *split_rows, int cap_split) {
  for (int b = 0; b < 4; b++) {
    bin_rows[b] = 0;
    bin_nnz[b] = 0;
  }
  if (m == 0)
    return 0;
  const int ntiles = port_analysis_ntiles(rowptr, m, T);
  const long long base = rowptr[0], end = rowptr[m];
  for (int t = 0; t <= ntiles; t++) {
    const long long pos = (long long)t * T;
    long long target = base + pos;
    if (target > end)
      target = end;
    int row, elem;
    unsigned char split = 0;
    if (t == 0) {
      row = 0;
      elem = (int)base;
    } else if (t == ntiles) {
      row = m;
      elem = (int)end;
    } else {
      row = lower_bound_merge(rowptr, m, base, pos);
      const long long cut = base + pos - (row - 1);
      elem = rowptr[row];
      if (cut < (long long)rowptr[row]) {
        const int len = rowptr[row] - rowptr[row - 1];
        if (len > medium_max) {
          split = 1;
          elem = (int)cut;
        }
      }
    }
    tile_row[t] = row;
    tile_elem[t] = elem;
    tile_split[t] = split;
    tile_part[t] = largest_less_equal(rowptr, 0, m + 1, target);
  }
  for (int t = 0; t < ntiles; t++)
    tile_maxlen[t] = 0;
  for (int r = 0; r < m; r++) {
    const int len = rowptr[r + 1] - rowptr[r];
    const int bin = len <= short_max ? 0 : (len <= medium_max ? 1 : (len <= T ? 2 : 3));
    if (row_bin)
      row_bin[r] = (unsigned char)bin;
    bin_rows[bin] += 1;
    bin_nnz[bin] += len;
    long long t = ((long long)rowptr[r] - base + r) / T;
    if (t > ntiles - 1)
      t = ntiles - 1;
    if (len > tile_maxlen[t])
      tile_maxlen[t] = len;
  }
  for (int t = 0; t < ntiles; t++) {
    const long long rows = tile_row[t + 1] - tile_row[t];
    const long long elems = tile_elem[t + 1] - tile_elem[t];
    const int skewed = tile_maxlen[t] > short_max && (long long)tile_maxlen[t] * rows > 2 * elems;
    if (tile_split[t] || tile_split[t + 1] || tile_maxlen[t] > medium_max || rows > 512 || skewed)
      tile_kind[t] = 2;
    else if (tile_maxlen[t] <= short_max)
      tile_kind[t] = 0;
    else
      tile_kind[t] = 1;
  }
  int nsplit = 0;
  for (int t = 1; t < ntiles; t++) {
    if (!tile_split[t])
      continue;
    const int r = tile_row[t] - 1;
    if (tile_row[t - 1] <= r) {
      if (split_rows && nsplit < cap_split) {
        long long t1 = ((long long)rowptr[r + 1] - 1 - base + r) / T;
        if (t1 > ntiles - 1)
          t1 = ntiles - 1;
        split_rows[nsplit] = r;
        split_rows[cap_split + nsplit] = t - 1;
        split_rows[2 * cap_split + nsplit] = (int)t1;
      }
      nsplit++;
    }
  }
  return nsplit;
}

/* Gather-coalescing statistic of analysis.cu:k_gather_stat: 4096 groups of 32 consecutive rows, middle element of every
 * non-empty row, out[0] = active lanes, out[1] = distinct (colindex >> 4) per group summed over the groups,
 * out[2] = non-zeros of the sampled rows, out[3] = those in rows longer than medium_max. */
void port_gather_stat(const int *rowptr, const int *col, int m, int medium_max, long long *out) {
  const int samples = 4096;
  out[0] = out[1] = out[2] = out[3] = 0;
  const long long span = m > 32 ? (long long)(m - 32) : 0;
  for (int w = 0; w < samples; w++) {
    int lines[32];
    int nl = 0;
    for (int lane = 0; lane < 32; lane++) {
      const long long r = (span * w) / samples + lane;
      if (r >= m)
        continue;
      const int s = rowptr[r], e = rowptr[r + 1];
      out[2] += e - s;
      if (e - s > medium_max)
        out[3] += e - s;
      if (e <= s)
        continue;
      const int line = col[s + ((e - s) >> 1)] >> 4;
      out[0] += 1;
      int seen = 0;
      for (int q = 0; q < nl; q++)
        if (lines[q] == line)
          seen = 1;
      if (!seen)
        lines[nl++] = line;
    }
    out[1] += nl;
  }
}

/* bounds[g] = lower_bound(rowptr, base + total * g / nshards), bounds[0] = 0, bounds[nshards] = m */
void port_shard_bounds(const int *rowptr, int m, int nshards, int *bounds) {
  if (m == 0) {
    for (int g = 0; g <= nshards; g++)
      bounds[g] = 0;
    return;
  }
  const long long base = rowptr[0];
  const long long total = (long long)rowptr[m] - base;
  for (int g = 0; g <= nshards; g++) {
    if (g == 0)
      bounds[g] = 0;
    else if (g == nshards)
      bounds[g] = m;
    else
      bounds[g] = lower_bound_rowptr(rowptr, m, base + (total * g) / nshards);
  }
}

/* Emulates the tile decomposition of the CUDA kernels on the CPU: every tile sums exactly the elements the CUDA
 * kernels would give it (owned rows, head fragment, tail fragment), split rows are recombined from the per-tile
 * partials like k_fixup does. Summation inside a row is left to right, so for matrices without split rows the result
 * equals port_host_spmv_axpby bit for bit; the tests use this to prove that the decomposition covers every element
 * of every row exactly once. Returns 0 on success, a negative code if a structural invariant is violated. */
int port_tiled_spmv(double alpha, double beta, const double *value, const int *rowptr, const int *colindex, int m,
                    int T, int medium_max, const int *tile_row, const int *tile_elem, const unsigned char *tile_split,
                    int ntiles, const double *x, double *y, double *partials /* [2*ntiles] */,
                    unsigned char *row_done /* [m] scratch */) {
  memset(row_done, 0, (size_t)m);
  const long long base = m > 0 ? rowptr[0] : 0;
  for (int t = 0; t < ntiles; t++) {
    const int r0 = tile_row[t], r1 = tile_row[t + 1];
    const int e0 = tile_elem[t], e1 = tile_elem[t + 1];
    if (e1 < e0 || r1 < r0)
      return -1;
    if (e1 - (e0 & ~3) > T + medium_max + 8 - 4)
      return -2; /* would overflow the shared-memory tile */
    const int has_tail = tile_split[t + 1] && r1 > r0;
    const int nrows = (r1 - r0) - (has_tail ? 1 : 0);
    partials[2 * t] = 0.0;
    partials[2 * t + 1] = 0.0;
    if (tile_split[t]) {
      const int hend = rowptr[r0] < e1 ? rowptr[r0] : e1;
      double s = 0;
      for (int j = e0; j < hend; j++)
        s += value[j] * x[colindex[j]];
      partials[2 * t] = s;
    } else if (r0 < m && rowptr[r0] != e0 && r1 > r0) {
      return -3; /* a clean boundary must start exactly at the first owned row */
    }
    if (has_tail) {
      double s = 0;
      for (int j = rowptr[r1 - 1]; j < e1; j++)
        s += value[j] * x[colindex[j]];
      partials[2 * t + 1] = s;
    }
    for (int r = r0; r < r0 + nrows; r++) {
      if (rowptr[r] < e0 || rowptr[r + 1] > e1)
        return -4; /* an owned row must lie inside the streamed range */
      if (row_done[r])
        return -5;
      double s = 0;
      for (int j = rowptr[r]; j < rowptr[r + 1]; j++)
        s += value[j] * x[colindex[j]];
      y[r] = alpha * s + beta * y[r];
      row_done[r] = 1;
    }
  }
  for (int t = 1; t < ntiles; t++) {
    if (!tile_split[t])
      continue;
    const int r = tile_row[t] - 1;
    if (tile_row[t - 1] <= r) {
      long long t1 = ((long long)rowptr[r + 1] - 1 - base + r) / T;
      if (t1 > ntiles - 1)
        t1 = ntiles - 1;
      double s = partials[2 * (t - 1) + 1];
      for (long long u = t; u <= t1; u++)
        s += partials[2 * u];
      if (row_done[r])
        return -6;
      y[r] = alpha * s + beta * y[r];
      row_done[r] = 1;
    }
  }
  for (int r = 0; r < m; r++)
    if (!row_done[r])
      return -7;
  return 0;
}

/* The reference's run-time strategy selector, restated (src/acc/hip-adaptive/adaptive.cpp:16-67): four samples of the
 * host row pointers decide which of its kernels multiplies the matrix. Integer arithmetic throughout, including the
 * average (bp_3 / m, :29) and the imbalance ratio of the two row halves (:34-35). Returns
 *   0 vector-row with two data blocks (:34-40)   1 adaptive line (:43-49)          2 adaptive line-enhance (:52-55)
 *   3 adaptive flat (:60-63)                     4 line-enhance (:66)
 * An empty half (division by zero in the reference) is reported as the two-block case. Used by the selector study in
 * DESIGN.md §7, which sets the reference's choice next to the tile kinds our analysis assigns. */
int port_adaptive_choice(const int *rowptr, int m) {
  if (m <= 0)
    return 4;
  const int bp_1 = rowptr[m / 2];
  const int bp_3 = rowptr[m];
  const int avg_nnz_per_row = bp_3 / m;
  const int nnz_block_0 = bp_1 - 0;
  const int nnz_block_1 = bp_3 - bp_1;
  if (nnz_block_0 == 0 || nnz_block_1 == 0) {
    if (nnz_block_0 != nnz_block_1)
      return 0;
  } else if ((nnz_block_1 > nnz_block_0 && nnz_block_1 / nnz_block_0 >= 4) ||
             (nnz_block_0 > nnz_block_1 && nnz_block_0 / nnz_block_1 >= 4)) {
    return 0;
  }
  if (avg_nnz_per_row <= 4)
    return 1;
  if (bp_3 <= 0xC00000)
    return 2;
  if (bp_3 > (1 << 23))
    return 3;
  return 4;
}

/* Staged-x form of the row kernels (spmv_acc_b200/csrc/analysis.cu:k_xstage_build), restated serially. For every row
 * block t with elements [tile_elem[t], tile_elem[t+1]): the sorted set of 128-byte lines of x it references
 * (line = colindex >> 4), the maximal runs of consecutive lines ("segments": first line and rank of that line in the
 * set), and for every element the 16-bit offset (rank of its line << 4) | (colindex & 15) into the staged copy of x.
 * This plays the role of the reference's x-remap kernels (src/acc/hip-thread-row/thread_row_block_x_remap.hpp:148-265,
 * which stage column indices and then x per block in LDS), decided once per matrix instead of per launch.
 * xdesc is 32 int32 per tile with the layout of struct XDesc: [0] nseg, [1] nlines, [2..17] first lines, [18..25] the
 * ranks as 16 uint16 packed little-endian, [26..31] zero. Returns the number of tiles that do not qualify (span of
 * lines >= span_max, more than lines_max lines, more than seg_max runs, or no elements); their xdesc[0] is -1 and
 * their lcol entries are left untouched. *max_lines receives the largest line count of the qualifying tiles. */
static int cmp_int(const void *a, const void *b) {
  const int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

int port_xstage(const int *colindex, const int *tile_elem, int ntiles, int elem_base, int span_max, int lines_max,
                int seg_max, unsigned short *lcol, int *xdesc, int *max_lines) {
  int failed = 0, best = 0;
  for (int t = 0; t < ntiles; t++) {
    int *xd = xdesc + 32 * (long long)t;
    for (int i = 0; i < 32; i++)
      xd[i] = 0;
    const int e0 = tile_elem[t], e1 = tile_elem[t + 1];
    const int cnt = e1 - e0;
    int ok = cnt > 0;
    int *lines = NULL;
    int nlines = 0, nseg = 0;
    if (ok) {
      lines = (int *)malloc(sizeof(int) * (size_t)cnt);
      for (int k = 0; k < cnt; k++)
        lines[k] = colindex[e0 + k] >> 4;
      qsort(lines, (size_t)cnt, sizeof(int), cmp_int);
      for (int k = 0; k < cnt; k++)
        if (k == 0 || lines[k] != lines[k - 1])
          lines[nlines++] = lines[k];
      for (int i = 0; i < nlines; i++)
        if (i == 0 || lines[i] != lines[i - 1] + 1)
          nseg++;
      ok = (lines[nlines - 1] - lines[0]) < span_max && nlines <= lines_max && nseg <= seg_max;
      /* too many runs: merge runs at most g missing lines apart, g = 1, 2, 4, ... 32 (k_xstage_build, merge path) */
      if (!ok && (lines[nlines - 1] - lines[0]) < span_max && nlines <= lines_max && nseg > seg_max && nseg <= 1024) {
        int pick = 0;
        for (int g = 1; g <= 32 && !pick; g <<= 1) {
          int c = 1;
          for (int i = 1; i < nlines; i++)
            c += (lines[i] - lines[i - 1] - 1) > g;
          if (c <= seg_max)
            pick = g;
        }
        if (pick) {
          unsigned short *moff = (unsigned short *)(xd + 18);
          int mline[64], mo[64], ms = 0, total = 0, start = lines[0];
          for (int i = 1; i <= nlines; i++)
            if (i == nlines || (lines[i] - lines[i - 1] - 1) > pick) {
              mline[ms] = start;
              mo[ms] = total;
              total += lines[i - 1] + 1 - start;
              ms++;
              if (i < nlines)
                start = lines[i];
            }
          if (total <= lines_max) {
            for (int j = 0; j < ms; j++) {
              xd[2 + j] = mline[j];
              moff[j] = (unsigned short)mo[j];
            }
            xd[0] = ms;
            xd[1] = total;
            if (total > best)
              best = total;
            for (int k = 0; k < cnt; k++) {
              const int c = colindex[e0 + k], l = c >> 4;
              int sg = 0;
              for (int j = 1; j < ms; j++)
                if (mline[j] <= l)
                  sg = j;
              lcol[(long long)e0 + k - elem_base] = (unsigned short)(((mo[sg] + (l - mline[sg])) << 4) | (c & 15));
            }
            free(lines);
            continue;
          }
        }
      }
    }
    if (!ok) {
      xd[0] = -1;
      failed++;
      free(lines);
      continue;
    }
    unsigned short *off = (unsigned short *)(xd + 18);
    int s = 0;
    for (int i = 0; i < nlines; i++)
      if (i == 0 || lines[i] != lines[i - 1] + 1) {
        xd[2 + s] = lines[i];
        off[s] = (unsigned short)i;
        s++;
      }
    xd[0] = nseg;
    xd[1] = nlines;
    if (nlines > best)
      best = nlines;
    for (int k = 0; k < cnt; k++) {
      const int c = colindex[e0 + k], l = c >> 4;
      int lo = 0, hi = nlines - 1; /* rank of l in the sorted set */
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lines[mid] < l)
          lo = mid + 1;
        else
          hi = mid;
      }
      lcol[(long long)e0 + k - elem_base] = (unsigned short)((lo << 4) | (c & 15));
    }
    free(lines);
  }
  if (max_lines)
    *max_lines = best;
  return failed;
}
