/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 *
 * Plain-C restatement of the reference's CPU SpMV and of its result checker.
 *   port_host_spmv_axpby  follows cli/verification.cpp:56-66  (y[i] = alpha * sum + beta * y[i], serial
 *                          left-to-right accumulation from 0, in place on y)
 *   port_host_spmv_ax     follows cli/verification.cpp:68-78  (y[i] = sum)
 *   port_verify_y         follows cli/verification.cpp:15-38  (|hy| <= 1e-12: abs err >= 1e-14 fails, else
 *                          rel err >= 1e-7 fails; reports max_error / first_failed_at / failed_count)
 *   port_verify           follows cli/verification.cpp:43-54  (rel err >= 1e-7 fails; returns the index of the
 *                          first failing row or -1; the reference prints instead of returning)
 * Pinning: tests/test_oracle.py checks these bit-for-bit against the reference's own functions compiled in place
 * from /root/reference into oracle/_ref/libref_oracle.so (see oracle/Makefile), and against the golden vectors
 * under tests/golden/ that were produced by that library (tests/golden/make_golden.py).
 *
 * port_row_bound is not from the reference: it evaluates the north-star tolerance
 *   |y - y_ref| <= tol * (|beta * y0_i| + |alpha| * sum_j |a_ij * x_j|)
 * next to the oracle result so the tests can apply it row by row.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference built with plain g++ -O2 on x86-64).
 */
#include <math.h>
#include <stdint.h>

void port_host_spmv_axpby(double alpha, double beta, const double *value, const int *rowptr, const int *colindex,
                          int m, int n, int nnz, const double *x, double *y) {
  (void)n;
  (void)nnz;
  for (int i = 0; i < m; i++) {
    double y0 = 0;
    for (int j = rowptr[i]; j < rowptr[i + 1]; j++) {
      y0 += value[j] * x[colindex[j]];
    }
    y[i] = alpha * y0 + beta * y[i];
  }
}

void port_host_spmv_ax(const double *value, const int *rowptr, const int *colindex, int m, int n, int nnz,
                       const double *x, double *y) {
  (void)n;
  (void)nnz;
  for (int i = 0; i < m; i++) {
    double y0 = 0;
    for (int j = rowptr[i]; j < rowptr[i + 1]; j++) {
      y0 += value[j] * x[colindex[j]];
    }
    y[i] = y0;
  }
}

typedef struct {
  double max_error;
  int first_failed_at;
  int failed_count;
} port_verify_result;

void port_verify_y(const double *dy, const double *hy, int n, port_verify_result *out) {
  int first_failed_at = -1;
  int failed_count = 0;
  double max_error = 0.0;
  for (int i = 0; i < n; i++) {
    const double err = fabs(dy[i] - hy[i]);
    if (err > max_error) {
      max_error = err;
    }
    if ((fabs(hy[i]) <= 1e-12 && err >= 1e-14) || (fabs(hy[i]) > 1e-12 && err / fabs(hy[i]) >= 1e-7)) {
      if (failed_count <= 0) {
        first_failed_at = i;
      }
      failed_count++;
    }
  }
  out->max_error = max_error;
  out->first_failed_at = first_failed_at;
  out->failed_count = failed_count;
}

int port_verify(const double *dy, const double *hy, int n) {
  for (int i = 0; i < n; i++) {
    if (fabs(dy[i] - hy[i]) / fabs(hy[i]) >= 1e-7) {
      return i;
    }
  }
  return -1;
}

/* bound[i] = |beta * y0[i]| + |alpha| * sum_j |value[j] * x[colindex[j]]|   (not reference code) */
void port_row_bound(double alpha, double beta, const double *value, const int *rowptr, const int *colindex, int m,
                    const double *x, const double *y0, double *bound) {
  for (int i = 0; i < m; i++) {
    double s = 0;
    for (int j = rowptr[i]; j < rowptr[i + 1]; j++) {
      s += fabs(value[j] * x[colindex[j]]);
    }
    bound[i] = fabs(beta * y0[i]) + fabs(alpha) * s;
  }
}

/* The reference's vector generator, cli/utils.hpp:46-56: min + (max-min) * (rand() % 100) / 101 with libc rand().
 * The caller seeds (the reference never calls srand, which equals srand(1)). */
#include <stdlib.h>
void port_generate_vector(int n, double *x) {
  for (int i = 0; i < n; i++) {
    x[i] = -1.0 + (1.0 - (-1.0)) * (double)(rand() % 100) / (double)(101);
  }
}
void port_srand(unsigned seed) { srand(seed); }
