// ORACLE — test infrastructure only. The reference's run-time strategy selector, compiled in place:
// /root/reference/src/acc/hip-adaptive/adaptive.cpp is built unmodified (oracle/Makefile, target ref-selector) with the
// five launchers it dispatches to defined here as recorders, so that `ref_adaptive_choice` returns which kernel the
// reference itself would run for a given row-pointer array. Used to pin oracle/analysis_port.c:port_adaptive_choice
// (the restatement the selector study quotes) against the reference's own code.
#include "selector_launchers.h"

void adaptive_sparse_spmv(int trans, const double alpha, const double beta, const csr_desc<int, double> h_csr_desc,
                          const csr_desc<int, double> d_csr_desc, const double *x, double *y);

static int g_choice = -1;

void adaptive_vec_row_sparse_spmv(int, int, int, const double, const double, const csr_desc<int, double>, const double *,
                                  double *) {
  g_choice = 0;
}
void adaptive_line_sparse_spmv(int, const double, const double, const csr_desc<int, double>, const double *, double *) {
  g_choice = 1;
}
void adaptive_enhance_sparse_spmv(int, const double, const double, const csr_desc<int, double>, const double *,
                                  double *) {
  g_choice = 2;
}
void adaptive_flat_sparse_spmv(int, int, int, const double, const double, const csr_desc<int, double>, const double *,
                               double *) {
  g_choice = 3;
}
void line_enhance_sparse_spmv(int, const double, const double, const csr_desc<int, double>, const double *, double *) {
  g_choice = 4;
}

// same numbering as port_adaptive_choice: 0 vector-row two blocks, 1 line, 2 line-enhance (adaptive), 3 flat, 4 line-enhance
extern "C" int ref_adaptive_choice(const int *rowptr, int m) {
  const csr_desc<int, double> h(m, m, rowptr[m], rowptr, nullptr, nullptr);
  g_choice = -1;
  adaptive_sparse_spmv(0, 1.0, 1.0, h, h, nullptr, nullptr);
  return g_choice;
}
