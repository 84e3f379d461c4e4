/* Test infrastructure. Declarations of the five launchers the reference's run-time selector
 * (src/acc/hip-adaptive/adaptive.cpp:16-67) dispatches to, standing in for the reference's strategy headers (which
 * pull in HIP device code). With this directory in front of the reference's src/acc on the include path, the
 * reference's adaptive.cpp compiles UNMODIFIED with g++; the launchers are defined in oracle/ref_selector_wrap.cpp and
 * only record which one the selector chose. Signatures follow the call sites in adaptive.cpp. */
#ifndef ORACLE_SELECTOR_LAUNCHERS_H
#define ORACLE_SELECTOR_LAUNCHERS_H
#include "api/types.h"

void adaptive_vec_row_sparse_spmv(int nnz_block_0, int nnz_block_1, int trans, const double alpha, const double beta,
                                  const csr_desc<int, double> d_csr_desc, const double *x, double *y);
void adaptive_line_sparse_spmv(int trans, const double alpha, const double beta, const csr_desc<int, double> d_csr_desc,
                               const double *x, double *y);
void adaptive_enhance_sparse_spmv(int trans, const double alpha, const double beta,
                                  const csr_desc<int, double> d_csr_desc, const double *x, double *y);
void adaptive_flat_sparse_spmv(int nnz_block_0, int nnz_block_1, int trans, const double alpha, const double beta,
                               const csr_desc<int, double> d_csr_desc, const double *x, double *y);
void line_enhance_sparse_spmv(int trans, const double alpha, const double beta, const csr_desc<int, double> d_csr_desc,
                              const double *x, double *y);
#endif
