// ORACLE — TEST INFRASTRUCTURE ONLY.
// extern "C" doors onto the reference's own C++ functions. This file is compiled TOGETHER WITH the reference
// sources, in place from /root/reference (never copied), into oracle/_ref/libref_oracle.so by oracle/Makefile.
//   host_spmv (both overloads)            cli/verification.cpp:56-78
//   verify_y<int,double>                  cli/verification.cpp:15-38
//   csr_adaptive_plus_analyze_imp         src/acc/hip-csr-adaptive-plus/csr_adaptive_plus_analyze.cpp:12-98
//   csr_mtx_reader / csr_binary_reader / matrix_market_reader + to_csr
//                                         cli/csr_mtx_reader.hpp, cli/csr_binary_reader.hpp,
//                                         cli/matrix_market_reader.hpp, cli/sparse_format.h:100-128
//   generate_vector / rand_double         cli/utils.hpp:46-56 (restated in the wrapper because utils.hpp needs HIP)
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "csr_binary_reader.hpp"
#include "csr_mtx_reader.hpp"
#include "matrix_market_reader.hpp"
#include "sparse_format.h"
#include "verification.h"

#include "hip-csr-adaptive-plus/csr_adaptive_plus_analyze.h"

static std::string g_ref_error;

extern "C" {

const char *ref_last_error() { return g_ref_error.c_str(); }

void ref_host_spmv_axpby(double alpha, double beta, const double *value, const int *rowptr, const int *colindex,
                         int m, int n, int nnz, const double *x, double *y) {
  host_spmv(alpha, beta, value, rowptr, colindex, m, n, nnz, x, y);
}

void ref_host_spmv_ax(const double *value, const int *rowptr, const int *colindex, int m, int n, int nnz,
                      const double *x, double *y) {
  host_spmv(value, rowptr, colindex, m, n, nnz, x, y);
}

struct ref_verify_result {
  double max_error;
  int first_failed_at;
  int failed_count;
};

void ref_verify_y(double *dy, double *hy, int n, ref_verify_result *out) {
  VerifyResult<int, double> r = verify_y<int, double>(dy, hy, n);
  out->max_error = r.max_error;
  out->first_failed_at = r.first_failed_at;
  out->failed_count = r.failed_count;
}

void ref_verify_print(double *dy, double *hy, int n) { verify(dy, hy, n); }

// THREADS_PER_BLOCK = 512, MIN_NNZ_PER_BLOCK = 2048 are the values csr_adaptive_plus_sparse_spmv uses
// (csr_adaptive_plus_spmv.cpp:134-137); vec selects the VEC_SIZE instantiation.
int ref_adaptive_plus_analyze(int m, int nnz, int min_nnz_per_block, int vec, const int *host_row_ptr,
                              int *break_points_out, int cap, int *first_block_of_row_out /* [m+1] */) {
  std::vector<int> bp;
  std::vector<int> first(m + 1, 0);
  int blocks = -1;
  switch (vec) {
  case 1: blocks = csr_adaptive_plus_analyze_imp<int, 512, 1>(m, nnz, min_nnz_per_block, bp, first, host_row_ptr, nullptr); break;
  case 2: blocks = csr_adaptive_plus_analyze_imp<int, 512, 2>(m, nnz, min_nnz_per_block, bp, first, host_row_ptr, nullptr); break;
  case 4: blocks = csr_adaptive_plus_analyze_imp<int, 512, 4>(m, nnz, min_nnz_per_block, bp, first, host_row_ptr, nullptr); break;
  case 8: blocks = csr_adaptive_plus_analyze_imp<int, 512, 8>(m, nnz, min_nnz_per_block, bp, first, host_row_ptr, nullptr); break;
  case 16: blocks = csr_adaptive_plus_analyze_imp<int, 512, 16>(m, nnz, min_nnz_per_block, bp, first, host_row_ptr, nullptr); break;
  case 32: blocks = csr_adaptive_plus_analyze_imp<int, 512, 32>(m, nnz, min_nnz_per_block, bp, first, host_row_ptr, nullptr); break;
  case 64: blocks = csr_adaptive_plus_analyze_imp<int, 512, 64>(m, nnz, min_nnz_per_block, bp, first, host_row_ptr, nullptr); break;
  default: return -1;
  }
  if ((int)bp.size() > cap)
    return -(int)bp.size();
  std::memcpy(break_points_out, bp.data(), sizeof(int) * bp.size());
  if (first_block_of_row_out)
    std::memcpy(first_block_of_row_out, first.data(), sizeof(int) * (m + 1));
  return blocks;
}

// ---- readers: two-phase (open -> query sizes -> copy out -> close) ----
struct ref_csr_handle {
  std::vector<double> val;
  std::vector<int> col, rowptr;
  std::vector<double> x;
  int rows = 0, cols = 0, nnz = 0;
};

static ref_csr_handle *from_raw(int rows, int cols, int nnz, const double *v, const int *c, const int *r,
                                const double *x) {
  ref_csr_handle *h = new ref_csr_handle();
  h->rows = rows;
  h->cols = cols;
  h->nnz = nnz;
  h->val.assign(v, v + nnz);
  h->col.assign(c, c + nnz);
  h->rowptr.assign(r, r + rows + 1);
  if (x)
    h->x.assign(x, x + cols);
  return h;
}

ref_csr_handle *ref_read_csr_text(const char *path) {
  try {
    csr_mtx_reader<int, double> rd{std::string(path)};
    rd.fill_mtx();
    rd.close_stream();
    double *v, *x;
    int *c, *r;
    rd.as_raw_ptr(v, c, r, x);
    return from_raw(rd.rows(), rd.cols(), rd.nnz(), v, c, r, x);
  } catch (const std::exception &e) {
    g_ref_error = e.what();
    return nullptr;
  }
}

ref_csr_handle *ref_read_bin2(const char *path) {
  csr_binary_reader<int32_t, double> rd;
  rd.load_mat(std::string(path));
  rd.close_stream();
  double *v;
  int *c, *r;
  rd.as_raw_ptr(v, c, r);
  if (!v || !c || !r)
    return nullptr;
  ref_csr_handle *h = from_raw(rd.rows(), rd.cols(), rd.nnz(), v, c, r, nullptr);
  delete[] v;
  delete[] c;
  delete[] r;
  return h;
}

ref_csr_handle *ref_read_mtx(const char *path) {
  try {
    matrix_market_reader<int, double> rd;
    matrix_market<int, double> mm = rd.load_mat(std::string(path));
    csr_mtx<int, double> csr = mm.to_csr();
    ref_csr_handle *h = from_raw(csr.rows, csr.cols, csr.nnz, csr.values, csr.col_index, csr.row_ptr, nullptr);
    delete[] csr.values;
    delete[] csr.col_index;
    delete[] csr.row_ptr;
    return h;
  } catch (const std::exception &e) {
    g_ref_error = e.what();
    return nullptr;
  }
}

void ref_csr_sizes(const ref_csr_handle *h, int *rows, int *cols, int *nnz, int *has_x) {
  *rows = h->rows;
  *cols = h->cols;
  *nnz = h->nnz;
  *has_x = h->x.empty() ? 0 : 1;
}

void ref_csr_copy(const ref_csr_handle *h, int *rowptr, int *col, double *val, double *x) {
  std::memcpy(rowptr, h->rowptr.data(), sizeof(int) * h->rowptr.size());
  std::memcpy(col, h->col.data(), sizeof(int) * h->col.size());
  std::memcpy(val, h->val.data(), sizeof(double) * h->val.size());
  if (x && !h->x.empty())
    std::memcpy(x, h->x.data(), sizeof(double) * h->x.size());
}

void ref_csr_free(ref_csr_handle *h) { delete h; }

// cli/utils.hpp:46-56 (utils.hpp itself pulls in HIP through HIP_CHECK users, so the two lines are restated)
void ref_generate_vector(int n, double *x) {
  for (int i = 0; i < n; i++) {
    x[i] = static_cast<double>(-1.0 + (1.0 - (-1.0)) * double(rand() % 100) / double((101)));
  }
}
void ref_srand(unsigned seed) { srand(seed); }

} // extern "C"
