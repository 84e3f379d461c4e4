// Synthetic CSR generators for the BASELINE.json configurations (measurement infrastructure, not the SpMV path).
// Each generator is restated in numpy in spmv_acc_b200/synth.py; tests/test_synth.py checks they agree bit for bit,
// so CPU-side oracles and the GPU see identical matrices without ever moving them over PCIe.
//
// Random numbers are counter based: hash(seed, a, b) -> 64 bits (splitmix64 finaliser applied twice), so any
// element can be generated independently on either side.
#include <cuda_runtime.h>
#include <stdint.h>

#define GEN_API extern "C" __attribute__((visibility("default")))

namespace {

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ unsigned long long hash3(unsigned long long seed, unsigned long long a,
                                                             unsigned long long b) {
  return mix64(mix64(seed * 0xD1342543DE82EF95ull + a) ^ (b * 0x2545F4914F6CDD1Dull));
}

// uniform in [-1, 1): 53 random bits
__host__ __device__ __forceinline__ double sym_unit(unsigned long long h) {
  return (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void k_vector(long long n, unsigned long long seed, double *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = sym_unit(hash3(seed, (unsigned long long)i, 0x5eedull));
}

// ---- 2D 5-point Laplacian on an NY x N grid (NY grid rows of N points), Dirichlet truncation, row = i*N + j ----
__global__ void k_s2d_counts(int N, int NY, long long r_lo, long long r_hi, int *__restrict__ counts) {
  for (long long r = r_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; r < r_hi;
       r += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(r / N), j = (int)(r % N);
    counts[r - r_lo] = 1 + (i > 0) + (i < NY - 1) + (j > 0) + (j < N - 1);
  }
}

__global__ void k_s2d_fill(int N, int NY, long long r_lo, long long r_hi, const int *__restrict__ rowptr,
                           int *__restrict__ col, double *__restrict__ val) {
  for (long long r = r_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; r < r_hi;
       r += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(r / N), j = (int)(r % N);
    int p = rowptr[r - r_lo];
    if (i > 0) { col[p] = (int)(r - N); val[p++] = -1.0; }
    if (j > 0) { col[p] = (int)(r - 1); val[p++] = -1.0; }
    col[p] = (int)r; val[p++] = 4.0;
    if (j < N - 1) { col[p] = (int)(r + 1); val[p++] = -1.0; }
    if (i < NY - 1) { col[p] = (int)(r + N); val[p++] = -1.0; }
  }
}

// ---- 3D 27-point averaging stencil on an N^3 grid, row = (z*N + y)*N + x, value 1/27, columns ascending ----
__device__ __forceinline__ int span1(int c, int N) { return 1 + (c > 0) + (c < N - 1); }

__global__ void k_s3d_counts(int N, long long r_lo, long long r_hi, int *__restrict__ counts) {
  for (long long r = r_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; r < r_hi;
       r += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(r % N), y = (int)((r / N) % N), z = (int)(r / ((long long)N * N));
    counts[r - r_lo] = span1(x, N) * span1(y, N) * span1(z, N);
  }
}

__global__ void k_s3d_fill(int N, long long r_lo, long long r_hi, const int *__restrict__ rowptr,
                           int *__restrict__ col, double *__restrict__ val) {
  const double w = 1.0 / 27.0;
  for (long long r = r_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; r < r_hi;
       r += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(r % N), y = (int)((r / N) % N), z = (int)(r / ((long long)N * N));
    int p = rowptr[r - r_lo];
    for (int dz = -1; dz <= 1; ++dz) {
      if (z + dz < 0 || z + dz >= N) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        if (y + dy < 0 || y + dy >= N) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          if (x + dx < 0 || x + dx >= N) continue;
          col[p] = (int)(((long long)(z + dz) * N + (y + dy)) * N + (x + dx));
          val[p++] = w;
        }
      }
    }
  }
}

// ---- uniform random: exactly k distinct columns per row, ascending ----
constexpr int kMaxK = 64;
__global__ void k_uniform_fill(long long r_lo, long long r_hi, int n, int k, unsigned long long seed,
                               int *__restrict__ col, double *__restrict__ val) {
  for (long long r = r_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; r < r_hi;
       r += (long long)gridDim.x * blockDim.x) {
    int c[kMaxK];
    int have = 0;
    for (unsigned long long t = 0; have < k; ++t) {
      const int cand = (int)((hash3(seed, (unsigned long long)r, t) >> 11) % (unsigned long long)n);
      int pos = 0;
      while (pos < have && c[pos] < cand) ++pos;
      if (pos < have && c[pos] == cand) continue; // duplicate: redraw with the next counter
      for (int q = have; q > pos; --q) c[q] = c[q - 1];
      c[pos] = cand;
      ++have;
    }
    const long long base = (r - r_lo) * k;
    for (int q = 0; q < k; ++q) {
      col[base + q] = c[q];
      val[base + q] = sym_unit(hash3(seed ^ 0xabcdefull, (unsigned long long)r, (unsigned long long)q));
    }
  }
}

// ---- R-MAT: edge e -> (row, col) by `scale` quadrant choices; key = row << 32 | col ----
__global__ void k_rmat_edges(int scale, long long nedges, double a, double b, double c, unsigned long long seed,
                             long long *__restrict__ keys) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nedges;
       e += (long long)gridDim.x * blockDim.x) {
    unsigned long long row = 0, colv = 0;
    for (int lvl = 0; lvl < scale; ++lvl) {
      const double u = (double)(hash3(seed, (unsigned long long)e, (unsigned long long)lvl) >> 11) *
                       (1.0 / 9007199254740992.0);
      const int quad = u < a ? 0 : (u < a + b ? 1 : (u < a + b + c ? 2 : 3));
      row = (row << 1) | (unsigned long long)(quad >> 1);
      colv = (colv << 1) | (unsigned long long)(quad & 1);
    }
    keys[e] = (long long)((row << 32) | colv);
  }
}

__global__ void k_rmat_finish(long long nnz, const long long *__restrict__ keys_sorted, unsigned long long seed,
                              int *__restrict__ col, double *__restrict__ val) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
    col[i] = (int)(keys_sorted[i] & 0xffffffffll);
    val[i] = sym_unit(hash3(seed ^ 0x1234567ull, (unsigned long long)i, 1ull));
  }
}

__global__ void k_rmat_rowptr(int m, long long nnz, const long long *__restrict__ keys_sorted,
                              int *__restrict__ rowptr) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= m; r += (long long)gridDim.x * blockDim.x) {
    const long long target = r << 32; // first key with row >= r
    long long lo = 0, hi = nnz;
    while (lo < hi) {
      const long long mid = lo + ((hi - lo) >> 1);
      if (keys_sorted[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = (int)lo;
  }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

inline int done(cudaStream_t) { return (int)cudaGetLastError(); }

} // namespace

GEN_API int spmv_b200_gen_vector(long long n, unsigned long long seed, double *d_out, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (n > 0) k_vector<<<grid_for(n), 256, 0, s>>>(n, seed, d_out);
  return done(s);
}

GEN_API int spmv_b200_gen_stencil2d_counts(int N, int NY, long long r_lo, long long r_hi, int *d_counts, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (r_hi > r_lo) k_s2d_counts<<<grid_for(r_hi - r_lo), 256, 0, s>>>(N, NY, r_lo, r_hi, d_counts);
  return done(s);
}

GEN_API int spmv_b200_gen_stencil2d_fill(int N, int NY, long long r_lo, long long r_hi, const int *d_rowptr, int *d_col,
                                         double *d_val, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (r_hi > r_lo) k_s2d_fill<<<grid_for(r_hi - r_lo), 256, 0, s>>>(N, NY, r_lo, r_hi, d_rowptr, d_col, d_val);
  return done(s);
}

GEN_API int spmv_b200_gen_stencil3d_counts(int N, long long r_lo, long long r_hi, int *d_counts, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (r_hi > r_lo) k_s3d_counts<<<grid_for(r_hi - r_lo), 256, 0, s>>>(N, r_lo, r_hi, d_counts);
  return done(s);
}

GEN_API int spmv_b200_gen_stencil3d_fill(int N, long long r_lo, long long r_hi, const int *d_rowptr, int *d_col,
                                         double *d_val, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (r_hi > r_lo) k_s3d_fill<<<grid_for(r_hi - r_lo), 256, 0, s>>>(N, r_lo, r_hi, d_rowptr, d_col, d_val);
  return done(s);
}

GEN_API int spmv_b200_gen_uniform_fill(long long r_lo, long long r_hi, int n, int k, unsigned long long seed,
                                       int *d_col, double *d_val, void *stream) {
  if (k < 1 || k > kMaxK || k > n) return -1;
  auto s = static_cast<cudaStream_t>(stream);
  if (r_hi > r_lo) k_uniform_fill<<<grid_for(r_hi - r_lo), 256, 0, s>>>(r_lo, r_hi, n, k, seed, d_col, d_val);
  return done(s);
}

GEN_API int spmv_b200_gen_rmat_edges(int scale, long long nedges, double a, double b, double c,
                                     unsigned long long seed, long long *d_keys, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (nedges > 0) k_rmat_edges<<<grid_for(nedges), 256, 0, s>>>(scale, nedges, a, b, c, seed, d_keys);
  return done(s);
}

GEN_API int spmv_b200_gen_rmat_finish(int m, long long nnz, const long long *d_keys_sorted, unsigned long long seed,
                                      int *d_rowptr, int *d_col, double *d_val, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (nnz > 0) k_rmat_finish<<<grid_for(nnz), 256, 0, s>>>(nnz, d_keys_sorted, seed, d_col, d_val);
  k_rmat_rowptr<<<grid_for((long long)m + 1), 256, 0, s>>>(m, nnz, d_keys_sorted, d_rowptr);
  return done(s);
}
