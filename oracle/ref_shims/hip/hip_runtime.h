// Compatibility header used ONLY to compile the reference's unchanged host drivers (cli/main.cpp, cli/utils.hpp,
// benchmark/utils/timer_utils.h) against the CUDA runtime when no HIP installation exists: it maps the HIP symbols
// those files use (enumerated by grep over cli/ and benchmark/) onto their CUDA equivalents. Not part of the product.
#ifndef SPMV_B200_HIP_RUNTIME_SHIM_H
#define SPMV_B200_HIP_RUNTIME_SHIM_H

#include <cuda_runtime.h>

typedef cudaError_t hipError_t;
typedef cudaEvent_t hipEvent_t;
typedef cudaStream_t hipStream_t;
#define hipSuccess cudaSuccess
#define hipMemcpyHostToDevice cudaMemcpyHostToDevice
#define hipMemcpyDeviceToHost cudaMemcpyDeviceToHost
#define hipMemcpyDeviceToDevice cudaMemcpyDeviceToDevice

static inline hipError_t hipSetDevice(int d) { return cudaSetDevice(d); }
static inline hipError_t hipMalloc(void **p, size_t n) { return cudaMalloc(p, n); }
static inline hipError_t hipFree(void *p) { return cudaFree(p); }
static inline hipError_t hipMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k) { return cudaMemcpy(d, s, n, k); }
static inline hipError_t hipMemset(void *d, int v, size_t n) { return cudaMemset(d, v, n); }
static inline hipError_t hipDeviceSynchronize() { return cudaDeviceSynchronize(); }
static inline const char *hipGetErrorString(hipError_t e) { return cudaGetErrorString(e); }
static inline hipError_t hipGetLastError() { return cudaGetLastError(); }
static inline hipError_t hipEventCreate(hipEvent_t *e) { return cudaEventCreate(e); }
static inline hipError_t hipEventRecord(hipEvent_t e, hipStream_t s = 0) { return cudaEventRecord(e, s); }
static inline hipError_t hipEventSynchronize(hipEvent_t e) { return cudaEventSynchronize(e); }
static inline hipError_t hipEventElapsedTime(float *ms, hipEvent_t a, hipEvent_t b) { return cudaEventElapsedTime(ms, a, b); }
static inline hipError_t hipEventDestroy(hipEvent_t e) { return cudaEventDestroy(e); }

#endif
