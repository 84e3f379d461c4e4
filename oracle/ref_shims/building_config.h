/* Hand-written stand-in for the CMake-generated building_config.h of the reference (src/building_config.h.in).
 * Only what the CPU-side sources need; used when compiling reference sources in place for oracle/_ref. */
#ifndef SPMV_BUILDING_CONFIG_H
#define SPMV_BUILDING_CONFIG_H
#define __WF_SIZE__ 32
constexpr int __WRAP_SIZE__ = __WF_SIZE__;
#define AVAILABLE_CU 148
#define KERNEL_STRATEGY_CUDA_B200
#define ARCH_NAME cuda
#endif
