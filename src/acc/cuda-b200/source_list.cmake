# this file is included from src/acc/CMakeLists.txt; current path is ${ACC_SRC_PATH} (= src/acc/cuda-b200).
set(CURRENT_ACC_CUDA_B200_SOURCE_DIR ${ACC_SRC_PATH})

set(ACC_HEADER_FILES ${ACC_HEADER_FILES}
        ${CURRENT_ACC_CUDA_B200_SOURCE_DIR}/cuda_b200_spmv.h
        )

set(ACC_SOURCE_FILES ${ACC_SOURCE_FILES}
        ${CURRENT_ACC_CUDA_B200_SOURCE_DIR}/cuda_b200_spmv.cpp
        )

# The kernels live in libspmv_b200.so (built by `python -m spmv_acc_b200.build`, sm_100a only).
# SPMV_B200_ROOT points at a checkout of the spmv-b200 repository.
if (NOT DEFINED SPMV_B200_ROOT)
    message(FATAL_ERROR "KERNEL_STRATEGY=CUDA_B200 needs -DSPMV_B200_ROOT=<path to the spmv-b200 checkout>")
endif ()
include_directories(${SPMV_B200_ROOT}/include)
find_package(CUDAToolkit REQUIRED)
# linked PUBLIC by the kernel library (src/acc/CMakeLists.txt), so the CLI and the benchmark get them transitively
set(SPMV_B200_LINK_LIBS ${SPMV_B200_ROOT}/spmv_acc_b200/lib/libspmv_b200.so CUDA::cudart)
