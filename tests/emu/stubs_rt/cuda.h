// TEST INFRASTRUCTURE -- stand-in for <cuda.h> when the ENGINE's host code (engine.cu, api.cpp) is built for the host
// (tests/test_cpu_engine_hostemu.py): the tensor map of tests/emu/stubs/cuda.h plus the driver types Batch::create names.
#pragma once
#include "../stubs/cuda.h"
typedef int CUresult;
constexpr CUresult CUDA_SUCCESS = 0;
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
