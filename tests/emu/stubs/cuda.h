// stand-in for <cuda.h> in host builds of the kernels (tests/emu): only the opaque tensor map type is needed
#pragma once
#include <cstdint>
struct alignas(64) CUtensorMap_st { uint64_t opaque[16]; };
typedef CUtensorMap_st CUtensorMap;
