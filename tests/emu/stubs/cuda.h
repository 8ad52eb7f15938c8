// stand-in for <cuda.h> in host builds of the kernels (tests/emu): the tensor map is a plain description of the tiled view
// (what cuTensorMapEncodeTiled is given in Batch::create) that the host stand-ins of the TMA loads walk
#pragma once
#include <cstdint>
struct alignas(64) CUtensorMap_st {
    uint8_t *base;
    uint32_t rank;
    uint32_t box[4];
    uint64_t dims[4];
    uint64_t strides[3];   // bytes, dimensions 1..rank-1 (dimension 0 is dense)
};
typedef CUtensorMap_st CUtensorMap;
