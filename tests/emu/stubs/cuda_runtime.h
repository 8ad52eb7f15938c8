// stand-in for <cuda_runtime.h> in host builds of the kernels (tests/emu): everything comes from warp_emu.hpp
#pragma once
