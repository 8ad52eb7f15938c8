// TEST INFRASTRUCTURE -- a CUDA runtime that is the host: device memory is host memory, streams and events are tokens (every
// "asynchronous" call completes before it returns, which is one valid order of what the streams and events allow), a kernel
// launch runs the kernel's source under warp_emu.hpp.  With it the engine's real host code -- Batch in engine.cu, the C-ABI in
// api.cpp -- runs on a machine without a GPU against the oracle (tests/test_cpu_engine_hostemu.py).  Never part of the product.
#pragma once
#include "warp_emu.hpp"
#include <cstdlib>
#include <cstring>
#include <cstdint>

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline const char *cudaGetErrorName(cudaError_t) { return "emu"; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated runtime error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->major = 10; p->minor = 0; p->multiProcessorCount = 2; return cudaSuccess; }

typedef struct EmuStream_ *cudaStream_t;
typedef struct EmuEvent_ *cudaEvent_t;
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventBlockingSync = 1;
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }

// memory: 256-byte aligned like cudaMalloc
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)std::aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc(p, n); }
constexpr unsigned cudaHostAllocPortable = 1, cudaHostRegisterPortable = 1;
template <typename T> inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; r++) std::memmove((uint8_t *)d + r * dp, (const uint8_t *)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
template <typename T> inline cudaError_t cudaMemcpyFromSymbol(void *d, const T &sym, size_t n) { std::memcpy(d, &sym, n); return cudaSuccess; }

// kernels: two resident blocks per "SM" (small persistent grids keep the emulation quick), attributes ignored
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }

// the one driver entry point the engine asks for: cuTensorMapEncodeTiled -> the plain description tests/emu/stubs/cuda.h holds
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess };
constexpr unsigned cudaEnableDefault = 0;
inline CUresult emuEncodeTiled(CUtensorMap *m, CUtensorMapDataType, cuuint32_t rank, void *base, const cuuint64_t *dims, const cuuint64_t *strides,
                               const cuuint32_t *box, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill) {
    std::memset(m, 0, sizeof *m);
    m->base = (uint8_t *)base;
    m->rank = rank;
    for (cuuint32_t d = 0; d < rank; d++) { m->box[d] = box[d]; m->dims[d] = dims[d]; }
    for (cuuint32_t d = 0; d + 1 < rank; d++) m->strides[d] = strides[d];
    return CUDA_SUCCESS;
}
inline cudaError_t cudaGetDriverEntryPoint(const char *, void **fn, unsigned, cudaDriverEntryPointQueryResult *) { *fn = (void *)&emuEncodeTiled; return cudaSuccess; }

// kernel<<<grid, block, smem, stream>>>(args) is rewritten to EMU_LAUNCH(kernel, grid, block, args) when engine.cu is prepared
namespace warp_emu {
inline void runGrid3(dim3 grid, dim3 block, const std::function<void()> &body) {
    gridDim.y = grid.y; gridDim.z = grid.z;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++) {
            // blockIdx is thread-local: the block's threads copy y / z from these
            gBlockY = y; gBlockZ = z;
            runGrid(grid.x, block.x, body);
        }
    gridDim.y = gridDim.z = 1;
}
}  // namespace warp_emu
#define EMU_LAUNCH(kernel, grid, block, ...) warp_emu::runGrid3(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); })
