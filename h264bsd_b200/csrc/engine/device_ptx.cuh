// device_ptx.cuh -- every line of inline PTX of the engine's ticketed kernels: acquire / release accesses, the global timer,
// mbarrier and TMA tile loads, dp4a and cvt.pack.  A classic include guard
// instead of #pragma once on purpose: tests/emu/ defines the guard and supplies host stand-ins for these few functions, so
// that the kernels built on them compile for the host as they are.
#ifndef B200_DEVICE_PTX_CUH
#define B200_DEVICE_PTX_CUH
#include <cstdint>
#include <cuda.h>

namespace b200 {

// ---- inter-warp completion flags (one 32-bit word per stream x macroblock) -----------------------
__device__ __forceinline__ uint32_t ldAcquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelease(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globalTimerNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- TMA / mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fenceMbarInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smemAddr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tmaLoad3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smemAddr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smemAddr(bar))
        : "memory");
}
__device__ __forceinline__ void tmaLoad4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smemAddr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smemAddr(bar))
        : "memory");
}

// ---- 1-D bulk copies (copy_bulk_kernel.cuh): addresses and size are multiples of 16 bytes ----------------------------
__device__ __forceinline__ void bulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulkStore(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smemAddr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// until at most N of this thread's most recent bulk groups are still reading their shared-memory source
template <int N>
__device__ __forceinline__ void bulkWaitRead() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// until all of this thread's bulk groups are complete (their writes performed)
__device__ __forceinline__ void bulkWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- integer SIMD used by the motion compensation (recon_kernel.cuh) ---------------------------------
// dp4a with unsigned pels and signed taps: the 6-tap filter (1,-5,20,20,-5,1) is two dot products
__device__ __forceinline__ int dp4aUS(uint32_t pels, int taps, int acc) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pels), "r"(taps), "r"(acc));
    return d;
}
// two-way dot product of signed 16-bit halves with signed bytes: acc + lo16(a) * byte0(b) + hi16(a) * byte1(b)
__device__ __forceinline__ int dp2aLoSS(uint32_t a, int b, int acc) {
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
// four ints saturated to bytes, p0 in the low byte: two cvt.pack (I2IP) instead of four clamps, shifts and ors
__device__ __forceinline__ uint32_t pack4sat(int p0, int p1, int p2, int p3) {
    uint32_t hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(p3), "r"(p2));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(p1), "r"(p0), "r"(hi));
    return r;
}

}  // namespace b200

#endif  // B200_DEVICE_PTX_CUH
