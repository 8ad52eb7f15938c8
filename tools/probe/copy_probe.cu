// copy_probe.cu -- what does the memory system give a pass-A-like copy?  Frames in the strip layout (a macroblock column is
// contiguous: 256 B of luma per macroblock, 128 B of chroma in a second region), every warp moves "chunks" of n vertically
// adjacent macroblocks from a reference frame to the current frame of the same stream.  Variants:
//   bulk    cp.async.bulk global -> shared -> global through a per-warp staging buffer (what passAKernel does)
//   regs    16-byte loads / stores through registers, 4 in flight per lane
//   memcpy  cudaMemcpyAsync device to device of the same bytes (the ceiling MEASURED_PEAKS.json quotes)
// and knobs: chunk length, warps per SM, with / without a 96-byte record read per macroblock (stride widthMbs * 96).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_probe copy_probe.cu && ./copy_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbarTryWait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void bulkStore(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smemAddr(src)), "r"(bytes) : "memory");
}

struct Params {
    uint8_t *pool;
    const uint8_t *recs;
    unsigned long long frameStride, offC;
    int nStreams, widthMbs, heightMbs, rowsY, rowsC, chunkRows, chunksPerCol, readRecs, sleepNs, colWalk;
    uint32_t totalChunks;
    uint32_t *ticket, *sink;
};

template <int kWarps, int kStage>
__global__ void __launch_bounds__(kWarps * 32) bulkCopy(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *stage = smem + (size_t)warp * (kStage * 384 + 128);
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage + kStage * 384);
    if (lane == 0) { mbarInit(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    uint32_t phase = 0, acc = 0;
    const uint32_t nWarps = gridDim.x * kWarps, chunksPerStream = (uint32_t)p.chunksPerCol * p.widthMbs;
    // colWalk: a ticket is a whole column, its chunks are walked one after the other by the same warp
    uint32_t chunk = (blockIdx.x * kWarps + warp) * (p.colWalk ? p.chunksPerCol : 1);
    while (chunk < p.totalChunks) {
        uint32_t next = 0;
        const bool lastOfCol = !p.colWalk || (chunk + 1) % p.chunksPerCol == 0;
        if (lane == 0 && lastOfCol) next = (atomicAdd(p.ticket, 1u) + nWarps) * (p.colWalk ? p.chunksPerCol : 1);
        if (!lastOfCol) next = chunk + 1;
        const uint32_t s = chunk / chunksPerStream, c2 = chunk - s * chunksPerStream;
        const int mbx = c2 / p.chunksPerCol, row0 = (c2 - mbx * p.chunksPerCol) * p.chunkRows;
        const int n = min(p.chunkRows, p.heightMbs - row0);
        if (p.readRecs) {
            for (int l = lane; l < n; l += 32) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(p.recs + ((size_t)s * p.widthMbs * p.heightMbs + (size_t)(row0 + l) * p.widthMbs + mbx) * 96);
                const uint4 hw = __ldg(reinterpret_cast<const uint4 *>(rw));
                acc += hw.x + hw.y + hw.z + __ldg(rw + 4) + __ldg(rw + 7) + __ldg(rw + 8);
            }
        }
        uint8_t *cur = p.pool + (unsigned long long)(s * 2 + 1) * p.frameStride, *ref = p.pool + (unsigned long long)(s * 2) * p.frameStride;
        const size_t offY = ((size_t)(mbx + 2) * p.rowsY + row0 * 16 + 32) * 16, offCc = p.offC + ((size_t)(mbx + 2) * p.rowsC + row0 * 8 + 16) * 16;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            mbarExpectTx(bar, 384u * n);
            bulkLoad(stage, ref + offY, 256u * n, bar);
            bulkLoad(stage + kStage * 256, ref + offCc, 128u * n, bar);
        }
        while (!mbarTryWait(bar, phase)) { if (p.sleepNs) __nanosleep(p.sleepNs); }
        phase ^= 1;
        if (lane == 0) {
            bulkStore(cur + offY, stage, 256u * n);
            bulkStore(cur + offCc, stage + kStage * 256, 128u * n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) *p.sink = acc;
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) regCopy(const Params p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    const uint32_t nWarps = gridDim.x * kWarps, chunksPerStream = (uint32_t)p.chunksPerCol * p.widthMbs;
    uint32_t chunk = blockIdx.x * kWarps + warp;
    while (chunk < p.totalChunks) {
        uint32_t next = 0;
        if (lane == 0) next = atomicAdd(p.ticket, 1u) + nWarps;
        const uint32_t s = chunk / chunksPerStream, c2 = chunk - s * chunksPerStream;
        const int mbx = c2 / p.chunksPerCol, row0 = (c2 - mbx * p.chunksPerCol) * p.chunkRows;
        const int n = min(p.chunkRows, p.heightMbs - row0);
        if (p.readRecs && lane < n) {
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(p.recs + ((size_t)s * p.widthMbs * p.heightMbs + (size_t)(row0 + lane) * p.widthMbs + mbx) * 96);
            const uint4 hw = __ldg(reinterpret_cast<const uint4 *>(rw));
            acc += hw.x + hw.y + hw.z + __ldg(rw + 4) + __ldg(rw + 7) + __ldg(rw + 8);
        }
        uint8_t *cur = p.pool + (unsigned long long)(s * 2 + 1) * p.frameStride;
        const long long delta = -(long long)p.frameStride;
        const size_t offY = ((size_t)(mbx + 2) * p.rowsY + row0 * 16 + 32) * 16, offCc = p.offC + ((size_t)(mbx + 2) * p.rowsC + row0 * 8 + 16) * 16;
        // luma: n * 16 units of 16 bytes, chroma n * 8
        for (int u = lane; u < n * 16; u += 128) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) if (u + 32 * j < n * 16) v[j] = __ldg(reinterpret_cast<const uint4 *>(cur + offY + delta) + u + 32 * j);
#pragma unroll
            for (int j = 0; j < 4; j++) if (u + 32 * j < n * 16) reinterpret_cast<uint4 *>(cur + offY)[u + 32 * j] = v[j];
        }
        for (int u = lane; u < n * 8; u += 128) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) if (u + 32 * j < n * 8) v[j] = __ldg(reinterpret_cast<const uint4 *>(cur + offCc + delta) + u + 32 * j);
#pragma unroll
            for (int j = 0; j < 4; j++) if (u + 32 * j < n * 8) reinterpret_cast<uint4 *>(cur + offCc)[u + 32 * j] = v[j];
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
    if (acc == 0x12345678u) *p.sink = acc;
}

int main() {
    Params p{};
    p.nStreams = 512; p.widthMbs = 120; p.heightMbs = 68;
    p.rowsY = 68 * 16 + 64; p.rowsC = 68 * 8 + 32;
    p.offC = (unsigned long long)124 * p.rowsY * 16;
    p.frameStride = (p.offC + (unsigned long long)124 * p.rowsC * 16 + 255) & ~255ull;
    const size_t poolBytes = (size_t)p.nStreams * 2 * p.frameStride, recBytes = (size_t)p.nStreams * 8160 * 96;
    CK(cudaMalloc(&p.pool, poolBytes)); CK(cudaMemset(p.pool, 7, poolBytes));
    uint8_t *recs; CK(cudaMalloc(&recs, recBytes)); CK(cudaMemset(recs, 0, recBytes)); p.recs = recs;
    CK(cudaMalloc(&p.ticket, 8)); p.sink = p.ticket + 1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double bytes = 2.0 * 384 * 8160 * p.nStreams;
    auto time = [&](const char *what, auto launch) -> int {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaMemset(p.ticket, 0, 8));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        printf("%-72s %.3f ms  %.0f GB/s\n", what, best, bytes / best / 1e6);
        return 0;
    };
    time("cudaMemcpyAsync D2D of the same number of bytes (read + write counted)", [&] { cudaMemcpyAsync(p.pool, p.pool + poolBytes / 2, (size_t)(bytes / 2), cudaMemcpyDeviceToDevice); });
    auto cfg = [&](int rows) { p.chunksPerCol = (p.heightMbs + rows - 1) / rows; p.chunkRows = (p.heightMbs + p.chunksPerCol - 1) / p.chunksPerCol; p.totalChunks = (uint32_t)p.chunksPerCol * p.widthMbs * p.nStreams; };
    char name[200];
#define BULK(W, S, ctasPerSm, rr, slp)                                                                                        \
    {                                                                                                                         \
        cfg(S); p.readRecs = rr; p.sleepNs = slp;                                                                             \
        const int smem = W * (S * 384 + 128);                                                                                 \
        CK(cudaFuncSetAttribute(bulkCopy<W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                          \
        snprintf(name, sizeof name, "bulk: %d warps x %d CTAs/SM, chunk %d rows (%d), records %d, sleep %d, colWalk %d", W, ctasPerSm, S, p.chunkRows, rr, slp, p.colWalk); \
        time(name, [&] { bulkCopy<W, S><<<148 * ctasPerSm, W * 32, smem>>>(p); });                                            \
    }
    p.colWalk = 0;
    BULK(4, 14, 5, 1, 160)
    BULK(4, 34, 2, 1, 160)
    BULK(4, 68, 1, 1, 160)
    BULK(1, 68, 4, 1, 160)
    BULK(1, 68, 3, 1, 160)
    BULK(1, 68, 2, 1, 160)
    BULK(1, 34, 4, 1, 160)
    BULK(1, 34, 8, 1, 160)
    BULK(2, 68, 2, 1, 160)
    p.colWalk = 1;
    BULK(4, 14, 5, 1, 160)
    BULK(4, 17, 4, 1, 160)
    BULK(4, 34, 2, 1, 160)
    p.colWalk = 0;
#define REGS(W, rows, ctasPerSm, rr)                                                                                          \
    {                                                                                                                         \
        cfg(rows); p.readRecs = rr;                                                                                           \
        snprintf(name, sizeof name, "regs: %d warps x %d CTAs/SM, chunk %d rows, records %d", W, ctasPerSm, p.chunkRows, rr); \
        time(name, [&] { regCopy<W><<<148 * ctasPerSm, W * 32>>>(p); });                                                      \
    }
    REGS(4, 14, 5, 1)
    REGS(8, 14, 4, 1)
    REGS(8, 14, 8, 1)
    REGS(8, 23, 8, 0)
    return 0;
}
