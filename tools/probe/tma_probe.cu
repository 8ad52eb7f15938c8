// bring-up probe: which TMA tile loads work for u8 planes (run on the GPU box)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../h264bsd_b200/csrc/engine/device_common.cuh"
using namespace b200;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k3(const __grid_constant__ CUtensorMap m, int x, int y, int z, uint8_t *out, int bytes) {
    __shared__ __align__(128) uint8_t buf[1024];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbarInit(&bar, 1); fenceMbarInit(); }
    __syncwarp();
    if (threadIdx.x == 0) { mbarExpectTx(&bar, bytes); tmaLoad3d(buf, &m, x, y, z, &bar); }
    mbarWait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = buf[i];
}
__global__ void k4(const __grid_constant__ CUtensorMap m, int x, int y, int z, int w, uint8_t *out, int bytes) {
    __shared__ __align__(128) uint8_t buf[1024];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbarInit(&bar, 1); fenceMbarInit(); }
    __syncwarp();
    if (threadIdx.x == 0) { mbarExpectTx(&bar, bytes); tmaLoad4d(buf, &m, x, y, z, w, &bar); }
    mbarWait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = buf[i];
}
int main(int argc, char **argv) {
    int ax = argc > 3 ? atoi(argv[1]) : 16, ay = argc > 3 ? atoi(argv[2]) : 8, az = argc > 3 ? atoi(argv[3]) : 1;
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int P = 704, R = 432, NF = 4; size_t fs = (size_t)P * R + 2 * 352 * 216; fs = (fs + 255) & ~255ull;
    uint8_t *pool; cudaMalloc(&pool, fs * NF);
    uint8_t *h = (uint8_t *)malloc(fs * NF);
    for (size_t i = 0; i < fs * NF; i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
    cudaMemcpy(pool, h, fs * NF, cudaMemcpyHostToDevice);
    uint8_t *out; cudaMalloc(&out, 1024); uint8_t ho[1024];
    for (int boxh = 21; boxh >= 21; boxh -= 13) {
        CUtensorMap m;
        cuuint64_t dims[3] = {P, R, NF}; cuuint64_t st[2] = {P, fs}; cuuint32_t box[3] = {32, (cuuint32_t)boxh, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pool, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode3 boxh=%d -> %d\n", boxh, (int)r);
        int coords[1][3] = {{ax, ay, az}};
        for (auto &c : coords) {
            k3<<<1, 32>>>(m, c[0], c[1], c[2], out, 32 * boxh);
            cudaError_t e = cudaDeviceSynchronize();
            printf(" k3 (%d,%d,%d): %s", c[0], c[1], c[2], cudaGetErrorString(e));
            if (e == cudaSuccess) {
                cudaMemcpy(ho, out, 1024, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int yy = 0; yy < boxh; yy++) for (int xx = 0; xx < 32; xx++) {
                    int gx = c[0] + xx, gy = c[1] + yy; uint8_t exp = (gx < P && gy < R) ? h[(size_t)c[2] * fs + (size_t)gy * P + gx] : 0;
                    bad += ho[yy * 32 + xx] != exp;
                }
                printf(" bad=%d", bad);
            }
            printf("\n");
            if (e != cudaSuccess) return 1;
        }
    }
    {
        CUtensorMap m;
        cuuint64_t dims[4] = {352, 216, 2, NF}; cuuint64_t st[3] = {352, 352 * 216, fs}; cuuint32_t box[4] = {16, 9, 2, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, pool + (size_t)P * R, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode4 -> %d\n", (int)r);
        k4<<<1, 32>>>(m, 5, 7, 0, 2, out, 288);
        cudaError_t e = cudaDeviceSynchronize();
        printf(" k4: %s", cudaGetErrorString(e));
        if (e == cudaSuccess) {
            cudaMemcpy(ho, out, 1024, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int pl = 0; pl < 2; pl++) for (int yy = 0; yy < 9; yy++) for (int xx = 0; xx < 16; xx++)
                bad += ho[pl * 144 + yy * 16 + xx] != h[2 * fs + (size_t)P * R + (size_t)pl * 352 * 216 + (size_t)(7 + yy) * 352 + 5 + xx];
            printf(" bad=%d", bad);
        }
        printf("\n");
    }
    return 0;
}
