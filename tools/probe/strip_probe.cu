// bring-up probe (run on the GPU box): does a tensor map whose strides are NOT ascending work?  The strip layout keeps a plane
// as [strip][row][16 bytes]; a TMA box {16, nx, nr} over dims (xin, strip, row) would land in shared memory as a raster window
// of pitch 16 * nx -- if cuTensorMapEncodeTiled / the hardware accept stride(strip) > stride(row).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../h264bsd_b200/csrc/engine/device_ptx.cuh"
using namespace b200;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k4(const __grid_constant__ CUtensorMap m, int x, int y, int z, int w, uint8_t *out, int bytes) {
    __shared__ __align__(128) uint8_t buf[4096];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbarInit(&bar, 1); fenceMbarInit(); }
    __syncwarp();
    if (threadIdx.x == 0) { mbarExpectTx(&bar, bytes); tmaLoad4d(buf, &m, x, y, z, w, &bar); }
    while (!mbarTryWait(&bar, 0)) {}
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = buf[i];
}
int main() {
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int strips = 44, rows = 432, NF = 3; const size_t stripBytes = (size_t)rows * 16, fs = ((size_t)strips * stripBytes + 255) & ~255ull;
    uint8_t *pool; cudaMalloc(&pool, fs * NF);
    uint8_t *h = (uint8_t *)malloc(fs * NF);
    for (size_t i = 0; i < fs * NF; i++) h[i] = (uint8_t)(i * 7 + (i >> 8) + (i >> 16));
    cudaMemcpy(pool, h, fs * NF, cudaMemcpyHostToDevice);
    uint8_t *out; cudaMalloc(&out, 4096); uint8_t ho[4096];
    int rc = 0;
    for (int nx = 1; nx <= 3; nx++) for (int nr : {9, 16, 21}) {
        CUtensorMap m;
        cuuint64_t dims[4] = {16, (cuuint64_t)strips, (cuuint64_t)rows, NF};
        cuuint64_t st[3] = {stripBytes, 16, fs};
        cuuint32_t box[4] = {16, (cuuint32_t)nx, (cuuint32_t)nr, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, pool, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode nx=%d nr=%d -> %d\n", nx, nr, (int)r);
        if (r != CUDA_SUCCESS) { rc = 1; continue; }
        const int coords[4][3] = {{5, 7, 1}, {0, 0, 0}, {41, 417, 2}, {43, 431, 2}};   // last ones run off the tensor (zero fill)
        for (auto &c : coords) {
            const int bytes = 16 * nx * nr;
            k4<<<1, 32>>>(m, 0, c[0], c[1], c[2], out, bytes);
            cudaError_t e = cudaDeviceSynchronize();
            printf("  strip %d row %d frame %d: %s", c[0], c[1], c[2], cudaGetErrorString(e));
            if (e != cudaSuccess) { printf("\n"); return 2; }
            cudaMemcpy(ho, out, 4096, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int yy = 0; yy < nr; yy++) for (int sx = 0; sx < nx; sx++) for (int xx = 0; xx < 16; xx++) {
                const int s = c[0] + sx, y = c[1] + yy;
                const uint8_t exp = (s < strips && y < rows) ? h[(size_t)c[2] * fs + (size_t)s * stripBytes + (size_t)y * 16 + xx] : 0;
                bad += ho[(yy * nx + sx) * 16 + xx] != exp;
            }
            printf(" bad=%d\n", bad);
            rc |= bad != 0;
        }
    }
    printf(rc ? "PROBE FAILED\n" : "PROBE OK: raster windows out of strip-major planes\n");
    return rc;
}
