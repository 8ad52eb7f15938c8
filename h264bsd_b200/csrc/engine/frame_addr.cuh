// frame_addr.cuh -- where a frame, a luma / chroma sample and a macroblock live in the frame pool (strip layout, see
// pool_geom.hpp), plus the clips.  No PTX, no CUDA runtime types: this header and the kernels that need nothing else
// (conceal_kernel.cuh) also compile for the host, where tests/emu runs them lane by lane against the CPU oracle.
#pragma once
#include <cstdint>
#include <cstddef>
#include "pool_geom.hpp"

namespace b200 {

__device__ __forceinline__ uint8_t *framePtr(uint8_t *pool, const PoolGeom &g, uint32_t frame) {
    return pool + (unsigned long long)frame * g.frameStride;
}
// sample (x, y) of the picture; x in [-32, W + 32), y in [-32, H + 32) reach into the replicated border
__device__ __forceinline__ uint8_t *lumaAt(uint8_t *frame, const PoolGeom &g, int x, int y) {
    const int xx = x + kPadY, yy = y + kPadY;
    return frame + ((size_t)(xx >> 4) * g.rowsY + yy) * 16 + (xx & 15);
}
__device__ __forceinline__ uint8_t *chromaAt(uint8_t *frame, const PoolGeom &g, int plane, int x, int y) {
    const int xx = x + kPadC, yy = y + kPadC;
    return frame + g.offC + ((size_t)(xx >> 3) * g.rowsC + yy) * 16 + plane * 8 + (xx & 7);
}
// the 256 luma bytes (16 rows x 16) and the 128 chroma bytes (8 rows x [8 Cb | 8 Cr]) of macroblock (mbx, mby)
__device__ __forceinline__ uint8_t *mbLuma(uint8_t *frame, const PoolGeom &g, int mbx, int mby) {
    return frame + ((size_t)(mbx + kPadMbs) * g.rowsY + mby * 16 + kPadY) * 16;
}
__device__ __forceinline__ uint8_t *mbChroma(uint8_t *frame, const PoolGeom &g, int mbx, int mby) {
    return frame + g.offC + ((size_t)(mbx + kPadMbs) * g.rowsC + mby * 8 + kPadC) * 16;
}

// row of macroblock address mb: floor(mb / widthMbs) without a division.  floor((2 mb + 1) / (2 w)) equals floor(mb / w), and
// with inv = ceil(2^31 / w) the product (2 mb + 1) * inv / 2^32 overshoots it by less than 2^-15 while the next integer is at
// least 1 / (2 w) away: exact for mb < 65536 and w < 16384 -- including w == 1, where ceil(2^32 / w) would not fit 32 bits.
__device__ __forceinline__ int mbRowOf(uint32_t mb, const PoolGeom &g) { return (int)__umulhi(2u * mb + 1u, g.invWidthMbs); }
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int clip3(int lo, int hi, int v) { return min(max(v, lo), hi); }

}  // namespace b200
