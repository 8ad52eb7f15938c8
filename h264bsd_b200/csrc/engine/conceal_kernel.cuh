// conceal_kernel.cuh -- spatial error concealment on the device (error path of the reconstruction).
// Needs nothing but the frame addressing helpers, warp shuffles and plain loads / stores, so the same source also compiles
// for the host: tests/emu/ runs it there, 32 threads standing in for the lanes of a warp, against the CPU oracle.
#pragma once
#include "frame_addr.cuh"

namespace b200 {

// =====================================================================================================
// Spatial error concealment (ConcealMb, h264bsd_conceal.c:266-600 with Transform :610-639): a macroblock that never arrived
// and has no reference picture to be copied from is estimated from the edge pels of its neighbours (those decoded, or
// concealed before it: the record's waitMask names them, B200_CN_*).  Per plane the edge pels are summed in four groups per
// side; from the sums come a DC value and the lowest horizontal and vertical frequency; the "transform" of those three
// gives a 4x4 grid of values, each repeated over a quarter of the block's width and height.
//
// Error path only.  The entries of a stream depend on each other in list order, so one warp walks a stream's list from the
// first entry to the last; streams are independent (one warp each).  Runs after every other reconstruction of the picture
// and before its filter.
// =====================================================================================================
constexpr int kConcealWarps = 4;

__global__ void __launch_bounds__(kConcealWarps * 32) concealKernel(const ReconParams p) {
    const PoolGeom &g = p.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t s = blockIdx.x * kConcealWarps + warp;
    if (s >= (uint32_t)g.nStreams) return;
    const StreamJob job = p.jobs[s];
    if (!job.nE) return;
    uint8_t *cur = framePtr(p.pool, g, s * (uint32_t)g.numSlots + job.curSlot);
    const uint16_t *list = job.orderE;
#pragma unroll 1
    for (uint32_t e = 0; e < job.nE; e++) {
        const uint32_t mb = __ldg(list + e);
        const uint32_t mask = __ldg(reinterpret_cast<const uint32_t *>(job.recs + mb) + 7) & 0xFu;   // waitMask
        const int mby = mbRowOf(mb, g), mbx = (int)(mb - (uint32_t)mby * g.widthMbs);
        const int hor = (int)(mask & 1u) + (int)((mask >> 1) & 1u), ver = (int)((mask >> 2) & 1u) + (int)((mask >> 3) & 1u);
#pragma unroll 1
        for (int pl = 0; pl < 3; pl++) {
            const int n = pl ? 8 : 16, grp = n >> 2;
            // where sample (x, y) of this plane lies (x, y relative to the macroblock; -1 and n reach into the neighbours)
            auto at = [&](int x, int y) -> uint8_t * {
                return pl ? chromaAt(cur, g, pl - 1, mbx * 8 + x, mby * 8 + y) : lumaAt(cur, g, mbx * 16 + x, mby * 16 + y);
            };
            // lanes 0..15: side = lane / 4 (above, below, left, right), group = lane % 4: the sum of that group's edge pels
            // (plain loads: the pels may have been written by this warp a moment ago)
            int sum = 0;
            {
                const int side = lane >> 2, gi = lane & 3;
                if (lane < 16 && ((mask >> side) & 1u)) {
                    for (int k = 0; k < grp; k++) {
                        const int t = gi * grp + k;
                        const uint8_t *q = side == 0 ? at(t, -1) : side == 1 ? at(t, n) : side == 2 ? at(-1, t) : at(n, t);
                        sum += *reinterpret_cast<const volatile uint8_t *>(q);
                    }
                }
            }
            int v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = __shfl_sync(0xffffffffu, sum, i);
            const int sa = v[0] + v[1] + v[2] + v[3], sb = v[4] + v[5] + v[6] + v[7];
            const int sl = v[8] + v[9] + v[10] + v[11], sr = v[12] + v[13] + v[14] + v[15];
            int dc = sa + sb + sl + sr;                                            // firstPhase[0]
            int fh = (v[0] + v[1] - v[2] - v[3]) + (v[4] + v[5] - v[6] - v[7]);    // firstPhase[1]
            int fv = (v[8] + v[9] - v[10] - v[11]) + (v[12] + v[13] - v[14] - v[15]);   // firstPhase[4]
            const int sh = pl ? 2 : 3;
            if (!hor && (mask & 4u) && (mask & 8u)) fh = (sl - sr) >> (sh + 2);
            else if (hor) fh >>= (sh + hor);
            if (!ver && (mask & 1u) && (mask & 2u)) fv = (sa - sb) >> (sh + 2);
            else if (ver) fv >>= (sh + ver);
            const int j = hor + ver;
            if (j == 1) dc >>= (sh + 1);
            else if (j == 2) dc >>= (sh + 2);
            else if (j == 3) dc = (21 * dc) >> (sh + 7);
            else dc >>= (sh + 3);
            // Transform(): value of grid cell (row r, column c) = column term + row term
            auto colTerm = [&](int c) { return c == 0 ? dc + fh : c == 1 ? dc + (fh >> 1) : c == 2 ? dc - (fh >> 1) : dc - fh; };
            auto rowTerm = [&](int r) { return r == 0 ? fv : r == 1 ? (fv >> 1) : r == 2 ? -(fv >> 1) : -fv; };
            if (pl == 0) {
                // 16 rows x 16 pels: lane = 2 * row + half, 8 pels = two grid cells of 4 pels
                const int row = lane >> 1, half = lane & 1;
                const int rt = rowTerm(row >> 2);
                const uint32_t c0 = (uint32_t)clip255(colTerm(half * 2) + rt) * 0x01010101u;
                const uint32_t c1 = (uint32_t)clip255(colTerm(half * 2 + 1) + rt) * 0x01010101u;
                *reinterpret_cast<uint2 *>(at(half * 8, row)) = make_uint2(c0, c1);
            } else {
                // 8 rows x 8 pels: lane = 4 * row + cell, 2 pels = one grid cell
                const int row = lane >> 2, cell = lane & 3;
                const uint32_t c0 = (uint32_t)clip255(colTerm(cell) + rowTerm(row >> 1));
                *reinterpret_cast<uint16_t *>(at(cell * 2, row)) = (uint16_t)(c0 * 0x0101u);
            }
        }
        __syncwarp();   // orders this entry's stores before the next entry's loads (same warp)
    }
}

}  // namespace b200
