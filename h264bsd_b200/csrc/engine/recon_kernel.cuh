// recon_kernel.cuh -- macroblock reconstruction: inverse zig-zag + dequant + IDCT, intra /
// inter prediction, add residual, write.  One warp per macroblock, persistent CTAs, tickets
// handed out in wavefront order so that a warp only ever waits on warps that already started.
//
// Restates, as sm_100a device code: h264bsd_transform.c:97-401 (residual),
// h264bsd_intra_prediction.c:478-1830, h264bsd_inter_prediction.c:361-482 +
// h264bsd_reconstruct.c:109-2367 (motion compensation), h264bsd_image.c:81-344 (write).
#pragma once
#include "device_common.cuh"

namespace b200 {

constexpr int kReconWarps = 8;
constexpr int kChunkA = 8;   // consecutive pass-A list entries per warp (TMA of entry i+1 overlaps the math of entry i)
constexpr int kChunkB = 8;   // most consecutive pass-B (wavefront) entries per warp task (ReconParams::chunkB <= kChunkB)


struct __align__(128) InterWarpSmem {
    uint8_t lumaWin[3][kLumaBoxW * kLumaBoxH + 16];           // 3 x 1024: two for the prefetch double buffer, [2] for the
    uint8_t chromaWin[3][2 * kChromaBoxW * kChromaBoxH + 64];  // 3 x 640   second partition of a two-partition step
    int16_t res[24][16];                                       // 768
    uint8_t pred[384];                                         // 384: multi-partition macroblocks only
    uint32_t meta[kChunkA][8];                                 // per entry: head words 0..3, refSlots, mv[0], mb address
    uint64_t mbar[3];
    uint32_t pad[10];
};
struct __align__(16) IntraWarpSmem {
    int16_t res[24][16];
    uint8_t itY[17][24];   // rows -1..15, cols -1..19 (+pad)
    uint8_t itC[2][9][12]; // rows -1..7, cols -1..7 (+pad)
    uint8_t stage[16];
};

// ---- residual -----------------------------------------------------------------------------------
// h264bsd_transform.c:58-59; class 0 = both coordinates even, 2 = both odd, 1 = mixed
__device__ __constant__ uint8_t cLevelScale[6][4] = {{10, 13, 16, 0}, {11, 14, 18, 0}, {13, 16, 20, 0}, {14, 18, 23, 0}, {16, 20, 25, 0}, {18, 23, 29, 0}};
__device__ __forceinline__ int levelScale(int qpMod, int cls) { return cLevelScale[qpMod][cls]; }

// h264bsdProcessBlock (transform.c:97-234): lev in zig-zag order -> out[16] raster residual
__device__ __forceinline__ bool idctBlock(const int16_t *lev, int qp, bool dcPreset, int dcValue, int *out) {
    int qpDiv = qp / 6, qpMod = qp - 6 * qpDiv;
    int s0 = levelScale(qpMod, 0) << qpDiv, s1 = levelScale(qpMod, 1) << qpDiv, s2 = levelScale(qpMod, 2) << qpDiv;
    int d[16];
    // zig-zag position -> raster: 0,1,4,8,5,2,3,6,9,12,13,10,7,11,14,15
    d[0] = lev[0] * s0;   d[1] = lev[1] * s1;   d[4] = lev[2] * s1;   d[8] = lev[3] * s0;
    d[5] = lev[4] * s2;   d[2] = lev[5] * s0;   d[3] = lev[6] * s1;   d[6] = lev[7] * s1;
    d[9] = lev[8] * s1;   d[12] = lev[9] * s1;  d[13] = lev[10] * s2; d[10] = lev[11] * s0;
    d[7] = lev[12] * s2;  d[11] = lev[13] * s1; d[14] = lev[14] * s1; d[15] = lev[15] * s2;
    if (dcPreset) d[0] = dcValue;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        int t0 = d[i] + d[i + 2], t1 = d[i] - d[i + 2];
        int t2 = (d[i + 1] >> 1) - d[i + 3], t3 = d[i + 1] + (d[i + 3] >> 1);
        d[i] = t0 + t3; d[i + 1] = t1 + t2; d[i + 2] = t1 - t2; d[i + 3] = t0 - t3;
    }
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int t0 = d[i] + d[i + 8], t1 = d[i] - d[i + 8];
        int t2 = (d[i + 4] >> 1) - d[i + 12], t3 = d[i + 4] + (d[i + 12] >> 1);
        out[i] = (t0 + t3 + 32) >> 6; out[i + 4] = (t1 + t2 + 32) >> 6;
        out[i + 8] = (t1 - t2 + 32) >> 6; out[i + 12] = (t0 - t3 + 32) >> 6;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) bad |= (unsigned)(out[i] + 512) > 1023u;
    return bad;
}

// h264bsdProcessLumaDc (transform.c:255-338); returns the element for DC-matrix raster index `pick`
__device__ __forceinline__ int lumaDcPick(const int16_t *lev, int qp, int pick) {
    int d[16];
    d[0] = lev[0]; d[1] = lev[1]; d[4] = lev[2]; d[8] = lev[3]; d[5] = lev[4]; d[2] = lev[5]; d[3] = lev[6]; d[6] = lev[7];
    d[9] = lev[8]; d[12] = lev[9]; d[13] = lev[10]; d[10] = lev[11]; d[7] = lev[12]; d[11] = lev[13]; d[14] = lev[14]; d[15] = lev[15];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        int t0 = d[i] + d[i + 2], t1 = d[i] - d[i + 2], t2 = d[i + 1] - d[i + 3], t3 = d[i + 1] + d[i + 3];
        d[i] = t0 + t3; d[i + 1] = t1 + t2; d[i + 2] = t1 - t2; d[i + 3] = t0 - t3;
    }
    int v[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int t0 = d[i] + d[i + 8], t1 = d[i] - d[i + 8], t2 = d[i + 4] - d[i + 12], t3 = d[i + 4] + d[i + 12];
        v[i] = t0 + t3; v[i + 4] = t1 + t2; v[i + 8] = t1 - t2; v[i + 12] = t0 - t3;
    }
    int sel = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) sel = (i == pick) ? v[i] : sel;
    int qpDiv = qp / 6, ls = levelScale(qp - 6 * qpDiv, 0);
    if (qp >= 12) return sel * (ls << (qpDiv - 2));
    return (sel * ls + (qpDiv == 1 ? 1 : 2)) >> (2 - qpDiv);
}

// h264bsdProcessChromaDc (transform.c:359-401): lev[0..3] of one plane, element `pick`
__device__ __forceinline__ int chromaDcPick(const int16_t *lev, int qp, int pick) {
    int qpDiv = qp / 6, ls = levelScale(qp - 6 * qpDiv, 0), shift = 1;
    if (qp >= 6) { ls <<= (qpDiv - 1); shift = 0; }
    int t0 = lev[0] + lev[2], t1 = lev[0] - lev[2], t2 = lev[1] - lev[3], t3 = lev[1] + lev[3];
    int v = pick == 0 ? t0 + t3 : pick == 1 ? t0 - t3 : pick == 2 ? t1 + t2 : t1 - t2;
    return (v * ls) >> shift;
}

// ---- motion compensation from the staged window -----------------------------------------------------
// window sample at picture position (xInt + dx, yInt + dy): `win` points at the sample (xInt-2, yInt-2) inside the box
__device__ __forceinline__ int W_(const uint8_t *win, int dx, int dy) { return win[(dy + 2) * kLumaBoxW + dx + 2]; }
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }
__device__ __forceinline__ int hsum(const uint8_t *win, int x, int y) {
    return tap6(W_(win, x - 2, y), W_(win, x - 1, y), W_(win, x, y), W_(win, x + 1, y), W_(win, x + 2, y), W_(win, x + 3, y));
}
__device__ __forceinline__ int vsum(const uint8_t *win, int x, int y) {
    return tap6(W_(win, x, y - 2), W_(win, x, y - 1), W_(win, x, y), W_(win, x, y + 1), W_(win, x, y + 2), W_(win, x, y + 3));
}
// clause 8.4.2.2.1; dispatch table of h264bsdPredictSamples (reconstruct.c:1848-1927)
__device__ __forceinline__ int lumaQpel(const uint8_t *win, int x, int y, int xf, int yf) {
    if ((xf | yf) == 0) return W_(win, x, y);
    if (yf == 0) {
        int b = clip255((hsum(win, x, y) + 16) >> 5);
        if (xf == 2) return b;
        return (b + W_(win, x + (xf >> 1), y) + 1) >> 1;
    }
    if (xf == 0) {
        int h = clip255((vsum(win, x, y) + 16) >> 5);
        if (yf == 2) return h;
        return (h + W_(win, x, y + (yf >> 1)) + 1) >> 1;
    }
    if (xf != 2 && yf != 2) {
        int b = clip255((hsum(win, x, y + (yf >> 1)) + 16) >> 5);
        int h = clip255((vsum(win, x + (xf >> 1), y) + 16) >> 5);
        return (b + h + 1) >> 1;
    }
    int j = clip255((tap6(hsum(win, x, y - 2), hsum(win, x, y - 1), hsum(win, x, y), hsum(win, x, y + 1),
                          hsum(win, x, y + 2), hsum(win, x, y + 3)) + 512) >> 10);
    if (xf == 2 && yf == 2) return j;
    if (xf == 2) {
        int b = clip255((hsum(win, x, y + (yf >> 1)) + 16) >> 5);
        return (j + b + 1) >> 1;
    }
    int h = clip255((vsum(win, x + (xf >> 1), y) + 16) >> 5);
    return (j + h + 1) >> 1;
}

// ---- Intra4x4 sample prediction (intra_prediction.c:1493-1830, clause 8.3.1.2) -----------------------
// A(i): above row sample i in -1..7 (i >= 4 replicated from 3 when above-right is unavailable); L(i): left column
template <typename FA, typename FL>
__device__ __forceinline__ int intra4x4Pel(int mode, int x, int y, bool avA, bool avB, FA A, FL L) {
    switch (mode) {
        case 0: return A(x);
        case 1: return L(y);
        case 2: {
            if (avA && avB) return (A(0) + A(1) + A(2) + A(3) + L(0) + L(1) + L(2) + L(3) + 4) >> 3;
            if (avA) return (L(0) + L(1) + L(2) + L(3) + 2) >> 2;
            if (avB) return (A(0) + A(1) + A(2) + A(3) + 2) >> 2;
            return 128;
        }
        case 3:
            if (x == 3 && y == 3) return (A(6) + 3 * A(7) + 2) >> 2;
            return (A(x + y) + 2 * A(x + y + 1) + A(x + y + 2) + 2) >> 2;
        case 4:
            if (x > y) return (A(x - y - 2) + 2 * A(x - y - 1) + A(x - y) + 2) >> 2;
            if (x < y) return ((y - x - 2 >= 0 ? L(y - x - 2) : A(-1)) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
            return (A(0) + 2 * A(-1) + L(0) + 2) >> 2;
        case 5: {
            int z = 2 * x - y, k = x - (y >> 1);
            if (z >= 0 && !(z & 1)) return (A(k - 1) + A(k) + 1) >> 1;
            if (z >= 0) return (A(k - 2) + 2 * A(k - 1) + A(k) + 2) >> 2;
            if (z == -1) return (L(0) + 2 * A(-1) + A(0) + 2) >> 2;
            return (L(y - 1) + 2 * L(y - 2) + (y - 3 >= 0 ? L(y - 3) : A(-1)) + 2) >> 2;
        }
        case 6: {
            int z = 2 * y - x, k = y - (x >> 1);
            if (z >= 0 && !(z & 1)) return ((k - 1 >= 0 ? L(k - 1) : A(-1)) + L(k) + 1) >> 1;
            if (z >= 0) return ((k - 2 >= 0 ? L(k - 2) : A(-1)) + 2 * (k - 1 >= 0 ? L(k - 1) : A(-1)) + L(k) + 2) >> 2;
            if (z == -1) return (L(0) + 2 * A(-1) + A(0) + 2) >> 2;
            return (A(x - 1) + 2 * A(x - 2) + A(x - 3) + 2) >> 2;
        }
        case 7: {
            int i = x + (y >> 1);
            return (y & 1) ? (A(i) + 2 * A(i + 1) + A(i + 2) + 2) >> 2 : (A(i) + A(i + 1) + 1) >> 1;
        }
        default: {
            int z = x + 2 * y, k = y + (x >> 1);
            if (z > 5) return L(3);
            if (z == 5) return (L(2) + 3 * L(3) + 2) >> 2;
            if (z & 1) return (L(k) + 2 * L(k + 1) + L(k + 2) + 2) >> 2;
            return (L(k) + L(k + 1) + 1) >> 1;
        }
    }
}

// Intra4x4 as a table: every predicted sample of every directional mode is one of three forms over the 13 edge samples
// E[0..3] = left column bottom-to-top, E[4] = corner, E[5..12] = above row incl. above-right (clause 8.3.1.2.1-8.3.1.2.9;
// generated and checked against the clause formulas by tools/gen_i4x4_table.py).  Entry [mode][y * 4 + x] =
// i0 | i1 << 8 | i2 << 16 | kind << 24; kind 0: E[i0], 1: (E[i0] + E[i1] + 1) >> 1, 2: (E[i0] + 2 E[i1] + E[i2] + 2) >> 2,
// 3: DC (mode 2, computed from sums).
__device__ const uint32_t gIntra4x4Table[9 * 16] = {
    0x00000005, 0x00000006, 0x00000007, 0x00000008, 0x00000005, 0x00000006, 0x00000007, 0x00000008, 0x00000005, 0x00000006, 0x00000007, 0x00000008, 0x00000005, 0x00000006, 0x00000007, 0x00000008,
    0x00000003, 0x00000003, 0x00000003, 0x00000003, 0x00000002, 0x00000002, 0x00000002, 0x00000002, 0x00000001, 0x00000001, 0x00000001, 0x00000001, 0x00000000, 0x00000000, 0x00000000, 0x00000000,
    0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000, 0x03000000,
    0x02070605, 0x02080706, 0x02090807, 0x020a0908, 0x02080706, 0x02090807, 0x020a0908, 0x020b0a09, 0x02090807, 0x020a0908, 0x020b0a09, 0x020c0b0a, 0x020a0908, 0x020b0a09, 0x020c0b0a, 0x020c0c0b,
    0x02050403, 0x02060504, 0x02070605, 0x02080706, 0x02040302, 0x02050403, 0x02060504, 0x02070605, 0x02030201, 0x02040302, 0x02050403, 0x02060504, 0x02020100, 0x02030201, 0x02040302, 0x02050403,
    0x01000504, 0x01000605, 0x01000706, 0x01000807, 0x02050403, 0x02060504, 0x02070605, 0x02080706, 0x02040302, 0x01000504, 0x01000605, 0x01000706, 0x02030201, 0x02050403, 0x02060504, 0x02070605,
    0x01000304, 0x02050403, 0x02040506, 0x02050607, 0x01000203, 0x02020304, 0x01000304, 0x02050403, 0x01000102, 0x02010203, 0x01000203, 0x02020304, 0x01000001, 0x02000102, 0x01000102, 0x02010203,
    0x01000605, 0x01000706, 0x01000807, 0x01000908, 0x02070605, 0x02080706, 0x02090807, 0x020a0908, 0x01000706, 0x01000807, 0x01000908, 0x01000a09, 0x02080706, 0x02090807, 0x020a0908, 0x020b0a09,
    0x01000203, 0x02010203, 0x01000102, 0x02000102, 0x01000102, 0x02000102, 0x01000001, 0x02000001, 0x01000001, 0x02000001, 0x00000000, 0x00000000, 0x00000000, 0x00000000, 0x00000000, 0x00000000,
};
// the reference decodes the sixteen 4x4 blocks one after the other (intra_prediction.c:701-833); blocks that do not
// depend on each other can share a step: two half-warps, ten steps
__device__ __constant__ int8_t cI4StepA[10] = {0, 1, 2, 3, 6, 7, 10, 11, 14, 15};
__device__ __constant__ int8_t cI4StepB[10] = {-1, -1, 4, 5, 8, 9, 12, 13, -1, -1};

// ---- shared by both passes --------------------------------------------------------------------------
struct MbHead {
    int mbType, qpY, qpC, flags;
    uint32_t mask, coefIndex;
};
__device__ __forceinline__ MbHead loadHead(const b200_mb_rec *rec) {
    const uint4 w = __ldg(reinterpret_cast<const uint4 *>(rec));  // bytes 0..15
    MbHead h;
    h.mbType = w.x & 0xFF; h.qpY = (w.x >> 8) & 0xFF; h.qpC = (w.x >> 16) & 0xFF; h.flags = w.x >> 24;
    h.mask = w.y; h.coefIndex = w.z;
    return h;
}

// residual of the whole macroblock into sm res[24][16] (lane b = block b); see mb_residual in oracle/px_oracle.c
__device__ __forceinline__ void mbResidual(const MbHead &h, const int16_t *coef, int16_t (*res)[16], int lane, uint32_t *errors) {
    const bool i16 = h.mbType >= B200_MB_I_16x16_FIRST;
    const uint32_t mask = h.mask;
    if (lane < 24) {
        const int nDc = ((mask >> 24) & 1) + ((mask >> 25) & 1);
        const bool coded = (mask >> lane) & 1;
        int16_t lev[16];
        if (coded) {
            const uint4 *src = reinterpret_cast<const uint4 *>(coef + (size_t)(nDc + __popc(mask & ((1u << lane) - 1u))) * 16);
            *reinterpret_cast<uint4 *>(lev) = __ldg(src);
            *reinterpret_cast<uint4 *>(lev + 8) = __ldg(src + 1);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) lev[i] = 0;
        }
        bool dcPreset = false;
        int dcVal = 0;
        if (lane < 16) {
            dcPreset = i16;
            if (i16 && (mask & B200_CM_LUMA_DC)) {
                int16_t dl[16];
                const uint4 *src = reinterpret_cast<const uint4 *>(coef);
                *reinterpret_cast<uint4 *>(dl) = __ldg(src);
                *reinterpret_cast<uint4 *>(dl + 8) = __ldg(src + 1);
                dcVal = lumaDcPick(dl, h.qpY, cBlkY[lane] * 4 + cBlkX[lane]);
            }
        } else {
            dcPreset = true;
            if (mask & B200_CM_CHROMA_DC) {
                int16_t dl[4];
                const uint2 *src = reinterpret_cast<const uint2 *>(coef + (size_t)((mask >> 24) & 1) * 16 + ((lane - 16) >> 2) * 4);
                *reinterpret_cast<uint2 *>(dl) = __ldg(src);
                dcVal = chromaDcPick(dl, h.qpC, lane & 3);
            }
        }
        int out[16];
        bool bad = false;
        if (coded || (dcPreset && dcVal != 0)) {
            bad = idctBlock(lev, lane < 16 ? h.qpY : h.qpC, dcPreset, dcVal, out);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) out[i] = 0;
        }
        if (bad) atomicAdd(errors, 1u);
        int16_t *dst = res[lane];
#pragma unroll
        for (int i = 0; i < 16; i += 2)
            *reinterpret_cast<uint32_t *>(dst + i) = (uint32_t)(uint16_t)out[i] | ((uint32_t)(uint16_t)out[i + 1] << 16);
    }
    __syncwarp();
}

// this lane's residuals: luma row r8 cols c8..c8+7, chroma plane cp row cr cols cc..cc+3
__device__ __forceinline__ void laneResidual(const int16_t (*res)[16], int lane, int *resY, int *resC) {
    const int r8 = lane >> 1, cp = lane >> 4, cr = (lane >> 1) & 7;
    const int by = r8 >> 2, ry = r8 & 3, bx = (lane & 1) * 2;
    const int16_t *ra = res[cRasterToBlk[by * 4 + bx]] + ry * 4;
    const int16_t *rb = res[cRasterToBlk[by * 4 + bx + 1]] + ry * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) { resY[i] = ra[i]; resY[4 + i] = rb[i]; }
    const int16_t *rc = res[16 + cp * 4 + (cr >> 2) * 2 + (lane & 1)] + (cr & 3) * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) resC[i] = rc[i];
}

// 8 consecutive bytes from an arbitrarily aligned shared-memory address
__device__ __forceinline__ uint2 lds8(const uint8_t *p) {
    const uint32_t a = smemAddr(p), sh = (a & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - (a & 3u));
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
__device__ __forceinline__ uint32_t lds4(const uint8_t *p) {
    const uint32_t a = smemAddr(p), sh = (a & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - (a & 3u));
    return __funnelshift_r(w[0], w[1], sh);
}


// ---- 8-wide luma prediction for one lane (row y, columns x0..x0+7 of a 16x16 partition) --------------------
constexpr int kTapsLo = 0x1414FB01;  // bytes (1, -5, 20, 20)
constexpr int kTapsHi = 0x000001FB;  // bytes (-5, 1, 0, 0)

// unclipped horizontal 6-tap sums for 8 outputs; rowp points at sample x0-2 of the row (13 samples are read)
__device__ __forceinline__ void hrow8(const uint8_t *rowp, int *hs) {
    const uint32_t a = smemAddr(rowp), sh0 = (a & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rowp - (a & 3u));
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    // aligned view: byte k of the row = byte (k + (a&3)) of (w0,w1,w2,w3)
    const uint32_t v0 = __funnelshift_r(w0, w1, sh0), v1 = __funnelshift_r(w1, w2, sh0), v2 = __funnelshift_r(w2, w3, sh0), v3 = w3 >> sh0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t lo = (k & 3) == 0 ? (k == 0 ? v0 : v1) : __funnelshift_r(k < 4 ? v0 : v1, k < 4 ? v1 : v2, 8 * (k & 3));
        const uint32_t hi = (k & 3) == 0 ? (k == 0 ? v1 : v2) : __funnelshift_r(k < 4 ? v1 : v2, k < 4 ? v2 : v3, 8 * (k & 3));
        hs[k] = dp4aUS(hi, kTapsHi, dp4aUS(lo, kTapsLo, 0));
    }
}
// unclipped vertical 6-tap sums for 8 outputs; colp points at sample (x0, y-2); rows are kLumaBoxW apart
__device__ __forceinline__ void vcol8(const uint8_t *colp, int *vs) {
    uint2 r[6];
#pragma unroll
    for (int t = 0; t < 6; t++) r[t] = lds8(colp + t * kLumaBoxW);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int sel = k & 3;
        const uint32_t a0 = k < 4 ? r[0].x : r[0].y, a1 = k < 4 ? r[1].x : r[1].y, a2 = k < 4 ? r[2].x : r[2].y;
        const uint32_t a3 = k < 4 ? r[3].x : r[3].y, a4 = k < 4 ? r[4].x : r[4].y, a5 = k < 4 ? r[5].x : r[5].y;
        // gather byte `sel` of four rows into one word
        const uint32_t t01 = __byte_perm(a0, a1, sel | ((4 + sel) << 4));
        const uint32_t t23 = __byte_perm(a2, a3, sel | ((4 + sel) << 4));
        const uint32_t lo = __byte_perm(t01, t23, 0x5410);
        const uint32_t hi = __byte_perm(a4, a5, sel | ((4 + sel) << 4)) & 0xFFFFu;
        vs[k] = dp4aUS(hi, kTapsHi, dp4aUS(lo, kTapsLo, 0));
    }
}
// four values -> four bytes with unsigned saturation, element 0 in the low byte (I2IP.U8.S32.SAT, two instructions)
// clip255((v + 16) >> 5) / clip255((v + 512) >> 10) of eight sums, packed
__device__ __forceinline__ uint2 pack8shift(const int *v, int rnd, int sh) {
    return make_uint2(pack4sat((v[0] + rnd) >> sh, (v[1] + rnd) >> sh, (v[2] + rnd) >> sh, (v[3] + rnd) >> sh),
                      pack4sat((v[4] + rnd) >> sh, (v[5] + rnd) >> sh, (v[6] + rnd) >> sh, (v[7] + rnd) >> sh));
}
__device__ __forceinline__ uint2 avg8(uint2 a, uint2 b) { return make_uint2(__vavgu4(a.x, b.x), __vavgu4(a.y, b.y)); }

// clause 8.4.2.2.1 for 8 horizontally adjacent samples: `win` points at window sample (xInt-2, yInt-2) (see W_);
// (x0, y) = position of the first sample inside the partition; the same arithmetic as lumaQpel, 8 at a time, on packed
// bytes: the rounded averages (a + b + 1) >> 1 are per-byte averages of clipped values
__device__ __forceinline__ uint2 lumaQpel8(const uint8_t *win, int x0, int y, int xf, int yf) {
    const bool jfam = (xf == 2 || yf == 2) && xf != 0 && yf != 0;
    if (!jfam) {
        const bool useH = xf != 0, useV = yf != 0;
        uint2 b = make_uint2(0, 0), h = make_uint2(0, 0);
        if (useH) {
            int t[8];
            hrow8(win + (y + 2 + (yf == 3 ? 1 : 0)) * kLumaBoxW + x0, t);
            b = pack8shift(t, 16, 5);
        }
        if (useV) {
            int t[8];
            vcol8(win + y * kLumaBoxW + x0 + 2 + (xf == 3 ? 1 : 0), t);
            h = pack8shift(t, 16, 5);
        }
        if (useH && useV) return avg8(b, h);
        if (useH) return xf == 2 ? b : avg8(b, lds8(win + (y + 2) * kLumaBoxW + x0 + 2 + (xf >> 1)));
        return yf == 2 ? h : avg8(h, lds8(win + (y + 2 + (yf >> 1)) * kLumaBoxW + x0 + 2));
    }
    int acc[8], bsel[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = 0; bsel[k] = 0; }
    const int brow = 2 + (yf == 3 ? 1 : 0);
#pragma unroll 1
    for (int t = 0; t < 6; t++) {
        int hs[8];
        hrow8(win + (y + t) * kLumaBoxW + x0, hs);
        const int c = (t == 0 || t == 5) ? 1 : (t == 1 || t == 4) ? -5 : 20;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            acc[k] += c * hs[k];
            if (t == brow) bsel[k] = hs[k];
        }
    }
    const uint2 j = pack8shift(acc, 512, 10);
    if (xf == 2 && yf == 2) return j;
    if (xf == 2) return avg8(j, pack8shift(bsel, 16, 5));
    int hv[8];
    vcol8(win + y * kLumaBoxW + x0 + 2 + (xf == 3 ? 1 : 0), hv);
    return avg8(j, pack8shift(hv, 16, 5));
}

// =====================================================================================================
// pass A: inter-predicted (and I_PCM) macroblocks.  They read only finished reference frames, so there
// is no ordering between them: plain grid, blockIdx.y = stream, a warp owns kChunkA consecutive entries
// of the stream's raster-ordered list and prefetches the next entry's reference window by TMA while it
// works on the current one.
// =====================================================================================================
struct InterInfo {
    uint32_t mb;
    MbHead h;
    bool single;     // one 16x16 partition (P_Skip / P_L0_16x16): window prefetched
    int mvx, mvy;
    int ox, cox;     // clamped window origins (bordered-plane coordinates), luma / chroma
};

__device__ __forceinline__ void issueWindow(InterWarpSmem &sm, int buf, const PoolGeom &g, const CUtensorMap *lumaMap, const CUtensorMap *chromaMap,
                                            int xInt, int yInt, int cxInt, int cyInt, uint32_t refFrame, int lane, int *oxOut, int *coxOut) {
    // clamp the window origin into the bordered plane: a window wholly outside the picture on an axis equals the
    // window at the clamped origin because the border is a replication; the box starts 16-byte aligned
    const int ox = clip3(-kPadY, g.W + kPadY - kLumaWin, xInt - 2) + kPadY;
    const int oy = clip3(-kPadY, g.H + kPadY - kLumaWin, yInt - 2) + kPadY;
    const int cox = clip3(-kPadC, g.W / 2 + kPadC - kChromaWin, cxInt) + kPadC;
    const int coy = clip3(-kPadC, g.H / 2 + kPadC - kChromaWin, cyInt) + kPadC;
    if (lane == 0) {
        fenceProxyAsync();
        mbarExpectTx(&sm.mbar[buf], kLumaBoxW * kLumaBoxH + 2 * kChromaBoxW * kChromaBoxH);
        tmaLoad3d(sm.lumaWin[buf], lumaMap, ox & ~15, oy, (int)refFrame, &sm.mbar[buf]);
        tmaLoad4d(sm.chromaWin[buf], chromaMap, cox & ~15, coy, 0, (int)refFrame, &sm.mbar[buf]);
    }
    *oxOut = ox;
    *coxOut = cox;
}

__global__ void __launch_bounds__(kReconWarps * 32, 3)
reconInterKernel(const ReconParams p, const __grid_constant__ CUtensorMap lumaMap, const __grid_constant__ CUtensorMap chromaMap) {
    extern __shared__ __align__(128) uint8_t interSmemRaw[];   // kReconWarps x InterWarpSmem (more than the 48 KB static limit)
    InterWarpSmem *smemAll = reinterpret_cast<InterWarpSmem *>(interSmemRaw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    InterWarpSmem &sm = smemAll[warp];
    if (lane == 0) {
        mbarInit(&sm.mbar[0], 1);
        mbarInit(&sm.mbar[1], 1);
        mbarInit(&sm.mbar[2], 1);
        fenceMbarInit();
    }
    __syncwarp();
    uint32_t phaseBits = 0;   // bit b = phase parity of window buffer b
    // persistent CTAs striding over virtual CTAs: v -> (stream, chunk of the stream's pass-A list); consecutive
    // virtual CTAs belong to the same stream, so neighbouring macroblocks are in flight together (L2 locality)
    for (uint32_t v = blockIdx.x; v < p.virtualCtasA; v += gridDim.x) {
    const uint32_t s = v / p.chunksA, chunk = v - s * p.chunksA;
    const StreamJob job = p.jobs[s];
    const uint32_t l0 = (chunk * kReconWarps + warp) * p.chunkA;   // index into the stream's pass-A entries that are not plain copies
    if (l0 >= job.nA) continue;
    const int n = min(p.chunkA, job.nA - l0);
    const uint32_t e0 = 2u * job.nR + job.nC + l0;               // the plain copies went to reconCopyKernel
    const uint32_t frameBase = s * (uint32_t)g.numSlots;
    uint8_t *cur = framePtr(p.pool, g, frameBase + job.curSlot);
    const int r8 = lane >> 1, c8 = (lane & 1) * 8;
    const int cp = lane >> 4, cr = (lane >> 1) & 7, cc = (lane & 1) * 4;

    // the chunk's records are fetched by the first n lanes in parallel (one dependent-load chain per chunk, not per MB).
    // (Barrier first: a lane may still be reading the previous chunk's records -- an I_PCM macroblock ends its turn without one.)
    __syncwarp();
    if (lane < n) {
        const uint32_t mb = __ldg(job.order + e0 + lane);
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(job.recs + mb);
        const uint4 hw = __ldg(reinterpret_cast<const uint4 *>(rw));
        const uint32_t refSlots = __ldg(rw + 4), mv0 = __ldg(rw + 8);
        uint32_t *m = sm.meta[lane];
        m[0] = hw.x; m[1] = hw.y; m[2] = hw.z; m[3] = hw.w; m[4] = refSlots; m[5] = mv0; m[6] = mb;
    }
    __syncwarp();
    auto prepare = [&](int i, int buf) -> InterInfo {
        InterInfo it;
        const uint32_t *m = sm.meta[i];
        it.mb = m[6];
        it.h.mbType = m[0] & 0xFF; it.h.qpY = (m[0] >> 8) & 0xFF; it.h.qpC = (m[0] >> 16) & 0xFF; it.h.flags = m[0] >> 24;
        it.h.mask = m[1]; it.h.coefIndex = m[2];
        it.single = it.h.mbType <= B200_MB_P_16x16;
        it.mvx = it.mvy = 0; it.ox = it.cox = 0;
        if (it.single) {
            const uint32_t mvv = m[5], refSlots = m[4];
            it.mvx = (int)(int16_t)(mvv & 0xFFFF); it.mvy = (int)(int16_t)(mvv >> 16);
            const int mby = mbRowOf(it.mb, g), mbx = (int)(it.mb - (uint32_t)mby * g.widthMbs);
            issueWindow(sm, buf, g, &lumaMap, &chromaMap, mbx * 16 + (it.mvx >> 2), mby * 16 + (it.mvy >> 2),
                        mbx * 8 + (it.mvx >> 3), mby * 8 + (it.mvy >> 3), frameBase + (refSlots & 0xFF), lane, &it.ox, &it.cox);
        }
        return it;
    };

    InterInfo nxt = prepare(0, 0);
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const int buf = i & 1;
        const InterInfo it = nxt;
        if (i + 1 < n) nxt = prepare(i + 1, buf ^ 1);   // the other buffer's previous user finished before this point
        const MbHead &h = it.h;
        const uint32_t mb = it.mb;
        const int mby = mbRowOf(mb, g), mbx = (int)(mb - (uint32_t)mby * g.widthMbs);
        const b200_mb_rec *rec = job.recs + mb;
        const int16_t *coef = job.coefs + (size_t)h.coefIndex * 16;
        uint8_t *dstY = lumaAt(cur, g, mbx * 16 + c8, mby * 16 + r8);
        uint8_t *dstC = chromaAt(cur, g, cp, mbx * 8 + cc, mby * 8 + cr);

        if (h.mbType == B200_MB_I_PCM) {
            // h264bsdWriteMacroblock (image.c:81-144): 384 raw bytes
            const uint8_t *src = reinterpret_cast<const uint8_t *>(coef);
            *reinterpret_cast<uint2 *>(dstY) = __ldg(reinterpret_cast<const uint2 *>(src + r8 * 16 + c8));
            *reinterpret_cast<uint32_t *>(dstC) = __ldg(reinterpret_cast<const uint32_t *>(src + 256 + cp * 64 + cr * 8 + cc));
            continue;
        }
        if (h.mask) mbResidual(h, coef, sm.res, lane, p.errors);   // into shared memory; read back after the prediction
        uint2 pv = make_uint2(0, 0);   // this lane's 8 luma prediction samples
        uint32_t pc = 0;               // and 4 chroma prediction samples
        // Partitions that are at least 8 wide (16x16, 16x8, 8x16, 8x8 sub-macroblocks) share one code path: every lane's
        // 8-sample luma span and 4-sample chroma span lie inside ONE partition, so the lane only has to pick that
        // partition's window, vector and origin.  Two partitions are staged at a time (windows `buf` and 2).
        uint32_t subTypes = 0;
        if (h.mbType >= B200_MB_P_8x8) subTypes = (__ldg(reinterpret_cast<const uint32_t *>(rec) + 3) >> 24) & 0xFF;
        const bool wide = h.mbType <= B200_MB_P_8x16 || subTypes == 0;
        if (wide) {
            const int rounds = it.single ? 1 : (h.mbType >= B200_MB_P_8x8 ? 2 : 1);
#pragma unroll 1
            for (int rd = 0; rd < rounds; rd++) {
                // per lane: window buffer, window origins, vector, position of the lane's spans inside the partition
                int bufL = buf, oxL = it.ox, mvxL = it.mvx, mvyL = it.mvy, lx = c8, ly = r8;
                int bufC = buf, coxC = it.cox, mvxC = it.mvx, mvyC = it.mvy, lcx = cc, lcy = cr;
                bool actL = true, actC = true;
                if (it.single) {
                    mbarWait(&sm.mbar[buf], (phaseBits >> buf) & 1u);
                    phaseBits ^= 1u << buf;
                } else {
                    // partitions A and B of this round (inter_prediction.c:361-482): first block, origin in pels
                    int blkA, blkB, pxB, pyA, pyB;
                    if (h.mbType == B200_MB_P_16x8) { blkA = 0; blkB = 8; pxB = 0; pyA = 0; pyB = 8; }
                    else if (h.mbType == B200_MB_P_8x16) { blkA = 0; blkB = 4; pxB = 8; pyA = 0; pyB = 0; }
                    else { blkA = 8 * rd; blkB = 8 * rd + 4; pxB = 8; pyA = pyB = 8 * rd; }
                    const uint32_t *rw = reinterpret_cast<const uint32_t *>(rec);
                    const uint32_t refSlots = __ldg(rw + 4), mvA = __ldg(rw + 8 + blkA), mvB = __ldg(rw + 8 + blkB);
                    const int ax = (int)(int16_t)(mvA & 0xFFFF), ay = (int)(int16_t)(mvA >> 16);
                    const int bx = (int)(int16_t)(mvB & 0xFFFF), by = (int)(int16_t)(mvB >> 16);
                    int oxA, coxA, oxB, coxB;
                    issueWindow(sm, buf, g, &lumaMap, &chromaMap, mbx * 16 + (ax >> 2), mby * 16 + pyA + (ay >> 2),
                                mbx * 8 + (ax >> 3), mby * 8 + (pyA >> 1) + (ay >> 3), frameBase + ((refSlots >> (8 * (blkA >> 2))) & 0xFF), lane, &oxA, &coxA);
                    issueWindow(sm, 2, g, &lumaMap, &chromaMap, mbx * 16 + pxB + (bx >> 2), mby * 16 + pyB + (by >> 2),
                                mbx * 8 + (pxB >> 1) + (bx >> 3), mby * 8 + (pyB >> 1) + (by >> 3), frameBase + ((refSlots >> (8 * (blkB >> 2))) & 0xFF), lane, &oxB, &coxB);
                    const bool inBL = h.mbType == B200_MB_P_16x8 ? r8 >= 8 : c8 == 8;
                    const bool inBC = h.mbType == B200_MB_P_16x8 ? cr >= 4 : cc == 4;
                    if (h.mbType >= B200_MB_P_8x8) { actL = (r8 >> 3) == rd; actC = (cr >> 2) == rd; }
                    bufL = inBL ? 2 : buf; oxL = inBL ? oxB : oxA; mvxL = inBL ? bx : ax; mvyL = inBL ? by : ay;
                    lx = c8 - (inBL ? pxB : 0); ly = r8 - (inBL ? pyB : pyA);
                    bufC = inBC ? 2 : buf; coxC = inBC ? coxB : coxA; mvxC = inBC ? bx : ax; mvyC = inBC ? by : ay;
                    lcx = cc - (inBC ? (pxB >> 1) : 0); lcy = cr - ((inBC ? pyB : pyA) >> 1);
                    mbarWait(&sm.mbar[buf], (phaseBits >> buf) & 1u);
                    phaseBits ^= 1u << buf;
                    mbarWait(&sm.mbar[2], (phaseBits >> 2) & 1u);
                    phaseBits ^= 4u;
                }
                if (actL) {
                    const int xf = mvxL & 3, yf = mvyL & 3;
                    const uint8_t *win = sm.lumaWin[bufL] + (oxL & 15);
                    if ((xf | yf) == 0) pv = lds8(win + (ly + 2) * kLumaBoxW + lx + 2);   // h264bsdFillBlock copy (reconstruct.c:1852)
                    else pv = lumaQpel8(win, lx, ly, xf, yf);
                }
                if (actC) {
                    const int cxf = mvxC & 7, cyf = mvyC & 7;
                    const uint8_t *cw = sm.chromaWin[bufC] + cp * (kChromaBoxW * kChromaBoxH) + lcy * kChromaBoxW + lcx + (coxC & 15);
                    if ((cxf | cyf) == 0) {
                        pc = lds4(cw);
                    } else {
                        // PredictChroma (reconstruct.c:415-475)
                        const uint2 ra = lds8(cw), rb = lds8(cw + kChromaBoxW);
                        const int w00 = (8 - cxf) * (8 - cyf), w01 = cxf * (8 - cyf), w10 = (8 - cxf) * cyf, w11 = cxf * cyf;
                        pc = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int A = (ra.x >> (8 * k)) & 0xFF, B = k < 3 ? (ra.x >> (8 * k + 8)) & 0xFF : ra.y & 0xFF;
                            const int Cc = (rb.x >> (8 * k)) & 0xFF, D = k < 3 ? (rb.x >> (8 * k + 8)) & 0xFF : rb.y & 0xFF;
                            pc |= (uint32_t)((w00 * A + w01 * B + w10 * Cc + w11 * D + 32) >> 6) << (8 * k);
                        }
                    }
                }
                if (rd + 1 < rounds) __syncwarp();   // the windows are overwritten by the next round's loads
            }
        } else {
            // sub-macroblocks with 8x4 / 4x8 / 4x4 partitions: one window per partition, sample by sample
            const uint32_t refSlots = __ldg(reinterpret_cast<const uint32_t *>(rec) + 4);
            const uint32_t *mvw = reinterpret_cast<const uint32_t *>(rec) + 8;
#pragma unroll 1
            for (int pi = 0; pi < 16; pi++) {
                int pw, ph;
                const int blk = pi;
                const int sub = (subTypes >> (2 * (pi >> 2))) & 3, j = pi & 3;
                if (sub == 0) { if (j) continue; pw = 8; ph = 8; }
                else if (sub == 1) { if (j & 1) continue; pw = 8; ph = 4; }
                else if (sub == 2) { if (j & 2) continue; pw = 4; ph = 8; }
                else { pw = 4; ph = 4; }
                const int px = cBlkX[blk] * 4, py = cBlkY[blk] * 4;
                const uint32_t mvv = __ldg(mvw + blk);
                const int mvx = (int)(int16_t)(mvv & 0xFFFF), mvy = (int)(int16_t)(mvv >> 16);
                const uint32_t refFrame = frameBase + ((refSlots >> (8 * (blk >> 2))) & 0xFF);
                int ox, cox;
                issueWindow(sm, buf, g, &lumaMap, &chromaMap, mbx * 16 + px + (mvx >> 2), mby * 16 + py + (mvy >> 2),
                            ((mbx * 16 + px) >> 1) + (mvx >> 3), ((mby * 16 + py) >> 1) + (mvy >> 3), refFrame, lane, &ox, &cox);
                mbarWait(&sm.mbar[buf], (phaseBits >> buf) & 1u);
                phaseBits ^= 1u << buf;
                const int xf = mvx & 3, yf = mvy & 3;
                const int lw = 31 - __clz(pw);
#pragma unroll 1
                for (int q = lane; q < pw * ph; q += 32) {
                    const int x = q & (pw - 1), y = q >> lw;
                    sm.pred[(py + y) * 16 + px + x] = (uint8_t)lumaQpel(sm.lumaWin[buf] + (ox & 15), x, y, xf, yf);
                }
                const int cw = pw >> 1, chh = ph >> 1, ncp = cw * chh, lcw = lw - 1;
                const int cxf = mvx & 7, cyf = mvy & 7;
#pragma unroll 1
                for (int q = lane; q < 2 * ncp; q += 32) {
                    const int pl = q >= ncp, qq = q - pl * ncp;
                    const int x = qq & (cw - 1), y = qq >> lcw;
                    const uint8_t *wp = sm.chromaWin[buf] + pl * (kChromaBoxW * kChromaBoxH) + y * kChromaBoxW + x + (cox & 15);
                    const int A = wp[0], B = wp[1], Cc = wp[kChromaBoxW], D = wp[kChromaBoxW + 1];
                    sm.pred[256 + pl * 64 + ((py >> 1) + y) * 8 + (px >> 1) + x] =
                        (uint8_t)(((8 - cxf) * (8 - cyf) * A + cxf * (8 - cyf) * B + (8 - cxf) * cyf * Cc + cxf * cyf * D + 32) >> 6);
                }
                __syncwarp();
            }
            pv = *reinterpret_cast<const uint2 *>(sm.pred + r8 * 16 + c8);
            pc = *reinterpret_cast<const uint32_t *>(sm.pred + 256 + cp * 64 + cr * 8 + cc);
        }
        // add residual + clip + store (h264bsdWriteOutputBlocks, image.c:172-344)
        if (h.mask) {
            int resY[8], resC[4];
            laneResidual(sm.res, lane, resY, resC);
            auto px = [](uint32_t w, int k) { return (int)((w >> (8 * k)) & 0xFF); };
            pv = make_uint2(pack4sat(px(pv.x, 0) + resY[0], px(pv.x, 1) + resY[1], px(pv.x, 2) + resY[2], px(pv.x, 3) + resY[3]),
                            pack4sat(px(pv.y, 0) + resY[4], px(pv.y, 1) + resY[5], px(pv.y, 2) + resY[6], px(pv.y, 3) + resY[7]));
            pc = pack4sat(px(pc, 0) + resC[0], px(pc, 1) + resC[1], px(pc, 2) + resC[2], px(pc, 3) + resC[3]);
        }
        *reinterpret_cast<uint2 *>(dstY) = pv;
        *reinterpret_cast<uint32_t *>(dstC) = pc;
        __syncwarp();
    }
    }  // virtual CTA loop
}

// =====================================================================================================
// pass B: intra-predicted macroblocks.  They read the unfiltered current picture, so an intra macroblock
// must come after its intra neighbours (the others were written by the copy pass and pass A).  CTAs take tickets; a
// ticket is kReconWarps warp tasks; warp task w is chunk w / nStreams of stream w % nStreams of the stream's
// wavefront-ordered list.  A warp works through its chunk in list order, so dependencies inside a chunk cost nothing,
// the warps of a CTA belong to different streams (they never wait for each other), and a warp only ever waits for
// chunks whose CTA took an earlier ticket.
// =====================================================================================================
__global__ void __launch_bounds__(kReconWarps * 32, 5) reconIntraKernel(const ReconParams p) {
    __shared__ IntraWarpSmem smemAll[kReconWarps];
    __shared__ uint32_t sI4Table[9 * 16];
    __shared__ uint32_t sTicket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    if (threadIdx.x == 0) sTicket = atomicAdd(p.ticket, 1u);
    if (threadIdx.x < 9 * 16) sI4Table[threadIdx.x] = gIntra4x4Table[threadIdx.x];
    __syncthreads();
    const uint32_t t = sTicket * kReconWarps + warp;
    const uint32_t chunk = t / (uint32_t)g.nStreams, s = t - chunk * (uint32_t)g.nStreams;
    if (chunk >= p.chunksB) return;
    const StreamJob job = p.jobs[s];
    const uint32_t e0 = chunk * p.chunkB;
    if (e0 >= job.nB) return;
    const int n = min(p.chunkB, job.nB - e0);
    IntraWarpSmem &sm = smemAll[warp];
    uint32_t *doneS = p.done + (size_t)s * g.nMbs;
    uint8_t *cur = framePtr(p.pool, g, s * (uint32_t)g.numSlots + job.curSlot);
    const int r8 = lane >> 1, c8 = (lane & 1) * 8;
    const int cp = lane >> 4, cr = (lane >> 1) & 7, cc = (lane & 1) * 4;

    // lane j < n fetches entry j's address and record head: one chain of dependent loads per chunk
    uint32_t mMb = 0, mMisc = 0;
    uint4 mHead = make_uint4(0, 0, 0, 0);
    if (lane < n) {
        mMb = __ldg(job.order + (2u * job.nR + job.nC + job.nA) + e0 + lane);
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(job.recs + mMb);
        mHead = __ldg(reinterpret_cast<const uint4 *>(rw));
        mMisc = (__ldg(rw + 5) & 0xFF) | ((__ldg(rw + 7) & 0xFF) << 8);
    }
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const uint32_t mb = __shfl_sync(0xffffffffu, mMb, i);
        const int mby = (int)(mb / (uint32_t)g.widthMbs), mbx = (int)(mb - (uint32_t)mby * g.widthMbs);
        const b200_mb_rec *rec = job.recs + mb;
        MbHead h;
        {
            const uint32_t hx = __shfl_sync(0xffffffffu, mHead.x, i);
            h.mbType = hx & 0xFF; h.qpY = (hx >> 8) & 0xFF; h.qpC = (hx >> 16) & 0xFF; h.flags = hx >> 24;
            h.mask = __shfl_sync(0xffffffffu, mHead.y, i);
            h.coefIndex = __shfl_sync(0xffffffffu, mHead.z, i);
        }
        const uint32_t misc = __shfl_sync(0xffffffffu, mMisc, i);   // intraChromaMode | waitMask << 8
        const int16_t *coef = job.coefs + (size_t)h.coefIndex * 16;
        uint8_t *dstY = lumaAt(cur, g, mbx * 16 + c8, mby * 16 + r8);
        uint8_t *dstC = chromaAt(cur, g, cp, mbx * 8 + cc, mby * 8 + cr);
        int resY[8], resC[4];
        if (h.mask) {
            mbResidual(h, coef, sm.res, lane, p.errors);
            laneResidual(sm.res, lane, resY, resC);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) resY[k] = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) resC[k] = 0;
        }
        const int flags = h.flags;
        const bool avA = flags & B200_MBF_AVAIL_A, avB = flags & B200_MBF_AVAIL_B;
        const bool avC = flags & B200_MBF_AVAIL_C, avD = flags & B200_MBF_AVAIL_D;
        // wait for the intra neighbours this macroblock reads (record byte 28: waitMask)
        const int waitMask = (misc >> 8) & 0xFF;
        {
            // a neighbour that is an earlier entry of this warp's own chunk needs no flag: program order + the warp barrier
            const int nmb = lane == 0 ? (int)mb - 1 : lane == 1 ? (int)mb - g.widthMbs : lane == 2 ? (int)mb - g.widthMbs + 1 : (int)mb - g.widthMbs - 1;
            bool mine = false;
#pragma unroll
            for (int j = 0; j < kChunkB - 1; j++) {
                const uint32_t mj = __shfl_sync(0xffffffffu, mMb, j);
                mine |= j < i && (int)mj == nmb;
            }
            if (lane < 4 && ((waitMask >> lane) & 1) && !mine) waitFlag(doneS + nmb, p.serial);
        }
        __syncwarp();
        // neighbouring pels (h264bsdGetNeighbourPels :545-614), straight from L2
        if (lane < 21) {
            const bool ok = lane == 0 ? avD : lane <= 16 ? avB : avC;
            sm.itY[0][lane] = ok ? __ldcg(lumaAt(cur, g, mbx * 16 - 1 + lane, mby * 16 - 1)) : 128;
        }
        if (lane < 16) sm.itY[1 + lane][0] = avA ? __ldcg(lumaAt(cur, g, mbx * 16 - 1, mby * 16 + lane)) : 128;
        if (lane < 18) {
            const int pl = lane >= 9, k = lane - pl * 9;
            const bool ok = k == 0 ? avD : avB;
            sm.itC[pl][0][k] = ok ? __ldcg(chromaAt(cur, g, pl, mbx * 8 - 1 + k, mby * 8 - 1)) : 128;
        }
        if (lane < 16) {
            const int pl = lane >> 3, k = lane & 7;
            sm.itC[pl][1 + k][0] = avA ? __ldcg(chromaAt(cur, g, pl, mbx * 8 - 1, mby * 8 + k)) : 128;
        }
        __syncwarp();

        if (h.mbType == B200_MB_I_4x4) {
            // h264bsdIntra4x4Prediction (:701-833): half-warp = block, lane = sample; edge samples straight from the tile
            const uint32_t *modew = reinterpret_cast<const uint32_t *>(rec) + 8;
            const int x = lane & 3, y = (lane >> 2) & 3;
#pragma unroll 1
            for (int st = 0; st < 10; st++) {
                const int b = lane < 16 ? cI4StepA[st] : cI4StepB[st];
                if (b >= 0) {
                    const int bx = cBlkX[b], by = cBlkY[b];
                    const int mode = (__ldg(modew + (b >> 2)) >> (8 * (b & 3))) & 0xFF;
                    const bool bA = bx ? true : avA, bB = by ? true : avB;
                    bool bC;   // above-right block available: decoded earlier (or the neighbouring macroblock's)
                    if (by == 0) bC = (bx == 3) ? avC : avB;
                    else if (bx == 3) bC = false;
                    else bC = cRasterToBlk[(by - 1) * 4 + bx + 1] < b;
                    const uint8_t *corner = &sm.itY[by * 4][bx * 4];   // E[4]; above row to its right, left column below it
                    auto E = [&](int i) -> int {
                        if (i <= 3) return corner[(4 - i) * 24];
                        int k = i - 4;                         // 0 = corner, 1..8 = above row
                        if (k > 4 && !bC) k = 4;
                        return corner[k];
                    };
                    int v;
                    if (mode == 2) {
                        const int sa = corner[1] + corner[2] + corner[3] + corner[4];
                        const int sl = corner[24] + corner[48] + corner[72] + corner[96];
                        v = (bA && bB) ? (sa + sl + 4) >> 3 : bA ? (sl + 2) >> 2 : bB ? (sa + 2) >> 2 : 128;
                    } else {
                        const uint32_t d = sI4Table[mode * 16 + y * 4 + x];
                        const int kind = d >> 24, e0 = E(d & 0xFF);
                        if (kind == 0) v = e0;
                        else if (kind == 1) v = (e0 + E((d >> 8) & 0xFF) + 1) >> 1;
                        else v = (e0 + 2 * E((d >> 8) & 0xFF) + E((d >> 16) & 0xFF) + 2) >> 2;
                    }
                    if (h.mask) v = clip255(v + sm.res[b][y * 4 + x]);
                    // the sample lies inside the block, every edge sample outside it, and the two blocks of a step do not
                    // touch each other's edges: no barrier between the reads above and this write
                    sm.itY[by * 4 + 1 + y][bx * 4 + 1 + x] = (uint8_t)v;
                }
                __syncwarp();
            }
            *reinterpret_cast<uint2 *>(dstY) = lds8(&sm.itY[1 + r8][1 + c8]);
        } else {
            // h264bsdIntra16x16Prediction (:627-687)
            const int mode = (h.mbType - B200_MB_I_16x16_FIRST) & 3;
            int pv[8];
            if (mode == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) pv[k] = sm.itY[0][1 + c8 + k];
            } else if (mode == 1) {
#pragma unroll
                for (int k = 0; k < 8; k++) pv[k] = sm.itY[1 + r8][0];
            } else if (mode == 2) {
                int sa = 0, sl = 0;
                for (int k = 0; k < 16; k++) { sa += sm.itY[0][1 + k]; sl += sm.itY[1 + k][0]; }
                const int v = (avA && avB) ? (sa + sl + 16) >> 5 : avA ? (sl + 8) >> 4 : avB ? (sa + 8) >> 4 : 128;
#pragma unroll
                for (int k = 0; k < 8; k++) pv[k] = v;
            } else {
                int Hh = 0, V = 0;
                for (int k = 0; k < 8; k++) {
                    Hh += (k + 1) * ((int)sm.itY[0][1 + 8 + k] - (int)sm.itY[0][1 + 6 - k]);
                    V += (k + 1) * ((int)sm.itY[1 + 8 + k][0] - (int)sm.itY[1 + 6 - k][0]);
                }
                const int a = 16 * ((int)sm.itY[16][0] + (int)sm.itY[0][16]);
                const int bb = (5 * Hh + 32) >> 6, cc2 = (5 * V + 32) >> 6;
#pragma unroll
                for (int k = 0; k < 8; k++) pv[k] = clip255((a + bb * (c8 + k - 7) + cc2 * (r8 - 7) + 16) >> 5);
            }
            uint32_t o0 = 0, o1 = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                o0 |= (uint32_t)clip255(pv[k] + resY[k]) << (8 * k);
                o1 |= (uint32_t)clip255(pv[4 + k] + resY[4 + k]) << (8 * k);
            }
            *reinterpret_cast<uint2 *>(dstY) = make_uint2(o0, o1);
        }
        // h264bsdIntraChromaPrediction (:845-915)
        {
            const int cmode = misc & 0xFF;
            const uint8_t(*tc)[12] = sm.itC[cp];
            int pv[4];
            if (cmode == 0) {
                const int bxq = lane & 1, byq = cr >> 2;
                int sa = 0, sl = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) { sa += tc[0][1 + bxq * 4 + k]; sl += tc[1 + byq * 4 + k][0]; }
                int v;
                if (bxq == byq) v = (avA && avB) ? (sa + sl + 4) >> 3 : avB ? (sa + 2) >> 2 : avA ? (sl + 2) >> 2 : 128;
                else if (bxq == 1) v = avB ? (sa + 2) >> 2 : avA ? (sl + 2) >> 2 : 128;
                else v = avA ? (sl + 2) >> 2 : avB ? (sa + 2) >> 2 : 128;
#pragma unroll
                for (int k = 0; k < 4; k++) pv[k] = v;
            } else if (cmode == 1) {
#pragma unroll
                for (int k = 0; k < 4; k++) pv[k] = tc[1 + cr][0];
            } else if (cmode == 2) {
#pragma unroll
                for (int k = 0; k < 4; k++) pv[k] = tc[0][1 + cc + k];
            } else {
                int Hh = 0, V = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    Hh += (k + 1) * ((int)tc[0][1 + 4 + k] - (int)tc[0][1 + 2 - k]);
                    V += (k + 1) * ((int)tc[1 + 4 + k][0] - (int)tc[1 + 2 - k][0]);
                }
                const int a = 16 * ((int)tc[8][0] + (int)tc[0][8]);
                const int bb = (17 * Hh + 16) >> 5, cc2 = (17 * V + 16) >> 5;
#pragma unroll
                for (int k = 0; k < 4; k++) pv[k] = clip255((a + bb * (cc + k - 3) + cc2 * (cr - 3) + 16) >> 5);
            }
            uint32_t oc = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) oc |= (uint32_t)clip255(pv[k] + resC[k]) << (8 * k);
            *reinterpret_cast<uint32_t *>(dstC) = oc;
        }
        // publish: the warp barrier orders every lane's stores before lane 0's release at gpu scope (cumulative; no extra fence)
        __syncwarp();
        if (lane == 0) stRelease(doneS + mb, p.serial);
    }
}

}  // namespace b200
