// recon_kernel.cuh -- macroblock reconstruction: inverse zig-zag + dequant + IDCT, intra /
// inter prediction, add residual, write.  One warp per macroblock.
//
// Restates, as sm_100a device code: h264bsd_transform.c:97-401 (residual),
// h264bsd_intra_prediction.c:478-1830, h264bsd_inter_prediction.c:361-482 +
// h264bsd_reconstruct.c:109-2367 (motion compensation), h264bsd_image.c:81-344 (write).
#pragma once
#include "device_common.cuh"

namespace b200 {

constexpr int kReconWarps = 8;    // warps per CTA of the intra pass
#ifndef B200_INTRA_MINBLOCKS
#define B200_INTRA_MINBLOCKS 5
#endif
#ifndef B200_PASSA_WARPS
#define B200_PASSA_WARPS 4
#endif
constexpr int kPassAWarps = B200_PASSA_WARPS;   // warps per CTA of pass A

// Tensor maps of pass A (Batch::create): strip-major planes -> raster windows (pool_geom.hpp).  A luma box is nx strips wide
// (16 nx pels) and 16 rows (integer vertical vector) or 21 rows (16 + 5 for the six-tap filter) high; a chroma box nx strips
// (8 nx pels of Cb and of Cr) by 8 or 9 rows.
struct PassAMaps {
    CUtensorMap luma[3][2];     // [nx - 1][0: 16 rows, 1: 21 rows]
    CUtensorMap chroma[2][2];   // [nx - 1][0: 8 rows,  1: 9 rows]
};

constexpr int kLumaBufBytes = 1024;     // >= 48 x 21
constexpr int kChromaBufBytes = 384;    // >= 32 x 9
constexpr int kCoefBufBytes = 896;      // >= 26 blocks of 32 bytes (I_PCM: 12)
#ifndef B200_STAGE_MBS
#define B200_STAGE_MBS 14
#endif
constexpr int kStageMbs = B200_STAGE_MBS;   // macroblocks of a chunk at most (Batch::create): what the copy staging buffer holds

struct __align__(128) PassAWarpBase {
    uint8_t luma[2][kLumaBufBytes];       // reference windows: the macroblock being computed / the next one (TMA, mbarrier double buffer)
    uint8_t chroma[2][kChromaBufBytes];
    uint8_t coef[2][kCoefBufBytes];       // the macroblock's levels (cp.async.bulk, same mbarrier)
    int16_t resY[16][16];                 // residual of the macroblock being computed, raster
    int16_t resC[2][8][8];
    uint64_t mbar[2];
    uint64_t mbarCopy;                    // the chunk's copies have arrived in `stage`
    uint8_t pad[104];
};
// first instance: the chunk's zero-motion copies on their way from the reference frame to the current one (bulk copies global ->
// shared -> global: luma of macroblock l at 256 l, chroma at 256 kStageMbs + 128 l)
struct __align__(128) PassAWarpSmem : PassAWarpBase {
    uint8_t stage[(kStageMbs * 384 + 127) / 128 * 128];
};
// second instance (it has no copies): prediction of sub-macroblocks with 8x4 / 4x8 / 4x4 partitions in the first 384 bytes, then a
// second pair of window buffers, then the round boxes
struct __align__(128) MultiWarpSmem : PassAWarpBase {
    uint8_t stage[4480];
};
static_assert(sizeof(PassAWarpBase) % 128 == 0 && sizeof(PassAWarpSmem) % 128 == 0 && sizeof(MultiWarpSmem) % 128 == 0,
              "per-warp shared memory keeps the TMA destinations 128-byte aligned");

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
// the same times 2^(qp / 6), by qp (transform.c:122-127): what a level is multiplied with
__device__ __constant__ uint16_t cScaleQp[52][4] = {{10, 13, 16, 0}, {11, 14, 18, 0}, {13, 16, 20, 0}, {14, 18, 23, 0}, {16, 20, 25, 0}, {18, 23, 29, 0}, {20, 26, 32, 0}, {22, 28, 36, 0}, {26, 32, 40, 0}, {28, 36, 46, 0}, {32, 40, 50, 0}, {36, 46, 58, 0}, {40, 52, 64, 0}, {44, 56, 72, 0}, {52, 64, 80, 0}, {56, 72, 92, 0}, {64, 80, 100, 0}, {72, 92, 116, 0}, {80, 104, 128, 0}, {88, 112, 144, 0}, {104, 128, 160, 0}, {112, 144, 184, 0}, {128, 160, 200, 0}, {144, 184, 232, 0}, {160, 208, 256, 0}, {176, 224, 288, 0}, {208, 256, 320, 0}, {224, 288, 368, 0}, {256, 320, 400, 0}, {288, 368, 464, 0}, {320, 416, 512, 0}, {352, 448, 576, 0}, {416, 512, 640, 0}, {448, 576, 736, 0}, {512, 640, 800, 0}, {576, 736, 928, 0}, {640, 832, 1024, 0}, {704, 896, 1152, 0}, {832, 1024, 1280, 0}, {896, 1152, 1472, 0}, {1024, 1280, 1600, 0}, {1152, 1472, 1856, 0}, {1280, 1664, 2048, 0}, {1408, 1792, 2304, 0}, {1664, 2048, 2560, 0}, {1792, 2304, 2944, 0}, {2048, 2560, 3200, 0}, {2304, 2944, 3712, 0}, {2560, 3328, 4096, 0}, {2816, 3584, 4608, 0}, {3328, 4096, 5120, 0}, {3584, 4608, 5888, 0}};

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
// A window lies in shared memory as a raster of `pitch` bytes per row (16, 32 or 48: the box is 1..3 strips wide).  G0 is
// the address of the partition's integer sample (0, 0) inside it; with a fractional vector component the window starts two
// samples before and ends three after the partition on that axis.
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }
__device__ __forceinline__ int hsum(const uint8_t *G0, int pitch, int x, int y) {
    const uint8_t *q = G0 + y * pitch + x;
    return tap6(q[-2], q[-1], q[0], q[1], q[2], q[3]);
}
__device__ __forceinline__ int vsum(const uint8_t *G0, int pitch, int x, int y) {
    const uint8_t *q = G0 + y * pitch + x;
    return tap6(q[-2 * pitch], q[-pitch], q[0], q[pitch], q[2 * pitch], q[3 * pitch]);
}
// clause 8.4.2.2.1; dispatch table of h264bsdPredictSamples (reconstruct.c:1848-1927), one sample
__device__ __forceinline__ int lumaQpel(const uint8_t *G0, int pitch, int x, int y, int xf, int yf) {
    if ((xf | yf) == 0) return G0[y * pitch + x];
    if (yf == 0) {
        int b = clip255((hsum(G0, pitch, x, y) + 16) >> 5);
        if (xf == 2) return b;
        return (b + G0[y * pitch + x + (xf >> 1)] + 1) >> 1;
    }
    if (xf == 0) {
        int h = clip255((vsum(G0, pitch, x, y) + 16) >> 5);
        if (yf == 2) return h;
        return (h + G0[(y + (yf >> 1)) * pitch + x] + 1) >> 1;
    }
    if (xf != 2 && yf != 2) {
        int b = clip255((hsum(G0, pitch, x, y + (yf >> 1)) + 16) >> 5);
        int h = clip255((vsum(G0, pitch, x + (xf >> 1), y) + 16) >> 5);
        return (b + h + 1) >> 1;
    }
    int j = clip255((tap6(hsum(G0, pitch, x, y - 2), hsum(G0, pitch, x, y - 1), hsum(G0, pitch, x, y), hsum(G0, pitch, x, y + 1),
                          hsum(G0, pitch, x, y + 2), hsum(G0, pitch, x, y + 3)) + 512) >> 10);
    if (xf == 2 && yf == 2) return j;
    if (xf == 2) {
        int b = clip255((hsum(G0, pitch, x, y + (yf >> 1)) + 16) >> 5);
        return (j + b + 1) >> 1;
    }
    int h = clip255((vsum(G0, pitch, x + (xf >> 1), y) + 16) >> 5);
    return (j + h + 1) >> 1;
}

// 8 consecutive bytes from an arbitrarily aligned shared-memory address
// (the low bits of a generic pointer into shared memory are those of the shared offset: no address-space conversion needed)
__device__ __forceinline__ uint2 lds8(const uint8_t *p) {
    const uint32_t a = (uint32_t)reinterpret_cast<uintptr_t>(p), sh = (a & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - (a & 3u));
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
__device__ __forceinline__ uint32_t lds4(const uint8_t *p) {
    const uint32_t a = (uint32_t)reinterpret_cast<uintptr_t>(p), sh = (a & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - (a & 3u));
    return __funnelshift_r(w[0], w[1], sh);
}

// ---- 8-wide luma prediction for one lane (row y, columns x0..x0+7 of a partition) --------------------
constexpr int kTapsLo = 0x1414FB01;  // bytes (1, -5, 20, 20)
constexpr int kTapsHi = 0x000001FB;  // bytes (-5, 1, 0, 0)

// horizontal 6-tap sums (+ acc0) for 8 outputs; rowp points at sample x0-2 of the row (13 samples are read).  The taps of
// output k are bytes k..k+5 of the row: two dot products, over bytes k..k+3 and k+4..k+7 (the last two times zero)
__device__ __forceinline__ void hrow8(const uint8_t *rowp, int *hs, int acc0) {
    const uint32_t a = (uint32_t)reinterpret_cast<uintptr_t>(rowp), sh0 = (a & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rowp - (a & 3u));
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    // aligned view: byte k of the row = byte (k + (a&3)) of (w0,w1,w2,w3)
    const uint32_t v0 = __funnelshift_r(w0, w1, sh0), v1 = __funnelshift_r(w1, w2, sh0), v2 = __funnelshift_r(w2, w3, sh0), v3 = w3 >> sh0;
    uint32_t q[12];   // q[k] = bytes k..k+3
    q[0] = v0; q[4] = v1; q[8] = v2;
#pragma unroll
    for (int k = 1; k < 4; k++) {
        q[k] = __funnelshift_r(v0, v1, 8 * k);
        q[4 + k] = __funnelshift_r(v1, v2, 8 * k);
        q[8 + k] = __funnelshift_r(v2, v3, 8 * k);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) hs[k] = dp4aUS(q[k + 4], kTapsHi, dp4aUS(q[k], kTapsLo, acc0));
}
// vertical 6-tap sums (+ acc0) for 8 outputs; colp points at sample (x0, y-2).  The six rows are read as 8-byte spans and
// transposed four columns at a time (byte permutes), so that a column's taps sit in one word for the dot products
__device__ __forceinline__ void vcol8(const uint8_t *colp, int pitch, int *vs, int acc0) {
    uint2 r[6];
#pragma unroll
    for (int t = 0; t < 6; t++) r[t] = lds8(colp + t * pitch);
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t a0 = half ? r[0].y : r[0].x, a1 = half ? r[1].y : r[1].x, a2 = half ? r[2].y : r[2].x;
        const uint32_t a3 = half ? r[3].y : r[3].x, a4 = half ? r[4].y : r[4].x, a5 = half ? r[5].y : r[5].x;
        const uint32_t t0 = __byte_perm(a0, a1, 0x5140), t1 = __byte_perm(a2, a3, 0x5140);
        const uint32_t t2 = __byte_perm(a0, a1, 0x7362), t3 = __byte_perm(a2, a3, 0x7362);
        const uint32_t c0 = __byte_perm(t0, t1, 0x5410), c1 = __byte_perm(t0, t1, 0x7632);
        const uint32_t c2 = __byte_perm(t2, t3, 0x5410), c3 = __byte_perm(t2, t3, 0x7632);
        // rows 4 and 5 of column k in the two low bytes (the two high bytes meet zero taps)
        vs[4 * half + 0] = dp4aUS(__byte_perm(a4, a5, 0x0040), kTapsHi, dp4aUS(c0, kTapsLo, acc0));
        vs[4 * half + 1] = dp4aUS(__byte_perm(a4, a5, 0x0051), kTapsHi, dp4aUS(c1, kTapsLo, acc0));
        vs[4 * half + 2] = dp4aUS(__byte_perm(a4, a5, 0x0062), kTapsHi, dp4aUS(c2, kTapsLo, acc0));
        vs[4 * half + 3] = dp4aUS(__byte_perm(a4, a5, 0x0073), kTapsHi, dp4aUS(c3, kTapsLo, acc0));
    }
}
// clip255(v >> sh) of eight sums, packed (the rounding constant is already in the sums)
__device__ __forceinline__ uint2 pack8shift(const int *v, int sh) {
    return make_uint2(pack4sat(v[0] >> sh, v[1] >> sh, v[2] >> sh, v[3] >> sh), pack4sat(v[4] >> sh, v[5] >> sh, v[6] >> sh, v[7] >> sh));
}
__device__ __forceinline__ uint2 avg8(uint2 a, uint2 b) { return make_uint2(__vavgu4(a.x, b.x), __vavgu4(a.y, b.y)); }

// clause 8.4.2.2.1 for 8 horizontally adjacent samples: (x0, y) = position of the first sample inside the partition; the
// same arithmetic as lumaQpel, 8 at a time, on packed bytes: the rounded averages (a + b + 1) >> 1 are per-byte averages of
// clipped values
__device__ __noinline__ uint2 lumaQpel8(const uint8_t *G0, int pitch, int x0, int y, int xf, int yf) {
    const uint8_t *at = G0 + y * pitch + x0;   // the lane's first integer sample
    if ((xf | yf) == 0) return lds8(at);       // h264bsdFillBlock copy (reconstruct.c:1852)
    const bool jfam = (xf == 2 || yf == 2) && xf != 0 && yf != 0;
    if (!jfam) {
        const bool useH = xf != 0, useV = yf != 0;
        uint2 b = make_uint2(0, 0), h = make_uint2(0, 0);
        if (useH) {
            int t[8];
            hrow8(at + (yf == 3 ? pitch : 0) - 2, t, 16);
            b = pack8shift(t, 5);
        }
        if (useV) {
            int t[8];
            vcol8(at - 2 * pitch + (xf == 3 ? 1 : 0), pitch, t, 16);
            h = pack8shift(t, 5);
        }
        if (useH && useV) return avg8(b, h);
        if (useH) return xf == 2 ? b : avg8(b, lds8(at + (xf >> 1)));
        return yf == 2 ? h : avg8(h, lds8(at + (yf >> 1) * pitch));
    }
    int acc[8], bsel[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = 512; bsel[k] = 0; }
    const int brow = 2 + (yf == 3 ? 1 : 0);
#pragma unroll 1
    for (int t = 0; t < 6; t++) {
        int hs[8];
        hrow8(at + (t - 2) * pitch - 2, hs, 0);
        const int c = (t == 0 || t == 5) ? 1 : (t == 1 || t == 4) ? -5 : 20;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            acc[k] += c * hs[k];
            if (t == brow) bsel[k] = hs[k] + 16;
        }
    }
    const uint2 j = pack8shift(acc, 10);
    if (xf == 2 && yf == 2) return j;
    if (xf == 2) return avg8(j, pack8shift(bsel, 5));
    int hv[8];
    vcol8(at - 2 * pitch + (xf == 3 ? 1 : 0), pitch, hv, 16);
    return avg8(j, pack8shift(hv, 5));
}

// ---- chroma: bilinear 1/8-pel (PredictChroma, reconstruct.c:415-475) ----------------------------------------
// A chroma window row in shared memory is [strip 0: 8 Cb | 8 Cr][strip 1: 8 Cb | 8 Cr]; pA points at the 8-byte group of the
// lane's plane that holds window column col0, sh = col0 & 7: five consecutive samples of the plane starting at col0
__device__ __forceinline__ void chromaRow5(const uint8_t *pA, uint32_t sh, uint32_t &lo, uint32_t &hi) {
    const uint2 A = *reinterpret_cast<const uint2 *>(pA), B = *reinterpret_cast<const uint2 *>(pA + 16);
    const bool up = sh >= 4u;
    const uint32_t w0 = up ? A.y : A.x, w1 = up ? B.x : A.y, w2 = up ? B.y : B.x, s = (sh & 3u) * 8u;
    lo = __funnelshift_r(w0, w1, s);
    hi = __funnelshift_r(w1, w2, s);
}
// four samples of plane cp: window columns cxo + lcx .. + 3, row lcy; the four weights times four pels are one dot product
__device__ __noinline__ uint32_t chromaPred4(const uint8_t *cbuf, int pitchC, int cxo, int cp, int lcx, int lcy, int cxf, int cyf) {
    const int col0 = cxo + lcx;
    const uint8_t *pA = cbuf + lcy * pitchC + (col0 >> 3) * 16 + cp * 8;
    const uint32_t sh = (uint32_t)col0 & 7u;
    uint32_t lo0, hi0;
    chromaRow5(pA, sh, lo0, hi0);
    if ((cxf | cyf) == 0) return lo0;
    uint32_t lo1, hi1;
    chromaRow5(pA + pitchC, sh, lo1, hi1);   // (with cyf == 0 this row may lie outside the box: it meets zero weights)
    const int wts = ((8 - cxf) * (8 - cyf)) | ((cxf * (8 - cyf)) << 8) | (((8 - cxf) * cyf) << 16) | ((cxf * cyf) << 24);
    const uint32_t s0 = __funnelshift_r(lo0, hi0, 24), s1 = __funnelshift_r(lo1, hi1, 24);
    const int v0 = dp4aUS(__byte_perm(lo0, lo1, 0x5410), wts, 32) >> 6, v1 = dp4aUS(__byte_perm(lo0, lo1, 0x6521), wts, 32) >> 6;
    const int v2 = dp4aUS(__byte_perm(lo0, lo1, 0x7632), wts, 32) >> 6, v3 = dp4aUS(__byte_perm(s0, s1, 0x5410), wts, 32) >> 6;
    return (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
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
    const int r8 = lane >> 1, cp = (lane >> 1) & 1, cr = lane >> 2;
    const int by = r8 >> 2, ry = r8 & 3, bx = (lane & 1) * 2;
    const int16_t *ra = res[cRasterToBlk[by * 4 + bx]] + ry * 4;
    const int16_t *rb = res[cRasterToBlk[by * 4 + bx + 1]] + ry * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) { resY[i] = ra[i]; resY[4 + i] = rb[i]; }
    const int16_t *rc = res[16 + cp * 4 + (cr >> 2) * 2 + (lane & 1)] + (cr & 3) * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) resC[i] = rc[i];
}

// =====================================================================================================
// pass A: every macroblock that does not look at the current picture -- inter-predicted ones, I_PCM, concealed copies.
// One kernel, in the order the macroblocks lie in memory: a warp task ("chunk") is up to 32 vertically adjacent macroblocks
// of one strip.  Lane l reads record l of the chunk and the warp sorts them by two ballots:
//   copies  P_Skip / P_L0_16x16 without residual and with a zero vector (60 % of the fixture's macroblocks) and concealed
//           copies: h264bsdPredictSamples degenerates to h264bsdFillBlock (reconstruct.c:1852) and h264bsdWriteOutputBlocks to
//           a store.  In the strip layout that is 384 contiguous bytes from the same place of the reference frame, and
//           neighbouring copies of a chunk continue each other: the warp moves all of them as one list of 16-byte units,
//           four loads in flight per lane before the first store;
//   inter   everything else with a vector: reference windows by TMA (one luma and one chroma box per partition, landing
//           in shared memory as rasters), levels by cp.async.bulk on the same mbarrier, both for macroblock i + 1 while
//           macroblock i is computed; residual by a 4-lanes-per-block transform with warp shuffles; 6-tap / bilinear
//           filters on packed bytes (dp4a); add, clip, and one 256-byte + one 128-byte store per macroblock.
// Copies wait on DRAM while inter macroblocks keep the issue slots busy -- in one kernel they overlap by themselves, and a
// window's halo is in L2 because the neighbouring macroblocks, whatever their kind, are in flight at the same time.
// =====================================================================================================

// ---- residual of an inter macroblock: 4 lanes per 4x4 block, 8 blocks per round, coded blocks only ---------------------
// h264bsdProcessBlock (transform.c:97-234) with the row transform inside a lane (lane r of a group owns row r of the block)
// and the column transform across the group's four lanes by two shuffle exchanges; h264bsdProcessChromaDc (:359-401) by
// lanes 16..23 first.  `cbuf` = the macroblock's levels in shared memory (b200_mb_rec layout: [chroma DC][coded blocks]).
__device__ __noinline__ void residualShfl(PassAWarpBase &sm, const uint8_t *cbuf, uint32_t mask, int qpY, int qpC, int lane, uint32_t *errors) {
    {   // every block that is not visited below has a zero residual
        uint4 *z = reinterpret_cast<uint4 *>(&sm.resY[0][0]);   // resY and resC are contiguous: 48 x 16 bytes
        const uint4 zero = make_uint4(0, 0, 0, 0);
        z[lane] = zero;
        if (lane < 16) z[lane + 32] = zero;
    }
    const int r = lane & 3, g4 = lane >> 2;
    // chroma DC: group g4 works on chroma block 16 + g4 in the third round and needs DC value g4 of the 2 x (2x2) transforms
    // (macroblock_layer.c:1371-1374); every lane of the group computes it
    int dcMine = 0;
    if (mask & B200_CM_CHROMA_DC) dcMine = chromaDcPick(reinterpret_cast<const int16_t *>(cbuf) + (g4 >> 2) * 4, qpC, g4 & 3);
    const uint32_t coded = mask & 0xFFFFFFu;
    // rounds with something to do: coded blocks, and for the chroma round a DC block (its values reach blocks without levels)
    const uint32_t active = coded | ((mask & B200_CM_CHROMA_DC) ? 0xFF0000u : 0u);
    const int nDc = (int)((mask >> 25) & 1u);
    const uint32_t zz = r == 0 ? 0x6510u : r == 1 ? 0xC742u : r == 2 ? 0xDB83u : 0xFEA9u;   // zig-zag positions of raster row r
    const int orow = ((r & 1) << 1) | (r >> 1);   // the row this lane holds after the column transform
    // this lane's two scale factors (columns 0, 2 / 1, 3 of its row), for a luma and for a chroma block
    const int sAY = cScaleQp[qpY][r & 1], sBY = cScaleQp[qpY][1 + (r & 1)], sAC = cScaleQp[qpC][r & 1], sBC = cScaleQp[qpC][1 + (r & 1)];
    // three rounds of eight blocks: luma 0..7, luma 8..15, chroma 16..23; group g4 takes block 8 round + g4; a round without an
    // active block is skipped
#pragma unroll 1
    for (int round = 0; round < 3; round++) {
        if (!((active >> (8 * round)) & 0xFFu)) continue;
        const int b = 8 * round + g4;
        const bool isCoded = (coded >> b) & 1u, valid = isCoded || (round == 2 && dcMine != 0);
        const int sA = round < 2 ? sAY : sAC, sB = round < 2 ? sBY : sBC;
        int d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        if (isCoded) {
            const int16_t *lev = reinterpret_cast<const int16_t *>(cbuf) + (nDc + __popc(coded & ((1u << b) - 1u))) * 16;
            d0 = lev[zz & 15u] * sA; d1 = lev[(zz >> 4) & 15u] * sB; d2 = lev[(zz >> 8) & 15u] * sA; d3 = lev[zz >> 12] * sB;
        }
        if (r == 0) {
            if (round == 2) d0 = dcMine;        // chroma: the DC comes from the 2x2 transform
            d0 += 32;                           // the DC term reaches all sixteen outputs with weight 1: rounding for the final >> 6
        }
        // row transform
        const int t0 = d0 + d2, t1 = d0 - d2, t2 = (d1 >> 1) - d3, t3 = d1 + (d3 >> 1);
        int e[4] = {t0 + t3, t1 + t2, t1 - t2, t0 - t3};
        // column transform: rows 0/2 and 1/3 meet first (f0 = E0 + E2 in lane 0, f1 = E0 - E2 in lane 2, f2 = (E1 >> 1) - E3 in
        // lane 1, f3 = E1 + (E3 >> 1) in lane 3), then 0/3 and 1/2 (rows 0 and 3 out of lanes 0 and 3, rows 1 and 2 out of lanes 2 and 1)
        int o[4];
        uint32_t range = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int pv = __shfl_xor_sync(0xffffffffu, e[c], 2);
            const int h = (r & 1) ? (e[c] >> 1) : e[c];
            const int f = (r == 2 ? -h : h) + (r == 1 ? -pv : pv);
            const int q = __shfl_xor_sync(0xffffffffu, f, 3);
            o[c] = ((r & 1) ? q - f : f + q) >> 6;
            range |= (uint32_t)(o[c] + 512);    // inside [-512, 511] exactly when no bit above the tenth is set
        }
        if (valid) {
            if (range >> 10) atomicAdd(errors, 1u);
            int16_t *dst;
            if (b < 16) dst = &sm.resY[(((b >> 1) & 1) + ((b >> 3) & 1) * 2) * 4 + orow][((b & 1) + ((b >> 2) & 1) * 2) * 4];
            else dst = &sm.resC[(b - 16) >> 2][((b >> 1) & 1) * 4 + orow][(b & 1) * 4];
            *reinterpret_cast<uint2 *>(dst) = make_uint2(((uint32_t)o[0] & 0xFFFFu) | ((uint32_t)o[1] << 16), ((uint32_t)o[2] & 0xFFFFu) | ((uint32_t)o[3] << 16));
        }
    }
    __syncwarp();
}

// The centre positions (xf and yf both fractional, one of them a half: clause 8.4.2.2.1 "j" and its quarter-sample neighbours)
// of a 16x16 partition, all lanes together: the 21 x 16 horizontal six-tap sums are computed ONCE (42 row halves over 32 lanes)
// and shared through shared memory -- `scratch` = the residual arrays, not yet in use when the prediction is made -- instead of
// six row halves per lane; the vertical six-tap sum over them is three two-way dot products per sample (dp2a over int16 pairs).
// G0 = integer sample (0, 0) of the partition in the window.  Ends with all reads of `scratch` done by this lane only: the
// caller puts a warp barrier before the residual overwrites it.
__device__ __noinline__ uint2 lumaCentre8(int16_t *scratch, const uint8_t *G0, int pitch, int lane, int xf, int yf) {
    int16_t (*H)[16] = reinterpret_cast<int16_t (*)[16]>(scratch);   // H[r] = sums of picture row r - 2; 21 rows, 672 bytes
    const int r8 = lane >> 1, c8 = (lane & 1) * 8;
#pragma unroll 1
    for (int task = lane; task < 42; task += 32) {
        const int hr = task >> 1, hc = (task & 1) * 8;
        int hs[8];
        hrow8(G0 + (hr - 2) * pitch + hc - 2, hs, 0);
        *reinterpret_cast<uint4 *>(&H[hr][hc]) =
            make_uint4(((uint32_t)hs[0] & 0xFFFFu) | ((uint32_t)hs[1] << 16), ((uint32_t)hs[2] & 0xFFFFu) | ((uint32_t)hs[3] << 16),
                       ((uint32_t)hs[4] & 0xFFFFu) | ((uint32_t)hs[5] << 16), ((uint32_t)hs[6] & 0xFFFFu) | ((uint32_t)hs[7] << 16));
    }
    __syncwarp();
    uint4 R[6];
#pragma unroll
    for (int t = 0; t < 6; t++) R[t] = *reinterpret_cast<const uint4 *>(&H[r8 + t][c8]);
    int acc[8];
    auto col2 = [&](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, int &lo, int &hi) {
        // rows t, t + 1 of a column side by side: taps (1, -5), (20, 20), (-5, 1)
        lo = dp2aLoSS(__byte_perm(a4, a5, 0x5410), 0x01FB, dp2aLoSS(__byte_perm(a2, a3, 0x5410), 0x1414, dp2aLoSS(__byte_perm(a0, a1, 0x5410), 0xFB01, 512)));
        hi = dp2aLoSS(__byte_perm(a4, a5, 0x7632), 0x01FB, dp2aLoSS(__byte_perm(a2, a3, 0x7632), 0x1414, dp2aLoSS(__byte_perm(a0, a1, 0x7632), 0xFB01, 512)));
    };
    col2(R[0].x, R[1].x, R[2].x, R[3].x, R[4].x, R[5].x, acc[0], acc[1]);
    col2(R[0].y, R[1].y, R[2].y, R[3].y, R[4].y, R[5].y, acc[2], acc[3]);
    col2(R[0].z, R[1].z, R[2].z, R[3].z, R[4].z, R[5].z, acc[4], acc[5]);
    col2(R[0].w, R[1].w, R[2].w, R[3].w, R[4].w, R[5].w, acc[6], acc[7]);
    const uint2 j = pack8shift(acc, 10);
    if (xf == 2 && yf == 2) return j;
    if (xf == 2) {
        const uint4 B = yf == 3 ? R[3] : R[2];   // the horizontal half-sample row next to j
        int bs[8];
        bs[0] = (int)(int16_t)(B.x & 0xFFFFu) + 16; bs[1] = ((int)B.x >> 16) + 16; bs[2] = (int)(int16_t)(B.y & 0xFFFFu) + 16; bs[3] = ((int)B.y >> 16) + 16;
        bs[4] = (int)(int16_t)(B.z & 0xFFFFu) + 16; bs[5] = ((int)B.z >> 16) + 16; bs[6] = (int)(int16_t)(B.w & 0xFFFFu) + 16; bs[7] = ((int)B.w >> 16) + 16;
        return avg8(j, pack8shift(bs, 5));
    }
    int hv[8];
    vcol8(G0 + (r8 - 2) * pitch + c8 + (xf == 3 ? 1 : 0), pitch, hv, 16);
    return avg8(j, pack8shift(hv, 5));
}

// window of a partition at picture position (px, py), size w x h, vector (mvx, mvy): clamp the origin into the bordered plane (a
// window wholly outside the picture on an axis equals the window at the clamped origin because the border is a replication,
// SURVEY 7.2) and pick the boxes.  geom = xo | nx << 4 | cxo << 8 | nxC << 12.
struct WindowGeom {
    int strip, row, map;       // luma box: first strip, first row, tensor map (nx - 1) * 2 + (21 rows ? 1 : 0)
    int stripC, rowC, mapC;    // chroma box
    uint32_t bytes, geom;
};
__device__ __forceinline__ WindowGeom windowGeom(int W, int H, int px, int py, int w, int h, int mvx, int mvy) {
    const int xf = mvx & 3, yf = mvy & 3, nc = w + (xf ? 5 : 0), nr = h + (yf ? 5 : 0);
    const int x0 = clip3(-kPadY, W + kPadY - nc, px + (mvx >> 2) - (xf ? 2 : 0)) + kPadY;
    const int y0 = clip3(-kPadY, H + kPadY - nr, py + (mvy >> 2) - (yf ? 2 : 0)) + kPadY;
    const int xo = x0 & 15, nx = (xo + nc + 15) >> 4;
    const int cxf = mvx & 7, cyf = mvy & 7, ncC = (w >> 1) + (cxf ? 1 : 0), nrC = (h >> 1) + (cyf ? 1 : 0);
    const int cx0 = clip3(-kPadC, W / 2 + kPadC - ncC, (px >> 1) + (mvx >> 3)) + kPadC;
    const int cy0 = clip3(-kPadC, H / 2 + kPadC - nrC, (py >> 1) + (mvy >> 3)) + kPadC;
    const int cxo = cx0 & 7, nxC = (cxo + ncC + 7) >> 3;
    WindowGeom r;
    r.strip = x0 >> 4; r.row = y0; r.map = (nx - 1) * 2 + (yf ? 1 : 0);
    r.stripC = cx0 >> 3; r.rowC = cy0; r.mapC = (nxC - 1) * 2 + (cyf ? 1 : 0);
    r.bytes = (uint32_t)(16 * nx * (yf ? 21 : 16) + 16 * nxC * (cyf ? 9 : 8));
    r.geom = (uint32_t)xo | ((uint32_t)nx << 4) | ((uint32_t)cxo << 8) | ((uint32_t)nxC << 12);
    return r;
}
// one window into (dstL, dstC), completion on `bar`, issued by lane 0 (sub-macroblock partitions below 8x8 only)
__device__ __noinline__ uint32_t issueWindowFn(uint8_t *dstL, uint8_t *dstC, uint64_t *bar, const PassAMaps *maps, int W, int H, int px, int py,
                                               int wh, int mvx, int mvy, uint32_t refFrame, int lane) {
    const WindowGeom wg = windowGeom(W, H, px, py, wh & 0xFF, wh >> 8, mvx, mvy);
    if (lane == 0) {
        fenceProxyAsync();
        mbarExpectTx(bar, wg.bytes);
        tmaLoad4d(dstL, &maps->luma[0][0] + wg.map, 0, wg.strip, wg.row, (int)refFrame, bar);
        tmaLoad4d(dstC, &maps->chroma[0][0] + wg.mapC, 0, wg.stripC, wg.rowC, (int)refFrame, bar);
    }
    return wg.geom;
}

#ifndef B200_PASSA_MINBLOCKS
#define B200_PASSA_MINBLOCKS 5
#endif

// add residual + clip + store of a macroblock (h264bsdWriteOutputBlocks, image.c:172-344): lane = 8 luma samples (bytes 8 lane..
// of the macroblock's 256) + 4 chroma samples (bytes 4 lane.. of its 128); a pel leaves its word and meets its residual in one
// dot product
__device__ __forceinline__ void addResidualStore(const PassAWarpBase &sm, uint32_t mask, uint2 pv, uint32_t pc, uint8_t *mbY, uint8_t *mbC, int lane) {
    if (mask) {
        const uint4 ra = reinterpret_cast<const uint4 *>(&sm.resY[0][0])[lane];
        const uint2 rc = *reinterpret_cast<const uint2 *>(&sm.resC[(lane >> 1) & 1][lane >> 2][(lane & 1) * 4]);
        auto lo = [](uint32_t w) { return (int)(int16_t)(w & 0xFFFFu); };
        auto hi = [](uint32_t w) { return (int)w >> 16; };
        pv = make_uint2(pack4sat(dp4aUS(pv.x, 0x00000001, lo(ra.x)), dp4aUS(pv.x, 0x00000100, hi(ra.x)),
                                 dp4aUS(pv.x, 0x00010000, lo(ra.y)), dp4aUS(pv.x, 0x01000000, hi(ra.y))),
                        pack4sat(dp4aUS(pv.y, 0x00000001, lo(ra.z)), dp4aUS(pv.y, 0x00000100, hi(ra.z)),
                                 dp4aUS(pv.y, 0x00010000, lo(ra.w)), dp4aUS(pv.y, 0x01000000, hi(ra.w))));
        pc = pack4sat(dp4aUS(pc, 0x00000001, lo(rc.x)), dp4aUS(pc, 0x00000100, hi(rc.x)),
                      dp4aUS(pc, 0x00010000, lo(rc.y)), dp4aUS(pc, 0x01000000, hi(rc.y)));
    }
    *reinterpret_cast<uint2 *>(mbY + lane * 8) = pv;
    *reinterpret_cast<uint32_t *>(mbC + lane * 4) = pc;
}

// wait for bulk copies that take microseconds: back off instead of spinning in the issue slots the other warps compute with
__device__ __forceinline__ void mbarWaitSleeping(uint64_t *bar, uint32_t parity) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!mbarTryWait(bar, parity)) {
        __nanosleep(160);
        if ((++spins & 255u) == 0) {
            const unsigned long long now = globalTimerNs();
            if (!t0) t0 = now;
            else if (now - t0 > kWatchdogNs) { atomicAdd(&gWatchdog[1], 1u); break; }
        }
    }
}

// First instance: the copies, the macroblocks with one partition (P_Skip / P_L0_16x16) and I_PCM -- 94 % of a typical P picture's
// pass-A macroblocks, through the short code path.  The macroblocks with several partitions (16x8, 8x16, 8x8 and below) are put
// on a list (stream << 16 | macroblock address, any order) for the second instance, passAMultiKernel.
__global__ void __launch_bounds__(kPassAWarps * 32, B200_PASSA_MINBLOCKS)
passAKernel(const ReconParams p, const __grid_constant__ PassAMaps maps) {
    extern __shared__ __align__(128) uint8_t interSmemRaw[];   // kPassAWarps x PassAWarpSmem (more than the 48 KB static limit)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    PassAWarpSmem &sm = reinterpret_cast<PassAWarpSmem *>(interSmemRaw)[warp];
    if (lane == 0) {
        mbarInit(&sm.mbar[0], 1);
        mbarInit(&sm.mbar[1], 1);
        mbarInit(&sm.mbarCopy, 1);
        fenceMbarInit();
    }
    __syncwarp();
    uint32_t phaseBits = 0;   // bit b = phase parity of mbarrier b, bit 2 = of the copy barrier
    const uint32_t nWarps = gridDim.x * kPassAWarps;
    const uint32_t chunksPerStream = p.chunksPerCol * (uint32_t)g.widthMbs;
    // this lane's spans inside a macroblock: 8 luma samples (row r8, columns c8..c8+7 = bytes 8 lane.. of the 256), 4 chroma
    // samples (plane cp, row cr, columns cc..cc+3 = bytes 4 lane.. of the 128)
    const int r8 = lane >> 1, c8 = (lane & 1) * 8;
    const int cr = lane >> 2, cp = (lane >> 1) & 1, cc = (lane & 1) * 4;

    // Lane l fetches the words of record l of a chunk -- head, reference slots, word 7 (concealment), first vector -- one chunk
    // ahead: while a chunk is worked on, the records of the warp's next one are on their way (ticket -> job -> records is a
    // chain of three dependent memory round trips otherwise)
    uint32_t fPos = 0, fS = 0, fW0 = 0, fMask = 0, fCoef = 0, fRef = 0, fW7 = 0, fMv = 0;
    auto fetch = [&](uint32_t c) {
        if (c >= p.totalChunks) return;
        const uint32_t s = c / chunksPerStream, c2 = c - s * chunksPerStream;
        const uint32_t mbx = c2 / p.chunksPerCol, row0 = (c2 - mbx * p.chunksPerCol) * p.chunkRows;
        fS = s;
        fPos = mbx | (row0 << 16);
        if ((int)(row0 + (uint32_t)lane) < g.heightMbs && (uint32_t)lane < p.chunkRows) {
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(p.jobs[s].recs + (size_t)(row0 + (uint32_t)lane) * g.widthMbs + mbx);
            const uint4 hw = __ldg(reinterpret_cast<const uint4 *>(rw));
            fRef = __ldg(rw + 4);
            fW7 = __ldg(rw + 7);
            fMv = __ldg(rw + 8);
            fW0 = hw.x; fMask = hw.y; fCoef = hw.z;
        }
    };

    // a warp's first chunk is its own number, the following ones come from a ticket counter (asked for two chunks ahead)
    uint32_t chunk = blockIdx.x * kPassAWarps + warp;
    fetch(chunk);
    uint32_t nextChunk = 0;
    if (lane == 0) nextChunk = atomicAdd(p.ticketA, 1u) + nWarps;
    nextChunk = __shfl_sync(0xffffffffu, nextChunk, 0);
    while (chunk < p.totalChunks) {
        uint32_t ticket2 = 0;
        if (lane == 0) ticket2 = atomicAdd(p.ticketA, 1u) + nWarps;
        const uint32_t s = fS;
        const int mbx = (int)(fPos & 0xFFFFu), row0 = (int)(fPos >> 16);
        const int n = min((int)p.chunkRows, g.heightMbs - row0);
        const uint32_t mW0 = fW0, mMask = fMask, mCoef = fCoef, mRef = fRef, mMv = fMv, w7 = fW7;
        fetch(nextChunk);
        const StreamJob job = p.jobs[s];
        const uint32_t frameBase = s * (uint32_t)g.numSlots;
        uint8_t *cur = framePtr(p.pool, g, frameBase + job.curSlot);
        uint8_t *lbase = mbLuma(cur, g, mbx, row0), *cbase = mbChroma(cur, g, mbx, row0);   // macroblock l of the chunk: + 256 l / + 128 l

        bool isCopy = false, isInter = false, isMulti = false;
        if (lane < n) {
            const uint32_t type = mW0 & 0xFFu;
            // a concealed macroblock carries the state the filter wants to see (Intra4x4); its pels are a copy of the reference
            // picture (no neighbours to wait for) or come from concealKernel (h264bsd_b200_tape.h)
            const bool concealed = ((mW0 >> 24) & B200_MBF_CONCEALED) && type == B200_MB_I_4x4;
            if (concealed) isCopy = (w7 & 0xFFu) == 0;
            else if (type <= B200_MB_P_16x16 && mMask == 0 && mMv == 0) isCopy = true;
            else if (type <= B200_MB_P_16x16 || type == B200_MB_I_PCM) isInter = true;
            else isMulti = type <= B200_MB_P_8x8REF0;
        }
        const uint32_t copyMask = __ballot_sync(0xffffffffu, isCopy);
        uint32_t interMask = __ballot_sync(0xffffffffu, isInter);
        const uint32_t multiMask = __ballot_sync(0xffffffffu, isMulti);
        if (multiMask) {
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(p.multiCount, (uint32_t)__popc(multiMask));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (isMulti) p.multiList[at + (uint32_t)__popc(multiMask & ((1u << lane) - 1u))] = (s << 16) | (uint32_t)((row0 + lane) * g.widthMbs + mbx);
        }

        // ---- copies: every run of vertically adjacent copies with the same reference frame is 256 n contiguous bytes of luma and
        // 128 n of chroma in the strip layout.  The lane of a run's first macroblock sends them through the staging buffer with
        // bulk copies -- no registers, no issue slots, and the inter macroblocks below are computed while they fly
        int runLen = 0;
        bulkWaitRead<0>();   // the staging buffer is free again once this lane's stores of the previous chunk have read it ...
        fenceProxyAsync();   // (... and what the previous chunk's computed macroblocks kept in their idle slots is out of the way)
        __syncwarp();        // ... and every other lane's
        if (copyMask) {
            // where the reference frame of this lane's macroblock lies relative to the current frame, in 256-byte units (a frame
            // stride is a multiple of 256)
            const int myDelta = ((int)(mRef & 0xFFu) - (int)job.curSlot) * (int)(g.frameStride >> 8);
            const int prevDelta = __shfl_up_sync(0xffffffffu, myDelta, 1);
            const bool cont = isCopy && lane > 0 && ((copyMask >> (lane - 1)) & 1u) && prevDelta == myDelta;
            const uint32_t contMask = __ballot_sync(0xffffffffu, cont);
            if (isCopy && !cont) runLen = __ffs(~((contMask >> lane) >> 1));   // 1 + the continuations that follow
            if (lane == 0) mbarExpectTx(&sm.mbarCopy, 384u * (uint32_t)__popc(copyMask));
            __syncwarp();
            if (runLen) {
                const long long delta = (long long)myDelta * 256;
                bulkLoad(sm.stage + lane * 256, lbase + lane * 256 + delta, 256u * (uint32_t)runLen, &sm.mbarCopy);
                bulkLoad(sm.stage + kStageMbs * 256 + lane * 128, cbase + lane * 128 + delta, 128u * (uint32_t)runLen, &sm.mbarCopy);
            }
        }

        // Every lane works out where the windows of ITS macroblock lie, all macroblocks of the chunk at once instead of one after
        // the other when their turn comes.
        //   gX = luma strip | chroma strip << 16     gY = luma row | chroma row << 16     gM = luma map | chroma map << 3 | window bytes << 8
        // for the loads, and for the lanes that compute
        //   gC = offset of integer sample (0, 0) in the luma window (7 bits) | luma pitch / 16 << 7 | chroma pitch / 16 << 9 |
        //        first chroma column << 11 | mvx & 7 << 14 | mvy & 7 << 17
        uint32_t gX = 0, gY = 0, gM = 0, gC = 0;
        if (isInter && (mW0 & 0xFFu) <= B200_MB_P_16x16) {
            const int mvx = (int)(int16_t)(mMv & 0xFFFFu), mvy = (int)(int16_t)(mMv >> 16);
            const WindowGeom wg = windowGeom(g.W, g.H, mbx * 16, (row0 + lane) * 16, 16, 16, mvx, mvy);
            gX = (uint32_t)wg.strip | ((uint32_t)wg.stripC << 16);
            gY = (uint32_t)wg.row | ((uint32_t)wg.rowC << 16);
            gM = (uint32_t)wg.map | ((uint32_t)wg.mapC << 3) | (wg.bytes << 8);
            const uint32_t nx = (wg.geom >> 4) & 3u, nxC = (wg.geom >> 12) & 3u;
            gC = ((wg.geom & 15u) + ((mvy & 3) ? 32u * nx : 0u) + ((mvx & 3) ? 2u : 0u)) | (nx << 7) | (nxC << 9) | (((wg.geom >> 8) & 7u) << 11) |
                 ((uint32_t)(mvx & 7) << 14) | ((uint32_t)(mvy & 7) << 17);
        }

        // A macroblock that is computed leaves what its turn will need in ITS staging slot (copies never touch it): the lanes
        // that compute fetch it with one broadcast load, lane 0 a second one with what the loads need
        if (isInter) {
            uint4 *box = reinterpret_cast<uint4 *>(sm.stage + lane * 256);
            box[0] = make_uint4(mW0, mMask, gC, 0u);
            box[1] = make_uint4(gX, gY, gM | ((mRef & 0xFFu) << 24), mCoef);
        }
        __syncwarp();

        // ---- what is staged for the macroblock whose turn comes next ----------------------------------------------------
        uint32_t nW0 = 0, nMask = 0, nGeom = 0;
        int nL = 0;
        // stage macroblock l of the chunk into buffer `buf`: its levels and its windows, issued by lane 0 from uniform registers
        // (issued by the record's own lane, a lane the compiler cannot name, every load becomes a loop that looks for the lane)
        auto prepare = [&](int l, int buf) {
            nL = l;
            const uint4 *box = reinterpret_cast<const uint4 *>(sm.stage + l * 256);
            const uint4 a = box[0];
            nW0 = a.x; nMask = a.y; nGeom = a.z;
            if (lane == 0) {
                const uint4 q = box[1];
                const uint32_t type = nW0 & 0xFFu;
                const uint32_t coefBytes = type == B200_MB_I_PCM ? 384u : 32u * (uint32_t)__popc(nMask & 0x3FFFFFFu);
                mbarExpectTx(&sm.mbar[buf], ((q.z >> 8) & 0xFFFFu) + coefBytes);
                if (type != B200_MB_I_PCM) {
                    const int refFrame = (int)(frameBase + (q.z >> 24));
                    tmaLoad4d(sm.luma[buf], &maps.luma[0][0] + (q.z & 7u), 0, (int)(q.x & 0xFFFFu), (int)(q.y & 0xFFFFu), refFrame, &sm.mbar[buf]);
                    tmaLoad4d(sm.chroma[buf], &maps.chroma[0][0] + ((q.z >> 3) & 3u), 0, (int)(q.x >> 16), (int)(q.y >> 16), refFrame, &sm.mbar[buf]);
                }
                if (coefBytes) bulkLoad(sm.coef[buf], job.coefs + (size_t)q.w * 16, coefBytes, &sm.mbar[buf]);
            }
        };
        if (interMask) prepare(__ffs(interMask) - 1, 0);

        // ---- inter macroblocks, one after the other: the next one is staged while this one is computed --------------------------
        int it = 0;
#pragma unroll 1
        while (interMask) {
            interMask &= interMask - 1;
            const int buf = it & 1;
            it++;
            const int l = nL;
            const uint32_t w0 = nW0, mask = nMask, geom = nGeom;
            const uint32_t type = w0 & 0xFFu;
            if (interMask) prepare(__ffs(interMask) - 1, buf ^ 1);
            mbarWait(&sm.mbar[buf], (phaseBits >> buf) & 1u);
            phaseBits ^= 1u << buf;
            if (type == B200_MB_I_PCM) {
                // h264bsdWriteMacroblock (image.c:81-144): 384 raw bytes, 256 Y then 64 Cb then 64 Cr
                const uint8_t *src = sm.coef[buf];
                *reinterpret_cast<uint2 *>(lbase + l * 256 + lane * 8) = *reinterpret_cast<const uint2 *>(src + lane * 8);
                *reinterpret_cast<uint32_t *>(cbase + l * 128 + lane * 4) = *reinterpret_cast<const uint32_t *>(src + 256 + cp * 64 + cr * 8 + cc);
                __syncwarp();
                continue;
            }
            // prediction first: the centre positions borrow the residual arrays
            const int cxf = (int)((geom >> 14) & 7u), cyf = (int)((geom >> 17) & 7u), xf = cxf & 3, yf = cyf & 3;
            const int pitch = (int)((geom >> 3) & 0x30u), pitchC = (int)((geom >> 5) & 0x30u);
            const uint8_t *G0 = sm.luma[buf] + (geom & 127u);
            const bool centre = (xf == 2 || yf == 2) && xf != 0 && yf != 0;
            const uint2 pv = centre ? lumaCentre8(&sm.resY[0][0], G0, pitch, lane, xf, yf) : lumaQpel8(G0, pitch, c8, r8, xf, yf);
            const uint32_t pc = chromaPred4(sm.chroma[buf], pitchC, (int)((geom >> 11) & 7u), cp, cc, cr, cxf, cyf);
            if (mask) {
                __syncwarp();
                residualShfl(sm, sm.coef[buf], mask, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF, lane, p.errors);
            }
            addResidualStore(sm, mask, pv, pc, lbase + l * 256, cbase + l * 128, lane);
            __syncwarp();
        }

        // ---- the copies have arrived: on to the current frame ------------------------------------------------------------------
        if (copyMask) {
            mbarWaitSleeping(&sm.mbarCopy, (phaseBits >> 2) & 1u);
            phaseBits ^= 4u;
            if (runLen) {
                bulkStore(lbase + lane * 256, sm.stage + lane * 256, 256u * (uint32_t)runLen);
                bulkStore(cbase + lane * 128, sm.stage + kStageMbs * 256 + lane * 128, 128u * (uint32_t)runLen);
                bulkCommit();
            }
        }
        chunk = nextChunk;
        nextChunk = __shfl_sync(0xffffffffu, ticket2, 0);
    }
    bulkWaitAll();   // this lane's last stores still read shared memory
}

// Second instance: the macroblocks with several partitions (inter_prediction.c:361-482), from the list the first instance made.
// A warp takes eight list entries at a time (lane e < 8 fetches entry e's record, one batch ahead) and works through them as a
// pipeline of ROUNDS: a round is two partitions that are at least 8 wide -- the two of a 16x8 / 8x16 macroblock, the upper or the
// lower two 8x8 sub-macroblocks -- so that every lane's 8-sample luma span and 4-sample chroma span lie inside ONE of them and
// the lane only has to pick that partition's window, vector and origin.  The lanes that hold the records work out, all at once,
// what every round of the batch needs -- where its windows lie, what to expect, what to compute with -- and leave it in a box in
// shared memory; after that a round costs the warp a few broadcast loads, and lane 0 issues its loads from uniform registers.
// The four windows of a round land in one of two buffer pairs (the second one lives in `stage`: this instance has no copies) on
// the pair's mbarrier, together with the macroblock's levels in its first round; the round after is staged before this one is
// waited for.  A sub-macroblock with 8x4 / 4x8 / 4x4 partitions makes its macroblock one round without windows: its partitions
// are fetched one by one when its turn has come.
constexpr int kMultiBatch = 8;
struct __align__(16) RoundBox {
    uint32_t gxA, gyA, gxB, gyB;            // luma strip | chroma strip << 16, luma row | chroma row << 16 of the two windows
    uint32_t maps, refA, refB, bytes;       // luma map A | chroma map A << 4 | luma B << 8 | chroma B << 12; reference frames; bytes to expect
    uint32_t coefLo, coefHi, coefBytes, flags;   // the macroblock's levels (first round); flags: round | last << 1 | small partitions << 2 | level buffer << 3
    uint32_t geomAB, fracs, w0, mask;       // WindowGeom.geom of A | B << 16; mvA.x & 7 | mvA.y & 7 << 3 | mvB.x & 7 << 6 | mvB.y & 7 << 9; record head
    uint32_t pos, dstLo, dstHi, dstC;       // mbx | mby << 16; the macroblock's luma in the current frame, its chroma relative to that
                                            // (small partitions: gxA, gyA = the record, where they read their vectors)
};
static_assert(sizeof(RoundBox) == 80 && 384 + 2 * kLumaBufBytes + 2 * kChromaBufBytes + 2 * kMultiBatch * sizeof(RoundBox) <= sizeof(MultiWarpSmem::stage),
              "the second window pair, the prediction of small partitions and the round boxes share `stage`");
__global__ void __launch_bounds__(kPassAWarps * 32, 5)
passAMultiKernel(const ReconParams p, const __grid_constant__ PassAMaps maps) {
    extern __shared__ __align__(128) uint8_t interSmemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    MultiWarpSmem &sm = reinterpret_cast<MultiWarpSmem *>(interSmemRaw)[warp];
    if (lane == 0) {
        mbarInit(&sm.mbar[0], 1);
        mbarInit(&sm.mbar[1], 1);
        fenceMbarInit();
    }
    __syncwarp();
    uint32_t phaseBits = 0;
    const uint32_t nWarps = gridDim.x * kPassAWarps;
    const uint32_t count = __ldcg(p.multiCount), nBatches = (count + kMultiBatch - 1) / kMultiBatch;
    const int r8 = lane >> 1, c8 = (lane & 1) * 8;
    const int cr = lane >> 2, cp = (lane >> 1) & 1, cc = (lane & 1) * 4;
    uint8_t *pred = sm.stage;
    RoundBox *boxes = reinterpret_cast<RoundBox *>(sm.stage + 384 + 2 * kLumaBufBytes + 2 * kChromaBufBytes);
    auto winL = [&](int pair, int w) -> uint8_t * { return pair ? sm.stage + 384 + w * kLumaBufBytes : sm.luma[w]; };
    auto winC = [&](int pair, int w) -> uint8_t * { return pair ? sm.stage + 384 + 2 * kLumaBufBytes + w * kChromaBufBytes : sm.chroma[w]; };

    // lane e < 8: entry e of a batch -- where the macroblock lies, its record's head, reference slots, first vector of each quadrant
    uint32_t fEntry = 0xFFFFFFFFu, fW0 = 0, fMask = 0, fCoef = 0, fW3 = 0, fRef = 0, fMv = 0, fMv1 = 0, fMv2 = 0, fMv3 = 0;
    auto fetch = [&](uint32_t b) {
        fEntry = 0xFFFFFFFFu;
        const uint32_t e = b * kMultiBatch + (uint32_t)lane;
        if (b >= nBatches || lane >= kMultiBatch || e >= count) return;
        fEntry = __ldcg(p.multiList + e);
        const uint32_t s = fEntry >> 16, mb = fEntry & 0xFFFFu;
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(p.jobs[s].recs + mb);
        const uint4 hw = __ldg(reinterpret_cast<const uint4 *>(rw));
        fRef = __ldg(rw + 4);
        fMv = __ldg(rw + 8); fMv1 = __ldg(rw + 12); fMv2 = __ldg(rw + 16); fMv3 = __ldg(rw + 20);
        fW0 = hw.x; fMask = hw.y; fCoef = hw.z; fW3 = hw.w;
    };
    uint32_t batch = blockIdx.x * kPassAWarps + warp;
    fetch(batch);
    uint32_t nextBatch = 0;
    if (lane == 0) nextBatch = atomicAdd(p.ticketA + 1, 1u) + nWarps;
    nextBatch = __shfl_sync(0xffffffffu, nextBatch, 0);
    while (batch < nBatches) {
        uint32_t ticket2 = 0;
        if (lane == 0) ticket2 = atomicAdd(p.ticketA + 1, 1u) + nWarps;
        const uint32_t mEntry = fEntry, mW0 = fW0, mMask = fMask, mCoef = fCoef, mW3 = fW3, mRef = fRef;
        const uint32_t mMv = fMv, mMv1 = fMv1, mMv2 = fMv2, mMv3 = fMv3;
        fetch(nextBatch);

        // ---- the lanes that hold a record fill the boxes of its rounds --------------------------------------------------------
        const bool mine = mEntry != 0xFFFFFFFFu;
        const uint32_t type = mW0 & 0xFFu, subTypes = type >= B200_MB_P_8x8 ? mW3 >> 24 : 0u;
        const bool small = mine && subTypes != 0, two = mine && type >= B200_MB_P_8x8 && subTypes == 0;
        const uint32_t mineMask = __ballot_sync(0xffffffffu, mine), twoMask = __ballot_sync(0xffffffffu, two);
        const uint32_t below = (1u << lane) - 1u;
        const int nRounds = __popc(mineMask) + __popc(twoMask);
        if (mine) {
            const int seq = __popc(mineMask & below), firstRound = seq + __popc(twoMask & below);
            const uint32_t s = mEntry >> 16, mb = mEntry & 0xFFFFu;
            const StreamJob job = p.jobs[s];
            const int mby = mbRowOf(mb, g), mbx = (int)mb - mby * g.widthMbs;
            const uint32_t frameBase = s * (uint32_t)g.numSlots;
            const int16_t *coefSrc = job.coefs + (size_t)mCoef * 16;
            const uint32_t coefBytes = 32u * (uint32_t)__popc(mMask & 0x3FFFFFFu);
            const unsigned long long recBits = (unsigned long long)reinterpret_cast<uintptr_t>(job.recs + mb);
#pragma unroll 1
            for (int rd = 0; rd < (two ? 2 : 1); rd++) {
                RoundBox bx;
                bx.coefLo = (uint32_t)reinterpret_cast<uintptr_t>(coefSrc); bx.coefHi = (uint32_t)((unsigned long long)reinterpret_cast<uintptr_t>(coefSrc) >> 32);
                bx.coefBytes = rd == 0 ? coefBytes : 0u;
                bx.flags = (uint32_t)rd | ((!two || rd == 1) ? 2u : 0u) | (small ? 4u : 0u) | ((uint32_t)(seq & 1) << 3);
                bx.w0 = mW0; bx.mask = mMask;
                bx.pos = (uint32_t)mbx | ((uint32_t)mby << 16);
                {
                    uint8_t *frame = framePtr(p.pool, g, frameBase + job.curSlot);
                    uint8_t *dy = mbLuma(frame, g, mbx, mby);
                    bx.dstLo = (uint32_t)reinterpret_cast<uintptr_t>(dy); bx.dstHi = (uint32_t)((unsigned long long)reinterpret_cast<uintptr_t>(dy) >> 32);
                    bx.dstC = (uint32_t)(mbChroma(frame, g, mbx, mby) - dy);
                }
                if (!small) {
                    int qA, qB, pxB, pyA, pyB, pw, ph;
                    if (type == B200_MB_P_16x8) { qA = 0; qB = 2; pxB = 0; pyA = 0; pyB = 8; pw = 16; ph = 8; }
                    else if (type == B200_MB_P_8x16) { qA = 0; qB = 1; pxB = 8; pyA = 0; pyB = 0; pw = 8; ph = 16; }
                    else { qA = 2 * rd; qB = 2 * rd + 1; pxB = 8; pyA = pyB = 8 * rd; pw = 8; ph = 8; }
                    const uint32_t mvA = qA ? mMv2 : mMv, mvB = qB == 1 ? mMv1 : qB == 2 ? mMv2 : mMv3;
                    const WindowGeom wa = windowGeom(g.W, g.H, mbx * 16, mby * 16 + pyA, pw, ph, (int)(int16_t)(mvA & 0xFFFFu), (int)(int16_t)(mvA >> 16));
                    const WindowGeom wb = windowGeom(g.W, g.H, mbx * 16 + pxB, mby * 16 + pyB, pw, ph, (int)(int16_t)(mvB & 0xFFFFu), (int)(int16_t)(mvB >> 16));
                    bx.gxA = (uint32_t)wa.strip | ((uint32_t)wa.stripC << 16); bx.gyA = (uint32_t)wa.row | ((uint32_t)wa.rowC << 16);
                    bx.gxB = (uint32_t)wb.strip | ((uint32_t)wb.stripC << 16); bx.gyB = (uint32_t)wb.row | ((uint32_t)wb.rowC << 16);
                    bx.maps = (uint32_t)wa.map | ((uint32_t)wa.mapC << 4) | ((uint32_t)wb.map << 8) | ((uint32_t)wb.mapC << 12);
                    bx.refA = frameBase + ((mRef >> (8 * qA)) & 0xFFu); bx.refB = frameBase + ((mRef >> (8 * qB)) & 0xFFu);
                    bx.bytes = wa.bytes + wb.bytes + bx.coefBytes;
                    bx.geomAB = wa.geom | (wb.geom << 16);
                    bx.fracs = (mvA & 7u) | (((mvA >> 16) & 7u) << 3) | ((mvB & 7u) << 6) | (((mvB >> 16) & 7u) << 9);
                } else {
                    bx.gxA = (uint32_t)recBits; bx.gyA = (uint32_t)(recBits >> 32); bx.gxB = bx.gyB = bx.maps = 0;
                    bx.refA = mRef; bx.refB = frameBase;   // (the partitions' reference slots and the stream's first frame)
                    bx.bytes = bx.coefBytes;
                    bx.geomAB = 0; bx.fracs = subTypes;
                }
                boxes[firstRound + rd] = bx;
            }
        }
        __syncwarp();

        // ---- the rounds: stage the next one, compute this one -----------------------------------------------------------------
        auto stageRound = [&](int r) {
            const uint4 q0 = reinterpret_cast<const uint4 *>(&boxes[r])[0], q1 = reinterpret_cast<const uint4 *>(&boxes[r])[1];
            const uint4 q2 = reinterpret_cast<const uint4 *>(&boxes[r])[2];
            if (lane == 0 && q1.w) {
                const int sp = r & 1;
                uint64_t *bar = &sm.mbar[sp];
                mbarExpectTx(bar, q1.w);
                if (!(q2.w & 4u)) {
                    tmaLoad4d(winL(sp, 0), &maps.luma[0][0] + (q1.x & 15u), 0, (int)(q0.x & 0xFFFFu), (int)(q0.y & 0xFFFFu), (int)q1.y, bar);
                    tmaLoad4d(winC(sp, 0), &maps.chroma[0][0] + ((q1.x >> 4) & 15u), 0, (int)(q0.x >> 16), (int)(q0.y >> 16), (int)q1.y, bar);
                    tmaLoad4d(winL(sp, 1), &maps.luma[0][0] + ((q1.x >> 8) & 15u), 0, (int)(q0.z & 0xFFFFu), (int)(q0.w & 0xFFFFu), (int)q1.z, bar);
                    tmaLoad4d(winC(sp, 1), &maps.chroma[0][0] + ((q1.x >> 12) & 15u), 0, (int)(q0.z >> 16), (int)(q0.w >> 16), (int)q1.z, bar);
                }
                if (q2.z) bulkLoad(sm.coef[(q2.w >> 3) & 1u], reinterpret_cast<const void *>((uintptr_t)(((unsigned long long)q2.y << 32) | q2.x)), q2.z, bar);
            }
        };
        if (nRounds) stageRound(0);
        uint2 pv = make_uint2(0, 0);   // this lane's 8 luma prediction samples
        uint32_t pc = 0;               // and 4 chroma prediction samples
#pragma unroll 1
        for (int r = 0; r < nRounds; r++) {
            if (r + 1 < nRounds) stageRound(r + 1);
            const uint4 q2 = reinterpret_cast<const uint4 *>(&boxes[r])[2], q3 = reinterpret_cast<const uint4 *>(&boxes[r])[3];
            const uint4 q4 = reinterpret_cast<const uint4 *>(&boxes[r])[4];
            const int pair = r & 1, rd = (int)(q2.w & 1u);
            const uint32_t w0 = q3.z, mask = q3.w, geomAB = q3.x, fracs = q3.y;
            const uint32_t type = w0 & 0xFFu;
            const int mbx = (int)(q4.x & 0xFFFFu), mby = (int)(q4.x >> 16);
            if (reinterpret_cast<const uint4 *>(&boxes[r])[1].w) {   // (armed: something was asked for)
                mbarWait(&sm.mbar[pair], (phaseBits >> pair) & 1u);
                phaseBits ^= 1u << pair;
            }
            if (rd == 0) {
                pv = make_uint2(0, 0);
                pc = 0;
                if (mask) residualShfl(sm, sm.coef[(q2.w >> 3) & 1u], mask, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF, lane, p.errors);
            }
            if (!(q2.w & 4u)) {
                int pxB, pyA, pyB;
                if (type == B200_MB_P_16x8) { pxB = 0; pyA = 0; pyB = 8; }
                else if (type == B200_MB_P_8x16) { pxB = 8; pyA = 0; pyB = 0; }
                else { pxB = 8; pyA = pyB = 8 * rd; }
                const bool inBL = type == B200_MB_P_16x8 ? r8 >= 8 : c8 == 8;
                const bool inBC = type == B200_MB_P_16x8 ? cr >= 4 : cc == 4;
                const bool actL = type < B200_MB_P_8x8 || (r8 >> 3) == rd, actC = type < B200_MB_P_8x8 || (cr >> 2) == rd;
                if (actL) {
                    const uint32_t gm = inBL ? geomAB >> 16 : geomAB & 0xFFFFu, fr = inBL ? fracs >> 6 : fracs;
                    const int xf = (int)(fr & 3u), yf = (int)((fr >> 3) & 3u);
                    const int pitch = (int)((gm >> 4) & 3u) * 16;
                    const uint8_t *G0 = winL(pair, inBL ? 1 : 0) + (gm & 15u) + (yf ? 2 * pitch : 0) + (xf ? 2 : 0);
                    pv = lumaQpel8(G0, pitch, c8 - (inBL ? pxB : 0), r8 - (inBL ? pyB : pyA), xf, yf);
                }
                if (actC) {
                    const uint32_t gm = inBC ? geomAB >> 16 : geomAB & 0xFFFFu, fr = inBC ? fracs >> 6 : fracs;
                    pc = chromaPred4(winC(pair, inBC ? 1 : 0), (int)((gm >> 12) & 3u) * 16, (int)((gm >> 8) & 7u), cp,
                                     cc - (inBC ? (pxB >> 1) : 0), cr - ((inBC ? pyB : pyA) >> 1), (int)(fr & 7u), (int)((fr >> 3) & 7u));
                }
            } else {
                // sub-macroblocks with 8x4 / 4x8 / 4x4 partitions: one window per partition, sample by sample
                const uint4 q0 = reinterpret_cast<const uint4 *>(&boxes[r])[0];
                const uint32_t *rw = reinterpret_cast<const uint32_t *>((uintptr_t)(((unsigned long long)q0.y << 32) | q0.x));
                const uint4 q1 = reinterpret_cast<const uint4 *>(&boxes[r])[1];
                const uint32_t refSlots = q1.y, frameBase = q1.z, subTypes = fracs;
                uint8_t *wl = winL(pair, 0), *wc = winC(pair, 0);
#pragma unroll 1
                for (int pi = 0; pi < 16; pi++) {
                    int pw, ph;
                    const int sub = (subTypes >> (2 * (pi >> 2))) & 3, j = pi & 3;
                    if (sub == 0) { if (j) continue; pw = 8; ph = 8; }
                    else if (sub == 1) { if (j & 1) continue; pw = 8; ph = 4; }
                    else if (sub == 2) { if (j & 2) continue; pw = 4; ph = 8; }
                    else { pw = 4; ph = 4; }
                    const int px = cBlkX[pi] * 4, py = cBlkY[pi] * 4;
                    const uint32_t mvw = __ldg(rw + 8 + pi);
                    const int mvx = (int)(int16_t)(mvw & 0xFFFFu), mvy = (int)(int16_t)(mvw >> 16);
                    const uint32_t gm = issueWindowFn(wl, wc, &sm.mbar[pair], &maps, g.W, g.H, mbx * 16 + px, mby * 16 + py, pw | (ph << 8), mvx, mvy,
                                                      frameBase + ((refSlots >> (8 * (pi >> 2))) & 0xFFu), lane);
                    mbarWait(&sm.mbar[pair], (phaseBits >> pair) & 1u);
                    phaseBits ^= 1u << pair;
                    const int xf = mvx & 3, yf = mvy & 3, pitch = (int)((gm >> 4) & 3u) * 16;
                    const uint8_t *G0 = wl + (gm & 15u) + (yf ? 2 * pitch : 0) + (xf ? 2 : 0);
                    const int lw = 31 - __clz(pw);
#pragma unroll 1
                    for (int q = lane; q < pw * ph; q += 32) {
                        const int x = q & (pw - 1), y = q >> lw;
                        pred[(py + y) * 16 + px + x] = (uint8_t)lumaQpel(G0, pitch, x, y, xf, yf);
                    }
                    const int cw = pw >> 1, chh = ph >> 1, ncp = cw * chh, lcw = lw - 1;
                    const int cxf = mvx & 7, cyf = mvy & 7, pitchC = (int)((gm >> 12) & 3u) * 16, cxo = (int)((gm >> 8) & 7u);
#pragma unroll 1
                    for (int q = lane; q < 2 * ncp; q += 32) {
                        const int pl = q >= ncp, qq = q - pl * ncp;
                        const int x = qq & (cw - 1), y = qq >> lcw;
                        auto S = [&](int sx, int sy) -> int {
                            const int col = cxo + sx;
                            return wc[sy * pitchC + (col >> 3) * 16 + pl * 8 + (col & 7)];
                        };
                        // (a sample that meets a zero weight may lie outside the box: it is read, not used)
                        const int A = S(x, y), B = S(x + 1, y), Cc = S(x, y + 1), D = S(x + 1, y + 1);
                        pred[256 + pl * 64 + ((py >> 1) + y) * 8 + (px >> 1) + x] =
                            (uint8_t)(((8 - cxf) * (8 - cyf) * A + cxf * (8 - cyf) * B + (8 - cxf) * cyf * Cc + cxf * cyf * D + 32) >> 6);
                    }
                    __syncwarp();
                }
                pv = *reinterpret_cast<const uint2 *>(pred + r8 * 16 + c8);
                pc = *reinterpret_cast<const uint32_t *>(pred + 256 + cp * 64 + cr * 8 + cc);
            }
            if (q2.w & 2u) {
                uint8_t *dy = reinterpret_cast<uint8_t *>((uintptr_t)(((unsigned long long)q4.z << 32) | q4.y));
                addResidualStore(sm, mask, pv, pc, dy, dy + q4.w, lane);
            }
            __syncwarp();   // the pair's windows (and the residual) are free for the loads of the turn after next
        }
        batch = nextBatch;
        nextBatch = __shfl_sync(0xffffffffu, ticket2, 0);
    }
}

// =====================================================================================================
// pass B: intra-predicted macroblocks.  They read the unfiltered current picture, so an intra macroblock
// must come after its intra neighbours (the others were written by pass A).  A warp owns one macroblock row of one
// stream: it reads the row's record heads 32 at a time, picks the intra-predicted ones and works through them left to right,
// so the left neighbour is its own previous step; rows publish their progress in a per-row counter and a macroblock whose
// upper neighbours are intra-predicted waits until the row above is past them.  Tickets are handed out row-major (row y of
// every stream before row y + 1 of any), so the row above is normally through long before and a warp only ever waits for rows
// with earlier tickets.
// =====================================================================================================
__global__ void __launch_bounds__(kReconWarps * 32, B200_INTRA_MINBLOCKS) reconIntraKernel(const ReconParams p) {
    __shared__ IntraWarpSmem smemAll[kReconWarps];
    // Intra4x4 tables, made once per CTA.  sI4Off[c][mode * 16 + sample]: gIntra4x4Table with the edge indices turned into byte
    // offsets from the block's corner sample in the tile (left column (4 - i) * 24, corner 0, above row 1..8); c = 0 is for
    // blocks whose above-right block is not available: above-row samples 4..7 fall back to sample 3 (intra_prediction.c:762-767).
    // sI4Step[half][step]: the block a half-warp predicts in that step of the ten-step schedule, 31 = none, as
    //   block | x4 << 5 | y4 << 7 | aboveRight << 9   (aboveRight: 0 no, 1 yes, 2 = the macroblock's B, 3 = its C neighbour)
    __shared__ uint32_t sI4Off[2][9 * 16];
    __shared__ uint32_t sI4Step[2][10];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    for (int i = threadIdx.x; i < 2 * 9 * 16; i += blockDim.x) {
        const int c = i / (9 * 16);
        const uint32_t d = gIntra4x4Table[i - c * 9 * 16];
        uint32_t o = d & 0xFF000000u;
        for (int k = 0; k < 3; k++) {
            const int e = (d >> (8 * k)) & 0xFF;
            int off;
            if (e <= 3) off = (4 - e) * 24;
            else { off = e - 4; if (off > 4 && !c) off = 4; }
            o |= (uint32_t)off << (8 * k);
        }
        sI4Off[c][i - c * 9 * 16] = o;
    }
    if (threadIdx.x < 20) {
        const int hf = threadIdx.x / 10, st = threadIdx.x - hf * 10;
        const int b = hf ? cI4StepB[st] : cI4StepA[st];
        uint32_t v = 31;
        if (b >= 0) {
            const int bx = cBlkX[b], by = cBlkY[b];
            // above-right block: the neighbouring macroblock's, or decoded earlier inside this one (intra_prediction.c:730-767)
            const int ar = by == 0 ? (bx == 3 ? 3 : 2) : bx == 3 ? 0 : (cRasterToBlk[(by - 1) * 4 + bx + 1] < b ? 1 : 0);
            v = (uint32_t)b | ((uint32_t)bx << 5) | ((uint32_t)by << 7) | ((uint32_t)ar << 9);
        }
        sI4Step[hf][st] = v;
    }
    __syncthreads();
    IntraWarpSmem &sm = smemAll[warp];
    const int W = g.widthMbs;
    const uint32_t serial16 = p.serial & 0xFFFFu, totalRows = (uint32_t)g.heightMbs * (uint32_t)g.nStreams;
    // this lane's spans inside a macroblock, as in pass A: 8 luma samples = bytes 8 lane.. of the 256, 4 chroma samples =
    // bytes 4 lane.. of the 128
    const int r8 = lane >> 1, c8 = (lane & 1) * 8;
    const int cr = lane >> 2, cp = (lane >> 1) & 1, cc = (lane & 1) * 4;
    // tickets in order, taken when a warp is ready for the row (a ticket taken ahead of time would let later rows start -- and
    // spin -- before this one; measured)
    for (;;) {
    uint32_t thisTicket = 0;
    if (lane == 0) thisTicket = atomicAdd(p.ticket, 1u);
    thisTicket = __shfl_sync(0xffffffffu, thisTicket, 0);
    if (thisTicket >= totalRows) break;
    const int mby = (int)(thisTicket / (uint32_t)g.nStreams);
    const uint32_t s = thisTicket - (uint32_t)mby * (uint32_t)g.nStreams;
    const StreamJob job = p.jobs[s];
    if (job.nB) {
    uint32_t *rowMine = p.done + (size_t)s * g.heightMbs + mby;
    uint8_t *cur = framePtr(p.pool, g, s * (uint32_t)g.numSlots + job.curSlot);
    const b200_mb_rec *recsRow = job.recs + (size_t)mby * W;
    uint32_t seen = 0;      // macroblocks of the row above known to be through
    bool any = false;       // this row has had an intra-predicted macroblock (only such rows are ever waited for)
#pragma unroll 1
    for (int x0 = 0; x0 < W; x0 += 32) {
    // lane i reads the head of record x0 + i: which macroblocks of this stretch are intra-predicted
    uint32_t mMisc = 0;
    uint4 mHead = make_uint4(0, 0, 0, 0);
    bool isIntra = false;
    if (x0 + lane < W) {
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(recsRow + x0 + lane);
        mHead = __ldg(reinterpret_cast<const uint4 *>(rw));
        const uint32_t type = mHead.x & 0xFFu;
        // (a concealed macroblock carries Intra4x4 for the filter's sake; its pels come from pass A or concealKernel)
        isIntra = type > B200_MB_P_8x8REF0 && type != B200_MB_I_PCM && !((mHead.x >> 24) & B200_MBF_CONCEALED);
        if (isIntra) mMisc = (__ldg(rw + 5) & 0xFF) | ((__ldg(rw + 7) & 0xFF) << 8);
    }
    uint32_t todo = __ballot_sync(0xffffffffu, isIntra);
    any = any || todo != 0;
#pragma unroll 1
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const int mbx = x0 + i;
        const b200_mb_rec *rec = recsRow + mbx;
        MbHead h;
        {
            const uint32_t hx = __shfl_sync(0xffffffffu, mHead.x, i);
            h.mbType = hx & 0xFF; h.qpY = (hx >> 8) & 0xFF; h.qpC = (hx >> 16) & 0xFF; h.flags = hx >> 24;
            h.mask = __shfl_sync(0xffffffffu, mHead.y, i);
            h.coefIndex = __shfl_sync(0xffffffffu, mHead.z, i);
        }
        const uint32_t misc = __shfl_sync(0xffffffffu, mMisc, i);   // intraChromaMode | waitMask << 8
        const int16_t *coef = job.coefs + (size_t)h.coefIndex * 16;
        uint8_t *dstY = mbLuma(cur, g, mbx, mby) + lane * 8;
        uint8_t *dstC = mbChroma(cur, g, mbx, mby) + lane * 4;
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
        // wait for the intra neighbours this macroblock reads (record byte 28: waitMask).  The left one is this warp's own
        // previous step; the row above has to be past the above-left (x), above (x + 1) or above-right (x + 2) macroblock
        const int waitMask = (misc >> 8) & 0xFF;
        // Intra4x4: prediction modes of the sixteen blocks (record words 8..11), on their way while the row above is waited for
        uint32_t modeWord = 0;
        if (lane < 4 && h.mbType == B200_MB_I_4x4) modeWord = __ldg(reinterpret_cast<const uint32_t *>(rec) + 8 + lane);
        {
            const uint32_t need = (uint32_t)min(W, (waitMask & B200_MBF_AVAIL_C) ? mbx + 2 : (waitMask & B200_MBF_AVAIL_B) ? mbx + 1
                                                   : (waitMask & B200_MBF_AVAIL_D) ? mbx : 0);
            if (lane == 0 && seen < need) seen = waitRow(rowMine - 1, serial16, need);
            seen = __shfl_sync(0xffffffffu, seen, 0);
        }
        __syncwarp();
        // neighbouring pels (h264bsdGetNeighbourPels :545-614), straight from L2
        if (lane < 4) reinterpret_cast<uint32_t *>(sm.stage)[lane] = modeWord;
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
            // h264bsdIntra4x4Prediction (:701-833): half-warp = block, lane = sample; edge samples straight from the tile.  The
            // sixteen prediction modes (record words 8..11) were fetched with the neighbouring pels and lie in sm.stage
            const int smp = lane & 15, myOff = (1 + (smp >> 2)) * 24 + 1 + (smp & 3);   // this lane's sample from the block's corner
#pragma unroll 1
            for (int st = 0; st < 10; st++) {
                const uint32_t si = sI4Step[lane >> 4][st];
                const int b = si & 31;
                if (b != 31) {
                    const int bx = (si >> 5) & 3, by = (si >> 7) & 3, ar = (si >> 9) & 3;
                    const int mode = sm.stage[b];
                    uint8_t *corner = &sm.itY[by * 4][bx * 4];   // corner sample; above row to its right, left column below it
                    int v;
                    if (mode == 2) {
                        const bool bA = bx ? true : avA, bB = by ? true : avB;
                        const int sa = corner[1] + corner[2] + corner[3] + corner[4];
                        const int sl = corner[24] + corner[48] + corner[72] + corner[96];
                        v = (bA && bB) ? (sa + sl + 4) >> 3 : bA ? (sl + 2) >> 2 : bB ? (sa + 2) >> 2 : 128;
                    } else {
                        const bool bC = ar == 1 || (ar == 2 && avB) || (ar == 3 && avC);
                        const uint32_t d = sI4Off[bC ? 1 : 0][mode * 16 + smp];
                        const int e0 = corner[d & 0xFF], e1 = corner[(d >> 8) & 0xFF], e2 = corner[(d >> 16) & 0xFF];
                        const uint32_t kind = d >> 24;
                        v = kind == 0 ? e0 : kind == 1 ? (e0 + e1 + 1) >> 1 : (e0 + 2 * e1 + e2 + 2) >> 2;
                    }
                    if (h.mask) v = clip255(v + sm.res[b][smp]);
                    // the sample lies inside the block, every edge sample outside it, and the two blocks of a step do not
                    // touch each other's edges: no barrier between the reads above and this write
                    corner[myOff] = (uint8_t)v;
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
        // publish the row's progress -- everything before its next intra macroblock (or the end of this stretch) is through.  The
        // warp barrier orders every lane's stores before lane 0's release at gpu scope (cumulative; no extra fence)
        __syncwarp();
        if (lane == 0) stRelease(rowMine, (serial16 << 16) | (uint32_t)min(W, todo ? x0 + __ffs(todo) - 1 : x0 + 32));
    }
    if (any && lane == 0 && x0 + 32 >= W) stRelease(rowMine, (serial16 << 16) | (uint32_t)W);   // (a row that ends without one)
    }  // stretch of 32 macroblocks
    }  // a stream with intra-predicted macroblocks in this picture
    }  // row tickets
}

}  // namespace b200
