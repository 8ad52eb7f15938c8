// deblock_kernel.cuh -- in-loop deblocking filter (h264bsdFilterPicture, h264bsd_deblocking.c:575-640)
// as a dependency-flag wavefront: one warp per macroblock, tickets in x+2y order; a macroblock waits
// for its left, top and top-right neighbours (exactly the macroblocks whose filtering the reference's
// raster order puts before it and whose pels it reads or rewrites).
//
// Also here: replication of the picture border (what h264bsdFillBlock's coordinate clamp computes on
// the fly, reconstruct.c:2244-2367), YUV -> ARGB (h264bsd_decoder.c:1163-1370) and a frame compare.
#pragma once
#include "device_common.cuh"

namespace b200 {

constexpr int kDeblockWarps = 8;
constexpr int kBsChunk = 8;      // stage 1: consecutive macroblocks per warp
constexpr int kFilterChunk = 4;  // stage 2: consecutive tickets per warp (different streams)

struct DeblockParams {
    uint8_t *pool;
    PoolGeom g;
    const StreamJob *jobs;
    const uint16_t *order;
    uint32_t *done;
    uint32_t *ticket;
    uint32_t serial;
    uint32_t totalTickets;
    uint32_t *bsWords;         // nStreams * nMbs * 4 words: packed boundary strengths (stage 1 -> stage 2)
    uint8_t *work;             // nStreams * nMbs: 1 = the macroblock has a non-zero boundary strength
};

struct __align__(16) DeblockWarpSmem {
    uint8_t y[20][24];      // rows -4..15, cols -4..15 (+4 pad)
    uint8_t c[2][10][12];   // rows -2..7,  cols -4..7
    uint8_t bs[32];         // [0..15] vertical edge left of block (bx,by) at by*4+bx ; [16..31] horizontal edge above it
};

struct EdgeThr { int alpha, beta, idxA; };
__device__ __forceinline__ EdgeThr makeThr(int qp, int offA, int offB) {
    EdgeThr t;
    t.idxA = clip3(0, 51, qp + offA);
    t.alpha = cAlpha[t.idxA];
    t.beta = cBeta[clip3(0, 51, qp + offB)];
    return t;
}

// FilterVerLumaEdge / FilterHorLuma(Edge), deblocking.c:656-965: one line of samples, q0 at *q, p side at -step
__device__ __forceinline__ void filterLumaLine(uint8_t *q, int step, int bS, const EdgeThr &t) {
    const int p0 = q[-step], p1 = q[-2 * step], q0 = q[0], q1 = q[step];
    if (!(abs(p0 - q0) < t.alpha && abs(p1 - p0) < t.beta && abs(q1 - q0) < t.beta)) return;
    const int p2 = q[-3 * step], q2 = q[2 * step];
    if (bS < 4) {
        const int tc = cTc0[t.idxA][bS - 1];
        int tcx = tc;
        if (abs(p2 - p0) < t.beta) { q[-2 * step] = (uint8_t)(p1 + clip3(-tc, tc, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1)); tcx++; }
        if (abs(q2 - q0) < t.beta) { q[step] = (uint8_t)(q1 + clip3(-tc, tc, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1)); tcx++; }
        const int d = clip3(-tcx, tcx, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        q[-step] = (uint8_t)clip255(p0 + d);
        q[0] = (uint8_t)clip255(q0 - d);
    } else {
        const bool strong = abs(p0 - q0) < ((t.alpha >> 2) + 2);
        if (strong && abs(p2 - p0) < t.beta) {
            const int s = p1 + p0 + q0, p3 = q[-4 * step];
            q[-step] = (uint8_t)((p2 + 2 * s + q1 + 4) >> 3);
            q[-2 * step] = (uint8_t)((p2 + s + 2) >> 2);
            q[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + s + 4) >> 3);
        } else {
            q[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        }
        if (strong && abs(q2 - q0) < t.beta) {
            const int s = p0 + q0 + q1, q3 = q[3 * step];
            q[0] = (uint8_t)((p1 + 2 * s + q2 + 4) >> 3);
            q[step] = (uint8_t)((s + q2 + 2) >> 2);
            q[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + s + 4) >> 3);
        } else {
            q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
    }
}
// FilterVerChromaEdge / FilterHorChroma(Edge), deblocking.c:967-1146
__device__ __forceinline__ void filterChromaLine(uint8_t *q, int step, int bS, const EdgeThr &t) {
    const int p0 = q[-step], p1 = q[-2 * step], q0 = q[0], q1 = q[step];
    if (!(abs(p0 - q0) < t.alpha && abs(p1 - p0) < t.beta && abs(q1 - q0) < t.beta)) return;
    if (bS < 4) {
        const int tc = cTc0[t.idxA][bS - 1] + 1;
        const int d = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        q[-step] = (uint8_t)clip255(p0 + d);
        q[0] = (uint8_t)clip255(q0 - d);
    } else {
        q[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

struct RecView {
    const uint32_t *w;
    __device__ __forceinline__ int mbType() const { return __ldg(w) & 0xFF; }
    __device__ __forceinline__ int qpY() const { return (__ldg(w) >> 8) & 0xFF; }
    __device__ __forceinline__ bool intra() const { return mbType() > B200_MB_P_8x8REF0; }
    __device__ __forceinline__ uint32_t coded() const { return __ldg(w + 1); }
    __device__ __forceinline__ int refSlot(int q) const { return (__ldg(w + 4) >> (8 * q)) & 0xFF; }
    __device__ __forceinline__ uint32_t mv(int b) const { return __ldg(w + 8 + b); }
};

// EdgeBoundaryStrength (:395-411) / InnerBoundaryStrength (:332-355) for two non-intra 4x4 blocks
__device__ __forceinline__ int bsPair(const RecView &q, int qb, const RecView &p, int pb) {
    if (((q.coded() >> qb) & 1) || ((p.coded() >> pb) & 1)) return 2;
    const uint32_t mq = q.mv(qb), mp = p.mv(pb);
    const int dx = (int)(int16_t)(mq & 0xFFFF) - (int)(int16_t)(mp & 0xFFFF);
    const int dy = (int)(int16_t)(mq >> 16) - (int)(int16_t)(mp >> 16);
    if (q.refSlot(qb >> 2) != p.refSlot(pb >> 2) || abs(dx) >= 4 || abs(dy) >= 4) return 1;
    return 0;
}

// ---- stage 1: boundary strengths, embarrassingly parallel ----------------------------------------------------
// GetBoundaryStrengths (deblocking.c:1187-1379) for every macroblock: 32 strengths packed as nibbles into 16 bytes
// (word j = segments 8j..8j+7; segments 0..15 = vertical edge left of raster block, 16..31 = horizontal edge above it)
// plus one "has work" byte.  Two thirds of the macroblocks of a typical P picture have nothing to filter.
__global__ void __launch_bounds__(kDeblockWarps * 32) strengthKernel(const DeblockParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    const uint32_t chunksPerStream = ((uint32_t)g.nMbs + kDeblockWarps * kBsChunk - 1) / (kDeblockWarps * kBsChunk);
    const uint32_t total = chunksPerStream * (uint32_t)g.nStreams;
    for (uint32_t v = blockIdx.x; v < total; v += gridDim.x) {
        const uint32_t s = v / chunksPerStream, chunk = v - s * chunksPerStream;
        const StreamJob job = p.jobs[s];
        const uint32_t m0 = (chunk * kDeblockWarps + warp) * kBsChunk;
#pragma unroll 1
        for (uint32_t mb = m0; mb < min(m0 + kBsChunk, (uint32_t)g.nMbs); mb++) {
            const RecView cur{reinterpret_cast<const uint32_t *>(job.recs + mb)};
            const RecView lef{reinterpret_cast<const uint32_t *>(job.recs + mb - 1)};
            const RecView top{reinterpret_cast<const uint32_t *>(job.recs + mb - g.widthMbs)};
            const uint32_t w0 = __ldg(cur.w);
            const int flags = w0 >> 24;
            int bs = 0;
            if (flags & B200_MBF_FILTER_INNER) {   // GetMbFilteringFlags :289-320 (resolved on the host)
                const bool fLeft = flags & B200_MBF_FILTER_LEFT, fTop = flags & B200_MBF_FILTER_TOP;
                const int e = lane & 15, bx = e & 3, by = e >> 2;
                const int qb = cRasterToBlk[by * 4 + bx];
                const bool curIntra = (w0 & 0xFF) > B200_MB_P_8x8REF0;
                if (lane < 16) {
                    if (bx == 0) bs = !fLeft ? 0 : (curIntra || lef.intra()) ? 4 : bsPair(cur, qb, lef, cRasterToBlk[by * 4 + 3]);
                    else bs = curIntra ? 3 : bsPair(cur, qb, cur, cRasterToBlk[by * 4 + bx - 1]);
                } else {
                    if (by == 0) bs = !fTop ? 0 : (curIntra || top.intra()) ? 4 : bsPair(cur, qb, top, cRasterToBlk[12 + bx]);
                    else bs = curIntra ? 3 : bsPair(cur, qb, cur, cRasterToBlk[(by - 1) * 4 + bx]);
                }
            }
            uint32_t word = (uint32_t)bs << (4 * (lane & 7));
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            word |= __shfl_xor_sync(0xffffffffu, word, 4);
            const bool any = __ballot_sync(0xffffffffu, word != 0) != 0;
            const size_t idx = (size_t)s * g.nMbs + mb;
            if ((lane & 7) == 0) p.bsWords[idx * 4 + (lane >> 3)] = word;
            if (lane == 0) p.work[idx] = any ? 1 : 0;
        }
    }
}

// ---- stage 2: the filter proper, macroblocks with work only ---------------------------------------------------------
// Tickets in wavefront order (x + 2y ascending, streams interleaved); a CTA takes kDeblockWarps * kFilterChunk
// consecutive tickets, so a warp's consecutive tickets belong to different streams.  A macroblock waits for its left,
// top and top-right neighbours -- the macroblocks whose filtering the reference's raster order puts before it and whose
// pels it reads or rewrites -- but only for those that have work themselves (the others never touch a pel).
__global__ void __launch_bounds__(kDeblockWarps * 32) deblockKernel(const DeblockParams p) {
    __shared__ DeblockWarpSmem smemAll[kDeblockWarps];
    __shared__ uint32_t sBase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DeblockWarpSmem &sm = smemAll[warp];
    const PoolGeom &g = p.g;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sBase = atomicAdd(p.ticket, (uint32_t)(kDeblockWarps * kFilterChunk));
        __syncthreads();
        const uint32_t base = sBase;
        if (base >= p.totalTickets) break;
#pragma unroll 1
        for (uint32_t j = 0; j < kFilterChunk; j++) {
            const uint32_t t = base + warp * kFilterChunk + j;
            if (t >= p.totalTickets) break;
            const uint32_t k = t / (uint32_t)g.nStreams, s = t - k * (uint32_t)g.nStreams;
            const uint32_t mb = p.order[k];
            const size_t sIdx = (size_t)s * g.nMbs;
            if (!p.work[sIdx + mb]) continue;   // nothing to filter: this macroblock touches no pel
            const int mby = (int)__umulhi(mb, g.invWidthMbs), mbx = (int)(mb - (uint32_t)mby * g.widthMbs);
            const StreamJob job = p.jobs[s];
            uint32_t *doneS = p.done + sIdx;
            const RecView cur{reinterpret_cast<const uint32_t *>(job.recs + mb)};
            const RecView lef{reinterpret_cast<const uint32_t *>(job.recs + mb - 1)};
            const RecView top{reinterpret_cast<const uint32_t *>(job.recs + mb - g.widthMbs)};
            const uint32_t w0 = __ldg(cur.w);
            const int flags = w0 >> 24;
            const bool fLeft = flags & B200_MBF_FILTER_LEFT, fTop = flags & B200_MBF_FILTER_TOP;
            if (lane < 4) {
                const uint32_t wv = p.bsWords[(sIdx + mb) * 4 + lane];
#pragma unroll
                for (int i = 0; i < 8; i++) sm.bs[lane * 8 + i] = (uint8_t)((wv >> (4 * i)) & 15u);
            }
            if (lane < 3) {
                int nmb = -1;
                if (lane == 0 && mbx > 0) nmb = (int)mb - 1;
                if (lane == 1 && mby > 0) nmb = (int)mb - g.widthMbs;
                if (lane == 2 && mby > 0 && mbx < g.widthMbs - 1) nmb = (int)mb - g.widthMbs + 1;
                if (nmb >= 0 && p.work[sIdx + nmb]) waitFlag(doneS + nmb, p.serial);
            }
            __syncwarp();
            uint8_t *frame = framePtr(p.pool, g, s * (uint32_t)g.numSlots + job.curSlot);
            // stage 20x20 luma + 2x 10x12 chroma (incl. 4 / 2 pels of the left and upper neighbours) from L2
            for (int i = lane; i < 100; i += 32) {
                const int r = i / 5, wcol = i - r * 5;
                const uint32_t v = __ldcg(reinterpret_cast<const uint32_t *>(lumaAt(frame, g, mbx * 16 - 4 + wcol * 4, mby * 16 - 4 + r)));
                *reinterpret_cast<uint32_t *>(&sm.y[r][wcol * 4]) = v;
            }
            for (int i = lane; i < 60; i += 32) {
                const int pl = i / 30, jj = i - pl * 30, r = jj / 3, wcol = jj - r * 3;
                const uint32_t v = __ldcg(reinterpret_cast<const uint32_t *>(chromaAt(frame, g, pl, mbx * 8 - 4 + wcol * 4, mby * 8 - 2 + r)));
                *reinterpret_cast<uint32_t *>(&sm.c[pl][r][wcol * 4]) = v;
            }
            __syncwarp();
            // thresholds: GetLumaEdgeThresholds :1390-1458, GetChromaEdgeThresholds :1469-1541
            const uint32_t w3 = __ldg(cur.w + 3);
            const int offA = (int)(int8_t)(w3 & 0xFF), offB = (int)(int8_t)((w3 >> 8) & 0xFF), cqo = (int)(int8_t)((w3 >> 16) & 0xFF);
            const int qp = (w0 >> 8) & 0xFF;
            const int qpL = fLeft ? lef.qpY() : qp, qpT = fTop ? top.qpY() : qp;
            if (lane < 16) {
                const EdgeThr tIn = makeThr(qp, offA, offB), tL = makeThr((qp + qpL + 1) >> 1, offA, offB);
#pragma unroll 1
                for (int bx = 0; bx < 4; bx++) {   // all vertical edges of row `lane`, left to right
                    const int bs = sm.bs[(lane >> 2) * 4 + bx];
                    if (bs) filterLumaLine(&sm.y[4 + lane][4 + bx * 4], 1, bs, bx ? tIn : tL);
                }
            } else {
                const int pl = (lane - 16) >> 3, r = lane & 7;
                const int qc = cQpC[clip3(0, 51, qp + cqo)], qcL = cQpC[clip3(0, 51, qpL + cqo)];
                const EdgeThr tIn = makeThr(qc, offA, offB), tL = makeThr((qc + qcL + 1) >> 1, offA, offB);
#pragma unroll 1
                for (int ed = 0; ed < 2; ed++) {
                    const int bs = sm.bs[(r >> 1) * 4 + ed * 2];
                    if (bs) filterChromaLine(&sm.c[pl][2 + r][4 + ed * 4], 1, bs, ed ? tIn : tL);
                }
            }
            __syncwarp();
            if (lane < 16) {
                const EdgeThr tIn = makeThr(qp, offA, offB), tT = makeThr((qp + qpT + 1) >> 1, offA, offB);
#pragma unroll 1
                for (int by = 0; by < 4; by++) {
                    const int bs = sm.bs[16 + by * 4 + (lane >> 2)];
                    if (bs) filterLumaLine(&sm.y[4 + by * 4][4 + lane], 24, bs, by ? tIn : tT);
                }
            } else {
                const int pl = (lane - 16) >> 3, cx = lane & 7;
                const int qc = cQpC[clip3(0, 51, qp + cqo)], qcT = cQpC[clip3(0, 51, qpT + cqo)];
                const EdgeThr tIn = makeThr(qc, offA, offB), tT = makeThr((qc + qcT + 1) >> 1, offA, offB);
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    const int bs = sm.bs[16 + half * 8 + (cx >> 1)];
                    if (bs) filterChromaLine(&sm.c[pl][2 + half * 4][4 + cx], 12, bs, half ? tIn : tT);
                }
            }
            __syncwarp();
            // write back: own rows incl. the 4 columns of the left neighbour, then the 4 rows of the upper neighbour
            for (int i = lane; i < 80; i += 32) {
                const int r = i / 5, wcol = i - r * 5;
                if (wcol == 0 && mbx == 0) continue;
                *reinterpret_cast<uint32_t *>(lumaAt(frame, g, mbx * 16 - 4 + wcol * 4, mby * 16 + r)) =
                    *reinterpret_cast<const uint32_t *>(&sm.y[4 + r][wcol * 4]);
            }
            if (mby > 0 && lane < 16) {
                const int r = lane >> 2, wcol = lane & 3;
                *reinterpret_cast<uint32_t *>(lumaAt(frame, g, mbx * 16 + wcol * 4, mby * 16 - 4 + r)) =
                    *reinterpret_cast<const uint32_t *>(&sm.y[r][4 + wcol * 4]);
            }
            for (int i = lane; i < 48; i += 32) {
                const int pl = i / 24, jj = i - pl * 24, r = jj / 3, wcol = jj - r * 3;
                if (wcol == 0 && mbx == 0) continue;
                *reinterpret_cast<uint32_t *>(chromaAt(frame, g, pl, mbx * 8 - 4 + wcol * 4, mby * 8 + r)) =
                    *reinterpret_cast<const uint32_t *>(&sm.c[pl][2 + r][wcol * 4]);
            }
            if (mby > 0 && lane < 8) {
                const int pl = lane >> 2, r = (lane >> 1) & 1, wcol = lane & 1;
                *reinterpret_cast<uint32_t *>(chromaAt(frame, g, pl, mbx * 8 + wcol * 4, mby * 8 - 2 + r)) =
                    *reinterpret_cast<const uint32_t *>(&sm.c[pl][r][4 + wcol * 4]);
            }
            // publish: all lanes' stores happen-before the release by lane 0
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                stRelease(doneS + mb, p.serial);
            }
        }
    }
}

// ---- border replication -------------------------------------------------------------------------------
struct BorderParams {
    uint8_t *pool;
    PoolGeom g;
    const StreamJob *jobs;
};
// one warp per row of one plane (incl. border rows) of one stream's current frame
__global__ void __launch_bounds__(256) borderKernel(const BorderParams p) {
    const PoolGeom &g = p.g;
    const int lane = threadIdx.x & 31;
    const int rowsTotal = g.rowsY + 2 * g.rowsC;
    const long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= (long long)rowsTotal * g.nStreams) return;
    const int s = (int)(task / rowsTotal);
    int r = (int)(task - (long long)s * rowsTotal);
    uint8_t *frame = framePtr(p.pool, g, (uint32_t)s * g.numSlots + p.jobs[s].curSlot);
    uint8_t *plane;
    int w, h, pad, pitch;
    if (r < g.rowsY) { plane = frame; w = g.W; h = g.H; pad = kPadY; pitch = g.pitchY; }
    else {
        r -= g.rowsY;
        const int pl = r >= g.rowsC;
        if (pl) r -= g.rowsC;
        plane = frame + (pl ? g.offCr : g.offCb);
        w = g.W / 2; h = g.H / 2; pad = kPadC; pitch = g.pitchC;
    }
    const int sy = clip3(0, h - 1, r - pad);          // source picture row
    const uint8_t *src = plane + (size_t)(sy + pad) * pitch + pad;
    uint8_t *dst = plane + (size_t)r * pitch;
    const bool inside = (r - pad) == sy;
    const uint32_t lv = src[0] * 0x01010101u, rv = src[w - 1] * 0x01010101u;
    if (inside) {
        for (int i = lane; i < pad / 4; i += 32) reinterpret_cast<uint32_t *>(dst)[i] = lv;
        for (int i = (pad + w) / 4 + lane; i < pitch / 4; i += 32) reinterpret_cast<uint32_t *>(dst)[i] = rv;
    } else {
        for (int i = lane; i < pitch / 4; i += 32) {
            uint32_t v;
            const int x = i * 4 - pad;
            if (x < 0) v = lv;
            else if (x >= w) v = rv;
            else v = *reinterpret_cast<const uint32_t *>(src + x);
            reinterpret_cast<uint32_t *>(dst)[i] = v;
        }
    }
}

// ---- YUV -> 32-bit pixels (h264bsdConvertToRGBA/BGRA/YCbCrA, decoder.c:1163-1370) --------------------------
// mode 0: A<<24|B<<16|G<<8|R   1: A<<24|R<<16|G<<8|B   2: A<<24|Cr<<16|Cb<<8|Y ; nearest chroma, coded size
__global__ void __launch_bounds__(256) convertKernel(const uint8_t *yPlane, int pitchY, const uint8_t *cbPlane, const uint8_t *crPlane,
                                                     int pitchC, int W, int mode, uint32_t *out) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;  // four pels per thread
    const int y = blockIdx.y;
    if (x4 >= W) return;
    const uint32_t yv = *reinterpret_cast<const uint32_t *>(yPlane + (size_t)y * pitchY + x4);
    const uint32_t cbv = *reinterpret_cast<const uint16_t *>(cbPlane + (size_t)(y >> 1) * pitchC + (x4 >> 1));
    const uint32_t crv = *reinterpret_cast<const uint16_t *>(crPlane + (size_t)(y >> 1) * pitchC + (x4 >> 1));
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int l = (yv >> (8 * i)) & 0xFF, cb = (cbv >> (8 * (i >> 1))) & 0xFF, cr = (crv >> (8 * (i >> 1))) & 0xFF;
        if (mode == 2) {
            o[i] = 0xFF000000u | ((uint32_t)cr << 16) | ((uint32_t)cb << 8) | (uint32_t)l;
        } else {
            const int c = l - 16, d = cb - 128, e = cr - 128;
            const uint32_t r = (uint32_t)clip255((298 * c + 409 * e + 128) >> 8);
            const uint32_t gg = (uint32_t)clip255((298 * c - 100 * d - 208 * e + 128) >> 8);
            const uint32_t b = (uint32_t)clip255((298 * c + 516 * d + 128) >> 8);
            o[i] = mode == 0 ? (0xFF000000u | (b << 16) | (gg << 8) | r) : (0xFF000000u | (r << 16) | (gg << 8) | b);
        }
    }
    *reinterpret_cast<uint4 *>(out + (size_t)y * W + x4) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- compare frame `slot` of every stream with stream 0's (picture area only) -------------------------------
__global__ void __launch_bounds__(256) compareKernel(const uint8_t *pool, PoolGeom g, const uint32_t *slots, uint32_t *mismatch) {
    const int s = blockIdx.y + 1;
    const uint8_t *a = pool + (unsigned long long)(0 * g.numSlots + slots[0]) * g.frameStride;
    const uint8_t *b = pool + (unsigned long long)((unsigned)s * g.numSlots + slots[s]) * g.frameStride;
    const int wordsY = g.W / 4, wordsC = g.W / 8;
    const long long total = (long long)wordsY * g.H + 2ll * wordsC * (g.H / 2);
    uint32_t bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        size_t off;
        if (i < (long long)wordsY * g.H) {
            const int y = (int)(i / wordsY), x = (int)(i - (long long)y * wordsY);
            off = (size_t)(y + kPadY) * g.pitchY + kPadY + x * 4;
        } else {
            long long j = i - (long long)wordsY * g.H;
            const int pl = j >= (long long)wordsC * (g.H / 2);
            if (pl) j -= (long long)wordsC * (g.H / 2);
            const int y = (int)(j / wordsC), x = (int)(j - (long long)y * wordsC);
            off = (pl ? g.offCr : g.offCb) + (size_t)(y + kPadC) * g.pitchC + kPadC + x * 4;
        }
        bad += *reinterpret_cast<const uint32_t *>(a + off) != *reinterpret_cast<const uint32_t *>(b + off);
    }
    if (bad) atomicAdd(mismatch + s, bad);
}

}  // namespace b200
