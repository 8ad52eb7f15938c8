// deblock_kernel.cuh -- in-loop deblocking filter (h264bsdFilterPicture, h264bsd_deblocking.c:575-640)
// with the reference's raster order kept where it matters: half a warp walks a macroblock row left to right
// and waits for the row above to be two macroblocks ahead (the macroblocks whose filtering the reference's
// raster order puts before it and whose pels it reads or rewrites).
//
// Also here: replication of the picture border (what h264bsdFillBlock's coordinate clamp computes on
// the fly, reconstruct.c:2244-2367), YUV -> ARGB (h264bsd_decoder.c:1163-1370) and a frame compare.
#pragma once
#include "device_common.cuh"

namespace b200 {

constexpr int kDeblockWarps = 8;

struct DeblockParams {
    uint8_t *pool;
    PoolGeom g;
    const StreamJob *jobs;
    uint32_t *done;            // nStreams * heightMbs row-progress words of stage 2: serial << 16 | macroblocks finished
    uint32_t *ticket;
    uint32_t serial;
    uint32_t totalTickets;     // stage 2: heightMbs * ceil(nStreams / 2)
    uint32_t *bsWords;         // nStreams * nMbs * 4 words: packed boundary strengths (stage 1 -> stage 2)
    uint8_t *work;             // nStreams * nMbs: 1 = the macroblock has a non-zero boundary strength
    unsigned long long *workCount;   // running total of macroblocks with work (statistics for the roofline accounting)
};

struct __align__(16) DeblockWarpSmem {
    uint8_t y[20][32];      // rows -4..15; bytes 12..15 = cols -4..-1 (left neighbour), bytes 16..31 = cols 0..15
    uint8_t c[2][10][16];   // rows -2..7;  bytes  4..7  = cols -4..-1,                  bytes  8..15 = cols 0..7
};
// alpha / beta / tc0 / chroma-qp tables in shared memory: the lanes of a warp look them up at different indices
// (luma and chroma lanes, inner and macroblock edges), which serialises on the constant cache
struct DeblockTables {
    uint8_t alpha[52], beta[52], qpc[52], tc0[52][4];
};

struct EdgeThr { int alpha, beta, idxA; };
__device__ __forceinline__ EdgeThr makeThr(const DeblockTables &tb, int qp, int offA, int offB) {
    EdgeThr t;
    t.idxA = clip3(0, 51, qp + offA);
    t.alpha = tb.alpha[t.idxA];
    t.beta = tb.beta[clip3(0, 51, qp + offB)];
    return t;
}

// One line of samples across one edge, luma or chroma, in registers.
// FilterVerLumaEdge / FilterHorLuma(Edge) deblocking.c:656-965 and FilterVerChromaEdge / FilterHorChroma(Edge) :967-1146:
// the chroma filter is the luma filter without the p1/q1 updates (bS < 4, tc = tc0 + 1) resp. the luma filter's weak
// branch (bS == 4), so luma and chroma lanes run the same instructions.  Returns false when the line stays as it is.
struct EdgeLine { int p3, p2, p1, p0, q0, q1, q2, q3; };
__device__ __forceinline__ bool filterLine(EdgeLine &v, int bS, const EdgeThr &t, int tc0, bool chroma) {
    const int p0 = v.p0, p1 = v.p1, q0 = v.q0, q1 = v.q1;
    if (!(abs(p0 - q0) < t.alpha && abs(p1 - p0) < t.beta && abs(q1 - q0) < t.beta)) return false;
    const int p2 = v.p2, q2 = v.q2;
    const bool ap = !chroma && abs(p2 - p0) < t.beta, aq = !chroma && abs(q2 - q0) < t.beta;
    if (bS < 4) {
        const int tc = tc0 + (chroma ? 1 : (int)ap + (int)aq);
        const int avg = (p0 + q0 + 1) >> 1;
        if (ap) v.p1 = p1 + clip3(-tc0, tc0, (p2 + avg - (p1 << 1)) >> 1);
        if (aq) v.q1 = q1 + clip3(-tc0, tc0, (q2 + avg - (q1 << 1)) >> 1);
        const int d = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        v.p0 = clip255(p0 + d);
        v.q0 = clip255(q0 - d);
    } else {
        const bool strong = abs(p0 - q0) < ((t.alpha >> 2) + 2);
        if (strong && ap) {
            const int sum = p1 + p0 + q0;
            v.p0 = (p2 + 2 * sum + q1 + 4) >> 3;
            v.p1 = (p2 + sum + 2) >> 2;
            v.p2 = (2 * v.p3 + 3 * p2 + sum + 4) >> 3;
        } else {
            v.p0 = (2 * p1 + p0 + q1 + 2) >> 2;
        }
        if (strong && aq) {
            const int sum = p0 + q0 + q1;
            v.q0 = (p1 + 2 * sum + q2 + 4) >> 3;
            v.q1 = (sum + q2 + 2) >> 2;
            v.q2 = (2 * v.q3 + 3 * q2 + sum + 4) >> 3;
        } else {
            v.q0 = (2 * q1 + q0 + p1 + 2) >> 2;
        }
    }
    return true;
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

// ---- stage 1: boundary strengths, embarrassingly parallel ----------------------------------------------------
// GetBoundaryStrengths (deblocking.c:1187-1379) for every macroblock: 32 strengths packed as nibbles into 16 bytes
// (word j = segments 8j..8j+7; segments 0..15 = vertical edge left of raster block, 16..31 = horizontal edge above it)
// plus one "has work" byte.  Two thirds of the macroblocks of a typical P picture have nothing to filter.
//
// Two paths.  A thread per macroblock decides the common case -- the macroblock and the neighbours it is filtered against
// are all "uniform" (P_Skip / P_L0_16x16 without coded luma blocks: one vector, one reference, no coefficients) -- from four
// words per record: every inner strength is 0 and each outer edge has ONE strength, 0 or 1.  The rest (intra, several
// partitions, coded blocks; a third of a P picture) goes through the general routine: half a warp per macroblock, lane e
// owns raster block e and computes the strength of the edge on its left and of the edge above it.
struct BsSide {
    uint32_t w0, coded, refs, mv;
};
__device__ __forceinline__ BsSide loadSide(const b200_mb_rec *rec, int blk) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rec);
    BsSide r;
    r.w0 = __ldg(w); r.coded = __ldg(w + 1); r.refs = __ldg(w + 4); r.mv = __ldg(w + 8 + blk);
    return r;
}
__device__ __forceinline__ bool mvFar(uint32_t a, uint32_t b) {
    const int dx = (int)(int16_t)(a & 0xFFFF) - (int)(int16_t)(b & 0xFFFF);
    const int dy = (int)(int16_t)(a >> 16) - (int)(int16_t)(b >> 16);
    return abs(dx) >= 4 || abs(dy) >= 4;
}
// EdgeBoundaryStrength (:395-411) / InnerBoundaryStrength (:332-355) for two non-intra 4x4 blocks
__device__ __forceinline__ int bsInter(const BsSide &q, int qb, const BsSide &p, int pb) {
    if (((q.coded >> qb) | (p.coded >> pb)) & 1u) return 2;
    const uint32_t rq = (q.refs >> (8 * (qb >> 2))) & 0xFF, rp = (p.refs >> (8 * (pb >> 2))) & 0xFF;
    return (rq != rp || mvFar(q.mv, p.mv)) ? 1 : 0;
}
// the general routine for one macroblock per half-warp (all 32 lanes take part; `write` = this half-warp's result counts)
__device__ __forceinline__ bool strengthGeneral(const DeblockParams &p, const PoolGeom &g, const b200_mb_rec *recs, uint32_t mb,
                                                size_t idx, bool write, int lane) {
    const int e = lane & 15, bx = e & 3, by = e >> 2;
    const int qb = cRasterToBlk[e];
    const int lb = cRasterToBlk[by * 4 + ((bx + 3) & 3)];   // block on the left (of the left neighbour when bx == 0)
    const int tb = cRasterToBlk[((by + 3) & 3) * 4 + bx];   // block above (of the upper neighbour when by == 0)
    const b200_mb_rec *rc = recs + mb;
    const BsSide cur = loadSide(rc, qb);
    const int flags = cur.w0 >> 24;
    const bool fLeft = flags & B200_MBF_FILTER_LEFT, fTop = flags & B200_MBF_FILTER_TOP;
    // GetMbFilteringFlags :289-320 was resolved on the host.  The neighbour records are fetched whether or not the edge is
    // filtered (their addresses must not depend on the flags just loaded); a record that does not exist (first macroblock /
    // first row) is replaced by the current one and ignored
    const BsSide lef = loadSide((bx == 0 && mb > 0) ? rc - 1 : rc, lb);
    const BsSide top = loadSide((by == 0 && mb >= (uint32_t)g.widthMbs) ? rc - g.widthMbs : rc, tb);
    const bool curIntra = (cur.w0 & 0xFF) > B200_MB_P_8x8REF0;
    int bv, bh;
    if (bx == 0) bv = !fLeft ? 0 : (curIntra || (lef.w0 & 0xFF) > B200_MB_P_8x8REF0) ? 4 : bsInter(cur, qb, lef, lb);
    else bv = curIntra ? 3 : bsInter(cur, qb, lef, lb);
    if (by == 0) bh = !fTop ? 0 : (curIntra || (top.w0 & 0xFF) > B200_MB_P_8x8REF0) ? 4 : bsInter(cur, qb, top, tb);
    else bh = curIntra ? 3 : bsInter(cur, qb, top, tb);
    if (!(flags & B200_MBF_FILTER_INNER)) bv = bh = 0;
    uint32_t wv = (uint32_t)bv << (4 * (lane & 7)), wh = (uint32_t)bh << (4 * (lane & 7));
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        wv |= __shfl_xor_sync(0xffffffffu, wv, d);
        wh |= __shfl_xor_sync(0xffffffffu, wh, d);
    }
    const uint32_t wv1 = __shfl_down_sync(0xffffffffu, wv, 8), wh1 = __shfl_down_sync(0xffffffffu, wh, 8);
    const bool any = (wv | wv1 | wh | wh1) != 0;
    if (e == 0 && write) {
        *reinterpret_cast<uint4 *>(p.bsWords + idx * 4) = make_uint4(wv, wv1, wh, wh1);
        p.work[idx] = any ? 1 : 0;
    }
    return any && e == 0 && write;
}

__global__ void __launch_bounds__(kDeblockWarps * 32, 1) strengthKernel(const DeblockParams p) {
    __shared__ unsigned sWork;
    if (threadIdx.x == 0) sWork = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PoolGeom &g = p.g;
    const uint32_t chunksPerStream = ((uint32_t)g.nMbs + kDeblockWarps * 32 - 1) / (kDeblockWarps * 32);
    const uint32_t total = chunksPerStream * (uint32_t)g.nStreams;
    unsigned nWork = 0;
    for (uint32_t v = blockIdx.x; v < total; v += gridDim.x) {
        const uint32_t s = v / chunksPerStream, chunk = v - s * chunksPerStream;
        const StreamJob job = p.jobs[s];
        const uint32_t m0 = (chunk * kDeblockWarps + warp) * 32u;
        if (m0 >= (uint32_t)g.nMbs) continue;
        const uint32_t mb = m0 + lane;
        const bool valid = mb < (uint32_t)g.nMbs;
        const size_t sBase = (size_t)s * g.nMbs;
        // thread per macroblock: head, coded mask, reference slots and first vector of this record and of the one above
        const uint32_t mbc = valid ? mb : (uint32_t)g.nMbs - 1;
        const uint32_t *cw = reinterpret_cast<const uint32_t *>(job.recs + mbc);
        const uint32_t *tw = mbc >= (uint32_t)g.widthMbs ? cw - 24 * g.widthMbs : cw;
        const uint32_t c0 = __ldg(cw), c1 = __ldg(cw + 1), c4 = __ldg(cw + 4), c8 = __ldg(cw + 8);
        const uint32_t t0 = __ldg(tw), t1 = __ldg(tw + 1), t4 = __ldg(tw + 4), t8 = __ldg(tw + 8);
        // the record on the left is the neighbouring lane's (lane 0: one more load)
        uint32_t l0 = __shfl_up_sync(0xffffffffu, c0, 1), l1 = __shfl_up_sync(0xffffffffu, c1, 1);
        uint32_t l4 = __shfl_up_sync(0xffffffffu, c4, 1), l8 = __shfl_up_sync(0xffffffffu, c8, 1);
        if (lane == 0 && mb > 0) { l0 = __ldg(cw - 24); l1 = __ldg(cw - 23); l4 = __ldg(cw - 20); l8 = __ldg(cw - 16); }
        auto uniform = [](uint32_t w0, uint32_t w1) { return (w0 & 0xFF) <= B200_MB_P_16x16 && (w1 & 0xFFFFu) == 0; };
        const int flags = c0 >> 24;
        const bool fLeft = flags & B200_MBF_FILTER_LEFT, fTop = flags & B200_MBF_FILTER_TOP;
        const bool fast = uniform(c0, c1) && (!fLeft || uniform(l0, l1)) && (!fTop || uniform(t0, t1));
        if (valid && fast) {
            uint32_t bl = 0, bt = 0;
            if (flags & B200_MBF_FILTER_INNER) {
                if (fLeft) bl = ((c4 ^ l4) & 0xFF) != 0 || mvFar(c8, l8);
                if (fTop) bt = ((c4 ^ t4) & 0xFF) != 0 || mvFar(c8, t8);
            }
            // left edge = segments 0, 4, 8, 12 (nibbles 0 and 4 of words 0 and 1); top edge = segments 16..19 (word 2)
            *reinterpret_cast<uint4 *>(p.bsWords + (sBase + mb) * 4) = make_uint4(bl * 0x00010001u, bl * 0x00010001u, bt * 0x1111u, 0u);
            p.work[sBase + mb] = (bl | bt) ? 1 : 0;
            nWork += (bl | bt);
        }
        // everything else, two macroblocks at a time
        uint32_t slow = __ballot_sync(0xffffffffu, valid && !fast);
        while (slow) {
            const int a = __ffs(slow) - 1;
            slow &= slow - 1;
            const int b = slow ? __ffs(slow) - 1 : -1;
            if (b >= 0) slow &= slow - 1;
            const uint32_t mbSel = m0 + (uint32_t)((lane < 16 || b < 0) ? a : b);
            nWork += strengthGeneral(p, g, job.recs, mbSel, sBase + mbSel, lane < 16 || b >= 0, lane) ? 1u : 0u;
        }
    }
    if (nWork) atomicAdd(&sWork, nWork);
    __syncthreads();
    if (threadIdx.x == 0 && sWork) atomicAdd(p.workCount, (unsigned long long)sWork);
}

// ---- stage 2: the filter proper, macroblocks with work only ---------------------------------------------------------
// The reference filters the macroblocks of a picture in raster order (deblocking.c:604-640); what that order means for the
// pels is: a macroblock comes after its left neighbour, and after its upper and upper-right neighbours (it reads and rewrites
// 3 pels beyond its left and upper edges, and the upper-right neighbour's left edge reaches the upper neighbour's columns
// 13..15).  Here a HALF-WARP owns one macroblock row of one stream and walks it left to right -- the left neighbour is its own
// previous step -- and rows publish their progress in a per-row counter: macroblock x of row y starts when row y - 1 has
// finished x + 2 macroblocks.  Tickets are handed out row-major (row y of every stream before row y + 1 of any), so the row
// above is normally finished long before: nobody spins.  The two halves of a warp take rows of two different streams.
//
// The macroblock and the 4 (2) pels of its left / upper neighbours are staged in shared memory by vector loads (in the strip
// layout 20 contiguous luma rows, 2 x 10 chroma half-rows).  Lane i of the half owns luma line i -- a row for the vertical
// edges, left to right, then a column for the horizontal edges, top to bottom, which is the order of FilterLuma
// (deblocking.c:1569-1623) because edges of one direction only interact along a line -- and afterwards chroma line i & 7 of
// plane i >> 3 (FilterChroma :1633-1745; chroma edges are luma edges 0 and 2).  A step is skipped when neither macroblock has a
// strength for it.
#ifndef B200_DEBLOCK_MINBLOCKS
#define B200_DEBLOCK_MINBLOCKS 3
#endif
__device__ __forceinline__ int bsNibble(uint32_t word, int idx) { return (int)((word >> (4 * idx)) & 15u); }

__global__ void __launch_bounds__(kDeblockWarps * 32, B200_DEBLOCK_MINBLOCKS) deblockKernel(const DeblockParams p) {
    __shared__ DeblockWarpSmem smemAll[kDeblockWarps][2];
    __shared__ DeblockTables tb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, li = lane & 15;
    DeblockWarpSmem &sm = smemAll[warp][half];
    const PoolGeom &g = p.g;
    if (threadIdx.x < 52) {
        tb.alpha[threadIdx.x] = cAlpha[threadIdx.x];
        tb.beta[threadIdx.x] = cBeta[threadIdx.x];
        tb.qpc[threadIdx.x] = cQpC[threadIdx.x];
        *reinterpret_cast<uint32_t *>(tb.tc0[threadIdx.x]) = *reinterpret_cast<const uint32_t *>(cTc0[threadIdx.x]);
    }
    // this lane's lines: luma row / column li; chroma plane cpl, row / column cl
    const int cpl = li >> 3, cl = li & 7;
    const int grpY = li >> 2, grpC = cl >> 1;   // which 4-line strength segment a line belongs to
    const uint32_t pairs = ((uint32_t)g.nStreams + 1u) / 2u, serial16 = p.serial & 0xFFFFu;
    const int W = g.widthMbs;
    const size_t stripY = (size_t)g.rowsY * 16, stripC = (size_t)g.rowsC * 16;
    __syncthreads();   // the tables
    for (;;) {
        // every warp takes its own tickets (no CTA barrier): ticket = (row, pair of streams), row-major
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(p.ticket, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= p.totalTickets) break;
        const int y = (int)(ticket / pairs);
        const uint32_t s0 = 2u * (ticket - (uint32_t)y * pairs);
        const bool live = s0 + half < (uint32_t)g.nStreams;     // (an odd stream count leaves the last half idle: it shadows half 0)
        const uint32_t s = live ? s0 + half : s0;
        const StreamJob job = p.jobs[s];
        uint8_t *frame = p.pool + (unsigned long long)(s * (uint32_t)g.numSlots + job.curSlot) * g.frameStride;
        const b200_mb_rec *recsRow = job.recs + (size_t)y * W;
        const size_t rowIdx = (size_t)s * g.nMbs + (size_t)y * W;
        uint32_t *rowMine = p.done + (size_t)s * g.heightMbs + y;
        uint32_t seen = 0;    // macroblocks the row above is known to have finished
#pragma unroll 1
        for (int x0 = 0; x0 < W; x0 += 16) {
            // lane i of the half gathers what macroblock x0 + i needs (work flag -> record, strengths, neighbours' qp): one chain
            // of memory latencies per 16 macroblocks
            uint32_t mW0 = 0, mW3 = 0, mQp = 0, mWork = 0;
            uint4 mBw = make_uint4(0, 0, 0, 0);
            {
                const int x = x0 + li;
                if (x < W && p.work[rowIdx + x]) {   // else nothing to filter: this macroblock touches no pel
                    const uint32_t *cw = reinterpret_cast<const uint32_t *>(recsRow + x);
                    mW0 = __ldg(cw); mW3 = __ldg(cw + 3);
                    mBw = __ldg(reinterpret_cast<const uint4 *>(p.bsWords + (rowIdx + x) * 4));
                    const int flags = mW0 >> 24;
                    // qp of the neighbours filtered against (never read otherwise: they may not exist)
                    const uint32_t qp = (mW0 >> 8) & 0xFF;
                    const uint32_t qpL = (flags & B200_MBF_FILTER_LEFT) ? (__ldg(cw - 24) >> 8) & 0xFF : qp;
                    const uint32_t qpT = (flags & B200_MBF_FILTER_TOP) ? (__ldg(cw - 24 * W) >> 8) & 0xFF : qp;
                    mQp = qpL | (qpT << 8);
                    mWork = 1;
                }
            }
            uint32_t todo = (__ballot_sync(0xffffffffu, mWork != 0) >> (16 * half)) & 0xFFFFu;
            // The macroblock's own 16 luma rows / 2 x 8 chroma rows (and the 4 columns of its left neighbour) do not depend on the
            // row above: they are fetched one macroblock ahead, into registers, while the current one is filtered.  When the next
            // macroblock with work is the right neighbour, its left columns are this macroblock's right columns: taken from the
            // tile in shared memory once it is finished.
            uint4 pfY = make_uint4(0, 0, 0, 0);     // lane li: luma row li
            uint2 pfC = make_uint2(0, 0);           // lane li: chroma plane li >> 3, row li & 7
            uint32_t pfYl = 0, pfCl = 0;            // the left neighbour's last 4 luma / chroma pels of those rows
            auto fetchOwn = [&](int x, bool withLeft) {
                const uint8_t *lum = mbLuma(frame, g, x, y) + li * 16, *chr = mbChroma(frame, g, x, y) + cl * 16 + cpl * 8;
                pfY = __ldcg(reinterpret_cast<const uint4 *>(lum));
                pfC = __ldcg(reinterpret_cast<const uint2 *>(chr));
                if (withLeft) {
                    pfYl = __ldcg(reinterpret_cast<const uint32_t *>(lum - stripY + 12));
                    pfCl = __ldcg(reinterpret_cast<const uint32_t *>(chr - stripC + 4));
                }
            };
            if (__any_sync(0xffffffffu, todo != 0)) fetchOwn(x0 + (todo ? __ffs(todo) - 1 : 0), true);
#pragma unroll 1
            while (__any_sync(0xffffffffu, todo != 0)) {
                const bool active = live && todo != 0;
                const int i = todo ? __ffs(todo) - 1 : 0;
                todo &= todo - 1;
                const int src = half * 16 + i, x = x0 + i;
                const uint32_t w0 = __shfl_sync(0xffffffffu, mW0, src), w3 = __shfl_sync(0xffffffffu, mW3, src);
                const uint32_t qpn = __shfl_sync(0xffffffffu, mQp, src);
                uint4 bw;
                bw.x = __shfl_sync(0xffffffffu, mBw.x, src); bw.y = __shfl_sync(0xffffffffu, mBw.y, src);
                bw.z = __shfl_sync(0xffffffffu, mBw.z, src); bw.w = __shfl_sync(0xffffffffu, mBw.w, src);
                if (!active) bw = make_uint4(0, 0, 0, 0);
                const int qp = (w0 >> 8) & 0xFF, qpL = qpn & 0xFF, qpT = (qpn >> 8) & 0xFF;
                // the tile: rows 0..15 from the registers filled one step ago ...
                *reinterpret_cast<uint4 *>(&sm.y[4 + li][16]) = pfY;
                *reinterpret_cast<uint32_t *>(&sm.y[4 + li][12]) = pfYl;
                *reinterpret_cast<uint2 *>(&sm.c[cpl][2 + cl][8]) = pfC;
                *reinterpret_cast<uint32_t *>(&sm.c[cpl][2 + cl][4]) = pfCl;
                // ... and the next macroblock's on their way
                const int iNext = todo ? __ffs(todo) - 1 : -1;
                const bool nextAdjacent = iNext == i + 1;
                if (__any_sync(0xffffffffu, iNext >= 0)) fetchOwn(x0 + (iNext >= 0 ? iNext : i), !nextAdjacent);
                // thresholds: GetLumaEdgeThresholds :1390-1458, GetChromaEdgeThresholds :1469-1541
                const int offA = (int)(int8_t)(w3 & 0xFF), offB = (int)(int8_t)((w3 >> 8) & 0xFF), cqo = (int)(int8_t)((w3 >> 16) & 0xFF);
                const int qcs = tb.qpc[clip3(0, 51, qp + cqo)], qcsL = tb.qpc[clip3(0, 51, qpL + cqo)], qcsT = tb.qpc[clip3(0, 51, qpT + cqo)];
                __syncwarp();

                // ---- vertical edges, luma then chroma: own rows only, nothing of the row above ----------------------------------
                // segment (grp, e) is nibble (grp & 1) * 4 + e of word grp >> 1
                const uint32_t vAny = bw.x | bw.y;   // (per lane: the halves differ)
#pragma unroll 1
                for (int pass = 0; pass < 2; pass++) {
                    const bool chroma = pass != 0;
                    const int qs = chroma ? qcs : qp, qsL = chroma ? qcsL : qpL;
                    const EdgeThr tIn = makeThr(tb, qs, offA, offB), tL = makeThr(tb, (qs + qsL + 1) >> 1, offA, offB);
                    const int grp = chroma ? grpC : grpY;
                    uint8_t *rowp = chroma ? &sm.c[cpl][2 + cl][8] : &sm.y[4 + li][16];    // sample (0, line)
                    const int estep = chroma ? 2 : 4;                            // pels between edge e and e + 1 (luma numbering)
                    const uint32_t vWord = (grp >> 1) ? bw.y : bw.x;
#pragma unroll 1
                    for (int e = 0; e < 4; e += (chroma ? 2 : 1)) {
                        // no strength on this edge in any row of either macroblock: skip (warp-uniform)
                        if (!__any_sync(0xffffffffu, ((vAny >> (4 * e)) & 0x000F000Fu) != 0)) continue;
                        const int bs = bsNibble(vWord, (grp & 1) * 4 + e);
                        if (bs) {
                            uint32_t *wp = reinterpret_cast<uint32_t *>(rowp + e * estep);
                            const uint32_t P = wp[-1], Q = wp[0];
                            EdgeLine v;
                            v.p3 = P & 0xFF; v.p2 = (P >> 8) & 0xFF; v.p1 = (P >> 16) & 0xFF; v.p0 = P >> 24;
                            v.q0 = Q & 0xFF; v.q1 = (Q >> 8) & 0xFF; v.q2 = (Q >> 16) & 0xFF; v.q3 = Q >> 24;
                            const EdgeThr &th = e ? tIn : tL;
                            if (filterLine(v, bs, th, tb.tc0[th.idxA][bs - 1], chroma)) {
                                wp[-1] = (uint32_t)v.p3 | ((uint32_t)(v.p2 & 0xFF) << 8) | ((uint32_t)(v.p1 & 0xFF) << 16) | ((uint32_t)v.p0 << 24);
                                wp[0] = (uint32_t)(v.q0 & 0xFF) | ((uint32_t)(v.q1 & 0xFF) << 8) | ((uint32_t)(v.q2 & 0xFF) << 16) | ((uint32_t)v.q3 << 24);
                            }
                        }
                    }
                }
                // ---- the row above: needed by the top edge only.  It must have finished macroblocks x and x + 1 ------------------
                const bool topEdge = (bw.z & 0xFFFFu) != 0;      // a strength on the macroblock's top edge (segments 16..19)
                const uint32_t need = (active && topEdge && y > 0) ? (uint32_t)min(x + 2, W) : 0u;
                if (li == 0 && seen < need) seen = waitRow(rowMine - 1, serial16, need);
                seen = __shfl_sync(0xffffffffu, seen, half * 16);
                __syncwarp();     // (orders the other lanes' loads of the upper rows after lane 0's acquire)
                uint8_t *lum = mbLuma(frame, g, x, y), *chr = mbChroma(frame, g, x, y);
                if (topEdge) {
                    // the upper neighbour's last 4 luma rows (lanes 0..3) and last 2 chroma rows of both planes (lanes 4..7)
                    if (li < 4) *reinterpret_cast<uint4 *>(&sm.y[li][16]) = __ldcg(reinterpret_cast<const uint4 *>(lum - (4 - li) * 16));
                    else if (li < 8) {
                        const int pl = (li >> 1) & 1, r = li & 1;
                        *reinterpret_cast<uint2 *>(&sm.c[pl][r][8]) = __ldcg(reinterpret_cast<const uint2 *>(chr - (2 - r) * 16 + pl * 8));
                    }
                }
                __syncwarp();
                // ---- horizontal edges, luma then chroma: segment 16 + e * 4 + grp is nibble (e & 1) * 4 + grp of word 2 + (e >> 1) ----
#pragma unroll 1
                for (int pass = 0; pass < 2; pass++) {
                    const bool chroma = pass != 0;
                    const int qs = chroma ? qcs : qp, qsT = chroma ? qcsT : qpT;
                    const EdgeThr tIn = makeThr(tb, qs, offA, offB), tT = makeThr(tb, (qs + qsT + 1) >> 1, offA, offB);
                    const int grp = chroma ? grpC : grpY;
                    uint8_t *colp = chroma ? &sm.c[cpl][2][8 + cl] : &sm.y[4][16 + li];    // sample (line, 0)
                    const int pitch = chroma ? 16 : 32;
                    const int estep = chroma ? 2 : 4;
#pragma unroll 1
                    for (int e = 0; e < 4; e += (chroma ? 2 : 1)) {
                        const uint32_t hWord = (e >> 1) ? bw.w : bw.z;
                        if (!__any_sync(0xffffffffu, ((hWord >> ((e & 1) * 16)) & 0xFFFFu) != 0)) continue;   // no strength on this edge in any column
                        const int bs = bsNibble(hWord, (e & 1) * 4 + grp);
                        if (bs) {
                            uint8_t *q = colp + e * estep * pitch;
                            EdgeLine v;
                            v.p0 = q[-pitch]; v.p1 = q[-2 * pitch]; v.q0 = q[0]; v.q1 = q[pitch];
                            if (chroma) { v.p2 = v.p3 = v.q2 = v.q3 = 0; }   // outside the chroma tile for the top edge; never used
                            else { v.p2 = q[-3 * pitch]; v.p3 = q[-4 * pitch]; v.q2 = q[2 * pitch]; v.q3 = q[3 * pitch]; }
                            const EdgeLine o = v;
                            const EdgeThr &th = e ? tIn : tT;
                            if (filterLine(v, bs, th, tb.tc0[th.idxA][bs - 1], chroma)) {
                                q[-pitch] = (uint8_t)v.p0; q[0] = (uint8_t)v.q0;
                                if (v.p1 != o.p1) q[-2 * pitch] = (uint8_t)v.p1;
                                if (v.q1 != o.q1) q[pitch] = (uint8_t)v.q1;
                                if (v.p2 != o.p2) q[-3 * pitch] = (uint8_t)v.p2;
                                if (v.q2 != o.q2) q[2 * pitch] = (uint8_t)v.q2;
                            }
                        }
                    }
                }
                __syncwarp();
                // write back: own rows incl. the 4 columns of the left neighbour, and -- where the top edge was filtered -- the
                // 4 (2) rows of the upper neighbour
                if (active) {
                    *reinterpret_cast<uint4 *>(lum + li * 16) = *reinterpret_cast<const uint4 *>(&sm.y[4 + li][16]);
                    *reinterpret_cast<uint2 *>(chr + cl * 16 + cpl * 8) = *reinterpret_cast<const uint2 *>(&sm.c[cpl][2 + cl][8]);
                    if (x > 0) {
                        *reinterpret_cast<uint32_t *>(lum + li * 16 - stripY + 12) = *reinterpret_cast<const uint32_t *>(&sm.y[4 + li][12]);
                        *reinterpret_cast<uint32_t *>(chr + cl * 16 + cpl * 8 - stripC + 4) = *reinterpret_cast<const uint32_t *>(&sm.c[cpl][2 + cl][4]);
                    }
                    if (topEdge) {
                        if (li < 4) *reinterpret_cast<uint4 *>(lum - (4 - li) * 16) = *reinterpret_cast<const uint4 *>(&sm.y[li][16]);
                        else if (li < 8) {
                            const int pl = (li >> 1) & 1, r = li & 1;
                            *reinterpret_cast<uint2 *>(chr - (2 - r) * 16 + pl * 8) = *reinterpret_cast<const uint2 *>(&sm.c[pl][r][8]);
                        }
                    }
                }
                // the right neighbour comes next: its left columns are this macroblock's columns 12..15 (7..4 of chroma) as they
                // are now
                if (nextAdjacent) {
                    pfYl = *reinterpret_cast<const uint32_t *>(&sm.y[4 + li][28]);
                    pfCl = *reinterpret_cast<const uint32_t *>(&sm.c[cpl][2 + cl][12]);
                }
                // publish the row's progress: everything before the half's next macroblock with work (or the end of this group)
                // is finished.  The warp barrier orders every lane's stores before the release of lane 0 of each half, and a
                // release at gpu scope is cumulative (the same pattern as a CTA semaphore: barrier, then st.release by one thread)
                __syncwarp();
                if (active && li == 0) stRelease(rowMine, (serial16 << 16) | (uint32_t)min(todo ? x0 + __ffs(todo) - 1 : x0 + 16, W));
            }
            // a group without work (or whose last macroblock had none) still moves the row on
            __syncwarp();
            if (live && li == 0) stRelease(rowMine, (serial16 << 16) | (uint32_t)min(x0 + 16, W));
        }
    }
}

// ---- border replication -------------------------------------------------------------------------------
struct BorderParams {
    uint8_t *pool;
    PoolGeom g;
    const StreamJob *jobs;
};
// Warp tasks per stream (borderTasksPerStream): a "cap" task fills the 32 (16) rows above and below one picture strip with
// the strip's first / last picture row (one 16-byte row per lane and store); a "side" task fills 128 rows of one of the four
// border strips of a plane with each row's edge pel (rows above / below the picture take the corner pel).
__host__ __device__ inline int borderTasksPerStream(const PoolGeom &g) {
    return g.widthMbs + 4 * ((g.rowsY + 127) / 128) + 4 * ((g.rowsC + 127) / 128);
}
__global__ void __launch_bounds__(256) borderKernel(const BorderParams p) {
    const PoolGeom &g = p.g;
    const int lane = threadIdx.x & 31;
    const int sideY = (g.rowsY + 127) / 128, sideC = (g.rowsC + 127) / 128;
    const int perStream = g.widthMbs + 4 * sideY + 4 * sideC;
    const long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= (long long)perStream * g.nStreams) return;
    const int s = (int)(task / perStream);
    int t = (int)(task - (long long)s * perStream);
    uint8_t *frame = framePtr(p.pool, g, (uint32_t)s * g.numSlots + p.jobs[s].curSlot);
    if (t < g.widthMbs) {
        // cap: picture strip t
        uint8_t *ys = frame + (size_t)(t + kPadMbs) * g.rowsY * 16, *cs = frame + g.offC + (size_t)(t + kPadMbs) * g.rowsC * 16;
        const uint4 top = *reinterpret_cast<const uint4 *>(ys + (size_t)kPadY * 16);
        const uint4 bot = *reinterpret_cast<const uint4 *>(ys + (size_t)(kPadY + g.H - 1) * 16);
        *reinterpret_cast<uint4 *>(ys + (size_t)lane * 16) = top;
        *reinterpret_cast<uint4 *>(ys + (size_t)(kPadY + g.H + lane) * 16) = bot;
        const bool low = lane >= 16;
        const uint4 c = *reinterpret_cast<const uint4 *>(cs + (size_t)(low ? kPadC + g.H / 2 - 1 : kPadC) * 16);
        *reinterpret_cast<uint4 *>(cs + (size_t)(low ? kPadC + g.H / 2 + lane - 16 : lane) * 16) = c;
        return;
    }
    t -= g.widthMbs;
    const bool luma = t < 4 * sideY;
    if (!luma) t -= 4 * sideY;
    const int blocks = luma ? sideY : sideC, which = t / blocks, blk = t - which * blocks;   // which: 0, 1 left strips; 2, 3 right strips
    const int strip = which < 2 ? which : g.widthMbs + which;
    const int rows = luma ? g.rowsY : g.rowsC, pad = luma ? kPadY : kPadC, h = luma ? g.H : g.H / 2;
    const int edgeX = which < 2 ? 0 : (luma ? g.W : g.W / 2) - 1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r = blk * 128 + k * 32 + lane;
        if (r >= rows) break;
        const int y = clip3(0, h - 1, r - pad);
        uint4 v;
        if (luma) {
            const uint32_t w = *lumaAt(frame, g, edgeX, y) * 0x01010101u;
            v = make_uint4(w, w, w, w);
            *reinterpret_cast<uint4 *>(frame + ((size_t)strip * g.rowsY + r) * 16) = v;
        } else {
            const uint32_t cb = *chromaAt(frame, g, 0, edgeX, y) * 0x01010101u, cr = *chromaAt(frame, g, 1, edgeX, y) * 0x01010101u;
            v = make_uint4(cb, cb, cr, cr);
            *reinterpret_cast<uint4 *>(frame + g.offC + ((size_t)strip * g.rowsC + r) * 16) = v;
        }
    }
}

// ---- YUV -> 32-bit pixels (h264bsdConvertToRGBA/BGRA/YCbCrA, decoder.c:1163-1370) --------------------------
// mode 0: A<<24|B<<16|G<<8|R   1: A<<24|R<<16|G<<8|B   2: A<<24|Cr<<16|Cb<<8|Y ; nearest chroma, coded size
__device__ __forceinline__ uint32_t convertPel(int l, int cb, int cr, int mode) {
    if (mode == 2) return 0xFF000000u | ((uint32_t)cr << 16) | ((uint32_t)cb << 8) | (uint32_t)l;
    const int c = l - 16, d = cb - 128, e = cr - 128;
    const uint32_t r = (uint32_t)clip255((298 * c + 409 * e + 128) >> 8);
    const uint32_t gg = (uint32_t)clip255((298 * c - 100 * d - 208 * e + 128) >> 8);
    const uint32_t bb = (uint32_t)clip255((298 * c + 516 * d + 128) >> 8);
    return mode == 0 ? (0xFF000000u | (bb << 16) | (gg << 8) | r) : (0xFF000000u | (r << 16) | (gg << 8) | bb);
}
// planar I420 input (caller-owned host pictures, h264bsdConvertTo*): 4 pels per thread, or 8 when the width allows it
template <int PELS>
__global__ void __launch_bounds__(256) convertKernelT(const uint8_t *yPlane, int pitchY, const uint8_t *cbPlane, const uint8_t *crPlane,
                                                      int pitchC, int W, int mode, uint32_t *out) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * PELS;
    const int y = blockIdx.y;
    if (x0 >= W) return;
    uint32_t yw[PELS / 4], cbv, crv;
    if (PELS == 8) {
        const uint2 t = *reinterpret_cast<const uint2 *>(yPlane + (size_t)y * pitchY + x0);
        yw[0] = t.x; yw[PELS / 4 - 1] = t.y;
        cbv = *reinterpret_cast<const uint32_t *>(cbPlane + (size_t)(y >> 1) * pitchC + (x0 >> 1));
        crv = *reinterpret_cast<const uint32_t *>(crPlane + (size_t)(y >> 1) * pitchC + (x0 >> 1));
    } else {
        yw[0] = *reinterpret_cast<const uint32_t *>(yPlane + (size_t)y * pitchY + x0);
        cbv = *reinterpret_cast<const uint16_t *>(cbPlane + (size_t)(y >> 1) * pitchC + (x0 >> 1));
        crv = *reinterpret_cast<const uint16_t *>(crPlane + (size_t)(y >> 1) * pitchC + (x0 >> 1));
    }
    uint32_t o[PELS];
#pragma unroll
    for (int i = 0; i < PELS; i++)
        o[i] = convertPel((yw[i >> 2] >> (8 * (i & 3))) & 0xFF, (cbv >> (8 * (i >> 1))) & 0xFF, (crv >> (8 * (i >> 1))) & 0xFF, mode);
#pragma unroll
    for (int i = 0; i < PELS; i += 4)
        *reinterpret_cast<uint4 *>(out + (size_t)y * W + x0 + i) = make_uint4(o[i], o[i + 1], o[i + 2], o[i + 3]);
}
// a frame of the pool (strip layout) -> W x H pixels, 8 pels per thread (half a strip row: 8-byte luma load, 4 + 4 chroma
// bytes, two 16-byte stores).  blockIdx.z = picture of a batch: the frame `inStride` bytes further, output `outStride` pixels
// further.  With a cropping rectangle (cropW x cropH at (cropX, cropY), all even) only that part is written, densely.
__global__ void __launch_bounds__(256) convertFrameKernel(const uint8_t *frame0, PoolGeom g, int mode, uint32_t *out,
                                                          unsigned long long inStride, unsigned long long outStride,
                                                          int cropX, int cropY, int cropW, int cropH) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int y = blockIdx.y;
    if (x0 >= g.W) return;
    uint8_t *frame = const_cast<uint8_t *>(frame0) + blockIdx.z * inStride;
    out += blockIdx.z * outStride;
    const uint2 yv = *reinterpret_cast<const uint2 *>(lumaAt(frame, g, x0, y));
    const uint32_t cbv = *reinterpret_cast<const uint32_t *>(chromaAt(frame, g, 0, x0 >> 1, y >> 1));
    const uint32_t crv = *reinterpret_cast<const uint32_t *>(chromaAt(frame, g, 1, x0 >> 1, y >> 1));
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; i++)
        o[i] = convertPel(((i < 4 ? yv.x : yv.y) >> (8 * (i & 3))) & 0xFF, (cbv >> (8 * (i >> 1))) & 0xFF, (crv >> (8 * (i >> 1))) & 0xFF, mode);
    if (cropW == g.W && cropH == g.H) {
        uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)y * g.W + x0);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else if (y >= cropY && y < cropY + cropH) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (x0 + i >= cropX && x0 + i < cropX + cropW) out[(size_t)(y - cropY) * cropW + (x0 + i - cropX)] = o[i];
    }
}

// ---- strip layout <-> planar I420 / NV12 of the coded size (or of a cropping rectangle) ----------------------------------
// What h264bsdNextOutputPicture hands out (decoder.c:599-623) is contiguous planar I420: one thread moves two luma rows of
// a strip (a whole 32-byte sector in, two 16-byte row pieces out) or one chroma row (8 Cb + 8 Cr in, 8 bytes to each plane;
// nv12: 16 interleaved bytes to the one chroma plane).  blockIdx.y = stream: its current frame (slot from the job) -> staging +
// stream * outStride.  crop: cropX, cropY, cropW, cropH in luma pels, cropX a multiple of 16 and the rest even; the output
// planes then have cropW x cropH (cropW/2 x cropH/2) pels.
struct PackParams {
    const uint8_t *pool;
    PoolGeom g;
    const StreamJob *jobs;     // curSlot per stream (nullptr: slot `slot` of every stream)
    uint32_t slot;
    uint8_t *out;
    unsigned long long outStride;
    int cropX, cropY, cropW, cropH;
    int nv12;
};
__global__ void __launch_bounds__(256) packKernel(const PackParams p) {
    const PoolGeom &g = p.g;
    const uint32_t s = blockIdx.y;
    const uint32_t slot = p.jobs ? p.jobs[s].curSlot : p.slot;
    uint8_t *f = const_cast<uint8_t *>(p.pool) + (unsigned long long)(s * (uint32_t)g.numSlots + slot) * g.frameStride;
    uint8_t *out = p.out + (size_t)s * p.outStride;
    const int unitsY = g.widthMbs * (g.H / 2), unitsC = g.widthMbs * (g.H / 2);
    const size_t planeY = (size_t)p.cropW * p.cropH, planeC = planeY / 4;
    const int cx0 = p.cropX, cx1 = p.cropX + p.cropW;   // luma columns kept
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < unitsY + unitsC; i += gridDim.x * blockDim.x) {
        if (i < unitsY) {
            const int yp = i / g.widthMbs, mbx = i - yp * g.widthMbs, y = 2 * yp, x = mbx * 16;
            if (y < p.cropY || y >= p.cropY + p.cropH || x + 16 <= cx0 || x >= cx1) continue;
            const uint4 *src = reinterpret_cast<const uint4 *>(lumaAt(f, g, x, y));
            const uint4 a = src[0], b = src[1];
            uint8_t *d = out + (size_t)(y - p.cropY) * p.cropW + (x - cx0);
            if (x >= cx0 && x + 16 <= cx1 && (p.cropW & 15) == 0) {
                *reinterpret_cast<uint4 *>(d) = a;
                *reinterpret_cast<uint4 *>(d + p.cropW) = b;
            } else {
                const uint8_t *ab = reinterpret_cast<const uint8_t *>(&a), *bb = reinterpret_cast<const uint8_t *>(&b);
                for (int k = 0; k < 16; k++)
                    if (x + k >= cx0 && x + k < cx1) { d[k] = ab[k]; d[p.cropW + k] = bb[k]; }
            }
        } else {
            const int j = i - unitsY, yc = j / g.widthMbs, mbx = j - yc * g.widthMbs, x = mbx * 8;
            if (2 * yc < p.cropY || 2 * yc >= p.cropY + p.cropH || 2 * x + 16 <= cx0 || 2 * x >= cx1) continue;
            const uint4 v = *reinterpret_cast<const uint4 *>(chromaAt(f, g, 0, x, yc));   // 8 Cb | 8 Cr
            const int cw = p.cropW / 2, ccx0 = cx0 / 2, row = yc - p.cropY / 2;
            const uint8_t *vb = reinterpret_cast<const uint8_t *>(&v);
            if (p.nv12) {
                uint8_t *d = out + planeY + (size_t)row * p.cropW + 2 * (x - ccx0);
                for (int k = 0; k < 8; k++)
                    if (x + k >= ccx0 && x + k < ccx0 + cw) { d[2 * k] = vb[k]; d[2 * k + 1] = vb[8 + k]; }
            } else if (x >= ccx0 && x + 8 <= ccx0 + cw && (cw & 7) == 0) {
                *reinterpret_cast<uint2 *>(out + planeY + (size_t)row * cw + (x - ccx0)) = make_uint2(v.x, v.y);
                *reinterpret_cast<uint2 *>(out + planeY + planeC + (size_t)row * cw + (x - ccx0)) = make_uint2(v.z, v.w);
            } else {
                for (int k = 0; k < 8; k++)
                    if (x + k >= ccx0 && x + k < ccx0 + cw) {
                        out[planeY + (size_t)row * cw + (x + k - ccx0)] = vb[k];
                        out[planeY + planeC + (size_t)row * cw + (x + k - ccx0)] = vb[8 + k];
                    }
            }
        }
    }
}
// the reverse, coded size only (test hook: put an I420 picture into a frame slot)
__global__ void __launch_bounds__(256) unpackKernel(uint8_t *frame, PoolGeom g, const uint8_t *in) {
    const int unitsY = g.widthMbs * g.H, unitsC = g.widthMbs * (g.H / 2);
    const size_t planeY = (size_t)g.W * g.H, planeC = planeY / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < unitsY + unitsC; i += gridDim.x * blockDim.x) {
        if (i < unitsY) {
            const int y = i / g.widthMbs, mbx = i - y * g.widthMbs;
            uint4 v;
            memcpy(&v, in + (size_t)y * g.W + mbx * 16, 16);
            *reinterpret_cast<uint4 *>(lumaAt(frame, g, mbx * 16, y)) = v;
        } else {
            const int j = i - unitsY, yc = j / g.widthMbs, mbx = j - yc * g.widthMbs;
            uint2 cb, cr;
            memcpy(&cb, in + planeY + (size_t)yc * (g.W / 2) + mbx * 8, 8);
            memcpy(&cr, in + planeY + planeC + (size_t)yc * (g.W / 2) + mbx * 8, 8);
            *reinterpret_cast<uint4 *>(chromaAt(frame, g, 0, mbx * 8, yc)) = make_uint4(cb.x, cb.y, cr.x, cr.y);
        }
    }
}

// ---- compare frame `slot` of every stream with stream 0's (picture area only) -------------------------------
__global__ void __launch_bounds__(256) compareKernel(const uint8_t *pool, PoolGeom g, const uint32_t *slots, uint32_t *mismatch) {
    const int s = blockIdx.y + 1;
    const uint8_t *a = pool + (unsigned long long)(0 * g.numSlots + slots[0]) * g.frameStride;
    const uint8_t *b = pool + (unsigned long long)((unsigned)s * g.numSlots + slots[s]) * g.frameStride;
    const int unitsY = g.widthMbs * g.H, unitsC = g.widthMbs * (g.H / 2);   // 16-byte strip rows inside the picture
    uint32_t bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < unitsY + unitsC; i += gridDim.x * blockDim.x) {
        size_t off;
        if (i < unitsY) {
            const int strip = i / g.H, row = i - strip * g.H;
            off = ((size_t)(strip + kPadMbs) * g.rowsY + row + kPadY) * 16;
        } else {
            const int j = i - unitsY, strip = j / (g.H / 2), row = j - strip * (g.H / 2);
            off = g.offC + ((size_t)(strip + kPadMbs) * g.rowsC + row + kPadC) * 16;
        }
        const uint4 x = *reinterpret_cast<const uint4 *>(a + off), y = *reinterpret_cast<const uint4 *>(b + off);
        bad += (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    if (bad) atomicAdd(mismatch + s, bad);
}

}  // namespace b200
