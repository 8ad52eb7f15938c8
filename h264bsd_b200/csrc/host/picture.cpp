// picture.cpp -- see picture.hpp
#include "picture.hpp"
#include "cavlc.hpp"
#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace b200 {

namespace {
// 4x4 luma block order inside a macroblock (clause 6.4.3; h264bsdBlockX/Y intra_prediction.c:86-89)
const uint8_t kBlkX[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
const uint8_t kBlkY[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};
const uint8_t kZ[4][4] = {{0, 1, 4, 5}, {2, 3, 6, 7}, {8, 9, 12, 13}, {10, 11, 14, 15}};  // [y][x]

// Table 9-4, coded_block_pattern for chroma_format_idc 1/2: codeNum -> {Intra, Inter}
const uint8_t kCbp[48][2] = {
    {47, 0},  {31, 16}, {15, 1},  {0, 2},   {23, 4},  {27, 8},  {29, 32}, {30, 3},  {7, 5},   {11, 10},
    {13, 12}, {14, 15}, {39, 47}, {43, 7},  {45, 11}, {46, 13}, {16, 14}, {3, 6},   {5, 9},   {10, 31},
    {12, 35}, {19, 37}, {21, 42}, {26, 44}, {28, 33}, {35, 34}, {37, 36}, {42, 40}, {44, 39}, {1, 43},
    {2, 45},  {4, 46},  {8, 17},  {17, 18}, {18, 20}, {20, 24}, {24, 19}, {6, 21},  {9, 26},  {22, 28},
    {25, 23}, {32, 27}, {33, 29}, {34, 30}, {36, 22}, {40, 25}, {38, 38}, {41, 41}};

// Table 8-15: QPc as a function of qPI (h264bsdQpC, h264bsd_util.c:53-55)
const uint8_t kQpC[52] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17,
                          18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 32, 33,
                          34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};

inline bool isInterType(uint32_t t) { return t <= B200_MB_P_8x8REF0; }
inline int numMbPart(uint32_t t) { return (t == B200_MB_P_16x16 || t == B200_MB_P_SKIP) ? 1 : (t == B200_MB_P_16x8 || t == B200_MB_P_8x16) ? 2 : 4; }
inline int numSubMbPart(uint32_t s) { return s == 0 ? 1 : s == 3 ? 4 : 2; }
}  // namespace

extern "C" uint32_t b200_cbp_probe(uint32_t codeNum, int intra) { return codeNum < 48 ? kCbp[codeNum][intra ? 0 : 1] : 0xFFFFFFFFu; }

struct PictureState::MbSyntax {
    uint32_t mbType;
    uint32_t cbp;
    int32_t qpDelta;
    uint8_t prevFlag[16], remMode[16];
    uint32_t chromaMode;
    uint32_t refIdx[4];
    int16_t mvd[4][2];
    uint32_t subType[4];
    int16_t subMvd[4][4][2];
    uint8_t totalCoeff[27];
    uint32_t sumAbs[27];    // per block: sum of level magnitudes ([24] luma DC, [25] Cb DC, [26] Cr DC); valid where coded
    uint32_t spill;         // bit i: level[i][0] was written by the 15-coefficient block before it (corrupt streams only:
                            // position 15 of an AC block is the next block's first entry, here as in the reference)
    uint32_t codedBlocks;   // bit i: level[i] holds levels (TotalCoeff != 0); bits 24 / 25 = B200_CM_LUMA_DC / B200_CM_CHROMA_DC
    alignas(16) int16_t level[26][16];  // [0..23] blocks, [24] luma DC, [25] chroma DC (Cb 0..3, Cr 4..7)
    uint8_t pcm[384];
    void clear() {
        // what a macroblock may read without having parsed it: the level arrays are cleared selectively after use, the
        // prediction syntax (prevFlag / remMode / mvd / subType / subMvd) is written for exactly the entries that are read
        cbp = 0; qpDelta = 0; chromaMode = 0; codedBlocks = 0; spill = 0;
        std::memset(refIdx, 0, sizeof refIdx);
        std::memset(totalCoeff, 0, sizeof totalCoeff);
    }
};

void PictureState::resize(uint32_t w, uint32_t h) {
    widthMbs = w; heightMbs = h; picSizeInMbs = w * h;
    ownSt_.clear();
    ownRecs_.clear();
    st = recs = nullptr;
    bound_ = false;
    aux.assign(picSizeInMbs, MbAux());
    sliceGroupMap.assign(picSizeInMbs, 0);
    coefs.clear();
    orderClass.assign(2 * (size_t)picSizeInMbs, 0);
    lateFixup_ = false;
    numIntraPred_ = numCopies_ = 0;
    sliceIdCounter = numDecodedMbs = lastMbAddr = 0;
}

void PictureState::bindOutput() {
    if (bound_) return;
    bound_ = true;
    b200_mb_rec *out = provider ? provider->pictureRecords(picSizeInMbs) : nullptr;
    if (!out) {
        if (ownRecs_.size() != picSizeInMbs) {
            ownRecs_.resize(picSizeInMbs);
            std::memset(ownRecs_.data(), 0, sizeof(b200_mb_rec) * picSizeInMbs);
        }
        out = ownRecs_.data();
    }
    st = recs = out;
}

void PictureState::splitState() {
    lateFixup_ = true;      // a macroblock decoded twice: classes and flags are settled by the full pass of finalizeRecords
    if (st != recs) return;
    ownSt_.assign(recs, recs + picSizeInMbs);
    st = ownSt_.data();
}

void PictureState::beginPicture() {
    bound_ = false;
    numDecodedMbs = 0;
    sliceIdCounter = 0;
    for (auto &a : aux) { a.sliceId = 0; a.decoded = 0; }
    coefs.clear();
    orderClass.resize(2 * (size_t)picSizeInMbs);
    lateFixup_ = false;
    numIntraPred_ = numCopies_ = 0;
    concealOrder.clear();
}

bool PictureState::allDecoded(bool redundant) const {
    if (!redundant) return numDecodedMbs == picSizeInMbs;
    uint32_t n = 0;
    for (const auto &a : aux) n += a.decoded ? 1 : 0;
    return n == picSizeInMbs;
}

uint32_t PictureState::nextMbAddress(uint32_t cur) const {
    uint32_t g = sliceGroupMap[cur];
    uint32_t i = cur + 1;
    while (i < picSizeInMbs && sliceGroupMap[i] != g) i++;
    return i == picSizeInMbs ? 0 : i;
}

// DetermineNc (h264bsd_macroblock_layer.c:810-870): average of left/above totalCoeff where available
int PictureState::nC(uint32_t mbAddr, uint32_t blk, const uint8_t *cur) const {
    int aMb, bMb;           // 0 current, 1 neighbour MB
    uint32_t aIdx, bIdx;
    if (blk < 16) {
        int x = kBlkX[blk], y = kBlkY[blk];
        if (x > 0) { aMb = 0; aIdx = kZ[y][x - 1]; } else { aMb = 1; aIdx = kZ[y][3]; }
        if (y > 0) { bMb = 0; bIdx = kZ[y - 1][x]; } else { bMb = 1; bIdx = kZ[3][x]; }
    } else {
        uint32_t i = blk & 3;
        if (i & 1) { aMb = 0; aIdx = blk - 1; } else { aMb = 1; aIdx = blk + 1; }
        if (i & 2) { bMb = 0; bIdx = blk - 2; } else { bMb = 1; bIdx = blk + 2; }
    }
    int n;
    if (!aMb && !bMb) {
        n = (cur[aIdx] + cur[bIdx] + 1) >> 1;
    } else if (!aMb) {
        n = cur[aIdx];
        int b = curNb_[1];
        if (b >= 0) n = (n + aux[b].totalCoeff[bIdx] + 1) >> 1;
    } else if (!bMb) {
        n = cur[bIdx];
        int a = curNb_[0];
        if (a >= 0) n = (n + aux[a].totalCoeff[aIdx] + 1) >> 1;
    } else {
        n = 0;
        bool haveA = false;
        int a = curNb_[0], b = curNb_[1];
        if (a >= 0) { n = aux[a].totalCoeff[aIdx]; haveA = true; }
        if (b >= 0) n = haveA ? (n + aux[b].totalCoeff[bIdx] + 1) >> 1 : aux[b].totalCoeff[bIdx];
    }
    return n;
}

// h264bsdDecodeMacroblockLayer (h264bsd_macroblock_layer.c:134-243)
bool PictureState::parseMacroblockLayer(BitReader &stream, MbSyntax &mb, uint32_t mbAddr, bool iSlice, uint32_t numRefIdxActive) {
    uint32_t v;
    int32_t s;
    mb.clear();
    BitReader br = stream;      // local copy: the position stays in a register; handed back before the residual and at the end
                                // (after an error the position is not used again)
    bool ok = br.ue(v);
    uint32_t off = iSlice ? 6 : 1;
    if (!ok || v + off > 31) return false;
    mb.mbType = v + off;
    if (mb.mbType == B200_MB_I_PCM) {
        while (!br.byteAligned()) {
            if (!br.get1(v) || v) return false;  // pcm_alignment_zero_bit
        }
        for (int i = 0; i < 384; i++) {
            if (!br.get(8, v)) return false;
            mb.pcm[i] = (uint8_t)v;
        }
        stream = br;
        return true;
    }
    bool inter = isInterType(mb.mbType);
    if (inter && numMbPart(mb.mbType) == 4) {
        // sub_mb_pred()
        for (int i = 0; i < 4; i++) {
            if (!br.ue(v) || v > 3) return false;
            mb.subType[i] = v;
        }
        if (numRefIdxActive > 1 && mb.mbType != B200_MB_P_8x8REF0) {
            for (int i = 0; i < 4; i++) {
                if (!br.te(v, numRefIdxActive > 2) || v >= numRefIdxActive) return false;
                mb.refIdx[i] = v;
            }
        }
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < numSubMbPart(mb.subType[i]); j++) {
                if (!br.se(s)) return false;
                mb.subMvd[i][j][0] = (int16_t)s;
                if (!br.se(s)) return false;
                mb.subMvd[i][j][1] = (int16_t)s;
            }
    } else if (inter) {
        int np = numMbPart(mb.mbType);
        if (numRefIdxActive > 1) {
            for (int j = 0; j < np; j++) {
                if (!br.te(v, numRefIdxActive > 2) || v >= numRefIdxActive) return false;
                mb.refIdx[j] = v;
            }
        }
        for (int j = 0; j < np; j++) {
            if (!br.se(s)) return false;
            mb.mvd[j][0] = (int16_t)s;
            if (!br.se(s)) return false;
            mb.mvd[j][1] = (int16_t)s;
        }
    } else {
        if (mb.mbType == B200_MB_I_4x4) {
            for (int i = 0; i < 16; i++) {
                if (!br.get1(v)) return false;
                mb.prevFlag[i] = (uint8_t)v;
                if (!v) {
                    if (!br.get(3, v)) return false;
                    mb.remMode[i] = (uint8_t)v;
                }
            }
        }
        if (!br.ue(v) || v > 3) return false;
        mb.chromaMode = v;
    }
    bool i16 = !inter && mb.mbType != B200_MB_I_4x4;
    if (!i16) {
        if (!br.ue(v) || v > 47) return false;
        mb.cbp = kCbp[v][mb.mbType == B200_MB_I_4x4 ? 0 : 1];
    } else {
        uint32_t t = mb.mbType - B200_MB_I_16x16_FIRST;
        uint32_t c = t >> 2;
        if (c > 2) c -= 3;
        mb.cbp = (t >= 12 ? 15u : 0u) + (c << 4);
    }
    if (mb.cbp || i16) {
        if (!br.se(s) || s < -26 || s > 25) return false;
        mb.qpDelta = s;
        stream = br;
        return parseResidual(stream, mb, mbAddr);
    }
    stream = br;
    return true;
}

// DecodeResidual (h264bsd_macroblock_layer.c:700-796)
bool PictureState::parseResidual(BitReader &br, MbSyntax &mb, uint32_t mbAddr) {
    bool i16 = !isInterType(mb.mbType) && mb.mbType != B200_MB_I_4x4;
    uint32_t cbp = mb.cbp;
    if (i16) {
        CavlcResult r = cavlcResidualBlock(br, mb.level[24], nC(mbAddr, 0, mb.totalCoeff), 16);
        if (r.totalCoeff < 0) return false;
        mb.totalCoeff[24] = (uint8_t)r.totalCoeff;
        mb.sumAbs[24] = r.sumAbs;
        if (r.totalCoeff) mb.codedBlocks |= B200_CM_LUMA_DC;
    }
    uint32_t blk = 0;
    for (int i8 = 0; i8 < 4; i8++) {
        if (cbp & (1u << i8)) {
            for (int j = 0; j < 4; j++, blk++) {
                int n = nC(mbAddr, blk, mb.totalCoeff);
                CavlcResult r = i16 ? cavlcResidualBlock(br, mb.level[blk] + 1, n, 15)
                                    : cavlcResidualBlock(br, mb.level[blk], n, 16);
                if (r.totalCoeff < 0) return false;
                mb.totalCoeff[blk] = (uint8_t)r.totalCoeff;
                mb.sumAbs[blk] = r.sumAbs;
                if (r.totalCoeff) mb.codedBlocks |= 1u << blk;
                if (i16 && (r.coeffMap >> 15)) mb.spill |= 1u << (blk + 1);
            }
        } else {
            blk += 4;
        }
    }
    uint32_t chroma = cbp >> 4;
    if (chroma & 3) {
        CavlcResult r = cavlcResidualBlock(br, mb.level[25], -1, 4);
        if (r.totalCoeff < 0) return false;
        mb.totalCoeff[25] = (uint8_t)r.totalCoeff;
        mb.sumAbs[25] = r.sumAbs;
        if (r.totalCoeff) mb.codedBlocks |= B200_CM_CHROMA_DC;
        r = cavlcResidualBlock(br, mb.level[25] + 4, -1, 4);
        if (r.totalCoeff < 0) return false;
        mb.totalCoeff[26] = (uint8_t)r.totalCoeff;
        mb.sumAbs[26] = r.sumAbs;
        if (r.totalCoeff) mb.codedBlocks |= B200_CM_CHROMA_DC;
    }
    if (chroma & 2) {
        for (blk = 16; blk < 24; blk++) {
            CavlcResult r = cavlcResidualBlock(br, mb.level[blk] + 1, nC(mbAddr, blk, mb.totalCoeff), 15);
            if (r.totalCoeff < 0) return false;
            mb.totalCoeff[blk] = (uint8_t)r.totalCoeff;
            mb.sumAbs[blk] = r.sumAbs;
            if (r.totalCoeff) mb.codedBlocks |= 1u << blk;
            if (r.coeffMap >> 15) {
                mb.spill |= 1u << (blk + 1);
                // out of the last chroma block the stray level lands in the Intra16x16 DC block, which is transformed later
                if (blk == 23) mb.sumAbs[24] += (uint32_t)std::abs((int)mb.level[24][0]);
            }
        }
    }
    return true;
}

// GetInterNeighbour (h264bsd_inter_prediction.c:963-988) addressed by 4x4 coordinates relative to
// the current macroblock; curZ = block index of the partition being predicted (blocks of the
// current macroblock that come later in decoding order are "not available", clause 6.4.11.7)
PictureState::NbMv PictureState::interNeighbour(uint32_t cur, int x, int y, int curZ) const {
    NbMv n{false, 0xFFFFFFFFu, {0, 0}};
    int nb;
    if (y >= 0 && x > 3) return n;  // right neighbour: never available
    if (x >= 0 && x <= 3 && y >= 0) {
        int z = kZ[y][x];
        if (z >= curZ) return n;
        nb = (int)cur;
    } else if (y < 0) {
        nb = x < 0 ? curNb_[3] : x > 3 ? curNb_[2] : curNb_[1];
    } else {
        nb = curNb_[0];
    }
    if (nb < 0) return n;   // not in the picture or in another slice
    n.avail = true;
    const b200_mb_rec &r = st[nb];
    if (isInterType(r.mbType)) {
        int z = kZ[y & 3][x & 3];
        n.refIdx = r.refIdx[z >> 2];
        n.mv[0] = r.u.mv[z][0];
        n.mv[1] = r.u.mv[z][1];
    }
    return n;
}

static int median3(int a, int b, int c) {
    int mx = std::max(a, std::max(b, c)), mn = std::min(a, std::min(b, c));
    return a + b + c - mx - mn;
}

// clause 8.4.1.3 (MvPrediction*, GetPredictionMv in h264bsd_inter_prediction.c:494-1026).
// dirHint: 0 median, 1 prefer B (16x8 upper), 2 prefer A (16x8 lower, 8x16 left), 3 prefer C (8x16 right)
bool PictureState::predictMv(uint32_t cur, int x, int y, int w, int h, uint32_t refIdx, int dirHint, int16_t out[2],
                             const NbMv *preA, const NbMv *preB) const {
    (void)h;
    int curZ = kZ[y][x];
    NbMv a = preA ? *preA : interNeighbour(cur, x - 1, y, curZ);
    NbMv b = preB ? *preB : interNeighbour(cur, x, y - 1, curZ);
    NbMv c = interNeighbour(cur, x + w, y - 1, curZ);
    if (!c.avail) c = interNeighbour(cur, x - 1, y - 1, curZ);
    if (dirHint == 1 && b.refIdx == refIdx) { out[0] = b.mv[0]; out[1] = b.mv[1]; return true; }
    if (dirHint == 2 && a.refIdx == refIdx) { out[0] = a.mv[0]; out[1] = a.mv[1]; return true; }
    if (dirHint == 3 && c.refIdx == refIdx) { out[0] = c.mv[0]; out[1] = c.mv[1]; return true; }
    if (b.avail || c.avail || !a.avail) {
        int isA = a.refIdx == refIdx, isB = b.refIdx == refIdx, isC = c.refIdx == refIdx;
        if (isA + isB + isC != 1) {
            out[0] = (int16_t)median3(a.mv[0], b.mv[0], c.mv[0]);
            out[1] = (int16_t)median3(a.mv[1], b.mv[1], c.mv[1]);
        } else if (isA) { out[0] = a.mv[0]; out[1] = a.mv[1]; }
        else if (isB) { out[0] = b.mv[0]; out[1] = b.mv[1]; }
        else { out[0] = c.mv[0]; out[1] = c.mv[1]; }
    } else {
        out[0] = a.mv[0]; out[1] = a.mv[1];
    }
    return true;
}

static inline bool mvInRange(int16_t hor, int16_t ver) {
    return (uint32_t)((int32_t)hor + 8192) < 16384u && (uint32_t)((int32_t)ver + 2048) < 4096u;
}

// motion vectors + reference slots of an inter macroblock (h264bsdInterPrediction's syntax half)
bool PictureState::deriveInter(MbSyntax &mb, uint32_t mbAddr, const Dpb &dpb) {
    b200_mb_rec &r = st[mbAddr];
    auto setMv = [&](int x0, int y0, int w, int h, int16_t hor, int16_t ver) {
        for (int yy = y0; yy < y0 + h; yy++)
            for (int xx = x0; xx < x0 + w; xx++) { r.u.mv[kZ[yy][xx]][0] = hor; r.u.mv[kZ[yy][xx]][1] = ver; }
    };
    r.subMbTypes = 0;
    switch (mb.mbType) {
        case B200_MB_P_16x16: {     // (P_Skip has its own path: finishSkip)
            const uint32_t refIdx = mb.refIdx[0];
            const NbMv a = interNeighbour(mbAddr, -1, 0, 0), b = interNeighbour(mbAddr, 0, -1, 0);
            int16_t p[2];
            predictMv(mbAddr, 0, 0, 4, 4, refIdx, 0, p, &a, &b);
            const int16_t mv[2] = {(int16_t)(mb.mvd[0][0] + p[0]), (int16_t)(mb.mvd[0][1] + p[1])};
            if (!mvInRange(mv[0], mv[1])) return false;
            int slot = dpb.refSlot(refIdx);
            if (slot < 0) return false;
            for (int z = 0; z < 16; z++) { r.u.mv[z][0] = mv[0]; r.u.mv[z][1] = mv[1]; }
            for (int q = 0; q < 4; q++) { r.refIdx[q] = (uint8_t)refIdx; r.refSlot[q] = (uint8_t)slot; }
            break;
        }
        case B200_MB_P_16x8: {
            for (int p = 0; p < 2; p++) {
                uint32_t refIdx = mb.refIdx[p];
                int16_t pr[2];
                predictMv(mbAddr, 0, 2 * p, 4, 2, refIdx, p == 0 ? 1 : 2, pr);
                int16_t hor = (int16_t)(mb.mvd[p][0] + pr[0]), ver = (int16_t)(mb.mvd[p][1] + pr[1]);
                if (!mvInRange(hor, ver)) return false;
                int slot = dpb.refSlot(refIdx);
                if (slot < 0) return false;
                setMv(0, 2 * p, 4, 2, hor, ver);
                r.refIdx[2 * p] = r.refIdx[2 * p + 1] = (uint8_t)refIdx;
                r.refSlot[2 * p] = r.refSlot[2 * p + 1] = (uint8_t)slot;
            }
            break;
        }
        case B200_MB_P_8x16: {
            for (int p = 0; p < 2; p++) {
                uint32_t refIdx = mb.refIdx[p];
                int16_t pr[2];
                predictMv(mbAddr, 2 * p, 0, 2, 4, refIdx, p == 0 ? 2 : 3, pr);
                int16_t hor = (int16_t)(mb.mvd[p][0] + pr[0]), ver = (int16_t)(mb.mvd[p][1] + pr[1]);
                if (!mvInRange(hor, ver)) return false;
                int slot = dpb.refSlot(refIdx);
                if (slot < 0) return false;
                setMv(2 * p, 0, 2, 4, hor, ver);
                r.refIdx[p] = r.refIdx[p + 2] = (uint8_t)refIdx;
                r.refSlot[p] = r.refSlot[p + 2] = (uint8_t)slot;
            }
            break;
        }
        default: {  // P_8x8 / P_8x8ref0
            for (int q = 0; q < 4; q++) {
                uint32_t refIdx = mb.refIdx[q];
                int slot = dpb.refSlot(refIdx);
                r.refIdx[q] = (uint8_t)refIdx;
                if (slot < 0) return false;
                r.refSlot[q] = (uint8_t)slot;
                r.subMbTypes |= (uint8_t)(mb.subType[q] << (2 * q));
                int qx = (q & 1) * 2, qy = (q >> 1) * 2;
                int sw = (mb.subType[q] == 0 || mb.subType[q] == 1) ? 2 : 1;
                int shh = (mb.subType[q] == 0 || mb.subType[q] == 2) ? 2 : 1;
                int n = numSubMbPart(mb.subType[q]);
                for (int j = 0; j < n; j++) {
                    int sx = qx, sy = qy;
                    if (mb.subType[q] == 1) sy += j;            // 8x4: stacked
                    else if (mb.subType[q] == 2) sx += j;       // 4x8: side by side
                    else if (mb.subType[q] == 3) { sx += j & 1; sy += j >> 1; }
                    int16_t pr[2];
                    predictMv(mbAddr, sx, sy, sw, shh, refIdx, 0, pr);
                    int16_t hor = (int16_t)(mb.subMvd[q][j][0] + pr[0]), ver = (int16_t)(mb.subMvd[q][j][1] + pr[1]);
                    if (!mvInRange(hor, ver)) return false;
                    setMv(sx, sy, sw, shh, hor, ver);
                }
            }
            break;
        }
    }
    return true;
}

// Intra availability, Intra4x4 mode derivation and the "mode needs an unavailable neighbour"
// checks of h264bsdIntra16x16Prediction / Intra4x4Prediction / IntraChromaPrediction
bool PictureState::deriveIntra(MbSyntax &mb, uint32_t mbAddr, bool constrainedIntra) {
    b200_mb_rec &r = st[mbAddr];
    auto availIntra = [&](int nb) {
        if (nb < 0) return false;
        if (constrainedIntra && isInterType(st[nb].mbType)) return false;
        return true;
    };
    int a = curNb_[0], b = curNb_[1], c = curNb_[2], d = curNb_[3];
    bool avA = availIntra(a), avB = availIntra(b), avC = availIntra(c), avD = availIntra(d);
    r.flags = (uint8_t)((avA ? B200_MBF_AVAIL_A : 0) | (avB ? B200_MBF_AVAIL_B : 0) |
                        (avC ? B200_MBF_AVAIL_C : 0) | (avD ? B200_MBF_AVAIL_D : 0));
    r.intraChromaMode = (uint8_t)mb.chromaMode;
    std::memset(&r.u, 0, sizeof r.u);
    if (mb.mbType == B200_MB_I_4x4) {
        for (int blk = 0; blk < 16; blk++) {
            int x = kBlkX[blk], y = kBlkY[blk];
            // neighbouring 4x4 blocks A (left) and B (above): MB + block index
            bool bA, bB, bD;
            int modeA = 2, modeB = 2;
            if (x > 0) { bA = true; modeA = r.u.intra.i4x4Mode[kZ[y][x - 1]]; }
            else { bA = avA; if (bA && st[a].mbType == B200_MB_I_4x4) { modeA = st[a].u.intra.i4x4Mode[kZ[y][3]]; } }
            if (y > 0) { bB = true; modeB = r.u.intra.i4x4Mode[kZ[y - 1][x]]; }
            else { bB = avB; if (bB && st[b].mbType == B200_MB_I_4x4) { modeB = st[b].u.intra.i4x4Mode[kZ[3][x]]; } }
            int mode;
            if (!(bA && bB)) mode = 2;
            else mode = std::min(modeA, modeB);
            if (!mb.prevFlag[blk]) mode = mb.remMode[blk] < mode ? mb.remMode[blk] : mb.remMode[blk] + 1;
            r.u.intra.i4x4Mode[blk] = (uint8_t)mode;
            // D (above-left) availability, N_D_4x4B (neighbour.c:92-98); C only matters for pels
            if (x == 0 && y == 0) bD = avD;
            else if (x == 0) bD = avA;
            else if (y == 0) bD = avB;
            else bD = true;
            switch (mode) {
                case 0: case 3: case 7: if (!bB) return false; break;
                case 1: case 8: if (!bA) return false; break;
                case 2: break;
                default: if (!bA || !bB || !bD) return false; break;  // 4,5,6
            }
        }
    } else {
        uint32_t predMode = (mb.mbType - B200_MB_I_16x16_FIRST) & 3;
        switch (predMode) {
            case 0: if (!avB) return false; break;
            case 1: if (!avA) return false; break;
            case 2: break;
            default: if (!avA || !avB || !avD) return false; break;
        }
    }
    switch (mb.chromaMode) {
        case 0: break;
        case 1: if (!avA) return false; break;
        case 2: if (!avB) return false; break;
        default: if (!avA || !avB || !avD) return false; break;
    }
    return true;
}

// ---- the [-512, 511] check of the inverse transform (h264bsdProcessBlock, h264bsd_transform.c:183-188,:205-206,:229-233) ----
// The reference runs the transform while it parses, and a residual sample outside the range makes the macroblock -- hence the
// slice -- corrupt: h264bsdMarkSliceCorrupted, concealment.  The pels are the GPU's business here, but which macroblocks count
// as decoded is not, so the host has to know the outcome.  Almost always a bound on the magnitudes settles it: every output is
// (sum of +-1 / +-1/2 weighted dequantised coefficients + 32) >> 6, so sum |level| * largest scale <= 32000 cannot leave the
// range.  Only a macroblock that fails the bound gets the exact transform below (same arithmetic as the kernels and the oracle).
namespace {
const uint8_t kZigzag4x4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
const int kLevelScale[6][3] = {{10, 13, 16}, {11, 14, 18}, {13, 16, 20}, {14, 18, 23}, {16, 20, 25}, {18, 23, 29}};
const uint8_t kScaleClass[16] = {0, 1, 0, 1, 1, 2, 1, 2, 0, 1, 0, 1, 1, 2, 1, 2};

// h264bsdProcessBlock: true if every residual sample of the block stays inside [-512, 511]
bool blockInRange(const int16_t *lev, int qp, bool dcPreset, int dcValue) {
    int d[16];
    const int qpDiv = qp / 6, qpMod = qp % 6;
    for (int i = 0; i < 16; i++) {
        const int r = kZigzag4x4[i];
        d[r] = lev[i] * (kLevelScale[qpMod][kScaleClass[r]] << qpDiv);
    }
    if (dcPreset) d[0] = dcValue;
    for (int i = 0; i < 16; i += 4) {
        const int t0 = d[i] + d[i + 2], t1 = d[i] - d[i + 2];
        const int t2 = (d[i + 1] >> 1) - d[i + 3], t3 = d[i + 1] + (d[i + 3] >> 1);
        d[i] = t0 + t3; d[i + 1] = t1 + t2; d[i + 2] = t1 - t2; d[i + 3] = t0 - t3;
    }
    for (int i = 0; i < 4; i++) {
        const int t0 = d[i] + d[i + 8], t1 = d[i] - d[i + 8];
        const int t2 = (d[i + 4] >> 1) - d[i + 12], t3 = d[i + 4] + (d[i + 12] >> 1);
        const int o[4] = {(t0 + t3 + 32) >> 6, (t1 + t2 + 32) >> 6, (t1 - t2 + 32) >> 6, (t0 - t3 + 32) >> 6};
        for (int k = 0; k < 4; k++)
            if ((unsigned)(o[k] + 512) > 1023u) return false;
    }
    return true;
}
// h264bsdProcessLumaDc (h264bsd_transform.c:255-338), dc[] in raster order of the 4x4 blocks
void lumaDcValues(const int16_t *lev, int qp, int dc[16]) {
    int d[16];
    const int qpDiv = qp / 6, ls = kLevelScale[qp % 6][0];
    for (int i = 0; i < 16; i++) d[kZigzag4x4[i]] = lev[i];
    for (int i = 0; i < 16; i += 4) {
        const int t0 = d[i] + d[i + 2], t1 = d[i] - d[i + 2], t2 = d[i + 1] - d[i + 3], t3 = d[i + 1] + d[i + 3];
        d[i] = t0 + t3; d[i + 1] = t1 + t2; d[i + 2] = t1 - t2; d[i + 3] = t0 - t3;
    }
    for (int i = 0; i < 4; i++) {
        const int t0 = d[i] + d[i + 8], t1 = d[i] - d[i + 8], t2 = d[i + 4] - d[i + 12], t3 = d[i + 4] + d[i + 12];
        const int v[4] = {t0 + t3, t1 + t2, t1 - t2, t0 - t3};
        for (int k = 0; k < 4; k++)
            dc[i + 4 * k] = qp >= 12 ? v[k] * (ls << (qpDiv - 2)) : (v[k] * ls + (qpDiv == 1 ? 1 : 2)) >> (2 - qpDiv);
    }
}
// h264bsdProcessChromaDc (h264bsd_transform.c:359-401), Cb in lev[0..3], Cr in lev[4..7]
void chromaDcValues(const int16_t *lev, int qp, int dc[8]) {
    int ls = kLevelScale[qp % 6][0], shift = 1;
    if (qp >= 6) { ls <<= (qp / 6 - 1); shift = 0; }
    for (int c = 0; c < 8; c += 4) {
        const int t0 = lev[c] + lev[c + 2], t1 = lev[c] - lev[c + 2], t2 = lev[c + 1] - lev[c + 3], t3 = lev[c + 1] + lev[c + 3];
        dc[c] = ((t0 + t3) * ls) >> shift;
        dc[c + 1] = ((t0 - t3) * ls) >> shift;
        dc[c + 2] = ((t1 + t2) * ls) >> shift;
        dc[c + 3] = ((t1 - t2) * ls) >> shift;
    }
}
}  // namespace

// ProcessResidual (h264bsd_macroblock_layer.c:1340-1421) as far as its error return goes
bool PictureState::residualInRange(const MbSyntax &mb, bool i16, int qpY, int qpC, uint32_t mask) const {
    // the bound
    const uint64_t sL = (uint64_t)29 << (qpY / 6), sC = (uint64_t)29 << (qpC / 6);
    const uint64_t dcL = (mask & B200_CM_LUMA_DC) ? (((uint64_t)mb.sumAbs[24] * ((uint64_t)18 << (qpY / 6))) >> 2) + 2 : 0;
    uint64_t dcCb = 0, dcCr = 0;
    if (mask & B200_CM_CHROMA_DC) {
        dcCb = (((uint64_t)mb.sumAbs[25] * ((uint64_t)18 << (qpC / 6))) >> 1) + 2;
        dcCr = (((uint64_t)mb.sumAbs[26] * ((uint64_t)18 << (qpC / 6))) >> 1) + 2;
        if (!mb.totalCoeff[25]) dcCb = 0;
        if (!mb.totalCoeff[26]) dcCr = 0;
    }
    uint64_t worst = std::max(dcL, std::max(dcCb, dcCr));
    for (uint32_t m = mask & 0xFFFFFFu; m; m &= m - 1) {
        const int b = __builtin_ctz(m);
        const uint64_t v = b < 16 ? dcL + mb.sumAbs[b] * sL : (b < 20 ? dcCb : dcCr) + mb.sumAbs[b] * sC;
        worst = std::max(worst, v);
    }
    if (worst <= 32000) return true;

    // the exact transform
    int lumaDc[16] = {0}, chromaDc[8] = {0};
    if (mask & B200_CM_LUMA_DC) lumaDcValues(mb.level[24], qpY, lumaDc);
    if (mask & B200_CM_CHROMA_DC) chromaDcValues(mb.level[25], qpC, chromaDc);
    static const int16_t zero[16] = {0};
    for (int b = 0; b < 24; b++) {
        const bool coded = (mask >> b) & 1;
        const bool dcPreset = b < 16 ? i16 : true;
        const int dcVal = b < 16 ? lumaDc[kBlkY[b] * 4 + kBlkX[b]] : chromaDc[b - 16];
        if (!coded && !(dcPreset && dcVal)) continue;
        if (!blockInRange(coded ? mb.level[b] : zero, b < 16 ? qpY : qpC, dcPreset, dcVal)) return false;
    }
    return true;
}

// Class of a macroblock (0 other pass-A, 1 zero-vector copy without residual, 4 pass-B: intra-predicted), its deblocking edge flags (GetMbFilteringFlags, deblocking.c:289-320) and, for an intra-predicted
// macroblock, the neighbours it has to wait for inside the intra pass -- settled right after the record is written, while it
// is in cache.  Everything it looks at is final at this point: a left / upper neighbour of the same slice has been decoded
// before this macroblock (curNb_), one that is decoded later belongs to a later slice.  Pictures where that does not hold
// (a macroblock decoded twice by redundant slices, corrupted slices, concealment) take the full pass of finalizeRecords.
inline void PictureState::classify(uint32_t a, b200_mb_rec &r) {
    uint8_t *cls = orderClass.data();
    const uint32_t idc = r.reserved0;
    uint8_t f = r.flags & (uint8_t)(B200_MBF_AVAIL_A | B200_MBF_AVAIL_B | B200_MBF_AVAIL_C | B200_MBF_AVAIL_D);
    if (idc != 1) {
        f |= B200_MBF_FILTER_INNER;
        if (curX_ && (idc != 2 || curNb_[0] >= 0)) f |= B200_MBF_FILTER_LEFT;
        if (a >= widthMbs && (idc != 2 || curNb_[1] >= 0)) f |= B200_MBF_FILTER_TOP;
    }
    uint8_t c = 0, w = 0;
    if (r.mbType > B200_MB_P_8x8REF0 && r.mbType != B200_MB_I_PCM) {
        c = 4;
        numIntraPred_++;
        if ((f & B200_MBF_AVAIL_A) && cls[a - 1] == 4) w |= B200_MBF_AVAIL_A;
        if ((f & B200_MBF_AVAIL_B) && cls[a - widthMbs] == 4) w |= B200_MBF_AVAIL_B;
        if ((f & B200_MBF_AVAIL_C) && cls[a - widthMbs + 1] == 4) w |= B200_MBF_AVAIL_C;
        if ((f & B200_MBF_AVAIL_D) && cls[a - widthMbs - 1] == 4) w |= B200_MBF_AVAIL_D;
    } else if (r.mbType <= B200_MB_P_16x16 && r.codedMask == 0 && (r.u.mv[0][0] | r.u.mv[0][1]) == 0) {
        c = 1;
        numCopies_++;
    }
    r.flags = f;
    r.waitMask = w;
    cls[a] = c;
}

// the syntax-level half of h264bsdDecodeMacroblock (h264bsd_macroblock_layer.c:965-1131)
bool PictureState::finishMacroblock(MbSyntax &mb, uint32_t mbAddr, int &qpY, const SliceHeader &sh, const Pps &pps, const Dpb &dpb) {
    b200_mb_rec &r = st[mbAddr];
    MbAux &ax = aux[mbAddr];
    const uint32_t type = mb.mbType;
    ax.decoded++;
    const bool first = ax.decoded == 1;
    // every byte of the record is set (it may live in memory nobody cleared).  SetMbParams (h264bsd_slice_data.c:254-273):
    // the slice id is set by the caller
    r.mbType = (uint8_t)type;
    r.reserved0 = (uint8_t)sh.disableDeblockingFilterIdc;
    r.filterOffsetA = (int8_t)sh.alphaOffset;
    r.filterOffsetB = (int8_t)sh.betaOffset;
    r.chromaQpIndexOffset = (int8_t)pps.chromaQpIndexOffset;
    r.coefIndex = (uint32_t)(coefs.size() / 16);
    r.waitMask = 0;
    r.reserved1[0] = r.reserved1[1] = r.reserved1[2] = 0;
    r.flags = 0;
    r.intraChromaMode = 0;
    r.subMbTypes = 0;

    // (P_Skip never comes here: finishSkip)
    if (type == B200_MB_I_PCM) {
        std::memset(r.refSlot, 0, 4);
        std::memset(r.refIdx, 0, 4);
        r.qpY = 0;
        r.qpC = kQpC[std::min(51, std::max(0, 0 + pps.chromaQpIndexOffset))];
        std::memset(ax.totalCoeff, 16, 24);
        std::memset(&r.u, 0, sizeof r.u);
        r.codedMask = 0xFFFFFFu;
        if (!first) { r.coefIndex = 0; return true; }
        std::memcpy(coefs.grow(12 * 16), mb.pcm, 384);
        if (recs != st) recs[mbAddr] = r;
        if (!lateFixup_) classify(mbAddr, recs[mbAddr]);
        return true;
    }

    std::memcpy(ax.totalCoeff, mb.totalCoeff, 27);
    if (mb.qpDelta) {
        qpY += mb.qpDelta;
        if (qpY < 0) qpY += 52;
        else if (qpY >= 52) qpY -= 52;
    }
    r.qpY = (uint8_t)qpY;
    r.qpC = kQpC[std::min(51, std::max(0, qpY + pps.chromaQpIndexOffset))];

    const bool inter = isInterType(type);
    const bool i16 = !inter && type != B200_MB_I_4x4;
    uint32_t mask = mb.codedBlocks;      // blocks with TotalCoeff != 0, collected by parseResidual
    if (!i16) mask &= ~B200_CM_LUMA_DC;
    r.codedMask = mask;
    if (mask && !residualInRange(mb, i16, qpY, r.qpC, mask)) return false;

    if (inter) {
        if (!deriveInter(mb, mbAddr, dpb)) return false;
    } else {
        std::memset(r.refSlot, 0, 4);
        std::memset(r.refIdx, 0, 4);
        if (!deriveIntra(mb, mbAddr, pps.constrainedIntraPred)) return false;
    }
    if (!first) { r.coefIndex = 0; return true; }

    // emit coefficients: [luma DC][chroma DC][coded blocks ascending]
    if (mask) {
        int16_t *dst = coefs.grow((size_t)__builtin_popcount(mask) * 16);
        if (mask & B200_CM_LUMA_DC) { std::memcpy(dst, mb.level[24], 32); dst += 16; }
        if (mask & B200_CM_CHROMA_DC) { std::memcpy(dst, mb.level[25], 16); std::memset(dst + 8, 0, 16); dst += 16; }
        for (uint32_t m = mask & 0xFFFFFFu; m; m &= m - 1) {
            const int b = __builtin_ctz(m);
            std::memcpy(dst, mb.level[b], 32);
            if (b >= 16 || i16) dst[0] = 0;     // AC block: its DC comes from the DC transform (a stray entry there is never used)
            dst += 16;
        }
    }
    if (recs != st) recs[mbAddr] = r;
    if (!lateFixup_) classify(mbAddr, recs[mbAddr]);
    return true;
}

// P_Skip, the most frequent macroblock: what finishMacroblock + deriveInter do for it, without the general machinery.
// slot0 = frame slot of reference index 0 in this slice's list (-1: no such picture -> error, inter_prediction.c:527-531).
inline bool PictureState::finishSkip(uint32_t mbAddr, int qpY, const SliceHeader &sh, const Pps &pps, int slot0) {
    b200_mb_rec &r = st[mbAddr];
    MbAux &ax = aux[mbAddr];
    ax.decoded++;
    const bool first = ax.decoded == 1;
    std::memset(ax.totalCoeff, 0, 27);
    r.mbType = B200_MB_P_SKIP;
    r.qpY = (uint8_t)qpY;
    r.qpC = kQpC[std::min(51, std::max(0, qpY + pps.chromaQpIndexOffset))];
    r.flags = 0;
    r.codedMask = 0;
    r.coefIndex = (uint32_t)(coefs.size() / 16);
    r.filterOffsetA = (int8_t)sh.alphaOffset;
    r.filterOffsetB = (int8_t)sh.betaOffset;
    r.chromaQpIndexOffset = (int8_t)pps.chromaQpIndexOffset;
    r.subMbTypes = 0;
    r.intraChromaMode = 0;
    r.reserved0 = (uint8_t)sh.disableDeblockingFilterIdc;
    r.waitMask = 0;
    r.reserved1[0] = r.reserved1[1] = r.reserved1[2] = 0;
    // clause 8.4.1.1: zero vector unless both neighbours A and B exist and neither is (reference 0, vector 0)
    uint32_t mvWord = 0;
    const int a = curNb_[0], b = curNb_[1];
    if (a >= 0 && b >= 0) {
        const b200_mb_rec &ra = st[a], &rb = st[b];
        // A: the 4x4 block right of the left edge in row 0 (block 5 of A); B: block below the top edge, column 0 (block 10 of B)
        uint32_t mvA, mvB;
        std::memcpy(&mvA, ra.u.mv[5], 4);
        std::memcpy(&mvB, rb.u.mv[10], 4);
        const bool interA = isInterType(ra.mbType), interB = isInterType(rb.mbType);
        const bool zero = (interA && ra.refIdx[1] == 0 && mvA == 0) || (interB && rb.refIdx[2] == 0 && mvB == 0);
        if (!zero) {
            NbMv na{true, interA ? ra.refIdx[1] : 0xFFFFFFFFu, {0, 0}}, nb{true, interB ? rb.refIdx[2] : 0xFFFFFFFFu, {0, 0}};
            if (interA) { na.mv[0] = ra.u.mv[5][0]; na.mv[1] = ra.u.mv[5][1]; }
            if (interB) { nb.mv[0] = rb.u.mv[10][0]; nb.mv[1] = rb.u.mv[10][1]; }
            int16_t p[2];
            predictMv(mbAddr, 0, 0, 4, 4, 0, 0, p, &na, &nb);
            if (!mvInRange(p[0], p[1])) return false;
            std::memcpy(&mvWord, p, 4);
        }
    }
    if (slot0 < 0) return false;
    uint32_t *mv = reinterpret_cast<uint32_t *>(&r.u);
    for (int z = 0; z < 16; z++) mv[z] = mvWord;
    std::memset(r.refIdx, 0, 4);
    std::memset(r.refSlot, slot0, 4);
    if (first && recs != st) recs[mbAddr] = r;
    if (first && !lateFixup_) classify(mbAddr, recs[mbAddr]);
    return true;
}

// h264bsdDecodeSliceData (h264bsd_slice_data.c:86-232)
SliceResult PictureState::decodeSlice(BitReader &br, const SliceHeader &sh, const Sps &sps, const Pps &pps, const Dpb &dpb) {
    (void)sps;
    static thread_local MbSyntax mb;
    static thread_local bool mbInit = false;
    if (!mbInit) { std::memset(&mb, 0, sizeof mb); mbInit = true; }

    bindOutput();
    failedMb_ = -1;
    if (sh.redundantPicCnt) splitState();   // a macroblock decoded again updates the state, not the record (first decode wins)
    uint32_t cur = sh.firstMb;
    uint32_t skipRun = 0;
    bool prevSkipped = false;
    sliceIdCounter++;
    lastMbAddr = 0;
    uint32_t mbCount = 0;
    int qpY = (int)pps.picInitQp + sh.sliceQpDelta;
    const bool iSlice = sh.isI();
    const uint16_t sid = (uint16_t)sliceIdCounter;
    uint32_t x = cur % widthMbs;
    const int slot0 = iSlice ? -1 : dpb.refSlot(0);
    bool more;
    do {
        MbAux &ax = aux[cur];
        if (!sh.redundantPicCnt && ax.decoded) return SliceResult::Error;
        ax.sliceId = sid;
        st[cur].sliceId = sid;
        {
            // neighbours A, B, C, D of this macroblock that exist and belong to the same slice (h264bsdInitMbNeighbours +
            // h264bsdIsNeighbourAvailable, neighbour.c:128-176, :370-382), resolved once per macroblock
            const bool up = cur >= widthMbs;
            const MbAux *am = aux.data();
            curX_ = x;
            curNb_[0] = (x && am[cur - 1].sliceId == sid) ? (int)cur - 1 : -1;
            curNb_[1] = (up && am[cur - widthMbs].sliceId == sid) ? (int)(cur - widthMbs) : -1;
            curNb_[2] = (up && x + 1 < widthMbs && am[cur - widthMbs + 1].sliceId == sid) ? (int)(cur - widthMbs + 1) : -1;
            curNb_[3] = (up && x && am[cur - widthMbs - 1].sliceId == sid) ? (int)(cur - widthMbs - 1) : -1;
        }
        if (!iSlice && !prevSkipped) {
            if (!br.ue(skipRun)) return SliceResult::Error;
            if (skipRun > picSizeInMbs - cur) return SliceResult::Error;
            if (skipRun) prevSkipped = true;
        }
        bool ok;
        if (skipRun) {
            skipRun--;
            ok = finishSkip(cur, qpY, sh, pps, slot0);
        } else {
            prevSkipped = false;
            if (!parseMacroblockLayer(br, mb, cur, iSlice, sh.numRefIdxL0Active)) {
                std::memset(mb.level, 0, sizeof mb.level);
                return SliceResult::Error;
            }
            ok = finishMacroblock(mb, cur, qpY, sh, pps, dpb);
            // the level arrays go back to all-zero for the next macroblock: only the blocks that received levels
            if (mb.mbType != B200_MB_I_PCM) {
                for (uint32_t m = mb.codedBlocks; m; m &= m - 1) std::memset(mb.level[__builtin_ctz(m)], 0, 32);
                for (uint32_t m = mb.spill; m; m &= m - 1) mb.level[__builtin_ctz(m)][0] = 0;
            }
        }
        if (!ok) {
            // the macroblock counts as decoded once more, but its record is half-made: see markSliceCorrupted
            failedMb_ = (int)cur;
            failedPrevDecoded_ = (uint8_t)(ax.decoded - 1);
            return SliceResult::Error;
        }
        if (ax.decoded == 1) mbCount++;
        more = skipRun || br.moreRbspData();
        if (iSlice) lastMbAddr = cur;
        const uint32_t next = nextMbAddress(cur);
        x = (next == cur + 1) ? (x + 1 == widthMbs ? 0 : x + 1) : next % widthMbs;
        cur = next;
        if (more && !cur) return SliceResult::Error;
    } while (more);
    if (numDecodedMbs + mbCount > picSizeInMbs) return SliceResult::Error;
    numDecodedMbs += mbCount;
    return SliceResult::Ok;
}

// h264bsdMarkSliceCorrupted (h264bsd_slice_data.c:298-354)
void PictureState::markSliceCorrupted(uint32_t firstMbInSlice, const Sps &sps) {
    lateFixup_ = true;
    uint32_t cur = firstMbInSlice;
    uint32_t sliceId = sliceIdCounter;
    if (lastMbAddr) {
        uint32_t i = lastMbAddr - 1, tmp = 0;
        while (i > cur) {
            if (aux[i].sliceId == sliceId) {
                tmp++;
                if (tmp >= std::max<uint32_t>(sps.widthMbs, 10)) break;
            }
            i--;
        }
        cur = i;
    }
    do {
        if (aux[cur].sliceId == sliceId && aux[cur].decoded) aux[cur].decoded--;
        else break;
        cur = nextMbAddress(cur);
    } while (cur);
    // An I slice that fails in the macroblock right after its first one: the walk above starts below the slice
    // (lastMbAddr - 1 < firstMbInSlice) and stops at once, so the reference leaves the FAILED macroblock marked as decoded;
    // nothing of it was written (h264bsd_macroblock_layer.c:1118-1130), it keeps whatever the frame buffer held.  There is no
    // such thing as "what the buffer held" here and the failed macroblock's record is half-made: it is given up like the rest
    // of its slice and concealed (documented deviation).
    if (failedMb_ >= 0) {
        if (aux[failedMb_].decoded > failedPrevDecoded_) aux[failedMb_].decoded = failedPrevDecoded_;
        failedMb_ = -1;
    }
}

void PictureState::finalizeRecords() {
    bindOutput();
    // Which pass a macroblock belongs to -- 0 / 1 pass A (inter, I_PCM, concealed copy; 1 = a zero-vector copy without residual),
    // 4 pass B (intra-predicted), 5 spatially concealed -- and its deblocking edge flags (GetMbFilteringFlags,
    // deblocking.c:289-320) were settled per macroblock by classify().  A picture with macroblocks decoded twice, corrupted
    // slices or concealment takes the pass over the records below instead (the slice ids are final only now).  The GPU reads the
    // records in raster order and sorts the macroblocks itself: the only list left is the concealment order.
    std::vector<uint8_t> &cls = orderClass;
    cls.resize(2 * (size_t)picSizeInMbs);
    uint32_t nB = numIntraPred_, nCopy = numCopies_;
    if (lateFixup_) {
        nB = nCopy = 0;
        uint32_t a = 0;
        for (uint32_t y = 0; y < heightMbs; y++)
            for (uint32_t x = 0; x < widthMbs; x++, a++) {
                b200_mb_rec &r = recs[a];
                const uint32_t idc = r.reserved0;
                uint8_t f = r.flags & (uint8_t)(B200_MBF_AVAIL_A | B200_MBF_AVAIL_B | B200_MBF_AVAIL_C | B200_MBF_AVAIL_D | B200_MBF_CONCEALED);
                if (idc != 1) {
                    f |= B200_MBF_FILTER_INNER;
                    if (x && (idc != 2 || aux[a - 1].sliceId == aux[a].sliceId)) f |= B200_MBF_FILTER_LEFT;
                    if (y && (idc != 2 || aux[a - widthMbs].sliceId == aux[a].sliceId)) f |= B200_MBF_FILTER_TOP;
                }
                r.flags = f;
                r.sliceId = aux[a].sliceId;
                if (st != recs) {
                    // a macroblock decoded again by a redundant slice: the filter sees the state of its LAST decode (st), with the
                    // edge flags that state asks for
                    b200_mb_rec &fr = st[a];
                    const uint32_t idcF = fr.reserved0;
                    uint8_t ff = fr.flags & (uint8_t)(B200_MBF_AVAIL_A | B200_MBF_AVAIL_B | B200_MBF_AVAIL_C | B200_MBF_AVAIL_D | B200_MBF_CONCEALED);
                    if (idcF != 1) {
                        ff |= B200_MBF_FILTER_INNER;
                        if (x && (idcF != 2 || aux[a - 1].sliceId == aux[a].sliceId)) ff |= B200_MBF_FILTER_LEFT;
                        if (y && (idcF != 2 || aux[a - widthMbs].sliceId == aux[a].sliceId)) ff |= B200_MBF_FILTER_TOP;
                    }
                    fr.flags = ff;
                    fr.sliceId = aux[a].sliceId;
                }
                uint8_t c = 0;
                if ((f & B200_MBF_CONCEALED) && r.mbType == B200_MB_I_4x4) {
                    // concealed: a zero-vector copy of the reference picture (pass A), or (class 5) spatial: concealKernel
                    c = r.waitMask == 0 ? 1 : 5;
                    nCopy += c == 1;
                    cls[a] = c;
                    continue;
                }
                r.waitMask = 0;
                if (r.mbType > B200_MB_P_8x8REF0 && r.mbType != B200_MB_I_PCM) { c = 4; nB++; }
                else if (r.mbType <= B200_MB_P_16x16 && r.codedMask == 0 && (r.u.mv[0][0] | r.u.mv[0][1]) == 0) { c = 1; nCopy++; }
                cls[a] = c;
            }
        // the neighbours an intra macroblock has to wait for inside the intra pass: the available ones that are intra-predicted
        // themselves
        for (a = 0; a < picSizeInMbs; a++) {
            if (cls[a] != 4) continue;
            b200_mb_rec &r = recs[a];
            uint8_t w = 0;
            if ((r.flags & B200_MBF_AVAIL_A) && cls[a - 1] == 4) w |= B200_MBF_AVAIL_A;
            if ((r.flags & B200_MBF_AVAIL_B) && cls[a - widthMbs] == 4) w |= B200_MBF_AVAIL_B;
            if ((r.flags & B200_MBF_AVAIL_C) && cls[a - widthMbs + 1] == 4) w |= B200_MBF_AVAIL_C;
            if ((r.flags & B200_MBF_AVAIL_D) && cls[a - widthMbs - 1] == 4) w |= B200_MBF_AVAIL_D;
            r.waitMask = w;
        }
    }
    numPassB = nB;
    numCopy = nCopy;
    // spatially concealed macroblocks, in concealment order
    numConceal = lateFixup_ ? (uint32_t)concealOrder.size() : 0;
    numPassA = picSizeInMbs - nB - numConceal;
    order.assign(concealOrder.begin(), concealOrder.begin() + numConceal);
}

// h264bsdConceal (h264bsd_conceal.c:124-262): records for the macroblocks of the picture that never arrived (or belonged to a
// slice marked corrupted), in the reference's order -- the row of the first decoded macroblock leftwards, then rightwards,
// the rows above it column by column upwards, the rows below in raster order.  What each record says is described in
// h264bsd_b200_tape.h ("Concealed macroblocks").  Returns the number of concealed macroblocks.
uint32_t PictureState::concealMissing(const Dpb &dpb, bool pSlice) {
    bindOutput();
    lateFixup_ = true;
    concealOrder.clear();
    // the reference picture with the smallest available index (:146-157); intraConcealmentFlag is never set in the reference
    int slot = -1;
    if (pSlice)
        for (uint32_t i = 0; i < 16 && slot < 0; i++) slot = dpb.refSlot(i);

    uint32_t first = 0;
    while (first < picSizeInMbs && !aux[first].decoded) first++;
    uint32_t n = 0;
    if (first == picSizeInMbs) {
        // nothing of the picture arrived (:172-201): previous picture or grey, and no filtering at all
        for (uint32_t a = 0; a < picSizeInMbs; a++) {
            b200_mb_rec r;
            std::memset(&r, 0, sizeof r);
            r.flags = B200_MBF_CONCEALED;
            r.reserved0 = 1;
            r.qpY = st[a].qpY;   // (not looked at: the filter is off for every macroblock of the picture)
            r.qpC = st[a].qpC;
            r.coefIndex = (uint32_t)(coefs.size() / 16);
            if (slot >= 0) {
                r.mbType = B200_MB_P_SKIP;
                std::memset(r.refSlot, slot, 4);
            } else {
                r.mbType = B200_MB_I_PCM;
                r.codedMask = 0xFFFFFFu;
                std::memset(coefs.grow(12 * 16), 128, 384);
            }
            st[a] = r;
            if (recs != st) recs[a] = r;
            aux[a].decoded = 1;
        }
        return picSizeInMbs;
    }

    auto conceal = [&](uint32_t row, uint32_t col) {
        const uint32_t a = row * widthMbs + col;
        b200_mb_rec r;
        std::memset(&r, 0, sizeof r);
        r.mbType = B200_MB_I_4x4;           // ConcealMb :296-306: what the in-loop filter will see
        r.qpY = 40;
        r.qpC = kQpC[40];
        r.flags = B200_MBF_CONCEALED;
        r.reserved0 = 0;
        if (slot >= 0) {
            std::memset(r.refSlot, slot, 4);    // :320-341 zero-vector prediction from that picture
        } else {
            uint8_t w = 0;                      // :346-420 the neighbours that are decoded (or concealed) by now
            if (row && aux[a - widthMbs].decoded) w |= B200_CN_ABOVE;
            if (row != heightMbs - 1 && aux[a + widthMbs].decoded) w |= B200_CN_BELOW;
            if (col && aux[a - 1].decoded) w |= B200_CN_LEFT;
            if (col != widthMbs - 1 && aux[a + 1].decoded) w |= B200_CN_RIGHT;
            r.waitMask = w;
            r.coefIndex = (uint32_t)concealOrder.size();
            concealOrder.push_back((uint16_t)a);
        }
        st[a] = r;
        if (recs != st) recs[a] = r;
        aux[a].decoded = 1;
        n++;
    };
    const uint32_t row0 = first / widthMbs, col0 = first % widthMbs;
    for (uint32_t j = col0; j--;) conceal(row0, j);
    for (uint32_t j = col0 + 1; j < widthMbs; j++)
        if (!aux[row0 * widthMbs + j].decoded) conceal(row0, j);
    if (row0)
        for (uint32_t j = 0; j < widthMbs; j++)
            for (uint32_t i = row0; i--;) conceal(i, j);
    for (uint32_t i = row0 + 1; i < heightMbs; i++)
        for (uint32_t j = 0; j < widthMbs; j++)
            if (!aux[i * widthMbs + j].decoded) conceal(i, j);
    return n;
}

}  // namespace b200
