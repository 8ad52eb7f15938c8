// params.cpp -- see params.hpp
#include "params.hpp"
#include <algorithm>

namespace b200 {

static unsigned log2Exact(uint32_t v) {  // v is a power of two >= 16
    unsigned i = 0;
    while (v >> i) i++;
    return i - 1;
}

// h264bsdDecodeNalUnit (h264bsd_nal_unit.c:69-118)
bool parseNalHeader(BitReader &br, NalHeader &nal) {
    uint32_t f, t;
    if (!br.get1(f)) return false;  // forbidden_zero_bit: not checked (errors ignored)
    if (!br.get(2, nal.refIdc)) return false;
    if (!br.get(5, t)) return false;
    nal.type = t;
    if (t == 2 || t == 3 || t == 4) return false;  // data partitioning unsupported
    if ((t == NAL_SPS || t == NAL_PPS || t == NAL_SLICE_IDR) && nal.refIdc == 0) return false;
    if ((t == NAL_SEI || t == NAL_AUD || t == NAL_END_SEQ || t == NAL_END_STREAM || t == NAL_FILLER) &&
        nal.refIdc != 0)
        return false;
    return true;
}

// GetDpbSize (h264bsd_seq_param_set.c:380-470): MaxDPB (bytes) of Table A-1 / frame bytes, max 16
static uint32_t levelDpbSize(uint32_t picSizeInMbs, uint32_t levelIdc, bool &valid) {
    struct L { uint32_t level, maxDpbBytes, maxFs; };
    static const L kLevels[] = {
        {10, 152064, 99},     {11, 345600, 396},     {12, 912384, 396},     {13, 912384, 396},
        {20, 912384, 396},    {21, 1824768, 792},    {22, 3110400, 1620},   {30, 3110400, 1620},
        {31, 6912000, 3600},  {32, 7864320, 5120},   {40, 12582912, 8192},  {41, 12582912, 8192},
        {42, 34816 * 384, 8704}, {50, 42393600, 22080}, {51, 70778880, 36864}};
    valid = false;
    for (const L &l : kLevels) {
        if (l.level != levelIdc) continue;
        if (picSizeInMbs > l.maxFs) return 0;
        valid = true;
        return std::min<uint32_t>(l.maxDpbBytes / (std::max<uint32_t>(picSizeInMbs, 1u) * 384u), 16);
    }
    return 0;
}

static bool parseHrd(BitReader &br) {
    uint32_t cpbCnt, v;
    if (!br.ue(cpbCnt)) return false;
    cpbCnt++;
    if (cpbCnt > 32) return false;
    if (!br.get(4, v) || !br.get(4, v)) return false;
    for (uint32_t i = 0; i < cpbCnt; i++) {
        if (!br.ue(v) || v > 4294967294u) return false;
        if (!br.ue(v) || v > 4294967294u) return false;
        if (!br.get1(v)) return false;
    }
    if (!br.get(5, v) || !br.get(5, v) || !br.get(5, v) || !br.get(5, v)) return false;
    return true;
}

// h264bsdDecodeVuiParameters (h264bsd_vui.c:82-400)
static bool parseVui(BitReader &br, Vui &vui) {
    uint32_t v;
    vui = Vui();
    if (!br.get1(v)) return false;
    vui.aspectRatioPresent = v;
    if (v) {
        if (!br.get(8, vui.aspectRatioIdc)) return false;
        if (vui.aspectRatioIdc == 255) {
            if (!br.get(16, vui.sarWidth) || !br.get(16, vui.sarHeight)) return false;
        }
    }
    if (!br.get1(v)) return false;  // overscan_info_present_flag
    if (v && !br.get1(v)) return false;
    if (!br.get1(v)) return false;
    vui.videoSignalTypePresent = v;
    if (v) {
        if (!br.get(3, vui.videoFormat)) return false;
        if (!br.get1(v)) return false;
        vui.videoFullRange = v;
        if (!br.get1(v)) return false;
        vui.colourDescriptionPresent = v;
        if (v) {
            if (!br.get(8, vui.colourPrimaries) || !br.get(8, vui.transferCharacteristics) ||
                !br.get(8, vui.matrixCoefficients))
                return false;
        }
    }
    if (!br.get1(v)) return false;  // chroma_loc_info_present_flag
    if (v) {
        uint32_t a, b;
        if (!br.ue(a) || a > 5) return false;
        if (!br.ue(b) || b > 5) return false;
    }
    if (!br.get1(v)) return false;  // timing_info_present_flag
    if (v) {
        if (!br.skip(32) || !br.skip(32)) return false;
        if (!br.get1(v)) return false;
    }
    uint32_t nalHrd, vclHrd;
    if (!br.get1(nalHrd)) return false;
    if (nalHrd && !parseHrd(br)) return false;
    if (!br.get1(vclHrd)) return false;
    if (vclHrd && !parseHrd(br)) return false;
    if (nalHrd || vclHrd) {
        if (!br.get1(v)) return false;  // low_delay_hrd_flag
    }
    if (!br.get1(v)) return false;  // pic_struct_present_flag
    if (!br.get1(v)) return false;
    vui.bitstreamRestriction = v;
    if (v) {
        uint32_t a;
        if (!br.get1(a)) return false;
        if (!br.ue(a) || a > 16) return false;  // max_bytes_per_pic_denom
        if (!br.ue(a) || a > 16) return false;  // max_bits_per_mb_denom
        if (!br.ue(a) || a > 16) return false;  // log2_max_mv_length_horizontal
        if (!br.ue(a) || a > 16) return false;  // log2_max_mv_length_vertical
        if (!br.ue(vui.numReorderFrames)) return false;
        if (!br.ue(vui.maxDecFrameBuffering)) return false;
    } else {
        vui.numReorderFrames = 16;
        vui.maxDecFrameBuffering = 16;
    }
    return true;
}

// h264bsdDecodeSeqParamSet (h264bsd_seq_param_set.c:84-360).  Baseline syntax only: a
// High-profile SPS is read with the same field order (and so mis-parsed) exactly as there.
bool parseSps(BitReader &br, Sps &sps) {
    uint32_t v;
    sps = Sps();
    if (!br.get(8, sps.profileIdc)) return false;
    if (!br.get1(v) || !br.get1(v) || !br.get1(v)) return false;  // constraint_set0..2
    if (!br.get(5, v)) return false;                                // reserved_zero_5bits
    if (!br.get(8, sps.levelIdc)) return false;
    if (!br.ue(sps.id) || sps.id >= kMaxSps) return false;
    if (!br.ue(v) || v > 12) return false;
    sps.maxFrameNum = 1u << (v + 4);
    if (!br.ue(v) || v > 2) return false;
    sps.pocType = v;
    if (sps.pocType == 0) {
        if (!br.ue(v) || v > 12) return false;
        sps.maxPocLsb = 1u << (v + 4);
    } else if (sps.pocType == 1) {
        if (!br.get1(v)) return false;
        sps.deltaPicOrderAlwaysZero = v;
        if (!br.se(sps.offsetForNonRefPic) || !br.se(sps.offsetForTopToBottomField)) return false;
        uint32_t n;
        if (!br.ue(n) || n > 255) return false;
        sps.offsetForRefFrame.resize(n);
        for (uint32_t i = 0; i < n; i++)
            if (!br.se(sps.offsetForRefFrame[i])) return false;
    }
    if (!br.ue(sps.numRefFrames) || sps.numRefFrames > kMaxRefPics) return false;
    if (!br.get1(v)) return false;
    sps.gapsInFrameNumAllowed = v;
    if (!br.ue(v)) return false;
    sps.widthMbs = v + 1;
    if (!br.ue(v)) return false;
    sps.heightMbs = v + 1;
    // the picture size comes from an untrusted stream: checked in 64 bits against what the engine can address (16-bit macroblock
    // addresses, 16-bit row coordinates) before anything is sized by it.  Level 5.1 allows 36 864 macroblocks per frame; the
    // reference fails later, in its allocation (MEMORY_ALLOCATION_ERROR, h264bsd_storage.c:347-378), here the parameter set is
    // refused
    if ((uint64_t)sps.widthMbs * (uint64_t)sps.heightMbs > 65535ull || sps.widthMbs > 4000u || sps.heightMbs > 4000u) return false;
    if (!br.get1(v)) return false;
    if (!v) return false;           // frame_mbs_only_flag must be 1
    if (!br.get1(v)) return false;  // direct_8x8_inference_flag
    if (!br.get1(v)) return false;
    sps.cropping = v;
    if (sps.cropping) {
        if (!br.ue(sps.cropLeft) || !br.ue(sps.cropRight) || !br.ue(sps.cropTop) || !br.ue(sps.cropBottom))
            return false;
        if ((int32_t)sps.cropLeft > 8 * (int32_t)sps.widthMbs - ((int32_t)sps.cropRight + 1) ||
            (int32_t)sps.cropTop > 8 * (int32_t)sps.heightMbs - ((int32_t)sps.cropBottom + 1))
            return false;
    }
    bool valid;
    uint32_t dpb = levelDpbSize(sps.widthMbs * sps.heightMbs, sps.levelIdc, valid);
    if (!valid || sps.numRefFrames > dpb) dpb = sps.numRefFrames;
    sps.maxDpbSize = dpb;
    if (!br.get1(v)) return false;
    sps.vuiPresent = v;
    if (sps.vuiPresent) {
        if (!parseVui(br, sps.vui)) return false;
        if (sps.vui.bitstreamRestriction) {
            if (sps.vui.numReorderFrames > sps.vui.maxDecFrameBuffering ||
                sps.vui.maxDecFrameBuffering < sps.numRefFrames ||
                sps.vui.maxDecFrameBuffering > sps.maxDpbSize)
                return false;
            sps.maxDpbSize = std::max<uint32_t>(1, sps.vui.maxDecFrameBuffering);
        }
    }
    br.trailingBits();  // result ignored, as in the reference
    return true;
}

bool spsEqual(const Sps &a, const Sps &b) {
    if (a.profileIdc != b.profileIdc || a.levelIdc != b.levelIdc || a.maxFrameNum != b.maxFrameNum ||
        a.pocType != b.pocType || a.numRefFrames != b.numRefFrames ||
        a.gapsInFrameNumAllowed != b.gapsInFrameNumAllowed || a.widthMbs != b.widthMbs ||
        a.heightMbs != b.heightMbs || a.cropping != b.cropping || a.vuiPresent != b.vuiPresent)
        return false;
    if (a.pocType == 0) {
        if (a.maxPocLsb != b.maxPocLsb) return false;
    } else if (a.pocType == 1) {
        if (a.deltaPicOrderAlwaysZero != b.deltaPicOrderAlwaysZero ||
            a.offsetForNonRefPic != b.offsetForNonRefPic ||
            a.offsetForTopToBottomField != b.offsetForTopToBottomField ||
            a.offsetForRefFrame != b.offsetForRefFrame)
            return false;
    }
    if (a.cropping) {
        if (a.cropLeft != b.cropLeft || a.cropRight != b.cropRight || a.cropTop != b.cropTop ||
            a.cropBottom != b.cropBottom)
            return false;
    }
    return true;
}

// h264bsdDecodePicParamSet (h264bsd_pic_param_set.c:90-336)
bool parsePps(BitReader &br, Pps &pps) {
    uint32_t v;
    int32_t s;
    pps = Pps();
    if (!br.ue(pps.id) || pps.id >= kMaxPps) return false;
    if (!br.ue(pps.spsId) || pps.spsId >= kMaxSps) return false;
    if (!br.get1(v)) return false;
    if (v) return false;  // entropy_coding_mode_flag: CAVLC only
    if (!br.get1(v)) return false;
    pps.picOrderPresent = v;
    if (!br.ue(v)) return false;
    pps.numSliceGroups = v + 1;
    if (pps.numSliceGroups > kMaxSliceGroups) return false;
    if (pps.numSliceGroups > 1) {
        if (!br.ue(pps.sliceGroupMapType) || pps.sliceGroupMapType > 6) return false;
        if (pps.sliceGroupMapType == 0) {
            pps.runLength.resize(pps.numSliceGroups);
            for (uint32_t i = 0; i < pps.numSliceGroups; i++) {
                if (!br.ue(v)) return false;
                pps.runLength[i] = v + 1;
            }
        } else if (pps.sliceGroupMapType == 2) {
            pps.topLeft.resize(pps.numSliceGroups - 1);
            pps.bottomRight.resize(pps.numSliceGroups - 1);
            for (uint32_t i = 0; i + 1 < pps.numSliceGroups; i++) {
                if (!br.ue(pps.topLeft[i]) || !br.ue(pps.bottomRight[i])) return false;
            }
        } else if (pps.sliceGroupMapType >= 3 && pps.sliceGroupMapType <= 5) {
            if (!br.get1(v)) return false;
            pps.sliceGroupChangeDirection = v;
            if (!br.ue(v)) return false;
            pps.sliceGroupChangeRate = v + 1;
        } else if (pps.sliceGroupMapType == 6) {
            if (!br.ue(v)) return false;
            pps.picSizeInMapUnits = v + 1;
            static const unsigned kBits[8] = {0, 1, 2, 2, 3, 3, 3, 3};  // Ceil(Log2(num_slice_groups))
            unsigned nb = kBits[pps.numSliceGroups - 1];
            // every id costs at least one bit of the NAL: bound the allocation by what is left
            if ((uint64_t)pps.picSizeInMapUnits * nb > br.bitsLeft() + 64) return false;
            pps.sliceGroupId.resize(pps.picSizeInMapUnits);
            for (uint32_t i = 0; i < pps.picSizeInMapUnits; i++) {
                if (!br.get(nb, v)) v = 0xFFFFFFFFu;
                if (v >= pps.numSliceGroups) return false;
                pps.sliceGroupId[i] = v;
            }
        }
    }
    if (!br.ue(v) || v > 31) return false;
    pps.numRefIdxL0Active = v + 1;
    if (!br.ue(v) || v > 31) return false;  // num_ref_idx_l1_active_minus1
    if (!br.get1(v)) return false;
    if (v) return false;  // weighted_pred_flag
    if (!br.get(2, v)) return false;
    if (v > 2) return false;  // weighted_bipred_idc
    if (!br.se(s) || s < -26 || s > 25) return false;
    pps.picInitQp = (uint32_t)(s + 26);
    if (!br.se(s) || s < -26 || s > 25) return false;  // pic_init_qs
    if (!br.se(s) || s < -12 || s > 12) return false;
    pps.chromaQpIndexOffset = s;
    if (!br.get1(v)) return false;
    pps.deblockingFilterControlPresent = v;
    if (!br.get1(v)) return false;
    pps.constrainedIntraPred = v;
    if (!br.get1(v)) return false;
    pps.redundantPicCntPresent = v;
    br.trailingBits();
    return true;
}

bool checkPps(const Pps &pps, const Sps &sps) {
    uint32_t picSize = sps.widthMbs * sps.heightMbs;
    if (pps.numSliceGroups > 1) {
        if (pps.sliceGroupMapType == 0) {
            for (uint32_t r : pps.runLength)
                if (r > picSize) return false;
        } else if (pps.sliceGroupMapType == 2) {
            for (uint32_t i = 0; i + 1 < pps.numSliceGroups; i++) {
                if (pps.topLeft[i] > pps.bottomRight[i] || pps.bottomRight[i] >= picSize) return false;
                if (pps.topLeft[i] % sps.widthMbs > pps.bottomRight[i] % sps.widthMbs) return false;
            }
        } else if (pps.sliceGroupMapType > 2 && pps.sliceGroupMapType < 6) {
            if (pps.sliceGroupChangeRate > picSize) return false;
        } else if (pps.sliceGroupMapType == 6 && pps.picSizeInMapUnits < picSize) {
            return false;
        }
    }
    return true;
}

// ---- slice header ------------------------------------------------------------------------
static bool parseReordering(BitReader &br, SliceHeader &sh, uint32_t maxPicNum) {
    uint32_t v;
    if (!br.get1(v)) return false;
    sh.reorderingFlag = v;
    if (!sh.reorderingFlag) return true;
    uint32_t i = 0, cmd;
    do {
        if (i > sh.numRefIdxL0Active) return false;
        if (!br.ue(cmd) || cmd > 3) return false;
        sh.reorder[i].idc = cmd;
        if (cmd == 0 || cmd == 1) {
            if (!br.ue(v) || v >= maxPicNum) return false;
            sh.reorder[i].absDiffPicNum = v + 1;
        } else if (cmd == 2) {
            if (!br.ue(v)) return false;
            sh.reorder[i].longTermPicNum = v;
        }
        i++;
    } while (cmd != 3);
    return i != 1;  // a lone "end" command is an error there
}

static bool parseMarking(BitReader &br, SliceHeader &sh, bool idr, uint32_t numRefFrames) {
    uint32_t v;
    if (idr) {
        if (!br.get1(v)) return false;
        sh.noOutputOfPriorPics = v;
        if (!br.get1(v)) return false;
        sh.longTermReference = v;
        if (!numRefFrames && sh.longTermReference) return false;
        return true;
    }
    if (!br.get1(v)) return false;
    sh.adaptiveMarking = v;
    if (!sh.adaptiveMarking) return true;
    uint32_t i = 0, op, n4 = 0, n5 = 0, n6 = 0, n13 = 0;
    do {
        if (i > 2 * numRefFrames + 2) return false;
        if (!br.ue(op) || op > 6) return false;
        MmcoOp &m = sh.mmco[i];
        m.op = op;
        if (op == 1 || op == 3) {
            if (!br.ue(v)) return false;
            m.differenceOfPicNums = v + 1;
        }
        if (op == 2) {
            if (!br.ue(m.longTermPicNum)) return false;
        }
        if (op == 3 || op == 6) {
            if (!br.ue(m.longTermFrameIdx)) return false;
        }
        if (op == 4) {
            if (!br.ue(v) || v > numRefFrames) return false;
            m.maxLongTermFrameIdx = v == 0 ? kNoLongTermFrameIndices : v - 1;
            n4++;
        }
        if (op == 5) n5++;
        if (op >= 1 && op <= 3) n13++;
        if (op == 6) n6++;
        i++;
    } while (op != 0);
    if (n4 > 1 || n5 > 1 || n6 > 1 || (n13 && n5)) return false;
    return true;
}

// NumSliceGroupChangeCycleBits: Ceil(Log2(PicSizeInMapUnits / SliceGroupChangeRate + 1))
static unsigned changeCycleBits(uint32_t picSize, uint32_t rate) {
    uint32_t t = picSize / rate + ((picSize % rate) ? 2 : 1);  // value range [0, t-1]
    unsigned n = 0;
    while ((1u << n) < t) n++;
    return n;
}

bool parseSliceHeader(BitReader &br, SliceHeader &sh, const Sps &sps, const Pps &pps, const NalHeader &nal) {
    uint32_t v;
    int32_t s;
    sh = SliceHeader();
    uint32_t picSize = sps.widthMbs * sps.heightMbs;
    if (!br.ue(sh.firstMb) || sh.firstMb >= picSize) return false;
    if (!br.ue(sh.sliceType)) return false;
    if (!sh.isI() && (!sh.isP() || nal.isIdr() || !sps.numRefFrames)) return false;
    if (!br.ue(sh.ppsId) || sh.ppsId != pps.id) return false;
    if (!br.get(log2Exact(sps.maxFrameNum), sh.frameNum)) return false;
    if (nal.isIdr() && sh.frameNum != 0) return false;
    if (nal.isIdr()) {
        if (!br.ue(sh.idrPicId) || sh.idrPicId > 65535) return false;
    }
    if (sps.pocType == 0) {
        if (!br.get(log2Exact(sps.maxPocLsb), sh.pocLsb)) return false;
        if (pps.picOrderPresent && !br.se(sh.deltaPocBottom)) return false;
        if (nal.isIdr() && (sh.pocLsb > sps.maxPocLsb / 2 ||
                            std::min((int32_t)sh.pocLsb, (int32_t)sh.pocLsb + sh.deltaPocBottom) != 0))
            return false;
    }
    if (sps.pocType == 1 && !sps.deltaPicOrderAlwaysZero) {
        if (!br.se(sh.deltaPoc[0])) return false;
        if (pps.picOrderPresent && !br.se(sh.deltaPoc[1])) return false;
        if (nal.isIdr() &&
            std::min(sh.deltaPoc[0], sh.deltaPoc[0] + sps.offsetForTopToBottomField + sh.deltaPoc[1]) != 0)
            return false;
    }
    if (pps.redundantPicCntPresent) {
        if (!br.ue(sh.redundantPicCnt) || sh.redundantPicCnt > 127) return false;
    }
    if (sh.isP()) {
        if (!br.get1(v)) return false;
        if (v) {
            if (!br.ue(v) || v > 15) return false;
            sh.numRefIdxL0Active = v + 1;
        } else {
            if (pps.numRefIdxL0Active > 16) return false;
            sh.numRefIdxL0Active = pps.numRefIdxL0Active;
        }
        if (!parseReordering(br, sh, sps.maxFrameNum)) return false;
    }
    if (nal.refIdc != 0) {
        if (!parseMarking(br, sh, nal.isIdr(), sps.numRefFrames)) return false;
    }
    if (!br.se(s)) return false;
    sh.sliceQpDelta = s;
    s += (int32_t)pps.picInitQp;
    if (s < 0 || s > 51) return false;
    if (pps.deblockingFilterControlPresent) {
        if (!br.ue(sh.disableDeblockingFilterIdc) || sh.disableDeblockingFilterIdc > 2) return false;
        if (sh.disableDeblockingFilterIdc != 1) {
            if (!br.se(s) || s < -6 || s > 6) return false;
            sh.alphaOffset = s * 2;
            if (!br.se(s) || s < -6 || s > 6) return false;
            sh.betaOffset = s * 2;
        }
    }
    if (pps.numSliceGroups > 1 && pps.sliceGroupMapType >= 3 && pps.sliceGroupMapType <= 5) {
        unsigned nb = changeCycleBits(picSize, pps.sliceGroupChangeRate);
        if (!br.get(nb, sh.sliceGroupChangeCycle)) return false;
        uint32_t maxCycle = (picSize + pps.sliceGroupChangeRate - 1) / pps.sliceGroupChangeRate;
        if (sh.sliceGroupChangeCycle > maxCycle) return false;
    }
    return true;
}

// ---- peeks -------------------------------------------------------------------------------
static bool skipToFrameNum(BitReader &br) {
    uint32_t v;
    return br.ue(v) && br.ue(v) && br.ue(v);  // first_mb_in_slice, slice_type, pic_parameter_set_id
}
bool peekPpsId(BitReader br, uint32_t &ppsId) {
    uint32_t v;
    if (!br.ue(v) || !br.ue(v) || !br.ue(v)) return false;
    if (v >= kMaxPps) return false;
    ppsId = v;
    return true;
}
bool peekFrameNum(BitReader br, uint32_t maxFrameNum, uint32_t &frameNum) {
    if (!skipToFrameNum(br)) return false;
    return br.get(log2Exact(maxFrameNum), frameNum);
}
bool peekIdrPicId(BitReader br, uint32_t maxFrameNum, uint32_t &idrPicId) {
    uint32_t v;
    if (!skipToFrameNum(br) || !br.get(log2Exact(maxFrameNum), v)) return false;
    return br.ue(idrPicId);
}
bool peekPocLsb(BitReader br, const Sps &sps, bool idr, uint32_t &pocLsb) {
    uint32_t v;
    if (!skipToFrameNum(br) || !br.get(log2Exact(sps.maxFrameNum), v)) return false;
    if (idr && !br.ue(v)) return false;
    return br.get(log2Exact(sps.maxPocLsb), pocLsb);
}
bool peekDeltaPocBottom(BitReader br, const Sps &sps, bool idr, int32_t &delta) {
    uint32_t v;
    if (!skipToFrameNum(br) || !br.get(log2Exact(sps.maxFrameNum), v)) return false;
    if (idr && !br.ue(v)) return false;
    if (!br.get(log2Exact(sps.maxPocLsb), v)) return false;
    return br.se(delta);
}
bool peekDeltaPoc(BitReader br, const Sps &sps, bool idr, bool picOrderPresent, int32_t delta[2]) {
    uint32_t v;
    if (!skipToFrameNum(br) || !br.get(log2Exact(sps.maxFrameNum), v)) return false;
    if (idr && !br.ue(v)) return false;
    if (!br.se(delta[0])) return false;
    if (picOrderPresent && !br.se(delta[1])) return false;
    return true;
}
bool peekNoOutputOfPriorPics(BitReader br, const Sps &sps, const Pps &pps, bool idr, uint32_t &flag) {
    (void)idr;
    uint32_t v;
    int32_t s;
    if (!skipToFrameNum(br) || !br.get(log2Exact(sps.maxFrameNum), v)) return false;
    if (!br.ue(v)) return false;  // idr_pic_id
    if (sps.pocType == 0) {
        if (!br.get(log2Exact(sps.maxPocLsb), v)) return false;
        if (pps.picOrderPresent && !br.se(s)) return false;
    }
    if (sps.pocType == 1 && !sps.deltaPicOrderAlwaysZero) {
        if (!br.se(s)) return false;
        if (pps.picOrderPresent && !br.se(s)) return false;
    }
    if (pps.redundantPicCntPresent && !br.ue(v)) return false;
    return br.get1(flag);
}

// ---- slice group map (clause 8.2.2; h264bsd_slice_group_map.c:504-590) ---------------------
void buildSliceGroupMap(std::vector<uint32_t> &map, const Pps &pps, uint32_t cycle, uint32_t w, uint32_t h) {
    uint32_t picSize = w * h;
    map.assign(picSize, 0);
    if (pps.numSliceGroups == 1) return;
    uint32_t n = pps.numSliceGroups;
    uint32_t units0 = 0, upperLeft = 0;
    if (pps.sliceGroupMapType >= 3 && pps.sliceGroupMapType <= 5) {
        units0 = std::min<uint64_t>((uint64_t)cycle * pps.sliceGroupChangeRate, picSize);
        upperLeft = pps.sliceGroupChangeDirection ? picSize - units0 : units0;
    }
    switch (pps.sliceGroupMapType) {
        case 0: {  // interleaved
            uint32_t i = 0;
            do {
                for (uint32_t g = 0; g < n && i < picSize; i += pps.runLength[g++])
                    for (uint32_t j = 0; j < pps.runLength[g] && i + j < picSize; j++) map[i + j] = g;
            } while (i < picSize);
            break;
        }
        case 1:  // dispersed
            for (uint32_t i = 0; i < picSize; i++) map[i] = ((i % w) + (((i / w) * n) >> 1)) % n;
            break;
        case 2: {  // foreground + left-over
            for (uint32_t i = 0; i < picSize; i++) map[i] = n - 1;
            for (uint32_t g = n - 1; g-- > 0;) {
                uint32_t y0 = pps.topLeft[g] / w, x0 = pps.topLeft[g] % w;
                uint32_t y1 = pps.bottomRight[g] / w, x1 = pps.bottomRight[g] % w;
                for (uint32_t y = y0; y <= y1; y++)
                    for (uint32_t x = x0; x <= x1; x++) map[y * w + x] = g;
            }
            break;
        }
        case 3: {  // box-out
            for (uint32_t i = 0; i < picSize; i++) map[i] = 1;
            int dirFlag = pps.sliceGroupChangeDirection ? 1 : 0;
            int x = (int)(w - dirFlag) >> 1, y = (int)(h - dirFlag) >> 1;
            int left = x, top = y, right = x, bottom = y;
            int xDir = dirFlag - 1, yDir = dirFlag;
            for (uint32_t k = 0; k < units0;) {
                bool vacant = map[(uint32_t)y * w + (uint32_t)x] == 1;
                if (vacant) map[(uint32_t)y * w + (uint32_t)x] = 0;
                if (xDir == -1 && x == left) {
                    left = std::max(left - 1, 0); x = left; xDir = 0; yDir = 2 * dirFlag - 1;
                } else if (xDir == 1 && x == right) {
                    right = std::min(right + 1, (int)w - 1); x = right; xDir = 0; yDir = 1 - 2 * dirFlag;
                } else if (yDir == -1 && y == top) {
                    top = std::max(top - 1, 0); y = top; xDir = 1 - 2 * dirFlag; yDir = 0;
                } else if (yDir == 1 && y == bottom) {
                    bottom = std::min(bottom + 1, (int)h - 1); y = bottom; xDir = 2 * dirFlag - 1; yDir = 0;
                } else {
                    x += xDir; y += yDir;
                }
                if (vacant) k++;
            }
            break;
        }
        case 4:  // raster scan
            for (uint32_t i = 0; i < picSize; i++)
                map[i] = i < upperLeft ? (uint32_t)pps.sliceGroupChangeDirection
                                       : 1u - (uint32_t)pps.sliceGroupChangeDirection;
            break;
        case 5: {  // wipe
            uint32_t k = 0;
            for (uint32_t j = 0; j < w; j++)
                for (uint32_t i = 0; i < h; i++)
                    map[i * w + j] = (k++ < upperLeft) ? (uint32_t)pps.sliceGroupChangeDirection
                                                       : 1u - (uint32_t)pps.sliceGroupChangeDirection;
            break;
        }
        default:  // explicit
            for (uint32_t i = 0; i < picSize; i++) map[i] = pps.sliceGroupId[i];
            break;
    }
}

}  // namespace b200
