// params.hpp -- sequence / picture parameter sets and slice header of the host-side syntax
// decoder (Baseline subset, same acceptance rules as the reference:
// h264bsd_seq_param_set.c:84-360, h264bsd_vui.c:82-492, h264bsd_pic_param_set.c:90-336,
// h264bsd_slice_header.c:96-1512).
#pragma once
#include <cstdint>
#include <vector>
#include "bits.hpp"

namespace b200 {

constexpr uint32_t kMaxSps = 32;
constexpr uint32_t kMaxPps = 256;
constexpr uint32_t kMaxRefPics = 16;
constexpr uint32_t kMaxSliceGroups = 8;
constexpr uint32_t kNoLongTermFrameIndices = 0xFFFF;

enum NalType : uint32_t {
    NAL_SLICE = 1, NAL_SLICE_IDR = 5, NAL_SEI = 6, NAL_SPS = 7, NAL_PPS = 8,
    NAL_AUD = 9, NAL_END_SEQ = 10, NAL_END_STREAM = 11, NAL_FILLER = 12
};

struct NalHeader {
    uint32_t refIdc = 0;
    uint32_t type = 0;
    bool isIdr() const { return type == NAL_SLICE_IDR; }
};

struct Vui {
    bool aspectRatioPresent = false;
    uint32_t aspectRatioIdc = 0, sarWidth = 0, sarHeight = 0;
    bool videoSignalTypePresent = false;
    uint32_t videoFormat = 5;
    bool videoFullRange = false;
    bool colourDescriptionPresent = false;
    uint32_t colourPrimaries = 2, transferCharacteristics = 2, matrixCoefficients = 2;
    bool bitstreamRestriction = false;
    uint32_t numReorderFrames = 16, maxDecFrameBuffering = 16;
};

struct Sps {
    uint32_t profileIdc = 0, levelIdc = 0, id = 0;
    uint32_t maxFrameNum = 0;
    uint32_t pocType = 0, maxPocLsb = 0;
    bool deltaPicOrderAlwaysZero = false;
    int32_t offsetForNonRefPic = 0, offsetForTopToBottomField = 0;
    std::vector<int32_t> offsetForRefFrame;
    uint32_t numRefFrames = 0;
    bool gapsInFrameNumAllowed = false;
    uint32_t widthMbs = 0, heightMbs = 0;
    bool cropping = false;
    uint32_t cropLeft = 0, cropRight = 0, cropTop = 0, cropBottom = 0;
    bool vuiPresent = false;
    Vui vui;
    uint32_t maxDpbSize = 0;
};

struct Pps {
    uint32_t id = 0, spsId = 0;
    bool picOrderPresent = false;
    uint32_t numSliceGroups = 1, sliceGroupMapType = 0;
    std::vector<uint32_t> runLength, topLeft, bottomRight, sliceGroupId;
    bool sliceGroupChangeDirection = false;
    uint32_t sliceGroupChangeRate = 0, picSizeInMapUnits = 0;
    uint32_t numRefIdxL0Active = 1;
    uint32_t picInitQp = 26;
    int32_t chromaQpIndexOffset = 0;
    bool deblockingFilterControlPresent = false, constrainedIntraPred = false, redundantPicCntPresent = false;
};

struct ReorderCmd { uint32_t idc = 3, absDiffPicNum = 0, longTermPicNum = 0; };
struct MmcoOp { uint32_t op = 0, differenceOfPicNums = 0, longTermPicNum = 0, longTermFrameIdx = 0, maxLongTermFrameIdx = 0; };

struct SliceHeader {
    uint32_t firstMb = 0, sliceType = 0, ppsId = 0, frameNum = 0, idrPicId = 0;
    uint32_t pocLsb = 0;
    int32_t deltaPocBottom = 0, deltaPoc[2] = {0, 0};
    uint32_t redundantPicCnt = 0;
    uint32_t numRefIdxL0Active = 0;
    int32_t sliceQpDelta = 0;
    uint32_t disableDeblockingFilterIdc = 0;
    int32_t alphaOffset = 0, betaOffset = 0;  // already *2
    uint32_t sliceGroupChangeCycle = 0;
    bool reorderingFlag = false;
    ReorderCmd reorder[kMaxRefPics + 2];
    // dec_ref_pic_marking
    bool noOutputOfPriorPics = false, longTermReference = false, adaptiveMarking = false;
    MmcoOp mmco[2 * kMaxRefPics + 3];

    bool isP() const { return sliceType == 0 || sliceType == 5; }
    bool isI() const { return sliceType == 2 || sliceType == 7; }
    bool containsMmco5() const {
        if (!adaptiveMarking) return false;
        for (const MmcoOp &m : mmco) {
            if (m.op == 0) break;
            if (m.op == 5) return true;
        }
        return false;
    }
};

// each returns true on success
bool parseNalHeader(BitReader &br, NalHeader &nal);
bool parseSps(BitReader &br, Sps &sps);
bool parsePps(BitReader &br, Pps &pps);
bool spsEqual(const Sps &a, const Sps &b);  // h264bsdCompareSeqParamSets semantics (true = same)
bool checkPps(const Pps &pps, const Sps &sps);  // CheckPps (h264bsd_storage.c:800-850)
bool parseSliceHeader(BitReader &br, SliceHeader &sh, const Sps &sps, const Pps &pps, const NalHeader &nal);

// peeks used for access-unit boundary detection (h264bsd_slice_header.c:1000-1500); `br` is taken by value
bool peekPpsId(BitReader br, uint32_t &ppsId);
bool peekFrameNum(BitReader br, uint32_t maxFrameNum, uint32_t &frameNum);
bool peekIdrPicId(BitReader br, uint32_t maxFrameNum, uint32_t &idrPicId);
bool peekPocLsb(BitReader br, const Sps &sps, bool idr, uint32_t &pocLsb);
bool peekDeltaPocBottom(BitReader br, const Sps &sps, bool idr, int32_t &delta);
bool peekDeltaPoc(BitReader br, const Sps &sps, bool idr, bool picOrderPresent, int32_t delta[2]);
bool peekNoOutputOfPriorPics(BitReader br, const Sps &sps, const Pps &pps, bool idr, uint32_t &flag);

void buildSliceGroupMap(std::vector<uint32_t> &map, const Pps &pps, uint32_t sliceGroupChangeCycle,
                        uint32_t widthMbs, uint32_t heightMbs);

}  // namespace b200
