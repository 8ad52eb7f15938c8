// dpb.hpp -- decoded picture buffer bookkeeping of the host-side syntax decoder.
//
// Only the BOOKKEEPING lives here (reference marking, list construction, output order);
// the frames themselves are slots in the engine's HBM frame pool, named by index.
// Behaviour follows h264bsd_dpb.c (sliding window :900-960, MMCO :300-560, list init/reorder
// :180-290,:1049, gaps in frame_num :1230-1340, output :1390-1520) so that picture order,
// error returns and the slot a picture is reconstructed into match the reference decoder;
// picture order count follows h264bsd_pic_order_cnt.c:80-348.
#pragma once
#include <cstdint>
#include <vector>
#include "params.hpp"

namespace b200 {

enum class PicStatus : uint8_t { Unused = 0, NonExisting, ShortTerm, LongTerm };

struct DpbPic {
    int slot = 0;  // frame slot in the engine's pool (the reference keeps a data pointer here)
    int32_t picNum = 0;
    uint32_t frameNum = 0;
    int32_t poc = 0;
    PicStatus status = PicStatus::Unused;
    bool toBeDisplayed = false;
    uint32_t picId = 0, numErrMbs = 0;
    bool isIdr = false;
    uint32_t picIndex = 0;  // decode-order index of the picture held
};

struct OutPic {
    int slot = 0;
    uint32_t picId = 0, numErrMbs = 0;
    bool isIdr = false;
    uint32_t picIndex = 0;
};

class Dpb {
public:
    void init(uint32_t dpbSize, uint32_t maxRefFrames, uint32_t maxFrameNum, bool noReordering);
    bool initialized() const { return !buffer_.empty(); }
    void reset() { buffer_.clear(); outBuf_.clear(); }

    uint32_t numSlots() const { return dpbSize_ + 1; }
    bool noReordering() const { return noReordering_; }
    uint32_t flushed = 0;

    int allocateImage();  // slot the next picture is written to
    void initRefPicList();
    bool reorderRefPicList(const SliceHeader &sh, uint32_t currFrameNum, uint32_t numRefIdxActive);
    int refSlot(uint32_t refIdx) const;  // -1: no such (existing) reference picture
    bool checkGapsInFrameNum(uint32_t frameNum, bool isRefPic, bool gapsAllowed);
    // sh == nullptr: non-reference picture
    bool markDecRefPic(const SliceHeader *sh, int slot, uint32_t frameNum, int32_t poc, bool isIdr,
                       uint32_t picId, uint32_t numErrMbs, uint32_t picIndex);
    const OutPic *outputPicture();
    void flushOutput();
    uint32_t pendingOutputs() const { return numOut_ - outIndex_; }
    const OutPic &pendingOutput(uint32_t i) const { return outBuf_[outIndex_ + i]; }

private:
    std::vector<DpbPic> buffer_;  // 17 entries, sorted after every picture
    int list_[17];                // positions in buffer_, -1 = none; deliberately NOT cleared between pictures
    int currentOut_ = 0;          // position
    std::vector<OutPic> outBuf_;
    uint32_t numOut_ = 0, outIndex_ = 0;
    uint32_t maxRefFrames_ = 0, dpbSize_ = 0, maxFrameNum_ = 0, maxLongTermFrameIdx_ = kNoLongTermFrameIndices;
    uint32_t numRefFrames_ = 0, fullness_ = 0, prevRefFrameNum_ = 0;
    bool noReordering_ = false;

    static bool isRef(const DpbPic &p) { return p.status != PicStatus::Unused; }
    static bool isExisting(const DpbPic &p) { return p.status == PicStatus::ShortTerm || p.status == PicStatus::LongTerm; }
    static bool isShort(const DpbPic &p) { return p.status == PicStatus::NonExisting || p.status == PicStatus::ShortTerm; }
    static bool isLong(const DpbPic &p) { return p.status == PicStatus::LongTerm; }
    void setUnused(DpbPic &p) {
        p.status = PicStatus::Unused;
        numRefFrames_--;
        if (!p.toBeDisplayed) fullness_--;
    }
    int findPic(int32_t picNum, bool shortTerm) const;
    void setPicNums(uint32_t currFrameNum);
    bool slidingWindow();
    bool outputOne();
    void sortBuffer();
    bool mmco1(uint32_t currPicNum, uint32_t diff);
    bool mmco2(uint32_t longTermPicNum);
    bool mmco3(uint32_t currPicNum, uint32_t diff, uint32_t longTermFrameIdx);
    void mmco4(uint32_t maxLongTermFrameIdx);
    void mmco5();
    bool mmco6(uint32_t frameNum, int32_t poc, uint32_t longTermFrameIdx);
};

struct PocState {
    uint32_t prevPocLsb = 0;
    int32_t prevPocMsb = 0;
    uint32_t prevFrameNum = 0, prevFrameNumOffset = 0;
};
int32_t decodePicOrderCnt(PocState &poc, const Sps &sps, const SliceHeader &sh, const NalHeader &nal);

}  // namespace b200
