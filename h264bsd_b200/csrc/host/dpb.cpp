// dpb.cpp -- see dpb.hpp
#include "dpb.hpp"
#include <algorithm>

namespace b200 {

void Dpb::init(uint32_t dpbSize, uint32_t maxRefFrames, uint32_t maxFrameNum, bool noReordering) {
    maxLongTermFrameIdx_ = kNoLongTermFrameIndices;
    maxRefFrames_ = std::max<uint32_t>(maxRefFrames, 1);
    dpbSize_ = noReordering ? maxRefFrames_ : dpbSize;
    maxFrameNum_ = maxFrameNum;
    noReordering_ = noReordering;
    fullness_ = numRefFrames_ = prevRefFrameNum_ = 0;
    buffer_.assign(17, DpbPic());
    for (uint32_t i = 0; i < 17; i++) buffer_[i].slot = (int)i;
    for (int &l : list_) l = -1;
    outBuf_.assign(dpbSize_ + 1 + 17, OutPic());
    numOut_ = outIndex_ = 0;
    currentOut_ = (int)dpbSize_;
}

int Dpb::allocateImage() {
    currentOut_ = (int)dpbSize_;
    return buffer_[currentOut_].slot;
}

void Dpb::initRefPicList() {
    for (uint32_t i = 0; i < numRefFrames_; i++) list_[i] = (int)i;
}

int Dpb::findPic(int32_t picNum, bool shortTerm) const {
    for (uint32_t i = 0; i < maxRefFrames_; i++) {
        const DpbPic &p = buffer_[i];
        if ((shortTerm ? isShort(p) : isLong(p)) && p.picNum == picNum) return (int)i;
    }
    return -1;
}

void Dpb::setPicNums(uint32_t currFrameNum) {
    for (uint32_t i = 0; i < numRefFrames_; i++) {
        DpbPic &p = buffer_[i];
        if (isShort(p))
            p.picNum = p.frameNum > currFrameNum ? (int32_t)p.frameNum - (int32_t)maxFrameNum_ : (int32_t)p.frameNum;
    }
}

// clause 8.2.4.3 (h264bsd_dpb.c:180-290)
bool Dpb::reorderRefPicList(const SliceHeader &sh, uint32_t currFrameNum, uint32_t numRefIdxActive) {
    setPicNums(currFrameNum);
    if (!sh.reorderingFlag) return true;
    uint32_t refIdx = 0;
    uint32_t picNumPred = currFrameNum;
    for (uint32_t i = 0; sh.reorder[i].idc < 3; i++) {
        int32_t picNum;
        bool shortTerm;
        if (sh.reorder[i].idc < 2) {
            int32_t noWrap;
            if (sh.reorder[i].idc == 0) {
                noWrap = (int32_t)picNumPred - (int32_t)sh.reorder[i].absDiffPicNum;
                if (noWrap < 0) noWrap += (int32_t)maxFrameNum_;
            } else {
                noWrap = (int32_t)(picNumPred + sh.reorder[i].absDiffPicNum);
                if (noWrap >= (int32_t)maxFrameNum_) noWrap -= (int32_t)maxFrameNum_;
            }
            picNumPred = (uint32_t)noWrap;
            picNum = noWrap;
            if ((uint32_t)noWrap > currFrameNum) picNum -= (int32_t)maxFrameNum_;
            shortTerm = true;
        } else {
            picNum = (int32_t)sh.reorder[i].longTermPicNum;
            shortTerm = false;
        }
        int index = findPic(picNum, shortTerm);
        if (index < 0 || !isExisting(buffer_[index])) return false;
        for (uint32_t j = numRefIdxActive; j > refIdx; j--) list_[j] = list_[j - 1];
        list_[refIdx++] = index;
        uint32_t k = refIdx;
        for (uint32_t j = refIdx; j <= numRefIdxActive; j++)
            if (list_[j] != index) list_[k++] = list_[j];
    }
    return true;
}

int Dpb::refSlot(uint32_t refIdx) const {
    if (refIdx > 16 || list_[refIdx] < 0) return -1;
    const DpbPic &p = buffer_[list_[refIdx]];
    return isExisting(p) ? p.slot : -1;
}

bool Dpb::slidingWindow() {
    if (numRefFrames_ < maxRefFrames_) return true;
    int index = -1;
    int32_t picNum = 0;
    for (uint32_t i = 0; i < numRefFrames_; i++)
        if (isShort(buffer_[i]) && (buffer_[i].picNum < picNum || index == -1)) {
            index = (int)i;
            picNum = buffer_[i].picNum;
        }
    if (index < 0) return false;
    setUnused(buffer_[index]);
    return true;
}

bool Dpb::mmco1(uint32_t currPicNum, uint32_t diff) {
    int idx = findPic((int32_t)currPicNum - (int32_t)diff, true);
    if (idx < 0) return false;
    setUnused(buffer_[idx]);
    return true;
}
bool Dpb::mmco2(uint32_t longTermPicNum) {
    int idx = findPic((int32_t)longTermPicNum, false);
    if (idx < 0) return false;
    setUnused(buffer_[idx]);
    return true;
}
bool Dpb::mmco3(uint32_t currPicNum, uint32_t diff, uint32_t longTermFrameIdx) {
    if (maxLongTermFrameIdx_ == kNoLongTermFrameIndices || longTermFrameIdx > maxLongTermFrameIdx_) return false;
    for (uint32_t i = 0; i < maxRefFrames_; i++)
        if (isLong(buffer_[i]) && (uint32_t)buffer_[i].picNum == longTermFrameIdx) {
            setUnused(buffer_[i]);
            break;
        }
    int idx = findPic((int32_t)currPicNum - (int32_t)diff, true);
    if (idx < 0 || !isExisting(buffer_[idx])) return false;
    buffer_[idx].status = PicStatus::LongTerm;
    buffer_[idx].picNum = (int32_t)longTermFrameIdx;
    return true;
}
void Dpb::mmco4(uint32_t maxIdx) {
    maxLongTermFrameIdx_ = maxIdx;
    for (uint32_t i = 0; i < maxRefFrames_; i++)
        if (isLong(buffer_[i]) &&
            ((uint32_t)buffer_[i].picNum > maxIdx || maxLongTermFrameIdx_ == kNoLongTermFrameIndices))
            setUnused(buffer_[i]);
}
void Dpb::mmco5() {
    for (uint32_t i = 0; i < 16; i++)
        if (isRef(buffer_[i])) {
            buffer_[i].status = PicStatus::Unused;
            if (!buffer_[i].toBeDisplayed) fullness_--;
        }
    while (outputOne()) {}
    numRefFrames_ = 0;
    maxLongTermFrameIdx_ = kNoLongTermFrameIndices;
    prevRefFrameNum_ = 0;
}
bool Dpb::mmco6(uint32_t frameNum, int32_t poc, uint32_t longTermFrameIdx) {
    if (maxLongTermFrameIdx_ == kNoLongTermFrameIndices || longTermFrameIdx > maxLongTermFrameIdx_) return false;
    for (uint32_t i = 0; i < maxRefFrames_; i++)
        if (isLong(buffer_[i]) && (uint32_t)buffer_[i].picNum == longTermFrameIdx) {
            setUnused(buffer_[i]);
            break;
        }
    if (numRefFrames_ >= maxRefFrames_) return false;
    DpbPic &c = buffer_[currentOut_];
    c.frameNum = frameNum;
    c.picNum = (int32_t)longTermFrameIdx;
    c.poc = poc;
    c.status = PicStatus::LongTerm;
    c.toBeDisplayed = !noReordering_;
    numRefFrames_++;
    fullness_++;
    return true;
}

// output the not-yet-displayed picture with the smallest POC (h264bsd_dpb.c OutputPicture)
bool Dpb::outputOne() {
    if (noReordering_) return false;
    DpbPic *best = nullptr;
    int32_t poc = 0x7FFFFFFF;
    for (uint32_t i = 0; i <= dpbSize_; i++)
        if (buffer_[i].toBeDisplayed && buffer_[i].poc < poc) {
            best = &buffer_[i];
            poc = buffer_[i].poc;
        }
    if (!best) return false;
    if (numOut_ < outBuf_.size()) {
        OutPic &o = outBuf_[numOut_++];
        o.slot = best->slot; o.isIdr = best->isIdr; o.picId = best->picId; o.numErrMbs = best->numErrMbs;
        o.picIndex = best->picIndex;
    }
    best->toBeDisplayed = false;
    if (!isRef(*best)) fullness_--;
    return true;
}

// order: short-term refs by PicNum descending, long-term by LongTermPicNum ascending, then
// non-reference pictures still waiting for output, then free entries (ComparePictures)
static int comparePics(const DpbPic &a, const DpbPic &b) {
    auto ref = [](const DpbPic &p) { return p.status != PicStatus::Unused; };
    auto shortT = [](const DpbPic &p) { return p.status == PicStatus::NonExisting || p.status == PicStatus::ShortTerm; };
    if (!ref(a) && !ref(b)) {
        if (a.toBeDisplayed && !b.toBeDisplayed) return -1;
        if (!a.toBeDisplayed && b.toBeDisplayed) return 1;
        return 0;
    }
    if (!ref(b)) return -1;
    if (!ref(a)) return 1;
    if (shortT(a) && shortT(b)) return a.picNum > b.picNum ? -1 : (a.picNum < b.picNum ? 1 : 0);
    if (shortT(a)) return -1;
    if (shortT(b)) return 1;
    return a.picNum > b.picNum ? 1 : (a.picNum < b.picNum ? -1 : 0);
}

// same gap sequence (7,3,1) as the reference's ShellSort so that equal-ranked entries end up in
// the same positions and a picture lands in the same frame slot as there
void Dpb::sortBuffer() {
    uint32_t num = dpbSize_ + 1;
    for (uint32_t step = 7; step; step >>= 1)
        for (uint32_t i = step; i < num; i++) {
            DpbPic tmp = buffer_[i];
            uint32_t j = i;
            while (j >= step && comparePics(buffer_[j - step], tmp) > 0) {
                buffer_[j] = buffer_[j - step];
                j -= step;
            }
            buffer_[j] = tmp;
        }
}

bool Dpb::markDecRefPic(const SliceHeader *sh, int slot, uint32_t frameNum, int32_t poc, bool isIdr,
                        uint32_t picId, uint32_t numErrMbs, uint32_t picIndex) {
    DpbPic &c = buffer_[currentOut_];
    if (slot != c.slot) return false;
    bool ok = true;
    bool toBeDisplayed = !noReordering_;
    if (!sh) {
        c.status = PicStatus::Unused;
        c.frameNum = frameNum;
        c.picNum = (int32_t)frameNum;
        c.poc = poc;
        c.toBeDisplayed = toBeDisplayed;
        if (!noReordering_) fullness_++;
    } else if (isIdr) {
        numOut_ = outIndex_ = 0;
        mmco5();
        if (sh->noOutputOfPriorPics || noReordering_) numOut_ = outIndex_ = 0;
        if (sh->longTermReference) {
            c.status = PicStatus::LongTerm;
            maxLongTermFrameIdx_ = 0;
        } else {
            c.status = PicStatus::ShortTerm;
            maxLongTermFrameIdx_ = kNoLongTermFrameIndices;
        }
        c.frameNum = 0;
        c.picNum = 0;
        c.poc = 0;
        c.toBeDisplayed = toBeDisplayed;
        fullness_ = 1;
        numRefFrames_ = 1;
    } else {
        bool markedLong = false;
        if (sh->adaptiveMarking) {
            for (uint32_t i = 0; sh->mmco[i].op; i++) {
                const MmcoOp &m = sh->mmco[i];
                switch (m.op) {
                    case 1: ok = mmco1(frameNum, m.differenceOfPicNums); break;
                    case 2: ok = mmco2(m.longTermPicNum); break;
                    case 3: ok = mmco3(frameNum, m.differenceOfPicNums, m.longTermFrameIdx); break;
                    case 4: mmco4(m.maxLongTermFrameIdx); break;
                    case 5: mmco5(); frameNum = 0; break;
                    case 6:
                        ok = mmco6(frameNum, poc, m.longTermFrameIdx);
                        if (ok) markedLong = true;
                        break;
                    default: ok = false; break;
                }
                if (!ok) break;
            }
        } else {
            ok = slidingWindow();
        }
        if (!markedLong) {
            if (numRefFrames_ < maxRefFrames_) {
                c.frameNum = frameNum;
                c.picNum = (int32_t)frameNum;
                c.poc = poc;
                c.status = PicStatus::ShortTerm;
                c.toBeDisplayed = toBeDisplayed;
                fullness_++;
                numRefFrames_++;
            } else {
                ok = false;
            }
        }
    }
    c.isIdr = isIdr;
    c.picId = picId;
    c.numErrMbs = numErrMbs;
    c.picIndex = picIndex;
    if (noReordering_) {
        if (numOut_ < outBuf_.size()) {
            OutPic &o = outBuf_[numOut_++];
            o.slot = c.slot; o.isIdr = c.isIdr; o.picId = c.picId; o.numErrMbs = c.numErrMbs; o.picIndex = c.picIndex;
        }
    } else {
        while (fullness_ > dpbSize_)
            if (!outputOne()) break;
    }
    sortBuffer();
    return ok;
}

// clause 8.2.5.2 (h264bsd_dpb.c:1230-1340)
bool Dpb::checkGapsInFrameNum(uint32_t frameNum, bool isRefPic, bool gapsAllowed) {
    numOut_ = outIndex_ = 0;
    if (!gapsAllowed) return true;
    if (frameNum != prevRefFrameNum_ && frameNum != (prevRefFrameNum_ + 1) % maxFrameNum_) {
        uint32_t unused = (prevRefFrameNum_ + 1) % maxFrameNum_;
        int keepSlot = buffer_[dpbSize_].slot;
        do {
            setPicNums(unused);
            if (!slidingWindow()) return false;
            while (fullness_ >= dpbSize_)
                if (!outputOne()) break;
            DpbPic &p = buffer_[dpbSize_];
            p.status = PicStatus::NonExisting;
            p.frameNum = unused;
            p.picNum = (int32_t)unused;
            p.poc = 0;
            p.toBeDisplayed = false;
            fullness_++;
            numRefFrames_++;
            sortBuffer();
            unused = (unused + 1) % maxFrameNum_;
        } while (unused != frameNum);
        // do not reconstruct into a slot that is still waiting in the output queue
        for (uint32_t i = 0; i < numOut_; i++)
            if (outBuf_[i].slot == buffer_[dpbSize_].slot) {
                for (uint32_t k = 0; k < dpbSize_; k++)
                    if (buffer_[k].slot == keepSlot) {
                        buffer_[k].slot = buffer_[dpbSize_].slot;
                        buffer_[dpbSize_].slot = keepSlot;
                        break;
                    }
                break;
            }
    } else if (isRefPic && frameNum == prevRefFrameNum_) {
        return false;
    }
    if (isRefPic)
        prevRefFrameNum_ = frameNum;
    else if (frameNum != prevRefFrameNum_)
        prevRefFrameNum_ = (frameNum + maxFrameNum_ - 1) % maxFrameNum_;
    return true;
}

const OutPic *Dpb::outputPicture() {
    if (outIndex_ < numOut_) return &outBuf_[outIndex_++];
    return nullptr;
}

void Dpb::flushOutput() {
    if (buffer_.empty()) return;
    flushed = 1;
    while (outputOne()) {}
}

// clause 8.2.1 (h264bsd_pic_order_cnt.c:80-348); frames only: returns min(top, bottom)
int32_t decodePicOrderCnt(PocState &poc, const Sps &sps, const SliceHeader &sh, const NalHeader &nal) {
    bool mmco5 = sh.containsMmco5();
    int32_t picOrderCnt = 0;
    if (sps.pocType == 0) {
        if (nal.isIdr()) {
            poc.prevPocMsb = 0;
            poc.prevPocLsb = 0;
        }
        if (sh.pocLsb < poc.prevPocLsb && poc.prevPocLsb - sh.pocLsb >= sps.maxPocLsb / 2)
            picOrderCnt = poc.prevPocMsb + (int32_t)sps.maxPocLsb;
        else if (sh.pocLsb > poc.prevPocLsb && sh.pocLsb - poc.prevPocLsb > sps.maxPocLsb / 2)
            picOrderCnt = poc.prevPocMsb - (int32_t)sps.maxPocLsb;
        else
            picOrderCnt = poc.prevPocMsb;
        if (nal.refIdc) poc.prevPocMsb = picOrderCnt;
        picOrderCnt += (int32_t)sh.pocLsb;
        if (sh.deltaPocBottom < 0) picOrderCnt += sh.deltaPocBottom;
        if (nal.refIdc) {
            if (mmco5) {
                poc.prevPocMsb = 0;
                poc.prevPocLsb = sh.deltaPocBottom < 0 ? (uint32_t)(-sh.deltaPocBottom) : 0;
                picOrderCnt = 0;
            } else {
                poc.prevPocLsb = sh.pocLsb;
            }
        }
        return picOrderCnt;
    }
    uint32_t frameNumOffset;
    if (nal.isIdr())
        frameNumOffset = 0;
    else if (poc.prevFrameNum > sh.frameNum)
        frameNumOffset = poc.prevFrameNumOffset + sps.maxFrameNum;
    else
        frameNumOffset = poc.prevFrameNumOffset;
    if (sps.pocType == 1) {
        uint32_t n = (uint32_t)sps.offsetForRefFrame.size();
        uint32_t absFrameNum = n ? frameNumOffset + sh.frameNum : 0;
        if (nal.refIdc == 0 && absFrameNum > 0) absFrameNum--;
        int32_t expectedDeltaPerCycle = 0;
        for (int32_t o : sps.offsetForRefFrame) expectedDeltaPerCycle += o;
        if (absFrameNum > 0) {
            uint32_t cycleCnt = (absFrameNum - 1) / n, inCycle = (absFrameNum - 1) % n;
            picOrderCnt = (int32_t)cycleCnt * expectedDeltaPerCycle;
            for (uint32_t i = 0; i <= inCycle; i++) picOrderCnt += sps.offsetForRefFrame[i];
        }
        if (nal.refIdc == 0) picOrderCnt += sps.offsetForNonRefPic;
        picOrderCnt += sh.deltaPoc[0];
        if (sps.offsetForTopToBottomField + sh.deltaPoc[1] < 0)
            picOrderCnt += sps.offsetForTopToBottomField + sh.deltaPoc[1];
    } else {
        if (nal.isIdr())
            picOrderCnt = 0;
        else if (nal.refIdc == 0)
            picOrderCnt = 2 * (int32_t)(frameNumOffset + sh.frameNum) - 1;
        else
            picOrderCnt = 2 * (int32_t)(frameNumOffset + sh.frameNum);
    }
    if (!mmco5) {
        poc.prevFrameNumOffset = frameNumOffset;
        poc.prevFrameNum = sh.frameNum;
    } else {
        poc.prevFrameNumOffset = 0;
        poc.prevFrameNum = 0;
        picOrderCnt = 0;
    }
    return picOrderCnt;
}

}  // namespace b200
