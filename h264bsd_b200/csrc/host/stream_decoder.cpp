// stream_decoder.cpp -- see stream_decoder.hpp
#include "stream_decoder.hpp"
#include "cavlc.hpp"
#include <algorithm>
#include <cstring>

namespace b200 {

StreamDecoder::StreamDecoder(PictureSink *sink, bool noOutputReordering)
    : sink_(sink), noReorderingRequested_(noOutputReordering) {
    cavlcInit();
    pic_.provider = sink;   // the sink may lend the memory the records of a picture are built in
}

// h264bsdExtractNalUnit (h264bsd_byte_stream.c:81-237): same start-code scan and the same
// `readBytes`, but the payload is unescaped into a private buffer instead of in place.
bool StreamDecoder::extractNal(const uint8_t *p, uint32_t len, uint32_t *readBytes) {
    uint32_t initByteCount, zeroCount, size;
    bool hasEmulation = false, invalid = false;
    if (len > 3 && p[0] == 0 && p[1] == 0 && (p[2] & 0xFE) == 0) {
        uint32_t byteCount = 2;
        zeroCount = 2;
        const uint8_t *rd = p + 2;
        for (;;) {
            uint8_t byte = *rd++;
            byteCount++;
            if (byteCount == len) {
                *readBytes = len;
                return false;  // no start code prefix
            }
            if (!byte) zeroCount++;
            else if (byte == 1 && zeroCount >= 2) break;
            else zeroCount = 0;
        }
        initByteCount = byteCount;
        zeroCount = 0;
        for (;;) {
            uint8_t byte = *rd++;
            byteCount++;
            if (!byte) zeroCount++;
            if (byte == 0x03 && zeroCount == 2) hasEmulation = true;
            if (byte == 0x01 && zeroCount >= 2) {
                size = byteCount - initByteCount - zeroCount - 1;
                zeroCount -= std::min<uint32_t>(zeroCount, 3);
                break;
            } else if (byte) {
                if (zeroCount >= 3) invalid = true;
                zeroCount = 0;
            }
            if (byteCount == len) {
                size = byteCount - initByteCount - zeroCount;
                break;
            }
        }
    } else {
        initByteCount = 0;
        zeroCount = 0;
        size = len;
        hasEmulation = true;
    }
    *readBytes = size + initByteCount + zeroCount;
    if (invalid) return false;
    const uint8_t *src = p + initByteCount;
    nal_.resize(size);
    if (!hasEmulation) {
        std::memcpy(nal_.data(), src, size);
        return true;
    }
    uint32_t zc = 0, w = 0;
    for (uint32_t i = 0; i < size; i++) {
        uint8_t b = src[i];
        if (zc == 2 && b == 0x03) {
            if (i + 1 == size || src[i + 1] > 0x03) return false;
            zc = 0;
            continue;
        }
        if (zc == 2 && b <= 0x02) return false;
        zc = b ? 0 : zc + 1;
        nal_[w++] = b;
    }
    nal_.resize(w);
    return true;
}

bool StreamDecoder::storeSps(Sps &sps) {
    uint32_t id = sps.id;
    if (!sps_[id]) {
        sps_[id].reset(new Sps());
    } else if (id == activeSpsId_) {
        if (!spsEqual(sps, *activeSps_)) {
            activeSpsId_ = kMaxSps + 1;
            activePpsId_ = kMaxPps + 1;
            activeSps_ = nullptr;
            activePps_ = nullptr;
        } else {
            return true;
        }
    }
    *sps_[id] = sps;
    return true;
}

bool StreamDecoder::storePps(Pps &pps) {
    uint32_t id = pps.id;
    if (!pps_[id]) {
        pps_[id].reset(new Pps());
    } else if (id == activePpsId_) {
        if (pps.spsId != activeSpsId_) activePpsId_ = kMaxPps + 1;
    }
    *pps_[id] = pps;
    if (id == activePpsId_) activePps_ = pps_[id].get();
    return true;
}

bool StreamDecoder::validParamSets() const {
    for (uint32_t i = 0; i < kMaxPps; i++)
        if (pps_[i] && sps_[pps_[i]->spsId] && checkPps(*pps_[i], *sps_[pps_[i]->spsId])) return true;
    return false;
}

// h264bsdActivateParamSets (h264bsd_storage.c:297-420)
uint32_t StreamDecoder::activateParamSets(uint32_t ppsId, bool isIdr) {
    if (!pps_[ppsId] || !sps_[pps_[ppsId]->spsId]) return 1;
    if (!checkPps(*pps_[ppsId], *sps_[pps_[ppsId]->spsId])) return 1;
    auto takeNew = [&]() {
        activePpsId_ = ppsId;
        activePps_ = pps_[ppsId].get();
        activeSpsId_ = activePps_->spsId;
        activeSps_ = sps_[activeSpsId_].get();
        pendingActivation_ = true;
    };
    if (activePpsId_ == kMaxPps) {
        takeNew();
    } else if (pendingActivation_) {
        pendingActivation_ = false;
        pic_.resize(activeSps_->widthMbs, activeSps_->heightMbs);
        bool noReorder = noReorderingRequested_ || activeSps_->pocType == 2 ||
                         (activeSps_->vuiPresent && activeSps_->vui.bitstreamRestriction &&
                          !activeSps_->vui.numReorderFrames);
        dpb_.init(activeSps_->maxDpbSize, activeSps_->numRefFrames, activeSps_->maxFrameNum, noReorder);
        if (sink_ && !sink_->configure(activeSps_->widthMbs, activeSps_->heightMbs, dpb_.numSlots())) return 2;
    } else if (ppsId != activePpsId_) {
        if (pps_[ppsId]->spsId != activeSpsId_) {
            if (!isIdr) return 1;
            takeNew();
        } else {
            activePpsId_ = ppsId;
            activePps_ = pps_[ppsId].get();
        }
    }
    return 0;
}

// h264bsdCheckAccessUnitBoundary (h264bsd_storage.c:626-790)
uint32_t StreamDecoder::checkAccessUnitBoundary(BitReader &br, const NalHeader &nu, bool &boundary) {
    boundary = false;
    if ((nu.type > 5 && nu.type < 12) || (nu.type > 12 && nu.type <= 18)) {
        boundary = true;
        return 0;
    }
    if (nu.type != NAL_SLICE && nu.type != NAL_SLICE_IDR) return 0;
    if (aub_.firstCall) {
        boundary = true;
        aub_.firstCall = false;
    }
    uint32_t ppsId;
    if (!peekPpsId(br, ppsId)) return 1;
    const Pps *pps = pps_[ppsId].get();
    if (!pps || !sps_[pps->spsId] ||
        (activeSpsId_ != kMaxSps && pps->spsId != activeSpsId_ && nu.type != NAL_SLICE_IDR))
        return 2;
    const Sps &sps = *sps_[pps->spsId];
    if (aub_.nuPrev.refIdc != nu.refIdc && (aub_.nuPrev.refIdc == 0 || nu.refIdc == 0)) boundary = true;
    if ((aub_.nuPrev.type == NAL_SLICE_IDR) != (nu.type == NAL_SLICE_IDR)) boundary = true;
    uint32_t frameNum;
    if (!peekFrameNum(br, sps.maxFrameNum, frameNum)) return 1;
    if (aub_.prevFrameNum != frameNum) {
        aub_.prevFrameNum = frameNum;
        boundary = true;
    }
    if (nu.type == NAL_SLICE_IDR) {
        uint32_t idrPicId;
        if (!peekIdrPicId(br, sps.maxFrameNum, idrPicId)) return 1;
        if (aub_.nuPrev.type == NAL_SLICE_IDR && aub_.prevIdrPicId != idrPicId) boundary = true;
        aub_.prevIdrPicId = idrPicId;
    }
    if (sps.pocType == 0) {
        uint32_t lsb;
        if (!peekPocLsb(br, sps, nu.type == NAL_SLICE_IDR, lsb)) return 1;
        if (aub_.prevPocLsb != lsb) {
            aub_.prevPocLsb = lsb;
            boundary = true;
        }
        if (pps->picOrderPresent) {
            int32_t d;
            if (!peekDeltaPocBottom(br, sps, nu.type == NAL_SLICE_IDR, d)) return 1;
            if (aub_.prevDeltaPocBottom != d) {
                aub_.prevDeltaPocBottom = d;
                boundary = true;
            }
        }
    } else if (sps.pocType == 1 && !sps.deltaPicOrderAlwaysZero) {
        int32_t d[2] = {0, 0};
        if (!peekDeltaPoc(br, sps, nu.type == NAL_SLICE_IDR, pps->picOrderPresent, d)) return 1;
        if (aub_.prevDeltaPoc[0] != d[0]) {
            aub_.prevDeltaPoc[0] = d[0];
            boundary = true;
        }
        if (pps->picOrderPresent && aub_.prevDeltaPoc[1] != d[1]) {
            aub_.prevDeltaPoc[1] = d[1];
            boundary = true;
        }
    }
    aub_.nuPrev = nu;
    return 0;
}

// the picReady tail of h264bsdDecode (h264bsd_decoder.c:473-511): hand the picture to the pixel
// engine (where the reference runs h264bsdFilterPicture), then POC + reference marking
void StreamDecoder::finishPicture() {
    pic_.finalizeRecords();
    b200_pic_hdr hdr;
    std::memset(&hdr, 0, sizeof hdr);
    hdr.widthMbs = pic_.widthMbs;
    hdr.heightMbs = pic_.heightMbs;
    hdr.curSlot = (uint32_t)currSlot_;
    hdr.numSlots = dpb_.numSlots();
    hdr.picIndex = picIndex_;
    hdr.isIdr = prevNal_.isIdr();
    hdr.isRef = prevNal_.refIdc != 0;
    hdr.numCoefBlocks = (uint32_t)(pic_.coefs.size() / 16);
    hdr.numErrMbs = numConcealedMbs_;
    hdr.picId = currentPicId_;
    hdr.numPassA = pic_.numPassA;
    hdr.numPassB = pic_.numPassB;
    hdr.numCopy = pic_.numCopy;
    hdr.orderOffset = hdr.reserved7 = 0;     // (the tape builder places the list)
    hdr.numConceal = pic_.numConceal;

    int32_t poc = decodePicOrderCnt(poc_, *activeSps_, sliceHeader_, prevNal_);
    if (validSliceInAccessUnit_) {
        dpb_.markDecRefPic(prevNal_.refIdc ? &sliceHeader_ : nullptr, currSlot_, sliceHeader_.frameNum, poc,
                           prevNal_.isIdr(), currentPicId_, numConcealedMbs_, picIndex_);
    }
    hdr.numOut = std::min<uint32_t>(dpb_.pendingOutputs(), 20);
    for (uint32_t i = 0; i < hdr.numOut; i++) {
        hdr.outSlot[i] = (uint8_t)dpb_.pendingOutput(i).slot;
        hdr.outPicIndex[i] = dpb_.pendingOutput(i).picIndex;
    }
    if (sink_) sink_->submitPicture(hdr, pic_.recs, pic_.coefs.data(), pic_.order.data(), pic_.st != pic_.recs ? pic_.st : nullptr);
    picIndex_++;
    pic_.beginPicture();  // h264bsdResetStorage
    picStarted_ = false;
    validSliceInAccessUnit_ = false;
}

uint32_t StreamDecoder::decode(const uint8_t *byteStrm, uint32_t len, uint32_t picId, uint32_t *readBytes) {
    bool picReady = false;
    if (prevBufNotFinished_ && byteStrm == prevBufPointer_) {
        *readBytes = prevBytesConsumed_;  // nal_ still holds the unit
    } else {
        if (!extractNal(byteStrm, len, readBytes)) return ERROR;
        prevBytesConsumed_ = *readBytes;
        prevBufPointer_ = byteStrm;
    }
    prevBufNotFinished_ = false;

    BitReader br(nal_.data(), nal_.size());
    NalHeader nal;
    if (!parseNalHeader(br, nal)) return ERROR;
    if (nal.type == 0 || nal.type >= 13) return RDY;

    bool boundary = false;
    uint32_t r = checkAccessUnitBoundary(br, nal, boundary);
    if (r) return r == 2 ? PARAM_SET_ERROR : ERROR;

    if (boundary) {
        if (picStarted_ && activeSps_) {
            if (pendingActivation_) return ERROR;
            // the previous picture never completed: fill what is missing (error path)
            bool pSlice = true;
            if (!validSliceInAccessUnit_) {
                currSlot_ = dpb_.allocateImage();
                dpb_.initRefPicList();
            } else {
                pSlice = sliceHeader_.isP();
            }
            numConcealedMbs_ += pic_.concealMissing(dpb_, pSlice);
            picReady = true;
            *readBytes = 0;
            prevBufNotFinished_ = true;
        } else {
            validSliceInAccessUnit_ = false;
        }
        skipRedundantSlices_ = false;
    }

    if (!picReady) {
        switch (nal.type) {
            case NAL_SPS: {
                Sps sps;
                if (!parseSps(br, sps)) return ERROR;
                storeSps(sps);
                break;
            }
            case NAL_PPS: {
                Pps pps;
                if (!parsePps(br, pps)) return ERROR;
                storePps(pps);
                break;
            }
            case NAL_SLICE_IDR:
            case NAL_SLICE: {
                if (skipRedundantSlices_) return RDY;
                picStarted_ = true;
                if (!validSliceInAccessUnit_) {  // start of picture
                    numConcealedMbs_ = 0;
                    currentPicId_ = picId;
                    uint32_t ppsId = 0;
                    peekPpsId(br, ppsId);
                    uint32_t spsId = activeSpsId_;
                    uint32_t a = activateParamSets(ppsId, nal.isIdr());
                    if (a) {
                        activePpsId_ = kMaxPps;
                        activePps_ = nullptr;
                        activeSpsId_ = kMaxSps;
                        activeSps_ = nullptr;
                        pendingActivation_ = false;
                        return a == 2 ? MEMALLOC_ERROR : PARAM_SET_ERROR;
                    }
                    if (spsId != activeSpsId_) {
                        const Sps *oldSps = oldSpsId_ < kMaxSps ? sps_[oldSpsId_].get() : nullptr;
                        const Sps *newSps = activeSps_;
                        uint32_t noOutputOfPriorPics = 1;
                        *readBytes = 0;
                        prevBufNotFinished_ = true;
                        bool ok = false;
                        if (nal.isIdr()) ok = peekNoOutputOfPriorPics(br, *newSps, *activePps_, true, noOutputOfPriorPics);
                        if (!ok || noOutputOfPriorPics || dpb_.noReordering() || !oldSps ||
                            oldSps->widthMbs != newSps->widthMbs || oldSps->heightMbs != newSps->heightMbs ||
                            oldSps->maxDpbSize != newSps->maxDpbSize)
                            dpb_.flushed = 0;
                        else
                            dpb_.flushOutput();
                        oldSpsId_ = activeSpsId_;
                        return HDRS_RDY;
                    }
                }
                if (pendingActivation_) return ERROR;
                SliceHeader sh;
                if (!parseSliceHeader(br, sh, *activeSps_, *activePps_, nal)) return ERROR;
                if (!validSliceInAccessUnit_) {
                    if (!nal.isIdr()) {
                        if (!dpb_.checkGapsInFrameNum(sh.frameNum, nal.refIdc != 0, activeSps_->gapsInFrameNumAllowed))
                            return ERROR;
                    }
                    currSlot_ = dpb_.allocateImage();
                }
                sliceHeader_ = sh;
                validSliceInAccessUnit_ = true;
                prevNal_ = nal;
                buildSliceGroupMap(pic_.sliceGroupMap, *activePps_, sh.sliceGroupChangeCycle, activeSps_->widthMbs,
                                   activeSps_->heightMbs);
                dpb_.initRefPicList();
                if (!dpb_.reorderRefPicList(sliceHeader_, sliceHeader_.frameNum, sliceHeader_.numRefIdxL0Active))
                    return ERROR;
                if (pic_.decodeSlice(br, sliceHeader_, *activeSps_, *activePps_, dpb_) != SliceResult::Ok) {
                    pic_.markSliceCorrupted(sliceHeader_.firstMb, *activeSps_);
                    return ERROR;
                }
                if (pic_.allDecoded(sliceHeader_.redundantPicCnt != 0)) {
                    picReady = true;
                    skipRedundantSlices_ = true;
                }
                break;
            }
            default:
                break;  // SEI and the rest: not decoded (h264bsd_decoder.c:464-470)
        }
    }

    if (picReady) {
        finishPicture();
        return PIC_RDY;
    }
    return RDY;
}

}  // namespace b200
