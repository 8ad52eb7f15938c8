// stream_decoder.hpp -- one H.264 Baseline decoder instance, host side.
//
// This is the control logic of h264bsdDecode (h264bsd_decoder.c:152-515) and of the storage /
// parameter-set / access-unit machinery behind it (h264bsd_storage.c), re-implemented so that
// the public API keeps its exact call/return contract while the pixel work is handed, one
// finished picture at a time, to a PictureSink (the B200 engine, or a test sink).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include "params.hpp"
#include "dpb.hpp"
#include "picture.hpp"
#include "h264bsd_b200_tape.h"

namespace b200 {

// same values as the enum in the public header (h264bsd_decoder.h:45-52)
enum DecodeResult : uint32_t { RDY = 0, PIC_RDY, HDRS_RDY, ERROR, PARAM_SET_ERROR, MEMALLOC_ERROR };

class PictureSink : public RecordProvider {
public:
    virtual ~PictureSink() {}
    // parameter sets activated: frame geometry and number of frame slots are now known
    virtual bool configure(uint32_t widthMbs, uint32_t heightMbs, uint32_t numSlots) = 0;
    // a complete picture; `recs` (widthMbs*heightMbs records; the memory pictureRecords() lent, if it lent any), `coefs` and
    // `order` are only valid during the call
    // filterRecs: nullptr, or a second record array for the in-loop filter alone (b200_pic_hdr.filterRecOffset)
    virtual bool submitPicture(const b200_pic_hdr &hdr, const b200_mb_rec *recs, const int16_t *coefs, const uint16_t *order,
                               const b200_mb_rec *filterRecs) = 0;
};

class StreamDecoder {
public:
    StreamDecoder(PictureSink *sink, bool noOutputReordering);

    uint32_t decode(const uint8_t *byteStrm, uint32_t len, uint32_t picId, uint32_t *readBytes);
    const OutPic *nextOutput() { return dpb_.outputPicture(); }
    void flushBuffer() { dpb_.flushOutput(); }

    const Sps *activeSps() const { return activeSps_; }
    bool validParamSets() const;
    uint32_t picSizeInMbs() const { return pic_.picSizeInMbs; }
    uint32_t picturesSubmitted() const { return picIndex_; }

private:
    bool extractNal(const uint8_t *p, uint32_t len, uint32_t *readBytes);
    uint32_t checkAccessUnitBoundary(BitReader &br, const NalHeader &nal, bool &boundary);  // 0 ok, 1 nok, 2 param set error
    uint32_t activateParamSets(uint32_t ppsId, bool isIdr);                                 // 0 ok, 1 nok
    bool storeSps(Sps &sps);
    bool storePps(Pps &pps);
    void finishPicture();

    PictureSink *sink_;
    bool noReorderingRequested_;

    std::unique_ptr<Sps> sps_[kMaxSps];
    std::unique_ptr<Pps> pps_[kMaxPps];
    uint32_t oldSpsId_ = 0, activePpsId_ = kMaxPps, activeSpsId_ = kMaxSps;
    Pps *activePps_ = nullptr;
    Sps *activeSps_ = nullptr;
    bool pendingActivation_ = false;

    bool skipRedundantSlices_ = false, picStarted_ = false, validSliceInAccessUnit_ = false;
    uint32_t numConcealedMbs_ = 0, currentPicId_ = 0;
    int currSlot_ = 0;

    Dpb dpb_;
    PocState poc_;
    PictureState pic_;

    struct {
        NalHeader nuPrev;
        uint32_t prevFrameNum = 0, prevIdrPicId = 0, prevPocLsb = 0;
        int32_t prevDeltaPocBottom = 0, prevDeltaPoc[2] = {0, 0};
        bool firstCall = true;
    } aub_;

    NalHeader prevNal_;
    SliceHeader sliceHeader_;  // last successfully decoded

    // NAL payload with emulation prevention removed (the reference strips it in place)
    std::vector<uint8_t> nal_;
    bool prevBufNotFinished_ = false;
    const uint8_t *prevBufPointer_ = nullptr;
    uint32_t prevBytesConsumed_ = 0;

    uint32_t picIndex_ = 0;
};

}  // namespace b200
