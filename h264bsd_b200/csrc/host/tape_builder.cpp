// tape_builder.cpp -- parse a whole Annex-B stream into a host-resident tape
// (include/h264bsd_b200_tape.h).  This is the "record" half of record-then-replay used by the
// batched engine (pre-parsed work-lists) and by the CPU tests; it drives the
// same StreamDecoder as the legacy h264bsdDecode() entry point, with the decode loop of
// posix/test_h264bsd.c:146-177.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "stream_decoder.hpp"
#include "h264bsd_b200.h"

namespace b200 {

class TapeSink : public PictureSink {
public:
    std::vector<b200_pic_hdr> pics;
    std::vector<uint8_t> recs;
    std::vector<uint8_t> coefs;
    std::vector<uint16_t> order;
    uint32_t widthMbs = 0, heightMbs = 0, numSlots = 0;
    bool configure(uint32_t w, uint32_t h, uint32_t slots) override {
        widthMbs = w; heightMbs = h; numSlots = std::max(numSlots, slots);
        return true;
    }
    bool submitPicture(const b200_pic_hdr &hdr, const b200_mb_rec *r, const int16_t *c, const uint16_t *o) override {
        b200_pic_hdr h = hdr;
        h.mbRecOffset = recs.size();
        h.coefOffset = coefs.size();
        size_t nrec = (size_t)hdr.widthMbs * hdr.heightMbs * sizeof(b200_mb_rec);
        recs.insert(recs.end(), (const uint8_t *)r, (const uint8_t *)r + nrec);
        size_t ncoef = (size_t)hdr.numCoefBlocks * B200_COEF_BLOCK_BYTES;
        coefs.insert(coefs.end(), (const uint8_t *)c, (const uint8_t *)c + ncoef);
        order.insert(order.end(), o, o + (size_t)hdr.widthMbs * hdr.heightMbs);
        pics.push_back(h);
        return true;
    }
};

}  // namespace b200

extern "C" b200_tape *h264bsdB200ParseStream(const uint8_t *stream, size_t len, uint32_t noOutputReordering) {
    using namespace b200;
    TapeSink sink;
    StreamDecoder dec(&sink, noOutputReordering != 0);
    std::vector<uint32_t> outputs;
    b200_tape *t = (b200_tape *)std::calloc(1, sizeof(b200_tape));
    if (!t) return nullptr;
    const uint8_t *p = stream;
    size_t left = len;
    while (left > 0) {
        uint32_t rb = 0;
        uint32_t chunk = left > 0x7FFFFFFFu ? 0x7FFFFFFFu : (uint32_t)left;
        uint32_t r = dec.decode(p, chunk, 0, &rb);
        p += rb;
        left -= rb;
        if (r == PIC_RDY) {
            while (const OutPic *o = dec.nextOutput()) outputs.push_back(o->picIndex);
        } else if (r == HDRS_RDY) {
            const Sps *sps = dec.activeSps();
            if (sps) {
                t->cropFlag = sps->cropping;
                if (sps->cropping) {
                    t->cropLeft = 2 * sps->cropLeft;
                    t->cropWidth = 16 * sps->widthMbs - 2 * (sps->cropLeft + sps->cropRight);
                    t->cropTop = 2 * sps->cropTop;
                    t->cropHeight = 16 * sps->heightMbs - 2 * (sps->cropTop + sps->cropBottom);
                }
                t->videoRange = sps->vuiPresent && sps->vui.videoSignalTypePresent && sps->vui.videoFullRange;
                t->matrixCoefficients = (sps->vuiPresent && sps->vui.videoSignalTypePresent &&
                                         sps->vui.colourDescriptionPresent) ? sps->vui.matrixCoefficients : 2;
            }
        } else if (r == ERROR || r == PARAM_SET_ERROR || r == MEMALLOC_ERROR) {
            t->status = r;
            break;
        }
    }
    if (!t->status) {
        dec.flushBuffer();
        while (const OutPic *o = dec.nextOutput()) outputs.push_back(o->picIndex);
    }
    t->numPics = (uint32_t)sink.pics.size();
    t->widthMbs = sink.widthMbs;
    t->heightMbs = sink.heightMbs;
    t->numSlots = sink.numSlots;
    t->mbRecBytes = sink.recs.size();
    t->coefBytes = sink.coefs.size();
    t->pics = (b200_pic_hdr *)std::malloc(sizeof(b200_pic_hdr) * (sink.pics.size() + 1));
    t->mbRecs = (uint8_t *)std::malloc(sink.recs.size() + 64);
    t->coefs = (uint8_t *)std::malloc(sink.coefs.size() + 64);
    t->outputPicIndex = (uint32_t *)std::malloc(sizeof(uint32_t) * (outputs.size() + 1));
    t->mbOrder = (uint16_t *)std::malloc(sizeof(uint16_t) * (sink.order.size() + 1));
    if (!t->pics || !t->mbRecs || !t->coefs || !t->outputPicIndex || !t->mbOrder) {
        h264bsdB200FreeTape(t);
        return nullptr;
    }
    std::memcpy(t->pics, sink.pics.data(), sizeof(b200_pic_hdr) * sink.pics.size());
    std::memcpy(t->mbRecs, sink.recs.data(), sink.recs.size());
    std::memcpy(t->coefs, sink.coefs.data(), sink.coefs.size());
    std::memcpy(t->mbOrder, sink.order.data(), sizeof(uint16_t) * sink.order.size());
    std::memcpy(t->outputPicIndex, outputs.data(), sizeof(uint32_t) * outputs.size());
    t->numOutputs = (uint32_t)outputs.size();
    return t;
}

extern "C" void h264bsdB200FreeTape(b200_tape *t) {
    if (!t) return;
    std::free(t->pics);
    std::free(t->mbRecs);
    std::free(t->coefs);
    std::free(t->mbOrder);
    std::free(t->outputPicIndex);
    std::free(t);
}
