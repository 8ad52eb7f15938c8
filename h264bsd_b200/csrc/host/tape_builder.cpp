// tape_builder.cpp -- parse a whole Annex-B stream into a host-resident tape
// (include/h264bsd_b200_tape.h).  This is the "record" half of record-then-replay used by the
// batched engine (pre-parsed work-lists) and by the CPU tests; it drives the
// same StreamDecoder as the legacy h264bsdDecode() entry point, with the decode loop of
// posix/test_h264bsd.c:146-177.
//
// A tape can be re-used for the next stream (h264bsdB200ReparseStream): its arrays keep their capacity,
// so steady-state parsing touches no fresh pages and a page-locked tape stays page-locked.
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <thread>
#include <vector>
#include <vector>
#include "stream_decoder.hpp"
#include "h264bsd_b200.h"

namespace b200 {

// growth of one tape array (malloc'd: the tape owns it)
template <typename T>
static bool ensure(T *&p, uint64_t &capBytes, uint64_t needBytes) {
    if (needBytes <= capBytes) return true;
    uint64_t cap = capBytes ? capBytes : (1u << 20);
    while (cap < needBytes) cap += cap / 2 + (1u << 20);
    void *q = std::realloc(p, cap + 64);
    if (!q) return false;
    p = (T *)q;
    capBytes = cap;
    return true;
}

class TapeSink : public PictureSink {
public:
    explicit TapeSink(b200_tape *t) : t_(t) {}
    bool ok = true, repin = false, sizeChanged = false;
    bool configure(uint32_t w, uint32_t h, uint32_t slots) override {
        // a tape holds pictures of one size (the batched engine lays out its frame pool once): a stream that activates a
        // sequence parameter set with another size ends the tape here (B200_TAPE_SIZE_CHANGE); such streams are for the
        // h264bsdDecode API, which re-allocates on H264BSD_HDRS_RDY like the reference
        if (t_->numPics && (w != t_->widthMbs || h != t_->heightMbs)) { sizeChanged = true; return false; }
        t_->widthMbs = w; t_->heightMbs = h;
        if (slots > t_->numSlots) t_->numSlots = slots;
        return true;
    }
    // the records of the next picture are built in place at the end of the tape's record array
    b200_mb_rec *pictureRecords(uint32_t nMbs) override {
        const size_t nrec = (size_t)nMbs * sizeof(b200_mb_rec);
        if (t_->pinned == 1 && t_->mbRecBytes + nrec > t_->capRecs) {
            h264bsdB200UnpinTape(t_);   // never realloc a page-locked block
            repin = true;
        }
        if (!ensure(t_->mbRecs, t_->capRecs, t_->mbRecBytes + nrec)) { ok = false; return nullptr; }
        return reinterpret_cast<b200_mb_rec *>(t_->mbRecs + t_->mbRecBytes);
    }
    bool submitPicture(const b200_pic_hdr &hdr, const b200_mb_rec *r, const int16_t *c, const uint16_t *o, const b200_mb_rec *filterRecs) override {
        const size_t nMbs = (size_t)hdr.widthMbs * hdr.heightMbs;
        const size_t nrec = nMbs * sizeof(b200_mb_rec), ncoef = (size_t)hdr.numCoefBlocks * B200_COEF_BLOCK_BYTES, nord = (size_t)hdr.numConceal * 2;
        uint64_t orderBytes = (uint64_t)t_->numOrder * 2;
        uint64_t picBytes = (uint64_t)t_->numPics * sizeof(b200_pic_hdr);
        const bool inPlace = t_->mbRecs && reinterpret_cast<const uint8_t *>(r) == t_->mbRecs + t_->mbRecBytes;
        if (t_->pinned == 1 && ((!inPlace && t_->mbRecBytes + nrec > t_->capRecs) || t_->coefBytes + ncoef > t_->capCoefs || orderBytes + nord + 2 > t_->capOrder)) {
            h264bsdB200UnpinTape(t_);   // never realloc a page-locked block
            repin = true;
        }
        if ((!inPlace && !ensure(t_->mbRecs, t_->capRecs, t_->mbRecBytes + nrec)) || !ensure(t_->coefs, t_->capCoefs, t_->coefBytes + ncoef) ||
            !ensure(t_->mbOrder, t_->capOrder, orderBytes + nord + 2) || !ensure(t_->pics, t_->capPics, picBytes + sizeof(b200_pic_hdr))) {
            ok = false;
            return false;
        }
        b200_pic_hdr h = hdr;
        h.mbRecOffset = t_->mbRecBytes;
        h.coefOffset = t_->coefBytes;
        h.orderOffset = t_->numOrder;
        if (!inPlace) std::memcpy(t_->mbRecs + t_->mbRecBytes, r, nrec);
        std::memcpy(t_->coefs + t_->coefBytes, c, ncoef);
        if (nord) std::memcpy((uint8_t *)t_->mbOrder + orderBytes, o, nord);
        t_->numOrder += hdr.numConceal;
        t_->mbRecBytes += nrec;
        if (filterRecs) {
            // the records the in-loop filter reads instead (rare: redundant slices decoded macroblocks a second time): right
            // behind the picture's own records
            if (t_->pinned == 1 && t_->mbRecBytes + nrec > t_->capRecs) {
                h264bsdB200UnpinTape(t_);
                repin = true;
            }
            if (!ensure(t_->mbRecs, t_->capRecs, t_->mbRecBytes + nrec)) { ok = false; return false; }
            std::memcpy(t_->mbRecs + t_->mbRecBytes, filterRecs, nrec);
            h.filterRecOffset = t_->mbRecBytes;
            t_->mbRecBytes += nrec;
        }
        t_->pics[t_->numPics] = h;
        t_->coefBytes += ncoef;
        t_->numPics++;
        return true;
    }
private:
    b200_tape *t_;
};

}  // namespace b200

static b200_tape *reparseStream(b200_tape *t, const uint8_t *stream, size_t len, uint32_t noOutputReordering);
extern "C" b200_tape *h264bsdB200ReparseStream(b200_tape *t, const uint8_t *stream, size_t len, uint32_t noOutputReordering) {
    try {
        return reparseStream(t, stream, len, noOutputReordering);
    } catch (const std::exception &) {      // std::bad_alloc / length_error: no exception leaves the C-ABI
        if (t) t->status = b200::MEMALLOC_ERROR;
        return t;
    }
}
static b200_tape *reparseStream(b200_tape *t, const uint8_t *stream, size_t len, uint32_t noOutputReordering) {
    using namespace b200;
    if (!t) {
        t = (b200_tape *)std::calloc(1, sizeof(b200_tape));
        if (!t) return nullptr;
    }
    // keep arrays + capacities (+ page-lock state), reset the content
    t->numPics = 0; t->widthMbs = t->heightMbs = t->numSlots = 0;
    t->cropFlag = t->cropLeft = t->cropWidth = t->cropTop = t->cropHeight = 0;
    t->videoRange = 0; t->matrixCoefficients = 2;
    t->mbRecBytes = t->coefBytes = 0;
    t->numOutputs = 0; t->status = 0; t->numOrder = 0;
    const uint64_t cap0[3] = {t->capRecs, t->capCoefs, t->capOrder};

    // bit 0 of the flags: no output reordering (h264bsdInit); bit 1: carry on after H264BSD_ERROR the way a player does (the
    // posix test program stops there, posix/test_h264bsd.c:171-173) -- what is missing from a picture is concealed at the next
    // access unit boundary (h264bsd_decoder.c:226-262)
    const bool resilient = (noOutputReordering & 2u) != 0;
    TapeSink sink(t);
    StreamDecoder dec(&sink, (noOutputReordering & 1u) != 0);
    std::vector<uint32_t> outputs;
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
        } else if (r == ERROR && resilient && rb) {
            continue;
        } else if (r == ERROR || r == PARAM_SET_ERROR || r == MEMALLOC_ERROR) {
            t->status = r;
            break;
        }
    }
    if (!sink.ok) t->status = MEMALLOC_ERROR;
    if (sink.sizeChanged) t->status = B200_TAPE_SIZE_CHANGE;
    if (!t->status) {
        dec.flushBuffer();
        while (const OutPic *o = dec.nextOutput()) outputs.push_back(o->picIndex);
    }
    uint64_t capOut = (uint64_t)t->capOutputs * sizeof(uint32_t);
    if (!ensure(t->outputPicIndex, capOut, (outputs.size() + 1) * sizeof(uint32_t))) {
        h264bsdB200FreeTape(t);
        return nullptr;
    }
    t->capOutputs = (uint32_t)(capOut / sizeof(uint32_t));
    if (!outputs.empty()) std::memcpy(t->outputPicIndex, outputs.data(), sizeof(uint32_t) * outputs.size());
    t->numOutputs = (uint32_t)outputs.size();
    (void)cap0;
    if (sink.repin) t->pinned = 2;   // the arrays moved: the caller may page-lock again
    return t;
}

extern "C" b200_tape *h264bsdB200ParseStream(const uint8_t *stream, size_t len, uint32_t noOutputReordering) {
    return h264bsdB200ReparseStream(nullptr, stream, len, noOutputReordering);
}

extern "C" void h264bsdB200FreeTape(b200_tape *t) {
    if (!t) return;
    if (t->pinned == 1) h264bsdB200UnpinTape(t);     // never free page-locked memory while it is registered
    std::free(t->pics);
    std::free(t->mbRecs);
    std::free(t->coefs);
    std::free(t->mbOrder);
    std::free(t->outputPicIndex);
    std::free(t);
}

// Parse `n` streams into `n` (re-used) tapes on `threads` host threads: the per-stream parse is serial by nature, streams are
// independent.  tapes[i] may be NULL (a new tape is made); returns the number of tapes that could not be built.
extern "C" int h264bsdB200ReparseStreams(b200_tape **tapes, uint32_t n, const uint8_t *const *streams, const size_t *lens,
                                         uint32_t noOutputReordering, uint32_t threads) {
    if (!tapes || !streams || !lens || !n) return -1;
    threads = threads ? (threads > n ? n : threads) : 1;
    std::atomic<uint32_t> next(0), failed(0);
    auto work = [&]() {
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            tapes[i] = h264bsdB200ReparseStream(tapes[i], streams[i], lens[i], noOutputReordering);
            if (!tapes[i]) failed.fetch_add(1);
        }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return (int)failed.load();
}

// The same, in the background: Begin returns at once, Wait joins and returns the number of failures.  The arrays passed to
// Begin must stay alive until Wait.
struct b200_parse_job {
    std::thread th;
    int failed = 0;
};
extern "C" b200_parse_job *h264bsdB200ReparseStreamsBegin(b200_tape **tapes, uint32_t n, const uint8_t *const *streams, const size_t *lens,
                                                         uint32_t noOutputReordering, uint32_t threads) {
    b200_parse_job *j = new (std::nothrow) b200_parse_job();
    if (!j) return nullptr;
    j->th = std::thread([=]() { j->failed = h264bsdB200ReparseStreams(tapes, n, streams, lens, noOutputReordering, threads); });
    return j;
}
extern "C" int h264bsdB200ReparseStreamsWait(b200_parse_job *j) {
    if (!j) return -1;
    j->th.join();
    const int f = j->failed;
    delete j;
    return f;
}
