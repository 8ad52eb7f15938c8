// picture.hpp -- per-picture macroblock state + slice-data decoding of the host-side syntax
// decoder.  Produces the tape records (include/h264bsd_b200_tape.h); touches no pel.
//
// Replaces the syntax half of h264bsd_slice_data.c:86-232 and h264bsd_macroblock_layer.c
// (:134-243 parse, :965-1131 driver) plus the motion-vector / intra-mode derivations that the
// reference performs inside its pixel functions (h264bsd_inter_prediction.c:494-1026,
// h264bsd_intra_prediction.c:627-833,:1886-1937).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <new>
#include <vector>
#include "bits.hpp"
#include "params.hpp"
#include "dpb.hpp"
#include "h264bsd_b200_tape.h"

namespace b200 {

// side state that is not part of the tape record (h264bsd_macroblock_layer.h:162-185)
struct MbAux {
    uint16_t sliceId = 0;
    uint8_t decoded = 0;
    uint8_t totalCoeff[27] = {0};
};

enum class SliceResult { Ok, Error };

// Who wants the records may lend the memory they are built in (the tape does: the records of a picture are then written
// once, in place, instead of state array -> record array -> tape).  nullptr = build them in the decoder's own arrays.
class RecordProvider {
public:
    virtual ~RecordProvider() {}
    virtual b200_mb_rec *pictureRecords(uint32_t nMbs) { (void)nMbs; return nullptr; }
};

class PictureState {
public:
    void resize(uint32_t widthMbs, uint32_t heightMbs);
    void beginPicture();  // h264bsdResetStorage: sliceId/decoded cleared, the rest persists
    uint32_t widthMbs = 0, heightMbs = 0, picSizeInMbs = 0;

    // st: per-MB state of the picture being decoded (last decode of each MB, incl. redundant slices); recs: the records of
    // the picture being built (first decode of each MB).  They are the SAME memory -- lent by the provider or ownRecs_ --
    // until a redundant slice shows up in the picture (then st moves to ownSt_).  Bound lazily per picture (bindOutput).
    b200_mb_rec *st = nullptr;
    b200_mb_rec *recs = nullptr;
    RecordProvider *provider = nullptr;
    std::vector<MbAux> aux;
    // coefficient pool of the picture being built, 16 int16 per block; grows without value-initialising what it hands out
    struct CoefPool {
        ~CoefPool() { std::free(p_); }
        const int16_t *data() const { return p_; }
        size_t size() const { return n_; }          // in int16
        void clear() { n_ = 0; }
        int16_t *grow(size_t count) {
            if (n_ + count > cap_) {
                size_t cap = cap_ ? cap_ : 4096;
                while (cap < n_ + count) cap *= 2;
                int16_t *q = static_cast<int16_t *>(std::realloc(p_, cap * sizeof(int16_t)));
                if (!q) throw std::bad_alloc();
                p_ = q; cap_ = cap;
            }
            int16_t *at = p_ + n_;
            n_ += count;
            return at;
        }
    private:
        int16_t *p_ = nullptr;
        size_t n_ = 0, cap_ = 0;
    } coefs;
    std::vector<uint16_t> order;    // concealment order of the picture being built (see b200_tape.mbOrder)
    uint32_t numPassA = 0, numPassB = 0, numCopy = 0, numConceal = 0;
    std::vector<uint16_t> concealOrder;   // spatially concealed macroblocks of the picture being built, concealment order
    std::vector<uint8_t> orderClass;   // scratch of finalizeRecords
    std::vector<uint32_t> sliceGroupMap;
    uint32_t sliceIdCounter = 0, numDecodedMbs = 0, lastMbAddr = 0;

    // decode one slice's macroblocks (h264bsdDecodeSliceData)
    SliceResult decodeSlice(BitReader &br, const SliceHeader &sh, const Sps &sps, const Pps &pps, const Dpb &dpb);
    void markSliceCorrupted(uint32_t firstMbInSlice, const Sps &sps);  // h264bsdMarkSliceCorrupted
    bool allDecoded(bool redundant) const;                             // h264bsdIsEndOfPicture
    // fill records of macroblocks that never arrived (error path; see DESIGN.md "concealment")
    uint32_t concealMissing(const Dpb &dpb, bool pSlice);
    // deblocking edge flags need the final slice ids of the whole picture (deblocking.c:289-320)
    void finalizeRecords();

private:
    struct MbSyntax;
    bool parseMacroblockLayer(BitReader &br, MbSyntax &mb, uint32_t mbAddr, bool iSlice, uint32_t numRefIdxActive);
    bool parseResidual(BitReader &br, MbSyntax &mb, uint32_t mbAddr);
    bool finishMacroblock(MbSyntax &mb, uint32_t mbAddr, int &qpY, const SliceHeader &sh, const Pps &pps, const Dpb &dpb);
    bool deriveInter(MbSyntax &mb, uint32_t mbAddr, const Dpb &dpb);
    bool deriveIntra(MbSyntax &mb, uint32_t mbAddr, bool constrainedIntra);
    void classify(uint32_t mbAddr, b200_mb_rec &rec);
    bool residualInRange(const MbSyntax &mb, bool i16, int qpY, int qpC, uint32_t mask) const;
    bool finishSkip(uint32_t mbAddr, int qpY, const SliceHeader &sh, const Pps &pps, int slot0);
    int nC(uint32_t mbAddr, uint32_t blk, const uint8_t *curTotalCoeff) const;
    uint32_t nextMbAddress(uint32_t cur) const;

    int mbA(uint32_t a) const { return (a % widthMbs) ? (int)a - 1 : -1; }
    int mbB(uint32_t a) const { return a >= widthMbs ? (int)(a - widthMbs) : -1; }
    int mbC(uint32_t a) const { return (a >= widthMbs && (a % widthMbs) < widthMbs - 1) ? (int)(a - widthMbs + 1) : -1; }
    int mbD(uint32_t a) const { return (a >= widthMbs && (a % widthMbs)) ? (int)(a - widthMbs - 1) : -1; }
    bool avail(uint32_t cur, int nb) const { return nb >= 0 && aux[nb].sliceId == aux[cur].sliceId; }
    void bindOutput();       // first touch of a picture: where do its records live
    void splitState();       // a redundant slice: st must no longer alias recs
    std::vector<b200_mb_rec> ownSt_, ownRecs_;
    bool bound_ = false;
    int curNb_[4] = {-1, -1, -1, -1};   // available neighbours A, B, C, D of the macroblock being decoded (decodeSlice)
    uint32_t curX_ = 0;                 // its column
    bool lateFixup_ = false;            // this picture needs the full pass of finalizeRecords (see classify)
    uint32_t numIntraPred_ = 0;         // intra-predicted macroblocks classified so far
    uint32_t numCopies_ = 0;            // zero-vector copies without residual classified so far
    int failedMb_ = -1;                 // macroblock whose derivation failed after it was counted as decoded (decodeSlice)
    uint8_t failedPrevDecoded_ = 0;     // ... and its decode count before that

    struct NbMv { bool avail; uint32_t refIdx; int16_t mv[2]; };
    NbMv interNeighbour(uint32_t cur, int x, int y, int curZ) const;
    bool predictMv(uint32_t cur, int x, int y, int w, int h, uint32_t refIdx, int dirHint, int16_t out[2],
                   const NbMv *preA = nullptr, const NbMv *preB = nullptr) const;
};

}  // namespace b200
