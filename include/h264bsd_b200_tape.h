/*
 * h264bsd_b200_tape.h -- the record-then-replay work-list ("tape") that separates the
 * host-side H.264 Baseline syntax decoder from the B200 pixel engine.
 *
 * Where the reference hands one parsed macroblock at a time straight to its pixel code
 * (h264bsd_slice_data.c:185 -> h264bsdDecodeMacroblock, h264bsd_macroblock_layer.c:965) and
 * one finished picture to the in-loop filter (h264bsd_decoder.c:475 -> h264bsdFilterPicture),
 * this engine RECORDS, per picture, one fixed-size record per macroblock (raster order) plus a pool of
 * packed int16 coefficient blocks, and the GPU REPLAYS whole pictures (of many streams).
 *
 * Everything syntax-level is finished on the host before a record is written: QP update
 * (macroblock_layer.c:1040-1046), chroma QP (:1401), motion-vector prediction
 * (h264bsd_inter_prediction.c:494-909), Intra4x4 mode derivation
 * (h264bsd_intra_prediction.c:1886-1937), neighbour availability incl. constrained-intra
 * masking (:644-655,:730-767), left/top edge filtering flags (h264bsd_deblocking.c:289-320)
 * and reference picture -> frame-slot resolution (h264bsd_dpb.c:847).  Everything that
 * touches a pel (inverse zig-zag, dequantisation, IDCT, prediction, add, in-loop filter)
 * is left to the consumer of the tape.
 *
 * Plain C, fixed-width types, no pointers inside records: the same bytes are consumed by
 * the CUDA kernels (h264bsd_b200/csrc), by the CPU oracle (oracle/px_oracle.c) and by tests.
 */
#ifndef H264BSD_B200_TAPE_H
#define H264BSD_B200_TAPE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mbType uses the reference's numbering (h264bsd_macroblock_layer.h:47-85) */
enum {
    B200_MB_P_SKIP = 0,
    B200_MB_P_16x16 = 1,
    B200_MB_P_16x8 = 2,
    B200_MB_P_8x16 = 3,
    B200_MB_P_8x8 = 4,
    B200_MB_P_8x8REF0 = 5,
    B200_MB_I_4x4 = 6,
    B200_MB_I_16x16_FIRST = 7, /* 7..30: pred mode = (t-7)&3, see macroblock_layer.c:924 */
    B200_MB_I_PCM = 31
};

/* b200_mb_rec.flags */
#define B200_MBF_AVAIL_A 0x01u /* left MB usable for intra prediction   */
#define B200_MBF_AVAIL_B 0x02u /* above MB                              */
#define B200_MBF_AVAIL_C 0x04u /* above-right MB                        */
#define B200_MBF_AVAIL_D 0x08u /* above-left MB                         */
#define B200_MBF_FILTER_LEFT 0x10u  /* filter the left MB edge  (deblocking.c:289-320) */
#define B200_MBF_FILTER_TOP 0x20u   /* filter the top MB edge                           */
#define B200_MBF_FILTER_INNER 0x40u /* filter inner edges (disable_deblocking_filter_idc != 1) */
#define B200_MBF_CONCEALED 0x80u    /* record synthesised for a macroblock missing from the stream (see below) */
/*
 * Concealed macroblocks (h264bsd_conceal.c:124-639).  When a picture ends with macroblocks missing (lost or corrupted
 * slices) the host writes their records the way h264bsdConceal leaves the reference's state: mbType = B200_MB_I_4x4,
 * qpY = 40, filter offsets and chromaQpIndexOffset 0, filtering enabled (ConcealMb :296-306) -- that is what the in-loop
 * filter sees.  How the pels are made:
 *   copy     (P slice and a reference picture exists, :320-341): waitMask == 0, refSlot[] = that picture, u = 0: pass A moves
 *            the co-located macroblock like any zero-vector copy;
 *   spatial  (otherwise, :346-600): waitMask = B200_CN_* bits of the neighbouring macroblocks whose edge pels enter the
 *            estimate (decoded or concealed earlier), coefIndex = position in the concealment order; the record is listed in
 *            b200_tape.mbOrder, in that order (every entry may read what the previous ones wrote).
 * A picture of which nothing arrived is copied from the reference / set to 128 with the filter off (:172-201): ordinary
 * P_Skip / I_PCM records with disable_deblocking_filter_idc = 1.
 */
#define B200_CN_ABOVE 0x01u
#define B200_CN_BELOW 0x02u
#define B200_CN_LEFT 0x04u
#define B200_CN_RIGHT 0x08u

/* b200_mb_rec.codedMask */
#define B200_CM_LUMA_DC (1u << 24)   /* Intra16x16 luma DC block present   */
#define B200_CM_CHROMA_DC (1u << 25) /* chroma DC (Cb4+Cr4) block present  */

#define B200_MB_REC_BYTES 96
#define B200_COEF_BLOCK_BYTES 32 /* 16 x int16, zig-zag order exactly as parsed */

/*
 * One macroblock.  96 bytes (three 32-byte sectors): D = 96 bytes of work-list per macroblock enter the roofline accounting.
 *
 * coefficient pool layout for this MB, starting at 32-byte block index `coefIndex`
 * (relative to the picture's pool):
 *     I_PCM:   12 blocks = the 384 raw samples (256 Y raster 16x16, 64 Cb, 64 Cr)
 *     else:    [luma DC block  if codedMask&B200_CM_LUMA_DC ]   16 levels, zig-zag order
 *              [chroma DC block if codedMask&B200_CM_CHROMA_DC]  Cb dc[0..3], Cr dc[0..3], 8 pad
 *              then one block per set bit b of codedMask&0xFFFFFF, ascending b
 *              (b = 0..15 luma in the standard's 4x4 block order, 16..19 Cb, 20..23 Cr);
 *              Intra16x16 / chroma AC blocks keep levels at zig-zag positions 1..15, [0] = 0
 *              (macroblock_layer.c:745-746,785-786).
 */
typedef struct b200_mb_rec {
    uint8_t mbType;           /*  0 */
    uint8_t qpY;              /*  1 luma QP after mb_qp_delta; 0 for I_PCM (macroblock_layer.c:996) */
    uint8_t qpC;              /*  2 h264bsdQpC[clip3(0,51,qpY+chromaQpIndexOffset)] */
    uint8_t flags;            /*  3 B200_MBF_* */
    uint32_t codedMask;       /*  4 bit b<24: totalCoeff[b] != 0 ; bits 24,25 see above */
    uint32_t coefIndex;       /*  8 */
    int8_t filterOffsetA;     /* 12 slice_alpha_c0_offset_div2*2 */
    int8_t filterOffsetB;     /* 13 slice_beta_offset_div2*2 */
    int8_t chromaQpIndexOffset; /* 14 */
    uint8_t subMbTypes;       /* 15 2 bits per 8x8 quadrant: 0 8x8, 1 8x4, 2 4x8, 3 4x4 */
    uint8_t refSlot[4];       /* 16 frame slot of the reference picture per 8x8 quadrant */
    uint8_t intraChromaMode;  /* 20 0 DC, 1 horizontal, 2 vertical, 3 plane */
    uint8_t reserved0;        /* 21 */
    uint16_t sliceId;         /* 22 (diagnostic; availability is already resolved in flags) */
    uint8_t refIdx[4];        /* 24 ref_idx_l0 per quadrant (diagnostic / MV-prediction state) */
    uint8_t waitMask;         /* 28 intra MBs: B200_MBF_AVAIL_* bits of the neighbours that are themselves intra-predicted in
                                    this picture (the only macroblocks an intra MB must wait for inside the intra pass) */
    uint8_t reserved1[3];     /* 29 */
    union {                   /* 32 */
        int16_t mv[16][2];    /* inter: {hor,ver} quarter-pel per 4x4 block, standard block order */
        struct {
            uint8_t i4x4Mode[16]; /* intra 4x4: final prediction mode per block */
            uint8_t pad[48];
        } intra;
    } u;
} b200_mb_rec;

/* One picture of one stream. */
typedef struct b200_pic_hdr {
    uint32_t widthMbs;
    uint32_t heightMbs;
    uint32_t curSlot;      /* frame slot this picture is reconstructed into */
    uint32_t numSlots;     /* frame slots the stream uses (dpbSize + 1) */
    uint32_t picIndex;     /* decode order index within the stream */
    uint32_t isIdr;
    uint32_t isRef;        /* nal_ref_idc != 0 */
    uint32_t numCoefBlocks;  /* 32-byte blocks in this picture's coefficient pool */
    uint64_t mbRecOffset;  /* byte offset of widthMbs*heightMbs records inside the tape's record area */
    uint64_t coefOffset;   /* byte offset of the coefficient pool inside the tape's coefficient area */
    uint32_t numErrMbs;    /* macroblocks synthesised (concealed) */
    uint32_t numOut;       /* pictures that become available for output after this one ... */
    uint8_t outSlot[20];   /* ... their frame slots, output order ... */
    uint32_t outPicIndex[20]; /* ... and the decode-order index of the picture held in that slot */
    uint32_t picId;        /* application picId (h264bsdDecode argument) */
    uint32_t numPassA;     /* macroblocks reconstructed without looking at the current picture: inter + I_PCM (+ concealed copies) */
    uint32_t numPassB;     /* intra-predicted macroblocks (read unfiltered neighbours of the current picture) */
    uint32_t numCopy;      /* of numPassA: one 16x16 partition, no residual, zero vector (and concealed copies) -- a plain copy of the
                              co-located macroblock of the reference frame (statistic; the GPU finds them itself) */
    uint32_t orderOffset;  /* where this picture's concealment order starts inside the tape's mbOrder (entries) */
    uint32_t reserved7;
    uint32_t numConceal;   /* spatially concealed macroblocks: the picture's entries in mbOrder, concealment order */
    uint32_t reserved5;
    uint64_t filterRecOffset; /* 0: the in-loop filter reads the records at mbRecOffset.  Else byte offset of a second record
                              array for the filter alone: a picture in which redundant slices decoded macroblocks a second
                              time keeps the pels of the first decode but is filtered with the state of the last one, as in
                              the reference (h264bsd_macroblock_layer.c:1003-1007,:1108-1111) */
} b200_pic_hdr;

/* tape.status when the stream switches to another picture size: the tape ends with the last picture of the old size
 * (a tape -- and the frame pool of the batched engine -- holds one size; h264bsdDecode handles such streams) */
#define B200_TAPE_SIZE_CHANGE 100u

/* A fully parsed stream in host memory (built by h264bsdB200ParseStream). */
typedef struct b200_tape {
    uint32_t numPics;
    uint32_t widthMbs;
    uint32_t heightMbs;
    uint32_t numSlots;
    uint32_t cropFlag, cropLeft, cropWidth, cropTop, cropHeight;
    uint32_t videoRange, matrixCoefficients;
    uint32_t reserved;
    uint64_t mbRecBytes;
    uint64_t coefBytes;
    b200_pic_hdr *pics;  /* numPics */
    uint8_t *mbRecs;     /* mbRecBytes */
    uint8_t *coefs;      /* coefBytes  */
    /* concealment order: the numConceal spatially concealed macroblocks of a picture (addresses) at pics[p].orderOffset, numOrder
     * entries in all; every entry may read
     * what the previous ones wrote.  Everything else the GPU sorts itself from the records, which lie in raster order: pass A takes
     * the inter / I_PCM / concealed-copy macroblocks in the order they lie in memory, pass B and the in-loop filter walk macroblock
     * rows (round 1 shipped four more sections here: zero-motion runs, single copies, other pass-A macroblocks, intra macroblocks in
     * wavefront order). */
    uint16_t *mbOrder;
    uint32_t numOutputs;        /* pictures in output order, incl. those drained by the final flush */
    uint32_t numOrder;          /* entries in mbOrder */
    uint32_t *outputPicIndex;   /* numOutputs decode-order indices */
    uint32_t status;            /* 0 ok, else the H264BSD_* code the parse stopped on, or B200_TAPE_SIZE_CHANGE */
    uint32_t pinned;            /* 0 pageable, 1 arrays page-locked (h264bsdB200PinTape), 2 page-lock stale after growth */
    /* allocation sizes of the arrays (they are kept when a tape is re-used, h264bsdB200ReparseStream) */
    uint64_t capRecs, capCoefs, capOrder, capPics;
    uint32_t capOutputs, reserved4;
} b200_tape;

#ifdef __cplusplus
}
#endif

#endif /* H264BSD_B200_TAPE_H */
