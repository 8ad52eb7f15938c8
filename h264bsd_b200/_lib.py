"""ctypes binding of libh264bsd_b200.so (built in-tree by h264bsd_b200/build.sh).

The library is the product; this module only declares its C-ABI (include/h264bsd_decoder.h,
include/h264bsd_b200.h, include/h264bsd_b200_tape.h).  If the shared object is missing the import
fails loudly -- there is no Python or CPU fallback for the pixel path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB: another build of the same library (kernel experiments); the product always loads the in-tree default
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libh264bsd_b200.so")

STORAGE_BYTES = 4648


class PicHdr(C.Structure):
    _fields_ = [
        ("widthMbs", C.c_uint32), ("heightMbs", C.c_uint32), ("curSlot", C.c_uint32), ("numSlots", C.c_uint32),
        ("picIndex", C.c_uint32), ("isIdr", C.c_uint32), ("isRef", C.c_uint32), ("numCoefBlocks", C.c_uint32),
        ("mbRecOffset", C.c_uint64), ("coefOffset", C.c_uint64), ("numErrMbs", C.c_uint32), ("numOut", C.c_uint32),
        ("outSlot", C.c_uint8 * 20), ("outPicIndex", C.c_uint32 * 20), ("picId", C.c_uint32), ("numPassA", C.c_uint32),
        ("numPassB", C.c_uint32), ("numCopy", C.c_uint32), ("orderOffset", C.c_uint32), ("reserved7", C.c_uint32),
        ("numConceal", C.c_uint32), ("reserved5", C.c_uint32), ("filterRecOffset", C.c_uint64),
    ]


class Tape(C.Structure):
    _fields_ = [
        ("numPics", C.c_uint32), ("widthMbs", C.c_uint32), ("heightMbs", C.c_uint32), ("numSlots", C.c_uint32),
        ("cropFlag", C.c_uint32), ("cropLeft", C.c_uint32), ("cropWidth", C.c_uint32), ("cropTop", C.c_uint32),
        ("cropHeight", C.c_uint32), ("videoRange", C.c_uint32), ("matrixCoefficients", C.c_uint32), ("reserved", C.c_uint32),
        ("mbRecBytes", C.c_uint64), ("coefBytes", C.c_uint64),
        ("pics", C.POINTER(PicHdr)), ("mbRecs", C.POINTER(C.c_uint8)), ("coefs", C.POINTER(C.c_uint8)),
        ("mbOrder", C.POINTER(C.c_uint16)),
        ("numOutputs", C.c_uint32), ("numOrder", C.c_uint32), ("outputPicIndex", C.POINTER(C.c_uint32)),
        ("status", C.c_uint32), ("pinned", C.c_uint32),
        ("capRecs", C.c_uint64), ("capCoefs", C.c_uint64), ("capOrder", C.c_uint64), ("capPics", C.c_uint64),
        ("capOutputs", C.c_uint32), ("reserved4", C.c_uint32),
    ]


# every symbol the two public headers declare (tests check that each one is exported)
LEGACY_SYMBOLS = [
    "h264bsdInit", "h264bsdDecode", "h264bsdShutdown", "h264bsdNextOutputPicture", "h264bsdNextOutputPictureRGBA",
    "h264bsdNextOutputPictureBGRA", "h264bsdNextOutputPictureYCbCrA", "h264bsdPicWidth", "h264bsdPicHeight",
    "h264bsdVideoRange", "h264bsdMatrixCoefficients", "h264bsdCroppingParams", "h264bsdSampleAspectRatio",
    "h264bsdCheckValidParamSets", "h264bsdFlushBuffer", "h264bsdProfile", "h264bsdAlloc", "h264bsdFree",
    "h264bsdConvertToRGBA", "h264bsdConvertToBGRA", "h264bsdConvertToYCbCrA",
]
BATCH_SYMBOLS = [
    "h264bsdB200ParseStream", "h264bsdB200ReparseStream", "h264bsdB200ReparseStreams", "h264bsdB200ReparseStreamsBegin", "h264bsdB200ReparseStreamsWait", "h264bsdB200FreeTape", "h264bsdB200DeviceCount", "h264bsdB200BatchCreate",
    "h264bsdB200BatchDestroy", "h264bsdB200BatchUploadTape", "h264bsdB200BatchReplicateTape", "h264bsdB200BatchUploadTapeRange", "h264bsdB200BatchUploadFence", "h264bsdB200BatchUploadTapesRange",
    "h264bsdB200BatchDecodePicture", "h264bsdB200BatchRun", "h264bsdB200BatchSync", "h264bsdB200BatchNumPics",
    "h264bsdB200BatchTimerStart", "h264bsdB200BatchTimerStop", "h264bsdB200BatchReadFrame", "h264bsdB200BatchWriteFrame",
    "h264bsdB200BatchConvertFrame", "h264bsdB200BatchConvertBench", "h264bsdB200BatchConvertBenchAll", "h264bsdB200BatchCompareStreams",
    "h264bsdB200BatchDebugStage", "h264bsdB200BatchIdctErrors", "h264bsdB200BatchDeblockWorkMbs", "h264bsdB200BatchWatchdog", "h264bsdB200BatchReadPictureAll", "h264bsdB200HostAlloc", "h264bsdB200HostFree",
    "h264bsdB200PinTape", "h264bsdB200UnpinTape", "h264bsdB200BatchKernelTiming", "h264bsdB200BatchKernelTimes", "h264bsdB200BatchLaunches", "h264bsdB200BatchH2DBytes",
    "h264bsdB200BatchD2HBytes", "h264bsdB200ParseUploadPoolCreate", "h264bsdB200ParseUploadPoolDestroy", "h264bsdB200BatchParseUploadBegin",
    "h264bsdB200BatchParseUploadWait", "h264bsdB200BatchReadPictureAllEx",
]

_lib = None


def load():
    """Load the native library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run h264bsd_b200/build.sh (or __graft_entry__.build()); "
            "the B200 engine has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u32, u32p, u8p = C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    # legacy API
    L.h264bsdInit.restype = u32; L.h264bsdInit.argtypes = [vp, u32]
    L.h264bsdDecode.restype = u32; L.h264bsdDecode.argtypes = [vp, vp, u32, u32, u32p]
    L.h264bsdShutdown.restype = None; L.h264bsdShutdown.argtypes = [vp]
    for n in ("h264bsdNextOutputPicture", "h264bsdNextOutputPictureRGBA", "h264bsdNextOutputPictureBGRA",
              "h264bsdNextOutputPictureYCbCrA"):
        getattr(L, n).restype = vp; getattr(L, n).argtypes = [vp, u32p, u32p, u32p]
    for n in ("h264bsdPicWidth", "h264bsdPicHeight", "h264bsdVideoRange", "h264bsdMatrixCoefficients",
              "h264bsdCheckValidParamSets", "h264bsdProfile"):
        getattr(L, n).restype = u32; getattr(L, n).argtypes = [vp]
    L.h264bsdCroppingParams.restype = None; L.h264bsdCroppingParams.argtypes = [vp, u32p, u32p, u32p, u32p, u32p]
    L.h264bsdSampleAspectRatio.restype = None; L.h264bsdSampleAspectRatio.argtypes = [vp, u32p, u32p]
    L.h264bsdFlushBuffer.restype = None; L.h264bsdFlushBuffer.argtypes = [vp]
    L.h264bsdAlloc.restype = vp; L.h264bsdAlloc.argtypes = []
    L.h264bsdFree.restype = None; L.h264bsdFree.argtypes = [vp]
    for n in ("h264bsdConvertToRGBA", "h264bsdConvertToBGRA", "h264bsdConvertToYCbCrA"):
        getattr(L, n).restype = None; getattr(L, n).argtypes = [u32, u32, vp, vp]
    # batch API
    L.h264bsdB200ParseStream.restype = C.POINTER(Tape); L.h264bsdB200ParseStream.argtypes = [vp, C.c_size_t, u32]
    L.h264bsdB200ReparseStream.restype = C.POINTER(Tape); L.h264bsdB200ReparseStream.argtypes = [C.POINTER(Tape), vp, C.c_size_t, u32]
    L.h264bsdB200ReparseStreams.restype = C.c_int
    L.h264bsdB200ReparseStreamsBegin.restype = C.c_void_p
    L.h264bsdB200ReparseStreamsBegin.argtypes = [C.POINTER(C.POINTER(Tape)), u32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), u32, u32]
    L.h264bsdB200ReparseStreamsWait.restype = C.c_int; L.h264bsdB200ReparseStreamsWait.argtypes = [C.c_void_p]
    L.h264bsdB200ReparseStreams.argtypes = [C.POINTER(C.POINTER(Tape)), u32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), u32, u32]
    L.h264bsdB200FreeTape.restype = None; L.h264bsdB200FreeTape.argtypes = [C.POINTER(Tape)]
    L.h264bsdB200DeviceCount.restype = C.c_int; L.h264bsdB200DeviceCount.argtypes = []
    L.h264bsdB200BatchCreate.restype = vp; L.h264bsdB200BatchCreate.argtypes = [C.c_int, u32, u32, u32, u32]
    L.h264bsdB200BatchDestroy.restype = None; L.h264bsdB200BatchDestroy.argtypes = [vp]
    L.h264bsdB200BatchUploadTape.restype = C.c_int; L.h264bsdB200BatchUploadTape.argtypes = [vp, u32, C.POINTER(Tape)]
    L.h264bsdB200BatchReplicateTape.restype = C.c_int; L.h264bsdB200BatchReplicateTape.argtypes = [vp, u32]
    L.h264bsdB200BatchUploadTapeRange.restype = C.c_int; L.h264bsdB200BatchUploadTapeRange.argtypes = [vp, u32, C.POINTER(Tape), u32, u32]
    L.h264bsdB200BatchUploadFence.restype = C.c_int; L.h264bsdB200BatchUploadFence.argtypes = [vp, u32]
    L.h264bsdB200BatchUploadTapesRange.restype = C.c_int; L.h264bsdB200BatchUploadTapesRange.argtypes = [vp, C.POINTER(C.POINTER(Tape)), u32, u32, u32]
    L.h264bsdB200BatchDecodePicture.restype = C.c_int; L.h264bsdB200BatchDecodePicture.argtypes = [vp, u32]
    L.h264bsdB200BatchRun.restype = C.c_int; L.h264bsdB200BatchRun.argtypes = [vp, u32, u32]
    L.h264bsdB200BatchSync.restype = C.c_int; L.h264bsdB200BatchSync.argtypes = [vp]
    L.h264bsdB200BatchNumPics.restype = u32; L.h264bsdB200BatchNumPics.argtypes = [vp]
    L.h264bsdB200BatchTimerStart.restype = C.c_int; L.h264bsdB200BatchTimerStart.argtypes = [vp]
    L.h264bsdB200BatchTimerStop.restype = C.c_int; L.h264bsdB200BatchTimerStop.argtypes = [vp, C.POINTER(C.c_float)]
    L.h264bsdB200BatchReadFrame.restype = C.c_int; L.h264bsdB200BatchReadFrame.argtypes = [vp, u32, u32, vp]
    L.h264bsdB200BatchWriteFrame.restype = C.c_int; L.h264bsdB200BatchWriteFrame.argtypes = [vp, u32, u32, vp]
    L.h264bsdB200BatchConvertFrame.restype = C.c_int; L.h264bsdB200BatchConvertFrame.argtypes = [vp, u32, u32, C.c_int, vp]
    L.h264bsdB200BatchConvertBench.restype = C.c_int
    L.h264bsdB200BatchConvertBenchAll.restype = C.c_int
    L.h264bsdB200BatchConvertBenchAll.argtypes = [vp, u32, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.h264bsdB200BatchConvertBench.argtypes = [vp, u32, u32, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.h264bsdB200BatchCompareStreams.restype = C.c_int; L.h264bsdB200BatchCompareStreams.argtypes = [vp, u32p]
    L.h264bsdB200BatchDebugStage.restype = C.c_int; L.h264bsdB200BatchDebugStage.argtypes = [vp, u32, C.c_int, C.c_int]
    L.h264bsdB200BatchIdctErrors.restype = u32; L.h264bsdB200BatchIdctErrors.argtypes = [vp]
    L.h264bsdB200BatchDeblockWorkMbs.restype = C.c_uint64; L.h264bsdB200BatchDeblockWorkMbs.argtypes = [vp]
    L.h264bsdB200BatchKernelTiming.restype = None; L.h264bsdB200BatchKernelTiming.argtypes = [vp, C.c_int]
    L.h264bsdB200BatchKernelTimes.restype = C.c_int; L.h264bsdB200BatchKernelTimes.argtypes = [vp, C.POINTER(C.c_float), u32p]
    L.h264bsdB200BatchReadPictureAll.restype = C.c_int; L.h264bsdB200BatchReadPictureAll.argtypes = [vp, u32, vp, C.c_size_t]
    L.h264bsdB200HostAlloc.restype = vp; L.h264bsdB200HostAlloc.argtypes = [C.c_size_t]
    L.h264bsdB200HostFree.restype = None; L.h264bsdB200HostFree.argtypes = [vp]
    L.h264bsdB200PinTape.restype = C.c_int; L.h264bsdB200PinTape.argtypes = [C.POINTER(Tape)]
    L.h264bsdB200UnpinTape.restype = None; L.h264bsdB200UnpinTape.argtypes = [C.POINTER(Tape)]
    L.h264bsdB200BatchWatchdog.restype = u32; L.h264bsdB200BatchWatchdog.argtypes = [vp, C.c_int]
    for n in ("h264bsdB200BatchLaunches", "h264bsdB200BatchH2DBytes", "h264bsdB200BatchD2HBytes"):
        getattr(L, n).restype = C.c_uint64; getattr(L, n).argtypes = [vp]
    L.h264bsdB200ParseUploadPoolCreate.restype = vp; L.h264bsdB200ParseUploadPoolCreate.argtypes = [u32]
    L.h264bsdB200ParseUploadPoolDestroy.restype = None; L.h264bsdB200ParseUploadPoolDestroy.argtypes = [vp]
    L.h264bsdB200BatchParseUploadBegin.restype = vp
    L.h264bsdB200BatchParseUploadBegin.argtypes = [vp, vp, u32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), u32]
    L.h264bsdB200BatchParseUploadWait.restype = C.c_int; L.h264bsdB200BatchParseUploadWait.argtypes = [vp]
    L.h264bsdB200BatchReadPictureAllEx.restype = C.c_int
    L.h264bsdB200BatchReadPictureAllEx.argtypes = [vp, u32, vp, C.c_size_t, u32, u32, u32, u32, C.c_int]
    _lib = L
    return L
