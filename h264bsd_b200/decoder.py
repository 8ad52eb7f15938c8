"""Host-side mirror of the reference's front-end wrapper for the preserved C API.

`H264bsdDecoder` has the shape of the reference's JS wrapper (wasm/h264bsd_decoder.js:36-339:
queueInput / decode / nextOutputPicture / outputPictureWidth / croppingParams, same return codes), so
tests read like the reference's own (wasm/test_node.js:36-55).  Every pel it returns was produced by
the GPU through libh264bsd_b200.so; there is no Python pixel code.
"""
import ctypes as C
import numpy as np
from . import _lib

RDY, PIC_RDY, HDRS_RDY, ERROR, PARAM_SET_ERROR, MEMALLOC_ERROR = range(6)
NO_INPUT = 1024


class H264bsdDecoder:
    RDY, PIC_RDY, HDRS_RDY, ERROR, PARAM_SET_ERROR, MEMALLOC_ERROR, NO_INPUT = RDY, PIC_RDY, HDRS_RDY, ERROR, PARAM_SET_ERROR, MEMALLOC_ERROR, NO_INPUT

    def __init__(self, no_output_reordering=False):
        self._L = _lib.load()
        self._storage = self._L.h264bsdAlloc()
        if not self._storage:
            raise MemoryError("h264bsdAlloc failed")
        if self._L.h264bsdInit(self._storage, 1 if no_output_reordering else 0) != 0:
            self._L.h264bsdFree(self._storage)
            self._storage = None
            raise RuntimeError("h264bsdInit failed (no usable CUDA device? the engine has no CPU fallback)")
        self._buf = None
        self._pos = 0
        self.onPictureReady = None
        self.onHeadersReady = None

    def release(self):
        if self._storage:
            self._L.h264bsdShutdown(self._storage)
            self._L.h264bsdFree(self._storage)
            self._storage = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def queueInput(self, data):
        rest = b"" if self._buf is None else bytes(self._buf[self._pos:])
        self._buf = (C.c_uint8 * (len(rest) + len(data))).from_buffer_copy(rest + bytes(data))
        self._pos = 0

    def inputBytesRemaining(self):
        return 0 if self._buf is None else len(self._buf) - self._pos

    def decode(self):
        if self.inputBytesRemaining() == 0:
            return NO_INPUT
        rb = C.c_uint32(0)
        ptr = C.addressof(self._buf) + self._pos
        ret = self._L.h264bsdDecode(self._storage, ptr, len(self._buf) - self._pos, 0, C.byref(rb))
        self._pos += rb.value
        if ret == PIC_RDY and callable(self.onPictureReady):
            self.onPictureReady()
        if ret == HDRS_RDY and callable(self.onHeadersReady):
            self.onHeadersReady()
        return ret

    def _meta(self):
        return C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)

    def nextOutputPicture(self):
        a, b, c = self._meta()
        p = self._L.h264bsdNextOutputPicture(self._storage, C.byref(a), C.byref(b), C.byref(c))
        if not p:
            return None
        n = self.outputPictureSizeBytes()
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)).copy()

    def _next_u32(self, fn):
        a, b, c = self._meta()
        p = fn(self._storage, C.byref(a), C.byref(b), C.byref(c))
        if not p:
            return None
        n = self.outputPictureWidth() * self.outputPictureHeight()
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,)).copy()

    def nextOutputPictureRGBA(self):
        return self._next_u32(self._L.h264bsdNextOutputPictureRGBA)

    def nextOutputPictureBGRA(self):
        return self._next_u32(self._L.h264bsdNextOutputPictureBGRA)

    def nextOutputPictureYCbCrA(self):
        return self._next_u32(self._L.h264bsdNextOutputPictureYCbCrA)

    def outputPictureWidth(self):
        return self._L.h264bsdPicWidth(self._storage) * 16

    def outputPictureHeight(self):
        return self._L.h264bsdPicHeight(self._storage) * 16

    def outputPictureSizeBytes(self):
        return self.outputPictureWidth() * self.outputPictureHeight() * 3 // 2

    def outputPictureSizeBytesRGBA(self):
        return self.outputPictureWidth() * self.outputPictureHeight() * 4

    def croppingParams(self):
        v = [C.c_uint32(0) for _ in range(5)]
        self._L.h264bsdCroppingParams(self._storage, *[C.byref(x) for x in v])
        if not v[0].value:
            return None
        return {"left": v[1].value, "width": v[2].value, "top": v[3].value, "height": v[4].value}

    def videoRange(self):
        return self._L.h264bsdVideoRange(self._storage)

    def flush(self):
        self._L.h264bsdFlushBuffer(self._storage)


def decode_stream(data, no_output_reordering=False):
    """posix/test_h264bsd.c:146-177 decode loop: returns the list of output frames (coded-size I420)."""
    d = H264bsdDecoder(no_output_reordering)
    frames = []
    d.queueInput(data)
    while d.inputBytesRemaining() > 0:
        r = d.decode()
        if r == PIC_RDY:
            while True:
                f = d.nextOutputPicture()
                if f is None:
                    break
                frames.append(f)
        elif r in (ERROR, PARAM_SET_ERROR, MEMALLOC_ERROR):
            d.release()
            raise RuntimeError(f"h264bsdDecode returned {r}")
    d.flush()
    while True:
        f = d.nextOutputPicture()
        if f is None:
            break
        frames.append(f)
    d.release()
    return frames
