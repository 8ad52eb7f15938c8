"""Batched engine: many independent streams per GPU (include/h264bsd_b200.h)."""
import ctypes as C
import numpy as np
from . import _lib


class ParsedStream:
    """A stream parsed on the host into a tape (h264bsdB200ParseStream)."""

    def __init__(self, data=None, no_output_reordering=False, resilient=False):
        self._L = _lib.load()
        self.ptr = None
        self.pinned = False
        self.status = None
        if data is not None:      # (None: an empty shell to be filled by reparse_many)
            self.reparse(data, no_output_reordering, resilient)

    def reparse(self, data, no_output_reordering=False, resilient=False):
        """parse another stream into the same tape (arrays and page-lock are kept).  resilient: carry on after decode
        errors; what is missing from a picture is concealed (B200_PARSE_RESILIENT)"""
        buf = data if isinstance(data, C.Array) else (C.c_uint8 * len(data)).from_buffer_copy(bytes(data))
        self.ptr = self._L.h264bsdB200ReparseStream(self.ptr, buf, len(buf), (1 if no_output_reordering else 0) | (2 if resilient else 0))
        if not self.ptr:
            raise MemoryError("h264bsdB200ParseStream failed")
        t = self.ptr.contents
        self.pinned = t.pinned == 1
        self.num_pics, self.width_mbs, self.height_mbs, self.num_slots = t.numPics, t.widthMbs, t.heightMbs, t.numSlots
        self.status = t.status
        self.rec_bytes, self.coef_bytes = t.mbRecBytes, t.coefBytes
        self.outputs = [t.outputPicIndex[i] for i in range(t.numOutputs)]
        self.pics = [t.pics[i] for i in range(t.numPics)]
        self.video_range = t.videoRange

    def _refresh(self):
        t = self.ptr.contents
        self.pinned = t.pinned == 1
        self.num_pics, self.width_mbs, self.height_mbs, self.num_slots = t.numPics, t.widthMbs, t.heightMbs, t.numSlots
        self.status = t.status
        self.rec_bytes, self.coef_bytes = t.mbRecBytes, t.coefBytes
        self.video_range = t.videoRange

    @staticmethod
    def reparse_many(parsed, data, threads, no_output_reordering=False):
        """re-parse `data` (one ctypes byte array shared by all, or a list of them) into every ParsedStream of `parsed` on
        `threads` native host threads (one C call, no interpreter lock); light refresh of the Python-side fields"""
        L = _lib.load()
        n = len(parsed)
        bufs = data if isinstance(data, (list, tuple)) else [data] * n
        tapes = (C.POINTER(_lib.Tape) * n)(*[p.ptr for p in parsed])
        ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
        lens = (C.c_size_t * n)(*[len(b) for b in bufs])
        bad = L.h264bsdB200ReparseStreams(tapes, n, ptrs, lens, 1 if no_output_reordering else 0, threads)
        if bad:
            raise MemoryError("h264bsdB200ReparseStreams failed for %d streams" % bad)
        for i, p in enumerate(parsed):
            p.ptr = tapes[i]
            p._refresh()

    @staticmethod
    def reparse_many_begin(parsed, data, threads, no_output_reordering=False):
        """reparse_many in the background (native threads only); returns a token for reparse_many_wait"""
        L = _lib.load()
        n = len(parsed)
        bufs = data if isinstance(data, (list, tuple)) else [data] * n
        tapes = (C.POINTER(_lib.Tape) * n)(*[p.ptr for p in parsed])
        ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
        lens = (C.c_size_t * n)(*[len(b) for b in bufs])
        job = L.h264bsdB200ReparseStreamsBegin(tapes, n, ptrs, lens, 1 if no_output_reordering else 0, threads)
        if not job:
            raise MemoryError("h264bsdB200ReparseStreamsBegin failed")
        return (job, parsed, tapes, ptrs, lens, bufs)     # keeps the arrays alive

    @staticmethod
    def reparse_many_wait(token):
        job, parsed, tapes = token[0], token[1], token[2]
        bad = _lib.load().h264bsdB200ReparseStreamsWait(job)
        if bad:
            raise MemoryError("h264bsdB200ReparseStreams failed for %d streams" % bad)
        for i, p in enumerate(parsed):
            p.ptr = tapes[i]
            p._refresh()

    @property
    def mbs_per_pic(self):
        return self.width_mbs * self.height_mbs

    @property
    def frame_bytes(self):
        return self.mbs_per_pic * 384

    def coded_blocks(self):
        return sum(p.numCoefBlocks for p in self.pics)

    def pin(self):
        """page-lock the tape for full-speed uploads"""
        if not self.pinned and self._L.h264bsdB200PinTape(self.ptr) == 0:
            self.pinned = True
        return self.pinned

    def close(self):
        if self.ptr:
            if self.pinned:
                self._L.h264bsdB200UnpinTape(self.ptr)
            self._L.h264bsdB200FreeTape(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    def __init__(self, n_streams, width_mbs, height_mbs, num_slots, device=0):
        self._L = _lib.load()
        self.h = self._L.h264bsdB200BatchCreate(device, n_streams, width_mbs, height_mbs, num_slots)
        if not self.h:
            raise RuntimeError("h264bsdB200BatchCreate failed (no usable CUDA device? the engine has no CPU fallback)")
        self.n_streams, self.width_mbs, self.height_mbs, self.num_slots = n_streams, width_mbs, height_mbs, num_slots
        self.frame_bytes = width_mbs * height_mbs * 384

    def _ck(self, r, what):
        if r != 0:
            raise RuntimeError(f"{what} failed")

    def upload(self, stream, parsed):
        self._ck(self._L.h264bsdB200BatchUploadTape(self.h, stream, parsed.ptr), "upload")

    def upload_range(self, stream, parsed, first_pic, num_pics):
        """streamed upload of pictures [first_pic, first_pic + num_pics) on the copy stream (first_pic == 0 first)"""
        self._ck(self._L.h264bsdB200BatchUploadTapeRange(self.h, stream, parsed.ptr, first_pic, num_pics), "upload_range")

    def upload_ranges(self, parsed_list, first_pic, num_pics):
        """upload_range for streams 0..len-1 in one call, then the fence for these pictures"""
        from . import _lib
        arr = (C.POINTER(_lib.Tape) * len(parsed_list))(*[p.ptr for p in parsed_list])
        self._ck(self._L.h264bsdB200BatchUploadTapesRange(self.h, arr, len(parsed_list), first_pic, num_pics), "upload_ranges")

    def upload_fence(self, through_pic):
        self._ck(self._L.h264bsdB200BatchUploadFence(self.h, through_pic), "upload_fence")

    def parse_upload_begin(self, pool, bufs, flags=0):
        """parse the bitstreams `bufs` (ctypes byte arrays, one per stream of this batch) on the pool's host threads and upload
        each work-list as it is finished (h264bsdB200BatchParseUploadBegin); returns a token for parse_upload_wait"""
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
        lens = (C.c_size_t * n)(*[len(b) for b in bufs])
        job = self._L.h264bsdB200BatchParseUploadBegin(self.h, pool, n, ptrs, lens, flags)
        if not job:
            raise MemoryError("h264bsdB200BatchParseUploadBegin failed")
        return (job, ptrs, lens, bufs)     # keeps the arrays alive

    def parse_upload_wait(self, token):
        bad = self._L.h264bsdB200BatchParseUploadWait(token[0])
        if bad:
            raise RuntimeError("parse + upload failed for %d streams" % bad)

    def read_picture_all_ex(self, k, dst, stride, crop=(0, 0, 0, 0), nv12=False):
        """picture k of every stream, cropped to crop = (x, y, w, h) (w == 0: coded size) and / or as NV12 -> dst + s * stride"""
        self._ck(self._L.h264bsdB200BatchReadPictureAllEx(self.h, k, dst, stride, crop[0], crop[1], crop[2], crop[3], int(nv12)), "read_picture_all_ex")

    def replicate(self, src=0):
        self._ck(self._L.h264bsdB200BatchReplicateTape(self.h, src), "replicate")

    def decode_picture(self, k):
        self._ck(self._L.h264bsdB200BatchDecodePicture(self.h, k), "decode_picture")

    def run(self, first, count):
        self._ck(self._L.h264bsdB200BatchRun(self.h, first, count), "run")

    def debug_stage(self, k, recon, deblock):
        self._ck(self._L.h264bsdB200BatchDebugStage(self.h, k, int(recon), int(deblock)), "debug_stage")

    def sync(self):
        self._ck(self._L.h264bsdB200BatchSync(self.h), "sync")

    def timer_start(self):
        self._ck(self._L.h264bsdB200BatchTimerStart(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self._L.h264bsdB200BatchTimerStop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def read_frame(self, stream, slot):
        out = np.empty(self.frame_bytes, np.uint8)
        self._ck(self._L.h264bsdB200BatchReadFrame(self.h, stream, slot, out.ctypes.data), "read_frame")
        return out

    def read_picture_all(self, k, dst_ptr, stride):
        self._ck(self._L.h264bsdB200BatchReadPictureAll(self.h, k, dst_ptr, stride), 'read_picture_all')

    def write_frame(self, stream, slot, frame):
        f = np.ascontiguousarray(frame, dtype=np.uint8)
        assert f.size == self.frame_bytes
        self._ck(self._L.h264bsdB200BatchWriteFrame(self.h, stream, slot, f.ctypes.data), "write_frame")

    def convert_frame(self, stream, slot, mode):
        out = np.empty(self.width_mbs * self.height_mbs * 256, np.uint32)
        self._ck(self._L.h264bsdB200BatchConvertFrame(self.h, stream, slot, mode, out.ctypes.data), "convert_frame")
        return out

    def convert_bench(self, stream, slot, mode, reps):
        ms = C.c_float(0)
        self._ck(self._L.h264bsdB200BatchConvertBench(self.h, stream, slot, mode, reps, C.byref(ms)), "convert_bench")
        return ms.value

    def convert_bench_all(self, slot, mode, reps):
        ms = C.c_float(0)
        self._ck(self._L.h264bsdB200BatchConvertBenchAll(self.h, slot, mode, reps, C.byref(ms)), "convert_bench_all")
        return ms.value

    def compare_streams(self, slots):
        arr = (C.c_uint32 * self.n_streams)(*slots)
        r = self._L.h264bsdB200BatchCompareStreams(self.h, arr)
        if r < 0:
            raise RuntimeError("compare_streams failed")
        return r

    def idct_errors(self):
        return self._L.h264bsdB200BatchIdctErrors(self.h)

    def kernel_timing(self, enable):
        self._L.h264bsdB200BatchKernelTiming(self.h, int(enable))

    def kernel_times(self):
        """({'recon': ms, 'deblock': ms, 'border': ms}, launches per stage) since the last call"""
        ms = (C.c_float * 6)()
        n = (C.c_uint32 * 6)()
        self._ck(self._L.h264bsdB200BatchKernelTimes(self.h, ms, n), 'kernel_times')
        keys = ('recon', 'deblock', 'border', 'recon_intra', 'strength', 'recon_multi')
        return ({k: ms[i] for i, k in enumerate(keys)}, {k: n[i] for i, k in enumerate(keys)})

    def deblock_work_mbs(self):
        return int(self._L.h264bsdB200BatchDeblockWorkMbs(self.h))

    def watchdog(self):
        return (self._L.h264bsdB200BatchWatchdog(self.h, 0), self._L.h264bsdB200BatchWatchdog(self.h, 1))

    def launches(self):
        return self._L.h264bsdB200BatchLaunches(self.h)

    def h2d_bytes(self):
        return self._L.h264bsdB200BatchH2DBytes(self.h)

    def d2h_bytes(self):
        return self._L.h264bsdB200BatchD2HBytes(self.h)

    def close(self):
        if self.h:
            self._L.h264bsdB200BatchDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
