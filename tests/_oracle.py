"""ctypes binding of the CPU oracle (oracle/px_oracle.c) and, when present, of the compiled reference
(oracle/_ref/libh264bsd_ref.so).  TEST INFRASTRUCTURE: imported only by tests/, tools/ debug scripts,
__graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libh264bsd_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libh264bsd_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_orc = None
_ref = None


def build_oracle():
    # make knows the dependencies (px_oracle.c and the tape header); a no-op when up to date
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)


def oracle():
    global _orc
    if _orc is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.px_create.restype = C.c_void_p; L.px_create.argtypes = [C.c_uint32] * 3
        L.px_destroy.restype = None; L.px_destroy.argtypes = [C.c_void_p]
        L.px_frame.restype = C.POINTER(C.c_uint8); L.px_frame.argtypes = [C.c_void_p, C.c_uint32]
        L.px_recon_picture.restype = C.c_int; L.px_recon_picture.argtypes = [C.c_void_p] * 4
        L.px_deblock_picture.restype = None; L.px_deblock_picture.argtypes = [C.c_void_p] * 3
        L.px_run_tape.restype = C.c_int; L.px_run_tape.argtypes = [C.c_void_p] * 3
        L.px_convert.restype = None; L.px_convert.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _orc = L
    return _orc


def reference():
    """The unmodified reference decoder compiled by oracle/Makefile, or None if not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        L = C.CDLL(REF_SO)
        L.ref_decode_stream.restype = C.c_int
        L.ref_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_convert.restype = None; L.ref_convert.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.ref_sizeof_storage.restype = C.c_size_t
        _ref = L
    return _ref


def stream_bytes(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


class OracleDecoder:
    """Steps a parsed tape picture by picture on the CPU."""

    def __init__(self, parsed):
        self.L = oracle()
        self.p = parsed
        self.ctx = self.L.px_create(parsed.width_mbs, parsed.height_mbs, parsed.num_slots)
        self.fb = parsed.frame_bytes
        t = parsed.ptr.contents
        self._recs = C.addressof(t.mbRecs.contents)
        self._coefs = C.addressof(t.coefs.contents) if t.coefBytes else 0
        self._pics = t.pics

    def frame(self, slot):
        return np.ctypeslib.as_array(self.L.px_frame(self.ctx, slot), shape=(self.fb,))

    def recon(self, k):
        h = self._pics[k]
        return self.L.px_recon_picture(self.ctx, C.byref(h), self._recs + h.mbRecOffset, self._coefs + h.coefOffset)

    def deblock(self, k):
        h = self._pics[k]
        self.L.px_deblock_picture(self.ctx, C.byref(h), self._recs + (h.filterRecOffset or h.mbRecOffset))

    def close(self):
        if self.ctx:
            self.L.px_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def oracle_run_tape(parsed, want_pre=False):
    L = oracle()
    n_out = len(parsed.outputs)
    post = np.zeros(n_out * parsed.frame_bytes, np.uint8)
    pre = np.zeros(parsed.num_pics * parsed.frame_bytes, np.uint8) if want_pre else None
    errs = L.px_run_tape(C.cast(parsed.ptr, C.c_void_p), post.ctypes.data, pre.ctypes.data if want_pre else None)
    return post, pre, errs


def oracle_convert(mode, width, height, yuv):
    out = np.empty(width * height, np.uint32)
    y = np.ascontiguousarray(yuv, dtype=np.uint8)
    oracle().px_convert(mode, width, height, y.ctypes.data, out.ctypes.data)
    return out
