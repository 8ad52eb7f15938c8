#!/usr/bin/env python
"""Writes tests/golden/synth_md5.json: for every seed of tests/synth_h264.py the md5 of the stream itself (so that a
change of the generator or of Python's `random` shows up as such) and the md5 of what the UNMODIFIED reference
decoder (oracle/_ref/libh264bsd_ref.so, built by oracle/Makefile from /root/reference) makes of it: all output
pictures at full coded size, output order, plus the pictures before the in-loop filter in decoding order.
Run in the container that has /root/reference; the tests then need neither it nor oracle/_ref."""
import ctypes as C
import hashlib
import json
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _oracle          # noqa: E402
import synth_h264       # noqa: E402

SEEDS = range(0, 160)
DAMAGED_SEEDS = range(0, 240)
# larger pictures, mostly still scenes: rows wider than a copy run (32 macroblocks), runs cut by slice and slice-group borders
LARGE = [(1, dict(W=45, H=18, still=True, pictures=4)), (2, dict(W=45, H=18, still=True, pictures=4)),
         (11, dict(W=40, H=23, still=True, pictures=3)), (12, dict(W=40, H=23, still=True, pictures=3)),
         (21, dict(W=64, H=4, still=True, pictures=5)), (32, dict(W=45, H=36, still=True, pictures=2)),
         (31, dict(W=45, H=36, pictures=2)), (51, dict(W=120, H=3, still=True, pictures=3)),
         # 1080p-size pictures (level 4): the list sizes, ticket counts and chunking of the real workload with slices / FMO
         (77, dict(W=120, H=68, still=True, pictures=2)), (78, dict(W=120, H=68, pictures=2))]


def reference_decode(data):
    """(number of output pictures or -1, frame bytes, output frames, pre-filter frames, decoded pictures)"""
    L = _oracle.reference()
    info = (C.c_uint32 * 8)()
    cap = 1 << 25
    post = np.zeros(cap, np.uint8)
    pre = np.zeros(cap, np.uint8)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    n = L.ref_decode_stream(buf, len(data), post.ctypes.data, cap, pre.ctypes.data, cap, None, 0, info)
    fb = info[0] * info[1] * 384
    assert max(n, 0) * fb <= cap and info[7] * fb <= cap
    return n, fb, post[:max(n, 0) * fb], pre[:info[7] * fb], int(info[7]), (int(info[0]), int(info[1]))


def reference_stream_info(data):
    """what the reference's getters say after the whole stream: h264bsdCroppingParams (flag, left, width, top, height),
    h264bsdVideoRange, h264bsdMatrixCoefficients"""
    L = _oracle.reference()
    info = (C.c_uint32 * 8)()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    L.ref_decode_stream(buf, len(data), None, 0, None, 0, None, 0, info)
    v = (C.c_uint32 * 4)()
    L.ref_video_info(v)
    return {"crop": [int(x) for x in info[2:7]], "video_range": int(v[0]), "matrix_coefficients": int(v[1])}


def reference_decode_resilient(data):
    """the same with a caller that carries on after H264BSD_ERROR (what a player does; posix/test_h264bsd.c exits): the
    reference marks slices corrupt and conceals what is missing at the next access unit boundary.  Also returns numErrMbs
    of every output picture."""
    L = _oracle.reference()
    L.ref_set_resilient.argtypes = [C.c_int]
    L.ref_err_mbs.restype = C.c_uint32
    L.ref_set_resilient(1)
    try:
        n, fb, post, pre, ndec, dims = reference_decode(data)
    finally:
        L.ref_set_resilient(0)
    errs = (C.c_uint32 * 4096)()
    ne = L.ref_err_mbs(errs, 4096)
    return n, fb, post, pre, ndec, dims, [int(v) for v in errs[:ne]]


def damaged_golden():
    """md5s of what the reference makes of the damaged streams (tests/synth_h264.py make_damaged_stream)"""
    out = {}
    for seed in DAMAGED_SEEDS:
        data = synth_h264.make_damaged_stream(seed)
        n, fb, post, pre, ndec, dims, errs = reference_decode_resilient(data)
        out[str(seed)] = {"stream_md5": hashlib.md5(data).hexdigest(), "width_mbs": dims[0], "height_mbs": dims[1],
                          "outputs": n, "decoded": ndec, "err_mbs": errs,
                          "post_md5": hashlib.md5(post.tobytes()).hexdigest(), "pre_md5": hashlib.md5(pre.tobytes()).hexdigest()}
    return out


def main():
    if _oracle.reference() is None:
        sys.exit("oracle/_ref/libh264bsd_ref.so is not built (make -C oracle ref)")
    path = os.path.join(_oracle.GOLDEN, "synth_damaged_md5.json")
    with open(path, "w") as f:
        json.dump(damaged_golden(), f, indent=0, sort_keys=True)
    print("wrote", path)
    out = {}
    for seed in SEEDS:
        data = synth_h264.make_stream(seed)
        n, fb, post, pre, ndec, dims = reference_decode(data)
        assert n >= 0, f"seed {seed}: the reference reports a decode error -- the generator wrote an invalid stream"
        out[str(seed)] = {"stream_md5": hashlib.md5(data).hexdigest(), "bytes": len(data), "width_mbs": dims[0],
                          "height_mbs": dims[1], "outputs": n, "decoded": ndec, "info": reference_stream_info(data),
                          "post_md5": hashlib.md5(post.tobytes()).hexdigest(), "pre_md5": hashlib.md5(pre.tobytes()).hexdigest()}
    for i, (seed, kw) in enumerate(LARGE):
        data = synth_h264.make_stream(seed, **kw)
        n, fb, post, pre, ndec, dims = reference_decode(data)
        assert n >= 0, f"large stream {i}: the reference reports a decode error"
        out[f"L{i}"] = {"stream_md5": hashlib.md5(data).hexdigest(), "bytes": len(data), "width_mbs": dims[0], "height_mbs": dims[1],
                        "outputs": n, "decoded": ndec, "seed": seed, "knobs": kw,
                        "post_md5": hashlib.md5(post.tobytes()).hexdigest(), "pre_md5": hashlib.md5(pre.tobytes()).hexdigest()}
    path = os.path.join(_oracle.GOLDEN, "synth_md5.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, len(out), "seeds")


if __name__ == "__main__":
    main()
