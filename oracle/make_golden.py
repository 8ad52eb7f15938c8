#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Regenerates tests/golden/md5.json from the UNMODIFIED reference decoder compiled
by oracle/Makefile (oracle/_ref/libh264bsd_ref.so).  Run in the build container (needs /root/reference for
`make -C oracle ref`); the JSON it writes is committed so that GPU-box tests need no reference.

Per stream: md5 of the concatenated output frames (coded size, picSizeInMbs*384 B each, output order), md5 of
the concatenated frames as they are handed to the in-loop filter (decode order), per-frame md5s of both, and
md5 of RGBA-then-BGRA conversion of the first two output pictures (h264bsdConvertToRGBA/BGRA)."""
import ctypes as C, hashlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _oracle

STREAMS = ["test_640x360.h264", "test_1920x1080.h264", "test_1920x1080_fullRange.h264"]

def main():
    ref = _oracle.reference()
    assert ref is not None, "build oracle/_ref first: make -C oracle ref"
    out = {}
    for name in STREAMS:
        data = _oracle.stream_bytes(name)
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        info = (C.c_uint32 * 8)()
        cap = 80 * 8160 * 384
        post = np.zeros(cap, np.uint8); pre = np.zeros(cap, np.uint8)
        n = ref.ref_decode_stream(buf, len(data), post.ctypes.data, cap, pre.ctypes.data, cap, None, 0, info)
        assert n > 0
        fb = info[0] * info[1] * 384
        W, H = info[0] * 16, info[1] * 16
        conv = hashlib.md5()
        for k in range(2):
            for mode in (0, 1):
                o = np.empty(W * H, np.uint32)
                frame = np.ascontiguousarray(post[k * fb:(k + 1) * fb])
                ref.ref_convert(mode, W, H, frame.ctypes.data, o.ctypes.data)
                conv.update(o.tobytes())
        out[name] = {
            "input_md5": hashlib.md5(data).hexdigest(), "pictures": n, "width_mbs": info[0], "height_mbs": info[1],
            "crop": [info[2], info[3], info[4], info[5], info[6]],
            "post_md5": hashlib.md5(post[:n * fb].tobytes()).hexdigest(),
            "pre_md5": hashlib.md5(pre[:info[7] * fb].tobytes()).hexdigest(),
            "post_frame_md5": [hashlib.md5(post[k * fb:(k + 1) * fb].tobytes()).hexdigest() for k in range(n)],
            "pre_frame_md5": [hashlib.md5(pre[k * fb:(k + 1) * fb].tobytes()).hexdigest() for k in range(info[7])],
            "rgba_bgra_first2_md5": conv.hexdigest(),
        }
        print(name, out[name]["post_md5"], out[name]["rgba_bgra_first2_md5"])
    with open(os.path.join(_oracle.GOLDEN, "md5.json"), "w") as f:
        json.dump(out, f, indent=1)

if __name__ == "__main__":
    main()
