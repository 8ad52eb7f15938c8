#!/usr/bin/env python
"""YUV -> BGRA kernel at batch size (BASELINE.json config 5): GB/s for one 1080p frame of every stream per launch"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264bsd_b200.batch import Batch, ParsedStream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ps = ParsedStream(open(os.path.join(ROOT, "tests/golden/test_1920x1080_fullRange.h264"), "rb").read())
b = Batch(n, ps.width_mbs, ps.height_mbs, ps.num_slots)
b.upload(0, ps); b.replicate(0)
b.run(0, 4); b.sync()
slot = ps.pics[3].curSlot
b.convert_bench_all(slot, 1, 2)
ms = b.convert_bench_all(slot, 1, 5) / 5
nbytes = n * ps.width_mbs * 16 * ps.height_mbs * 16 * 5.5
from h264bsd_b200 import _lib
L = _lib.load()
fb = ps.frame_bytes
host = L.h264bsdB200HostAlloc(fb * n)
b.read_picture_all(3, host, fb); b.sync()          # packKernel: strip layout -> planar I420 of every stream, then one D2H
print(json.dumps({"streams": n, "ms_per_launch": ms, "GB_per_s": nbytes / (ms / 1e3) / 1e9}))
