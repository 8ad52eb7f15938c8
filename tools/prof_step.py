#!/usr/bin/env python
"""one short replay for ncu: N streams of the 1080p tape, pictures [0, P)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264bsd_b200.batch import Batch, ParsedStream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ps = ParsedStream(open(os.path.join(ROOT, "tests/golden/test_1920x1080.h264"), "rb").read())
b = Batch(n, ps.width_mbs, ps.height_mbs, ps.num_slots)
b.upload(0, ps); b.replicate(0)
b.run(0, P); b.sync()
print("done", b.watchdog())
