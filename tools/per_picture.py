#!/usr/bin/env python
"""per-picture stage times of one replay: streams [pictures-to-print]; one JSON object per line"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264bsd_b200.batch import Batch, ParsedStream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ps = ParsedStream(open(os.path.join(ROOT, "tests/golden/test_1920x1080.h264"), "rb").read())
b = Batch(n, ps.width_mbs, ps.height_mbs, ps.num_slots)
b.upload(0, ps); b.replicate(0)
b.run(0, ps.num_pics); b.sync()
b.kernel_timing(True)
tot = {}
for p in range(ps.num_pics):
    b.run(p, 1); b.sync()
    st, _ = b.kernel_times()
    h = ps.pics[p]
    for k, v in st.items(): tot[k] = tot.get(k, 0.0) + v
    print(json.dumps({"pic": p, "nA": h.numPassA, "nB": h.numPassB, "coef": h.numCoefBlocks, **{k: round(v, 3) for k, v in st.items()}}))
print(json.dumps({"streams": n, "total_ms": {k: round(v, 2) for k, v in tot.items()}, "watchdog": b.watchdog()}))
