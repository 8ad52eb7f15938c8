#!/usr/bin/env python
"""stage timing of a short replay (no parity check): streams, pictures"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264bsd_b200.batch import Batch, ParsedStream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ps = ParsedStream(open(os.path.join(ROOT, "tests/golden/test_1920x1080.h264"), "rb").read())
b = Batch(n, ps.width_mbs, ps.height_mbs, ps.num_slots)
b.upload(0, ps); b.replicate(0)
b.run(0, ps.num_pics); b.sync()
b.timer_start()
for _ in range(reps): b.run(0, ps.num_pics)
ms_free = b.timer_stop()
b.kernel_timing(True)
b.timer_start()
for _ in range(reps): b.run(0, ps.num_pics)
ms = b.timer_stop()
st, cnt = b.kernel_times()
mbs = n * ps.num_pics * ps.mbs_per_pic * reps
print(json.dumps({"streams": n, "ms_per_pass_concurrent": ms_free / reps, "MB_per_s_concurrent": mbs / (ms_free / 1000), "ms_per_pass": ms / reps, "MB_per_s": mbs / (ms / 1000), "stage_ms_per_pass": {k: v / reps for k, v in st.items()},
                  "launches": cnt, "watchdog": b.watchdog()}))
