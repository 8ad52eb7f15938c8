#!/usr/bin/env python
"""G batches side by side on their own CUDA streams: streams are independent, so the 512-stream
batch can be cut into G groups that each run their own picture sequence; the block scheduler then places the filter of one group
next to pass A of another next to the copies of a third -- kernels whose limits differ (issue slots / DRAM latency / dependency
chains).  No engine change: G Batch objects, pictures issued round-robin (picture k of group g, then of group g + 1, ...),
optionally offset against each other by `stagger` pictures so that the groups are in different stages at any time.

usage: group_bench.py [total_streams=512] [groups=4] [passes=2] [stagger=0]
env:   B200_GRID_DIV=G caps every persistent grid at 1/G of the resident CTAs (try G = 1, 2, groups)
Prints one JSON line: macroblocks/s over all groups (host clock around `passes` passes, all groups synchronised on both sides)
and whether the last picture of the first and last stream of every group has the reference's md5."""
import sys, os, json, time, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264bsd_b200.batch import Batch, ParsedStream

total = int(sys.argv[1]) if len(sys.argv) > 1 else 512
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 4
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
stagger = int(sys.argv[4]) if len(sys.argv) > 4 else 0
per = total // groups
ps = ParsedStream(open(os.path.join(ROOT, "tests/golden/test_1920x1080.h264"), "rb").read())
gold = json.load(open(os.path.join(ROOT, "tests/golden/md5.json")))["test_1920x1080.h264"]["post_frame_md5"]
np_ = ps.num_pics
batches = []
for g in range(groups):
    b = Batch(per, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    b.replicate(0)
    batches.append(b)


def issue(n_passes):
    """pictures of all groups, round-robin; group g runs `stagger * g` pictures behind group 0"""
    steps = n_passes * np_
    for t in range(steps + stagger * (groups - 1)):
        for g, b in enumerate(batches):
            k = t - stagger * g
            if 0 <= k < steps:
                b.decode_picture(k % np_)


issue(1)                                    # warm-up pass (job tables, first-touch)
for b in batches:
    b.sync()
t0 = time.time()
issue(passes)
for b in batches:
    b.sync()
dt = time.time() - t0
ok = True
for b in batches:
    for s in (0, per - 1):
        ok = ok and hashlib.md5(b.read_frame(s, ps.pics[-1].curSlot).tobytes()).hexdigest() == gold[-1]
    ok = ok and b.watchdog() == (0, 0) and b.idct_errors() == 0
mbs = groups * per * np_ * ps.mbs_per_pic * passes
print(json.dumps({"streams": groups * per, "groups": groups, "stagger": stagger, "grid_div": os.environ.get("B200_GRID_DIV", "1"),
                  "ms_per_pass": dt / passes * 1e3, "MB_per_s": mbs / dt, "bit_exact": bool(ok),
                  "launches": sum(b.launches() for b in batches)}))
for b in batches:
    b.close()
