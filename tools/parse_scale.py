#!/usr/bin/env python
"""host parser scaling (no GPU): N native threads (h264bsdB200ReparseStreams, what bench.py's end-to-end leg uses) re-parsing
the 1080p stream into 4 tapes each, pageable memory.   usage: parse_scale.py [threads ...]"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264bsd_b200.batch import ParsedStream
data = open(os.path.join(ROOT, "tests/golden/test_1920x1080.h264"), "rb").read()
bits = (C.c_uint8 * len(data)).from_buffer_copy(data)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cpu.max n/a", e)
MBS = 595680
for nt in [int(a) for a in sys.argv[1:]] or [1, 8, 16]:
    n = 4 * nt
    tapes = [ParsedStream(bits) for _ in range(n)]
    best = 1e9
    for rep in range(3):
        t0 = time.time()
        ParsedStream.reparse_many(tapes, bits, nt)
        best = min(best, time.time() - t0)
    print("%d threads, %d streams: %.3f s -> %.1f M MB/s aggregate, %.3f s per stream and thread" % (nt, n, best, n * MBS / best / 1e6, best * nt / n), flush=True)
    for t in tapes:
        t.close()
