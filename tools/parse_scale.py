#!/usr/bin/env python
"""host parser scaling: N threads each re-parsing the 1080p stream into its own tape (pageable memory, no GPU)"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concurrent.futures import ThreadPoolExecutor
from h264bsd_b200.batch import ParsedStream
data = open(os.path.join(ROOT, "tests/golden/test_1920x1080.h264"), "rb").read()
bits = (C.c_uint8 * len(data)).from_buffer_copy(data)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cpu.max n/a", e)
for nt in [int(a) for a in sys.argv[1:]] or [1, 8, 32, 64, 128]:
    tapes = [ParsedStream(bits) for _ in range(nt)]
    pool = ThreadPoolExecutor(max_workers=nt)
    def one(i):
        t = time.time(); tapes[i].reparse(bits); return time.time() - t
    for rep in range(2):
        t0 = time.time(); r = list(pool.map(one, range(nt))); dt = time.time() - t0
    print(nt, "threads: wall %.3f s, thread mean %.3f max %.3f, aggregate %.1f M MB/s" % (dt, sum(r) / nt, max(r), nt * 595680 / dt / 1e6), flush=True)
    for t in tapes: t.close()
    pool.shutdown()
