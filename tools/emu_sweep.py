#!/usr/bin/env python
"""Randomised sweep of the engine's kernels run on the host (tests/emu): random synthetic / damaged streams, random tuning knobs
(chunkA, chunkB, copyRuns, filterChunk), 1..4 streams, 1..4 blocks -- every picture against the CPU oracle.  No GPU needed.
Because the emulated lanes of a warp are independent threads, a missing __syncwarp() shows up here as a mismatch.
usage: emu_sweep.py [rng seed] [trials]"""
import os
import random
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_h264                     # noqa: E402
import test_cpu_kernel_emu as T       # noqa: E402
from h264bsd_b200.batch import ParsedStream   # noqa: E402


def main():
    rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    emu = T.build_and_load()
    bad = done = 0
    for _ in range(trials):
        seed = rng.randrange(400)
        damaged = rng.random() < 0.3
        data = synth_h264.make_damaged_stream(seed) if damaged else synth_h264.make_stream(seed)
        ps = ParsedStream(data, resilient=damaged)
        if ps.status == 0 and ps.num_pics and ps.mbs_per_pic <= 36:
            knobs = (rng.randint(1, 8), rng.randint(1, 8), rng.randint(1, 16), rng.randint(1, 8))
            ns, blocks = rng.randint(1, 4), rng.randint(1, 4)
            try:
                T.run_engine_on_the_host(emu, ps, n_streams=ns, knobs=knobs, blocks=blocks, max_pics=5)
                done += 1
            except AssertionError as e:
                bad += 1
                print("FAIL seed", seed, "damaged", damaged, "knobs", knobs, "streams", ns, "blocks", blocks, str(e)[:120], flush=True)
        ps.close()
    print(f"{done} streams replayed, {bad} failed")


if __name__ == "__main__":
    main()
