#!/usr/bin/env python
"""GPU debugging aid: run every synthetic stream (tests/synth_h264.py) through the batched engine stage by stage and
print, for each stream that differs from the CPU oracle, where: picture, stage, macroblocks, their record fields and the
first differing sample.  usage: synth_gpu_report.py [first_seed last_seed]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle, synth_h264
from h264bsd_b200.batch import Batch, ParsedStream

REC = np.dtype([("mbType", "u1"), ("qpY", "u1"), ("qpC", "u1"), ("flags", "u1"), ("codedMask", "<u4"), ("coefIndex", "<u4"),
                ("foA", "i1"), ("foB", "i1"), ("cqo", "i1"), ("sub", "u1"), ("refSlot", "u1", 4), ("icm", "u1"), ("idc", "u1"),
                ("sliceId", "<u2"), ("refIdx", "u1", 4), ("wait", "u1"), ("r1", "u1", 3), ("mv", "<i2", (16, 2))])


def mb_diffs(got, want, W, H):
    """[(mb, plane, x, y, got, want)] first differing sample of every differing macroblock"""
    out = []
    ysz = W * H * 256
    gy, wy = got[:ysz].reshape(H * 16, W * 16), want[:ysz].reshape(H * 16, W * 16)
    gc, wc = got[ysz:].reshape(2, H * 8, W * 8), want[ysz:].reshape(2, H * 8, W * 8)
    for my in range(H):
        for mx in range(W):
            d = np.argwhere(gy[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16] != wy[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16])
            n = len(d)
            first = None
            if n:
                y, x = d[0]
                first = ("Y", int(x), int(y), int(gy[my * 16 + y, mx * 16 + x]), int(wy[my * 16 + y, mx * 16 + x]))
            for p in range(2):
                dc = np.argwhere(gc[p, my * 8:my * 8 + 8, mx * 8:mx * 8 + 8] != wc[p, my * 8:my * 8 + 8, mx * 8:mx * 8 + 8])
                if len(dc) and first is None:
                    y, x = dc[0]
                    first = ("Cb" if p == 0 else "Cr", int(x), int(y), int(gc[p, my * 8 + y, mx * 8 + x]), int(wc[p, my * 8 + y, mx * 8 + x]))
                n += len(dc)
            if n:
                out.append((my * W + mx, n) + first)
    return out


def main():
    lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 160)
    bad = 0
    for seed in range(lo, hi):
        ps = ParsedStream(synth_h264.make_stream(seed))
        if ps.status != 0:
            print(f"seed {seed}: parser status {ps.status}"); continue
        W, H = ps.width_mbs, ps.height_mbs
        t = ps.ptr.contents
        area = C.string_at(t.mbRecs, t.mbRecBytes)
        recs = np.concatenate([np.frombuffer(area, REC, count=W * H, offset=p.mbRecOffset) for p in ps.pics])
        orc = _oracle.OracleDecoder(ps)
        try:
            b = Batch(1, W, H, ps.num_slots)
        except Exception as e:
            print(f"seed {seed}: {W}x{H} slots {ps.num_slots}: Batch create failed: {e}"); bad += 1; continue
        b.upload(0, ps)
        reported = 0
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            for stage in ("recon", "deblock"):
                if stage == "recon":
                    b.debug_stage(k, True, False); orc.recon(k)
                else:
                    b.debug_stage(k, False, True); orc.deblock(k)
                got, want = b.read_frame(0, slot), orc.frame(slot)
                if not np.array_equal(got, want):
                    d = mb_diffs(got, want, W, H)
                    if reported < 3:
                        print(f"seed {seed}: {W}x{H} slots {ps.num_slots} pic {k}/{ps.num_pics} {stage}: {len(d)} MBs differ")
                        for (mb, n, pl, x, y, g, w) in d[:6]:
                            r = recs[k * W * H + mb]
                            print(f"   mb {mb} ({mb % W},{mb // W}) n={n} first {pl}({x},{y}) got {g} want {w} | type {r['mbType']} qp {r['qpY']}/{r['qpC']} flags {r['flags']:#x} "
                                  f"mask {r['codedMask']:#x} idc {r['idc']} offs {r['foA']},{r['foB']} sub {r['sub']:#x} slots {r['refSlot'].tolist()} mv0 {r['mv'][0].tolist()} slice {r['sliceId']}")
                    reported += 1
                    # continue from the oracle's state so that later pictures are judged on their own
                    b.write_frame(0, slot, want)
        wd = b.watchdog(); ie = b.idct_errors()
        if reported or wd != (0, 0) or ie:
            bad += 1
            print(f"seed {seed}: {reported} stage(s) differ, watchdog {wd}, idct errors {ie}", flush=True)
        b.close(); orc.close(); ps.close()
    print(f"{bad} of {hi - lo} streams differ")


if __name__ == "__main__":
    main()
