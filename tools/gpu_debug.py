#!/usr/bin/env python
"""GPU bring-up helper (run under gpurun): stage-isolated comparison of the CUDA engine with the CPU
oracle, picture by picture.  Prints where the first differences are and which macroblock types they hit."""
import sys, os, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle
from h264bsd_b200.batch import Batch, ParsedStream

REC_DT = np.dtype([('mbType','u1'),('qpY','u1'),('qpC','u1'),('flags','u1'),('codedMask','<u4'),('coefIndex','<u4'),
                   ('fa','i1'),('fb','i1'),('cqo','i1'),('sub','u1'),('refSlot','u1',4),('icm','u1'),('idc','u1'),
                   ('sliceId','<u2'),('refIdx','u1',4),('r1','u1',4),('mv','<i2',(16,2))])

def mb_diff(a, b, wm, hm):
    """set of macroblock indices whose pels differ between two coded-size I420 frames"""
    W, H = wm*16, hm*16
    bad = set()
    ya, yb = a[:W*H].reshape(H, W), b[:W*H].reshape(H, W)
    d = (ya != yb)
    if d.any():
        blk = d.reshape(hm, 16, wm, 16).any(axis=(1, 3))
        for y, x in zip(*np.nonzero(blk)): bad.add((int(y)*wm+int(x), 'Y'))
    for pl in range(2):
        o = W*H + pl*(W*H//4)
        ca, cb = a[o:o+W*H//4].reshape(H//2, W//2), b[o:o+W*H//4].reshape(H//2, W//2)
        d = (ca != cb)
        if d.any():
            blk = d.reshape(hm, 8, wm, 8).any(axis=(1, 3))
            for y, x in zip(*np.nonzero(blk)): bad.add((int(y)*wm+int(x), 'C'))
    return bad

def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('DBG_TIMEOUT', '40')), exit=True)
    name = sys.argv[1] if len(sys.argv) > 1 else 'test_640x360.h264'
    npics = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    data = _oracle.stream_bytes(name)
    ps = ParsedStream(data)
    wm, hm = ps.width_mbs, ps.height_mbs
    print(f"{name}: {ps.num_pics} pics {wm}x{hm} MBs slots {ps.num_slots}", flush=True)
    orc = _oracle.OracleDecoder(ps)
    b = Batch(1, wm, hm, ps.num_slots)
    b.upload(0, ps)
    t = ps.ptr.contents
    recs_all = np.ctypeslib.as_array(t.mbRecs, shape=(t.mbRecBytes,)).view(REC_DT)
    nmb = wm*hm
    tot_bad = 0
    for k in range(min(npics, ps.num_pics)):
        h = ps.pics[k]
        recs = recs_all[k*nmb:(k+1)*nmb]
        # state before picture k: every slot holds the oracle's frames
        for s in range(ps.num_slots): b.write_frame(0, s, orc.frame(s))
        orc.recon(k)
        pre = orc.frame(h.curSlot).copy()
        b.debug_stage(k, True, False); b.sync()
        if any(b.watchdog()): print("   watchdog after recon", b.watchdog(), flush=True)
        got = b.read_frame(0, h.curSlot)
        bad = mb_diff(got, pre, wm, hm)
        if bad:
            tot_bad += 1
            types = collections.Counter((int(recs['mbType'][m]), pl) for m, pl in bad)
            first = sorted(bad)[:6]
            print(f"pic {k} RECON: {len(bad)} bad (mb,plane); by (mbType,plane): {dict(types)}; first {first}", flush=True)
            m = first[0][0]
            W = wm*16
            y0, x0 = (m//wm)*16, (m % wm)*16
            ga = got[:W*hm*16].reshape(hm*16, W)[y0:y0+16, x0:x0+16]; pa = pre[:W*hm*16].reshape(hm*16, W)[y0:y0+16, x0:x0+16]
            print("   mb", m, "type", recs['mbType'][m], "mask", hex(recs['codedMask'][m]), "mv0", recs['mv'][m][0], "flags", hex(recs['flags'][m]))
            print("   got row0", ga[0].tolist()); print("   exp row0", pa[0].tolist())
            dd = np.nonzero(ga != pa); print("   diff positions (y,x) first:", list(zip(dd[0][:8].tolist(), dd[1][:8].tolist())))
        else:
            print(f"pic {k} RECON ok", flush=True)
        # deblock in isolation from the oracle's unfiltered picture
        b.write_frame(0, h.curSlot, pre)
        orc.deblock(k)
        post = orc.frame(h.curSlot).copy()
        import time
        t0 = time.time()
        b.debug_stage(k, False, True)
        b.sync()
        print('   deblock kernel wall', round(time.time() - t0, 4), 'watchdog', b.watchdog(), flush=True)
        got = b.read_frame(0, h.curSlot)
        bad = mb_diff(got, post, wm, hm)
        if bad:
            tot_bad += 1
            types = collections.Counter((int(recs['mbType'][m]), pl) for m, pl in bad)
            print(f"pic {k} DEBLOCK: {len(bad)} bad; by (mbType,plane): {dict(types)}; first {sorted(bad)[:6]}", flush=True)
        else:
            print(f"pic {k} DEBLOCK ok", flush=True)
    print("idct errors", b.idct_errors(), "watchdog", b.watchdog(), flush=True)
    # free-running decode of the whole stream
    b2 = Batch(1, wm, hm, ps.num_slots)
    b2.upload(0, ps)
    orc2 = _oracle.OracleDecoder(ps)
    nbad = 0
    for k in range(ps.num_pics):
        b2.decode_picture(k)
        orc2.recon(k); orc2.deblock(k)
        got = b2.read_frame(0, ps.pics[k].curSlot)
        if not np.array_equal(got, orc2.frame(ps.pics[k].curSlot)):
            nbad += 1
            if nbad <= 3: print(f"free-run pic {k}: mismatch", len(mb_diff(got, orc2.frame(ps.pics[k].curSlot).copy(), wm, hm)))
    print(f"free-run: {nbad} of {ps.num_pics} pictures differ; isolated stages with differences: {tot_bad}")

if __name__ == '__main__':
    main()
