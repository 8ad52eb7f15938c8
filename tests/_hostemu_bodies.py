"""TEST INFRASTRUCTURE: what tests/test_cpu_engine_hostemu.py runs, in a process of its own, against the host build of the engine
(B200_LIB points the bindings at tests/emu/_build/libh264bsd_b200_hostemu.so).  Small synthetic streams: the emulation is slow."""
import hashlib
import json
import os
import numpy as np
import _oracle
import synth_h264
from h264bsd_b200 import _lib
from h264bsd_b200.batch import Batch, ParsedStream
from h264bsd_b200.decoder import H264bsdDecoder, PIC_RDY, ERROR, PARAM_SET_ERROR, MEMALLOC_ERROR

assert "hostemu" in _lib.LIB_PATH, "these bodies are for the host build of the engine only"
GOLD = json.load(open(os.path.join(_oracle.GOLDEN, "synth_md5.json")))
DAMAGED = json.load(open(os.path.join(_oracle.GOLDEN, "synth_damaged_md5.json")))


def small(g, limit=40):
    return g["width_mbs"] * g["height_mbs"] <= limit


def batched():
    """Batch: upload, replicate, decode_picture, debug_stage, read_frame, read_picture_all, compare_streams"""
    seeds = [int(s) for s in sorted(GOLD, key=lambda s: (len(s), s)) if not s.startswith("L") and small(GOLD[s])][:3]
    seeds += [14, 48]     # pictures with filter-only records (macroblocks decoded twice by redundant slices)
    pics = filter_recs = 0
    for i, seed in enumerate(seeds):
        data = synth_h264.make_stream(seed)
        assert hashlib.md5(data).hexdigest() == GOLD[str(seed)]["stream_md5"]
        ps = ParsedStream(data)
        assert ps.status == 0
        orc = _oracle.OracleDecoder(ps)
        b = Batch(2, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        b.replicate(0)
        fb = ps.frame_bytes
        both = np.zeros(2 * fb, np.uint8)
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            if k % 2:
                b.debug_stage(k, True, False)
                orc.recon(k)
                assert np.array_equal(b.read_frame(1, slot), orc.frame(slot)), f"seed {seed}: reconstruction of picture {k}"
                b.debug_stage(k, False, True)
                orc.deblock(k)
            else:
                b.decode_picture(k)
                orc.recon(k)
                orc.deblock(k)
            assert np.array_equal(b.read_frame(1, slot), orc.frame(slot)), f"seed {seed}: picture {k}"
            assert b.compare_streams([slot, slot]) == 0, f"seed {seed}: the two instances differ at picture {k}"
            b.read_picture_all(k, both.ctypes.data, fb)
            b.sync()
            assert np.array_equal(both[:fb], orc.frame(slot)) and np.array_equal(both[fb:], orc.frame(slot)), f"seed {seed}: packed read-back of picture {k}"
            pics += 1
            filter_recs += ps.pics[k].filterRecOffset != 0
            if k == 1 and i < 2:
                # colour conversion (convertKernel) against the oracle's h264bsdConvertToRGBA / BGRA / YCbCrA
                W, H = ps.width_mbs * 16, ps.height_mbs * 16
                for mode in (0, 1, 2):
                    assert np.array_equal(b.convert_frame(1, slot, mode), _oracle.oracle_convert(mode, W, H, orc.frame(slot))), f"seed {seed}: conversion mode {mode}"
        assert b.idct_errors() == 0 and b.watchdog() == (0, 0), f"seed {seed}"
        b.close()
        if i in (1, 3):
            # the same stream again, work-lists uploaded in groups of two pictures on the upload stream (what bench.py's
            # end-to-end leg does): every picture of both instances against the frames the oracle holds at the end
            b = Batch(2, ps.width_mbs, ps.height_mbs, ps.num_slots)
            b.upload_ranges([ps, ps], 0, min(2, ps.num_pics))
            for g0 in range(0, ps.num_pics, 2):
                if g0 + 2 < ps.num_pics:
                    b.upload_ranges([ps, ps], g0 + 2, min(2, ps.num_pics - g0 - 2))
                for k in range(g0, min(g0 + 2, ps.num_pics)):
                    b.decode_picture(k)
            b.sync()
            last = ps.pics[ps.num_pics - 1].curSlot
            assert np.array_equal(b.read_frame(0, last), orc.frame(last)) and np.array_equal(b.read_frame(1, last), orc.frame(last)), f"seed {seed}: streamed upload"
            b.close()
        orc.close(); ps.close()
    # damaged streams: concealKernel, concealed copies, filter over concealed macroblocks
    concealed = 0
    for seed in [int(s) for s in sorted(DAMAGED, key=int) if small(DAMAGED[s]) and DAMAGED[s]["outputs"] > 0 and sum(DAMAGED[s]["err_mbs"]) > 0][:2]:
        ps = ParsedStream(synth_h264.make_damaged_stream(seed), resilient=True)
        if ps.status != 0 or ps.num_pics == 0:
            ps.close()
            continue
        orc = _oracle.OracleDecoder(ps)
        b = Batch(1, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            b.decode_picture(k)
            orc.recon(k)
            orc.deblock(k)
            assert np.array_equal(b.read_frame(0, slot), orc.frame(slot)), f"damaged seed {seed}: picture {k}"
            concealed += ps.pics[k].numErrMbs > 0
        assert b.watchdog() == (0, 0)
        b.close(); orc.close(); ps.close()
    assert pics >= 12 and concealed >= 2 and filter_recs >= 3, (pics, concealed, filter_recs)
    print(f"batched ok: {pics} pictures, {concealed} with concealment, {filter_recs} with filter-only records")


def legacy_decode(data, resilient):
    d = H264bsdDecoder(False)
    frames = []
    d.queueInput(data)
    while d.inputBytesRemaining() > 0:
        r = d.decode()
        if r == PIC_RDY:
            while (f := d.nextOutputPicture()) is not None:
                frames.append(f)
        elif r in (ERROR, PARAM_SET_ERROR, MEMALLOC_ERROR) and not resilient:
            raise RuntimeError(f"h264bsdDecode returned {r}")
    d.flush()
    while (f := d.nextOutputPicture()) is not None:
        frames.append(f)
    d.release()
    return frames


def legacy():
    """h264bsdInit / Decode / NextOutputPicture / Shutdown"""
    def digest(frames):
        h = hashlib.md5()
        for f in frames:
            h.update(np.ascontiguousarray(f).tobytes())
        return h.hexdigest()
    n = 0
    for seed in [int(s) for s in sorted(GOLD, key=lambda s: (len(s), s)) if not s.startswith("L") and small(GOLD[s])][:4] + [14, 18]:
        g = GOLD[str(seed)]
        frames = legacy_decode(synth_h264.make_stream(seed), False)
        assert len(frames) == g["outputs"] and digest(frames) == g["post_md5"], f"seed {seed}: output pictures differ from the reference"
        n += 1
    # h264bsdNextOutputPictureRGBA: the first output picture of a stream, converted on the device
    data = synth_h264.make_stream(0)
    first = legacy_decode(data, False)[0]
    d = H264bsdDecoder(False)
    d.queueInput(data)
    while d.decode() != PIC_RDY:
        pass
    rgba = d.nextOutputPictureRGBA()
    d.release()
    assert np.array_equal(rgba, _oracle.oracle_convert(0, GOLD["0"]["width_mbs"] * 16, GOLD["0"]["height_mbs"] * 16, first)), "NextOutputPictureRGBA"
    nd = 0
    for seed in [int(s) for s in sorted(DAMAGED, key=int) if small(DAMAGED[s]) and DAMAGED[s]["outputs"] > 0 and sum(DAMAGED[s]["err_mbs"]) > 0][:2]:
        g = DAMAGED[str(seed)]
        frames = legacy_decode(synth_h264.make_damaged_stream(seed), True)
        assert len(frames) == g["outputs"] and digest(frames) == g["post_md5"], f"damaged seed {seed}: output pictures differ from the reference"
        nd += 1
    assert n >= 4 and nd >= 2, (n, nd)
    print(f"legacy ok: {n} valid and {nd} damaged streams")


def sweep(first, count, limit=60):
    """many synthetic streams through the host build of the engine, two instances, stage by stage against the oracle
    (development aid: python -c 'import conftest, _hostemu_bodies as t; t.sweep(0, 100)' with B200_LIB set)"""
    seeds = [int(s) for s in sorted(GOLD, key=lambda s: (len(s), s)) if not s.startswith("L") and small(GOLD[s], limit)][first:first + count]
    pics = 0
    for seed in seeds:
        ps = ParsedStream(synth_h264.make_stream(seed))
        assert ps.status == 0
        orc = _oracle.OracleDecoder(ps)
        b = Batch(2, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        b.replicate(0)
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            b.debug_stage(k, True, False)
            orc.recon(k)
            assert np.array_equal(b.read_frame(1, slot), orc.frame(slot)), f"seed {seed}: reconstruction of picture {k}"
            b.debug_stage(k, False, True)
            orc.deblock(k)
            assert np.array_equal(b.read_frame(0, slot), orc.frame(slot)), f"seed {seed}: picture {k}"
            pics += 1
        assert b.idct_errors() == 0 and b.watchdog() == (0, 0), f"seed {seed}"
        b.close(); orc.close(); ps.close()
        print(f"seed {seed} ok", flush=True)
    print(f"sweep ok: {len(seeds)} streams, {pics} pictures")


def replay(ps, n_streams, max_pics=None, stages=True):
    """a tape through the host build of the engine next to the oracle, `n_streams` instances, every picture of every instance"""
    orc = _oracle.OracleDecoder(ps)
    b = Batch(n_streams, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    if n_streams > 1:
        b.replicate(0)
    n = ps.num_pics if max_pics is None else min(ps.num_pics, max_pics)
    for k in range(n):
        slot = ps.pics[k].curSlot
        if stages:
            b.debug_stage(k, True, False)
            orc.recon(k)
            for st in range(n_streams):
                assert np.array_equal(b.read_frame(st, slot), orc.frame(slot)), f"reconstruction of picture {k}, instance {st}"
            b.debug_stage(k, False, True)
            orc.deblock(k)
        else:
            b.decode_picture(k)
            orc.recon(k)
            orc.deblock(k)
        for st in range(n_streams):
            assert np.array_equal(b.read_frame(st, slot), orc.frame(slot)), f"picture {k}, instance {st}"
    assert b.idct_errors() == 0 and b.watchdog() == (0, 0)
    b.close(); orc.close()
    return n


def kernels(kind):
    """what the kernels see beyond the few streams of batched(): all macroblock types, several reference frames, vectors far outside
    the picture (valid); concealment and concealed copies (damaged); long columns of copies in pictures taller than a pass-A chunk
    and wider than a filter stretch (large); an encoder's pictures -- every interpolation position -- with three instances, an odd
    count for the filter's half-warp pairs (reference)"""
    pics = 0
    if kind == "valid":
        for seed in range(0, 26):
            ps = ParsedStream(synth_h264.make_stream(seed))
            if ps.status == 0 and ps.num_pics and ps.mbs_per_pic <= 40:
                pics += replay(ps, 2, max_pics=6)
            ps.close()
        assert pics >= 40, pics
    elif kind == "damaged":
        for seed in range(0, 40):
            ps = ParsedStream(synth_h264.make_damaged_stream(seed), resilient=True)
            if ps.status == 0 and ps.num_pics and ps.mbs_per_pic <= 40:
                pics += replay(ps, 2, max_pics=6, stages=False)
            ps.close()
        assert pics >= 40, pics
    elif kind == "large":
        for data, mp in ((synth_h264.make_stream(2, W=45, H=18, still=True, pictures=4), 3), (synth_h264.make_stream(21, W=64, H=4, still=True, pictures=5), 4),
                         (synth_h264.make_stream(12, W=3, H=35, still=True, pictures=4), 4)):
            ps = ParsedStream(data)
            assert ps.status == 0
            pics += replay(ps, 2, max_pics=mp, stages=False)
            ps.close()
        assert pics >= 10, pics
    else:
        ps = ParsedStream(_oracle.stream_bytes("test_640x360.h264"))
        pics += replay(ps, 3, max_pics=5, stages=False)
        ps.close()
    print(f"kernels ok: {kind}, {pics} pictures")
