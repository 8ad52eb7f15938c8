"""The shipped sources of two device kernels -- the spatial concealment kernel (conceal_kernel.cuh) and the copy pass
(copy_kernel.cuh) -- compiled for the host and run lane by lane (tests/emu/warp_emu.hpp: 32 threads per warp, a barrier per
shuffle) against the CPU oracle.  Checks a kernel's indexing and arithmetic where no GPU is at hand (the concealment path was
written without one); the GPU suite checks the same pictures on the hardware."""
import copy
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import _oracle
import synth_h264
from h264bsd_b200.batch import ParsedStream

ROOT = _oracle.ROOT
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libkernels_emu.so")


def build_and_load():
    """compile tests/emu/kernels_emu.cpp (the engine's kernel sources + warp_emu.hpp) with g++ and bind its entry points"""
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    # reconInterKernel's dynamic shared memory is `extern __shared__ ...[]`; with __shared__ standing for `static` on the host that
    # line -- and only that line -- has to read plain `extern` (the array itself is defined in kernels_emu.cpp)
    src = open(os.path.join(ROOT, "h264bsd_b200", "csrc", "engine", "recon_kernel.cuh")).read()
    line = "extern __shared__ __align__(128) uint8_t interSmemRaw[];"
    assert src.count(line) == 1
    with open(os.path.join(EMU_DIR, "_build", "recon_kernel_emu.cuh"), "w") as f:
        f.write(src.replace(line, "extern uint8_t interSmemRaw[];"))
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas",
                           "-I" + os.path.join(EMU_DIR, "stubs"), "-I" + EMU_DIR, "-I" + os.path.join(EMU_DIR, "_build"),
                           "-I" + os.path.join(ROOT, "h264bsd_b200", "csrc", "engine"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(EMU_DIR, "kernels_emu.cpp"), "-o", EMU_SO, "-lpthread"])
    L = C.CDLL(EMU_SO)
    L.emu_geom.argtypes = [C.c_uint32] * 3 + [C.POINTER(C.c_uint64)]
    L.emu_conceal.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_void_p, C.c_void_p] + [C.c_uint32] * 6
    L.emu_copy.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_void_p, C.c_void_p] + [C.c_uint32] * 5
    L.emu_deblock.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_void_p] + [C.c_uint32] * 3
    L.emu_deblock.restype = C.c_uint32
    L.emu_engine_create.argtypes = [C.c_uint32] * 4
    L.emu_engine_create.restype = C.c_void_p
    L.emu_engine_destroy.argtypes = [C.c_void_p]
    L.emu_engine_pool.argtypes = [C.c_void_p]
    L.emu_engine_pool.restype = C.POINTER(C.c_uint8)
    L.emu_engine_picture.argtypes = [C.c_void_p] * 5 + [C.c_uint32] * 6 + [C.c_int] * 2 + [C.c_uint32] * 5
    L.emu_engine_picture.restype = C.c_uint32
    return L


@pytest.fixture(scope="module")
def emu():
    return build_and_load()


def aligned_pool(nbytes, fill=128):
    """the frame pool comes from cudaMalloc (256-byte aligned); the copy kernel moves 16-byte vectors"""
    raw = np.full(nbytes + 256, fill, np.uint8)
    off = (-raw.ctypes.data) % 256
    return raw[off:off + nbytes]


def to_pool_with_border(frame, W, H, geom, slot, pool):
    """a finished frame as the engine keeps it: picture plus replicated border (borderKernel)"""
    pitchY, pitchC, rowsY, rowsC, offCb, offCr, stride, pads = geom
    padY, padC = pads & 0xFFFFFFFF, pads >> 32
    base = slot * stride
    Y = pool[base:base + pitchY * rowsY].reshape(rowsY, pitchY)
    Y[:, :W + 2 * padY] = np.pad(frame[:W * H].reshape(H, W), padY, mode="edge")
    for off, src in ((offCb, frame[W * H:W * H + W * H // 4]), (offCr, frame[W * H + W * H // 4:])):
        P = pool[base + off:base + off + pitchC * rowsC].reshape(rowsC, pitchC)
        P[:, :W // 2 + 2 * padC] = np.pad(src.reshape(H // 2, W // 2), padC, mode="edge")


N_STREAMS = 6      # two blocks of four warps, the second half empty


def to_pool(frame, W, H, geom, slot, pool):
    pitchY, pitchC, rowsY, rowsC, offCb, offCr, stride, pads = geom
    padY, padC = pads & 0xFFFFFFFF, pads >> 32
    base = slot * stride
    Y = pool[base:base + pitchY * rowsY].reshape(rowsY, pitchY)
    Y[padY:padY + H, padY:padY + W] = frame[:W * H].reshape(H, W)
    for off, src in ((offCb, frame[W * H:W * H + W * H // 4]), (offCr, frame[W * H + W * H // 4:])):
        P = pool[base + off:base + off + pitchC * rowsC].reshape(rowsC, pitchC)
        P[padC:padC + H // 2, padC:padC + W // 2] = src.reshape(H // 2, W // 2)


def from_pool(W, H, geom, slot, pool):
    pitchY, pitchC, rowsY, rowsC, offCb, offCr, stride, pads = geom
    padY, padC = pads & 0xFFFFFFFF, pads >> 32
    base = slot * stride
    Y = pool[base:base + pitchY * rowsY].reshape(rowsY, pitchY)[padY:padY + H, padY:padY + W]
    out = [Y.reshape(-1)]
    for off in (offCb, offCr):
        P = pool[base + off:base + off + pitchC * rowsC].reshape(rowsC, pitchC)[padC:padC + H // 2, padC:padC + W // 2]
        out.append(P.reshape(-1))
    return np.concatenate(out)


def test_conceal_kernel_source_matches_oracle_on_the_host(emu):
    checked = mbs = 0
    for seed in range(0, 240):
        ps = ParsedStream(synth_h264.make_damaged_stream(seed), resilient=True)
        if ps.status != 0 or not any(p.numConceal for p in ps.pics):
            ps.close()
            continue
        W, H = ps.width_mbs * 16, ps.height_mbs * 16
        g = (C.c_uint64 * 8)()
        emu.emu_geom(ps.width_mbs, ps.height_mbs, ps.num_slots, g)
        geom = [int(v) for v in g]
        pool = np.full(geom[6] * ps.num_slots * N_STREAMS, 128, np.uint8)
        orc = _oracle.OracleDecoder(ps)
        t = ps.ptr.contents
        order = C.cast(t.mbOrder, C.c_void_p).value
        for k in range(ps.num_pics):
            h = ps.pics[k]
            if h.numConceal:
                h0 = copy.copy(h)
                h0.numConceal = 0          # everything but the spatial estimates
                orc.L.px_recon_picture(orc.ctx, C.byref(h0), orc._recs + h.mbRecOffset, orc._coefs + h.coefOffset)
                for st in range(N_STREAMS):
                    to_pool(orc.frame(h.curSlot), W, H, geom, st * ps.num_slots + h.curSlot, pool)
                n_a = h.numPassA - h.numRunMbs - h.numCopy
                emu.emu_conceal(pool.ctypes.data, ps.width_mbs, ps.height_mbs, ps.num_slots, h.curSlot, orc._recs + h.mbRecOffset,
                                order + 2 * k * ps.mbs_per_pic, h.numRun, h.numCopy, n_a, h.numPassB, h.numConceal, N_STREAMS)
                orc.recon(k)
                for st in range(N_STREAMS):
                    assert np.array_equal(from_pool(W, H, geom, st * ps.num_slots + h.curSlot, pool), orc.frame(h.curSlot)), \
                        f"seed {seed}: picture {k}, stream {st}: concealKernel (emulated) differs from the oracle"
                checked += 1
                mbs += h.numConceal
            else:
                orc.recon(k)
            orc.deblock(k)
        orc.close()
        ps.close()
        if checked >= 40:
            break
    assert checked >= 20 and mbs >= 100, (checked, mbs)


def copy_list_mbs(order, k, h, nmb):
    """macroblock addresses in the run and single-copy sections of picture k's processing order"""
    o = np.ctypeslib.as_array(order, shape=((k + 1) * nmb,))[k * nmb:]
    mbs = []
    for i in range(h.numRun):
        mbs += list(range(int(o[2 * i]), int(o[2 * i]) + int(o[2 * i + 1])))
    mbs += [int(a) for a in o[2 * h.numRun:2 * h.numRun + h.numCopy]]
    return mbs


def mb_pixels(frame, W, H, mb):
    wm = W // 16
    x, y = (mb % wm) * 16, (mb // wm) * 16
    Y = frame[:W * H].reshape(H, W)[y:y + 16, x:x + 16]
    C2 = frame[W * H:].reshape(2, H // 2, W // 2)[:, y // 2:y // 2 + 8, x // 2:x // 2 + 8]
    return np.concatenate([Y.reshape(-1), C2.reshape(-1)])


BULK = 0x80000000   # copyRuns bit that selects the B200_COPY_BULK=1 launch sequence in the emulated engine
DEEP = 0x40000000   # ... reconCopyKernelDeep (B200_COPY_VARIANT=2; emu_copy only)


@pytest.mark.parametrize("kind,variant", [("still", "lanes"), ("damaged", "lanes"), ("still", "bulk"), ("damaged", "bulk"), ("still", "deep")])
def test_copy_kernel_source_matches_oracle_on_the_host(emu, kind, variant):
    """zero-motion runs, single integer-vector copies (vectors far outside the picture included) and -- in the damaged
    streams -- concealed macroblocks copied from the reference picture: every listed macroblock against the oracle, three
    streams on a two-block grid, runs per task as the engine's default and at its maximum.  variant "bulk": the experimental
    reconCopyBulkKernel (cp.async.bulk through shared memory, B200_COPY_BULK=1) moves the runs -- the emulation checks the
    16-byte alignment of every bulk copy and defers the stores until the wait that releases their staging buffer.  variant
    "deep": reconCopyKernelDeep (B200_COPY_VARIANT=2), the same body with the loads of four steps before the first store"""
    if kind == "still":
        streams = [synth_h264.make_stream(s, still=True, W=w, H=hh, pictures=3) for s, w, hh in ((3, 11, 4), (4, 40, 3), (6, 7, 6), (9, 37, 2))]
        streams += [synth_h264.make_stream(s) for s in range(0, 24 if variant == "lanes" else 8)]   # (emulation is slow)
        resilient = False
    else:
        streams = [synth_h264.make_damaged_stream(s) for s in range(0, 80 if variant == "lanes" else 24)]   # (emulation is slow)
        resilient = True
    n_streams = 3
    pics = mbs_checked = concealed_copies = 0
    for data in streams:
        ps = ParsedStream(data, resilient=resilient)
        if ps.status != 0 or ps.num_pics == 0:
            ps.close()
            continue
        W, H, nmb = ps.width_mbs * 16, ps.height_mbs * 16, ps.mbs_per_pic
        g = (C.c_uint64 * 8)()
        emu.emu_geom(ps.width_mbs, ps.height_mbs, ps.num_slots, g)
        geom = [int(v) for v in g]
        orc = _oracle.OracleDecoder(ps)
        t = ps.ptr.contents
        order = C.cast(t.mbOrder, C.c_void_p).value
        rec = np.dtype([("mbType", "u1"), ("pad0", "u1", 2), ("flags", "u1"), ("rest", "u1", 92)])
        recs = np.frombuffer(C.string_at(t.mbRecs, t.mbRecBytes), rec)
        for k in range(ps.num_pics):
            h = ps.pics[k]
            if h.numRun + h.numCopy:
                # the frame slots as they are when picture k is reconstructed: finished pictures, borders replicated
                pool = aligned_pool(geom[6] * ps.num_slots * n_streams)
                for st in range(n_streams):
                    for slot in range(ps.num_slots):
                        to_pool_with_border(orc.frame(slot), W, H, geom, st * ps.num_slots + slot, pool)
                copy_runs = (4 if (pics & 1) == 0 else 16) | {"lanes": 0, "bulk": BULK, "deep": DEEP}[variant]
                emu.emu_copy(pool.ctypes.data, ps.width_mbs, ps.height_mbs, ps.num_slots, h.curSlot, orc._recs + h.mbRecOffset,
                             order + 2 * k * nmb, h.numRun, h.numCopy, n_streams, copy_runs, 2)
                orc.recon(k)
                want = orc.frame(h.curSlot)
                listed = copy_list_mbs(t.mbOrder, k, h, nmb)
                for st in range(n_streams):
                    got = from_pool(W, H, geom, st * ps.num_slots + h.curSlot, pool)
                    for mb in listed:
                        assert np.array_equal(mb_pixels(got, W, H, mb), mb_pixels(want, W, H, mb)), \
                            f"picture {k}, stream {st}, macroblock {mb}: reconCopyKernel (emulated) differs from the oracle"
                pics += 1
                mbs_checked += len(listed)
                concealed_copies += int(sum(1 for mb in listed if recs[k * nmb + mb]["flags"] & 0x80))
            else:
                orc.recon(k)
            orc.deblock(k)
        orc.close()
        ps.close()
    assert pics >= (10 if variant == "lanes" else 6) and mbs_checked >= (300 if variant == "lanes" else 150), (pics, mbs_checked)
    if kind == "damaged":
        assert concealed_copies >= (20 if variant == "lanes" else 4), concealed_copies


@pytest.mark.parametrize("kind", ["valid", "damaged"])
def test_filter_kernels_source_match_oracle_on_the_host(emu, kind):
    """strengthKernel + deblockKernel (boundary strengths, then the ticketed wavefront filter with its flag waits) on whole
    pictures: multi-slice / FMO pictures with all three filter modes and offsets, and damaged pictures whose concealed
    macroblocks are filtered as Intra4x4 / QP 40.  Two streams, two blocks of eight concurrent warps."""
    if kind == "valid":
        streams = [(synth_h264.make_stream(s), False) for s in range(0, 40)]
    else:
        streams = [(synth_h264.make_damaged_stream(s), True) for s in range(0, 60)]
    n_streams = 2
    pics = concealed = 0
    for data, resilient in streams:
        ps = ParsedStream(data, resilient=resilient)
        if ps.status != 0 or ps.num_pics == 0 or ps.mbs_per_pic > 40:
            ps.close()
            continue
        W, H = ps.width_mbs * 16, ps.height_mbs * 16
        g = (C.c_uint64 * 8)()
        emu.emu_geom(ps.width_mbs, ps.height_mbs, ps.num_slots, g)
        geom = [int(v) for v in g]
        orc = _oracle.OracleDecoder(ps)
        for k in range(min(ps.num_pics, 4)):
            h = ps.pics[k]
            orc.recon(k)
            pool = aligned_pool(geom[6] * ps.num_slots * n_streams)
            for st in range(n_streams):
                to_pool(orc.frame(h.curSlot), W, H, geom, st * ps.num_slots + h.curSlot, pool)
            wd = emu.emu_deblock(pool.ctypes.data, ps.width_mbs, ps.height_mbs, ps.num_slots, h.curSlot,
                                 orc._recs + (h.filterRecOffset or h.mbRecOffset),     # the filter's own records where a picture has them
                                 n_streams, 8 if (pics & 1) else 3, 2)
            orc.deblock(k)
            assert wd == 0, "a flag wait ran into the watchdog"
            for st in range(n_streams):
                assert np.array_equal(from_pool(W, H, geom, st * ps.num_slots + h.curSlot, pool), orc.frame(h.curSlot)), \
                    f"picture {k}, stream {st}: strengthKernel + deblockKernel (emulated) differ from the oracle"
            pics += 1
            concealed += h.numErrMbs > 0
        orc.close()
        ps.close()
        if pics >= 60:
            break
    assert pics >= 30, pics
    if kind == "damaged":
        assert concealed >= 5, concealed


def run_engine_on_the_host(emu, ps, n_streams=2, knobs=(8, 1, 4, 8), blocks=2, max_pics=None, stages=False):
    """replay a tape through the emulated engine (every kernel of Batch::launchPicture, pool kept across pictures) next to the
    oracle; returns the number of pictures compared"""
    W, H, nmb = ps.width_mbs * 16, ps.height_mbs * 16, ps.mbs_per_pic
    g = (C.c_uint64 * 8)()
    emu.emu_geom(ps.width_mbs, ps.height_mbs, ps.num_slots, g)
    geom = [int(v) for v in g]
    eng = emu.emu_engine_create(ps.width_mbs, ps.height_mbs, ps.num_slots, n_streams)
    pool = np.ctypeslib.as_array(emu.emu_engine_pool(eng), shape=(geom[6] * ps.num_slots * n_streams,))
    orc = _oracle.OracleDecoder(ps)
    t = ps.ptr.contents
    order = C.cast(t.mbOrder, C.c_void_p).value
    n = ps.num_pics if max_pics is None else min(ps.num_pics, max_pics)
    try:
        for k in range(n):
            h = ps.pics[k]
            args = (eng, orc._recs + h.mbRecOffset, (orc._recs + h.filterRecOffset) if h.filterRecOffset else None,
                    orc._coefs + h.coefOffset, order + 2 * k * nmb, h.curSlot,
                    h.numRun, h.numCopy, h.numPassA - h.numRunMbs - h.numCopy, h.numPassB, h.numConceal)
            if stages:
                assert emu.emu_engine_picture(*args, 1, 0, *knobs, blocks) == 0
                orc.recon(k)
                for st in range(n_streams):
                    assert np.array_equal(from_pool(W, H, geom, st * ps.num_slots + h.curSlot, pool), orc.frame(h.curSlot)), \
                        f"picture {k}, stream {st}: reconstruction (emulated kernels) differs from the oracle"
                assert emu.emu_engine_picture(*args, 0, 1, *knobs, blocks) == 0
                orc.deblock(k)
            else:
                assert emu.emu_engine_picture(*args, 1, 1, *knobs, blocks) == 0, "watchdog or IDCT range error"
                orc.recon(k)
                orc.deblock(k)
            for st in range(n_streams):
                assert np.array_equal(from_pool(W, H, geom, st * ps.num_slots + h.curSlot, pool), orc.frame(h.curSlot)), \
                    f"picture {k}, stream {st}: the emulated engine differs from the oracle"
    finally:
        emu.emu_engine_destroy(eng)
        orc.close()
    return n


@pytest.mark.parametrize("kind", ["valid", "damaged", "large"])
def test_whole_engine_source_matches_oracle_on_the_host(emu, kind):
    """every kernel of the per-picture launch sequence -- copy pass, TMA-staged inter pass, ticketed intra pass, concealment,
    boundary strengths, wavefront filter, border -- as shipped, on the host, pictures chained through the frame pool the way
    the engine chains them: synthetic streams (all macroblock types, several reference frames, vectors far outside the
    picture), damaged streams (concealment), and the larger still-scene streams (long copy runs)"""
    if kind == "valid":
        jobs = [(synth_h264.make_stream(s), False, None, True) for s in range(0, 30)]
    elif kind == "damaged":
        jobs = [(synth_h264.make_damaged_stream(s), True, None, False) for s in range(0, 40)]
    else:
        jobs = [(synth_h264.make_stream(2, W=45, H=18, still=True, pictures=4), False, 3, False),
                (synth_h264.make_stream(21, W=64, H=4, still=True, pictures=5), False, 4, False)]
    pics = 0
    for i, (data, resilient, max_pics, stages) in enumerate(jobs):
        ps = ParsedStream(data, resilient=resilient)
        if ps.status == 0 and ps.num_pics and (kind == "large" or ps.mbs_per_pic <= 40):
            knobs = (8, 1, 4, 8) if i % 2 == 0 else (3, 2, 16, 2)       # chunkA, chunkB, copyRuns, filterChunk
            if i % 3 == 2:
                knobs = (knobs[0], knobs[1], knobs[2] | BULK, knobs[3])  # runs by reconCopyBulkKernel
            pics += run_engine_on_the_host(emu, ps, knobs=knobs, blocks=2 + i % 2, max_pics=max_pics or 6, stages=stages)
        ps.close()
    assert pics >= (5 if kind == "large" else 40), pics


def test_whole_engine_source_on_the_reference_stream(emu):
    """the first pictures of the reference's own test_640x360.h264 (an IDR picture and P pictures of an encoder: long zero-motion
    runs, every interpolation position) through the emulated engine, three streams on four blocks"""
    ps = ParsedStream(_oracle.stream_bytes("test_640x360.h264"))
    try:
        assert run_engine_on_the_host(emu, ps, n_streams=3, knobs=(8, 1, 4, 8), blocks=4, max_pics=6) == 6
    finally:
        ps.close()
