"""The shipped source of the spatial concealment kernel (h264bsd_b200/csrc/engine/conceal_kernel.cuh) compiled for the host
and run lane by lane (tests/emu/warp_emu.hpp: 32 threads per warp, a barrier per shuffle) against the CPU oracle, on the
pictures of the damaged streams that need it.  Checks the kernel's indexing and arithmetic where no GPU is at hand; the GPU
suite checks the same pictures on the hardware."""
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
EMU_SO = os.path.join(EMU_DIR, "_build", "libconceal_emu.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-I" + EMU_DIR,
                           "-I" + os.path.join(ROOT, "h264bsd_b200", "csrc", "engine"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(EMU_DIR, "conceal_emu.cpp"), "-o", EMU_SO, "-lpthread"])
    L = C.CDLL(EMU_SO)
    L.emu_geom.argtypes = [C.c_uint32] * 3 + [C.POINTER(C.c_uint64)]
    L.emu_conceal.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_void_p, C.c_void_p] + [C.c_uint32] * 6
    return L


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
