"""CPU suite, part 2: host logic and the C-ABI surface (no GPU, no compute calls into the engine)."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import _oracle
from h264bsd_b200 import _lib
from h264bsd_b200.batch import ParsedStream

ROOT = _oracle.ROOT


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(h264bsd[A-Za-z0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = _declared("h264bsd_decoder.h") + _declared("h264bsd_b200.h")
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert set(_lib.LEGACY_SYMBOLS) <= set(declared)
    assert set(_lib.BATCH_SYMBOLS) <= set(declared)


def test_storage_abi_size():
    # callers allocate storage_t themselves: 4648 bytes on LP64 (reference src/h264bsd_storage.h:75-152)
    hdr = open(os.path.join(ROOT, "include", "h264bsd_storage.h")).read()
    assert "#define H264BSD_STORAGE_BYTES 4648" in hdr
    assert _lib.STORAGE_BYTES == 4648


def test_init_fails_loudly_without_gpu():
    L = _lib.load()
    if L.h264bsdB200DeviceCount() > 0:
        pytest.skip("a GPU is visible")
    st = (C.c_uint8 * 4648)()
    assert L.h264bsdInit(C.addressof(st), 0) == 1  # HANTRO_NOK: no CPU pixel path to fall back to
    assert not L.h264bsdB200BatchCreate(0, 1, 4, 4, 2)


def test_record_layout():
    rec = np.dtype([('mbType', 'u1'), ('qpY', 'u1'), ('qpC', 'u1'), ('flags', 'u1'), ('codedMask', '<u4'), ('coefIndex', '<u4'),
                    ('fa', 'i1'), ('fb', 'i1'), ('cqo', 'i1'), ('sub', 'u1'), ('refSlot', 'u1', 4), ('icm', 'u1'), ('idc', 'u1'),
                    ('sliceId', '<u2'), ('refIdx', 'u1', 4), ('r1', 'u1', 4), ('mv', '<i2', (16, 2))])
    assert rec.itemsize == 96
    assert C.sizeof(_lib.PicHdr) == 176


def test_parse_360p_statistics():
    """stream facts of SURVEY.md section 6: 73 pictures = 2 IDR + 71 P, one slice per picture, 4 frame slots"""
    ps = ParsedStream(_oracle.stream_bytes("test_640x360.h264"))
    assert ps.status == 0 and ps.num_pics == 73 and ps.num_slots == 4
    assert [p.isIdr for p in ps.pics].count(1) == 2 and ps.pics[0].isIdr and ps.pics[40].isIdr
    assert all(p.isRef for p in ps.pics) and all(p.numErrMbs == 0 for p in ps.pics)
    t = ps.ptr.contents
    recs = np.ctypeslib.as_array(t.mbRecs, shape=(t.mbRecBytes,)).reshape(-1, 96)
    types = recs[:, 0]
    nmb = ps.mbs_per_pic
    assert (types[:nmb] >= 6).all()            # IDR: intra only
    assert (types[nmb:2 * nmb] <= 5).sum() > 0  # P picture has inter macroblocks
    assert (recs[:, 21] == 0).all()            # disable_deblocking_filter_idc 0 everywhere
    # coefficient pool accounting: one 32-byte block per coded block / DC block, 12 per I_PCM
    masks = recs[:, 4:8].copy().view('<u4')[:, 0]
    pop = np.array([bin(int(m) & 0x3FFFFFF).count("1") for m in masks[types != 31]])
    assert pop.sum() * 32 == t.coefBytes


def test_parse_truncated_stream_reports_error_or_partial():
    data = _oracle.stream_bytes("test_640x360.h264")
    ps = ParsedStream(data[:len(data) // 3])
    assert ps.num_pics < 73 and ps.num_pics > 5


def test_parse_garbage_does_not_crash():
    rng = np.random.default_rng(7)
    junk = bytes(rng.integers(0, 256, 20000, dtype=np.uint8))
    ps = ParsedStream(b"\x00\x00\x00\x01" + junk)
    assert ps.num_pics == 0
    # a valid stream with corrupted slice payload bytes: must terminate (error or concealed pictures), never hang
    data = bytearray(_oracle.stream_bytes("test_640x360.h264"))
    for i in range(5000, len(data), 997):
        data[i] ^= 0x5A
    ps = ParsedStream(bytes(data))
    assert ps.num_pics <= 80
