"""CPU suite, part 2: host logic and the C-ABI surface (no GPU, no compute calls into the engine)."""
import ctypes as C
import os
import re
import subprocess
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
    assert C.sizeof(_lib.PicHdr) == 200


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


# ---- CAVLC tables against the reference's own look-up arrays (needs the reference sources: build container only) ----
REF_CAVLC = "/root/reference/src/h264bsd_cavlc.c"
REF_VLC = "/root/reference/src/h264bsd_vlc.c"


def _ref_arrays(path):
    txt = open(path, encoding="latin1").read()
    out = {}
    for m in re.finditer(r"static const u(?:8|16|32) (\w+)\[\d*\]\s*=\s*\{([^}]*)\}", txt):
        out[m.group(1)] = [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(2))]
    return out


@pytest.mark.skipif(not os.path.exists(REF_CAVLC), reason="reference sources not mounted")
def test_cavlc_tables_match_reference():
    """every 16-bit prefix decodes to the same (length, TrailingOnes, TotalCoeff) / total_zeros / run_before as the
    reference's DecodeCoeffToken / DecodeTotalZeros / DecodeRunBefore (h264bsd_cavlc.c:400-700)"""
    A = _ref_arrays(REF_CAVLC)
    L = _lib.load()
    L.b200_cavlc_probe.restype = C.c_uint32
    L.b200_cavlc_probe.argtypes = [C.c_int, C.c_int, C.c_uint32]

    def ref_token(bits, nc):
        if nc < 0:
            v = A["coeffTokenMinus1_0"][bits >> 13] or A["coeffTokenMinus1_1"][bits >> 8]
        elif nc < 2:
            if bits >= 0x8000: v = 0x0001
            elif bits >= 0x0C00: v = A["coeffToken0_0"][bits >> 10]
            elif bits >= 0x0100: v = A["coeffToken0_1"][bits >> 6]
            elif bits >= 0x0020: v = A["coeffToken0_2"][(bits >> 2) - 8]
            else: v = A["coeffToken0_3"][bits]
        elif nc < 4:
            if bits >= 0x8000: v = 0x0002 if bits & 0x4000 else 0x0822
            elif bits >= 0x1000: v = A["coeffToken2_0"][bits >> 10]
            elif bits >= 0x0200: v = A["coeffToken2_1"][bits >> 7]
            else: v = A["coeffToken2_2"][bits >> 2]
        elif nc < 8:
            v = A["coeffToken4_0"][bits >> 10] or A["coeffToken4_1"][bits >> 6]
        else:
            v = A["coeffToken8"][bits >> 10]
        return (v & 0x1F, (v >> 5) & 0x3F, (v >> 11) & 0x1F) if v else None

    for nc in (-1, 0, 2, 4, 8):
        for bits in range(0, 1 << 16, 1 if nc in (0, 2) else 4):
            mine = L.b200_cavlc_probe(0, nc, bits)
            got = (mine & 31, (mine >> 5) & 3, mine >> 7) if mine & 31 else None
            assert got == ref_token(bits, nc), (nc, hex(bits))

    def ref_tz(bits9, tc):
        t = {1: None, 2: ("totalZeros_2", 3), 3: ("totalZeros_3", 3), 4: ("totalZeros_4", 4), 5: ("totalZeros_5", 4),
             6: ("totalZeros_6", 3), 7: ("totalZeros_7", 3), 8: ("totalZeros_8", 3), 9: ("totalZeros_9", 3),
             10: ("totalZeros_10", 4), 11: ("totalZeros_11", 5), 12: ("totalZeros_12", 5), 13: ("totalZeros_13", 6),
             14: ("totalZeros_14", 7)}
        if tc == 1:
            v = A["totalZeros_1_0"][bits9 >> 4] or A["totalZeros_1_1"][bits9]
        elif tc == 15:
            v = 0x11 if bits9 >> 8 else 0x01
        else:
            v = A[t[tc][0]][bits9 >> t[tc][1]]
        return (v & 0xF, v >> 4) if v else None

    for tc in range(1, 16):
        for b9 in range(512):
            mine = L.b200_cavlc_probe(1, tc, b9 << 7)
            got = (mine & 15, mine >> 4) if mine & 15 else None
            assert got == ref_tz(b9, tc), (tc, b9)

    def ref_tz_dc(bits9, tc):
        b = bits9 >> 6
        if b > 3: v = 0x01
        elif tc == 3: v = 0x11
        elif b > 1: v = 0x12
        elif tc == 2: v = 0x22
        elif b: v = 0x23
        else: v = 0x33
        return (v & 0xF, v >> 4)

    for tc in range(1, 4):
        for b9 in range(512):
            mine = L.b200_cavlc_probe(2, tc, b9 << 7)
            assert (mine & 15, mine >> 4) == ref_tz_dc(b9, tc), (tc, b9)

    def ref_run(bits11, zl):
        if zl <= 6:
            name, sh = {1: ("runBefore_1", 10), 2: ("runBefore_2", 9), 3: ("runBefore_3", 9), 4: ("runBefore_4", 8),
                        5: ("runBefore_5", 8), 6: ("runBefore_6", 8)}[zl]
            v = A[name][bits11 >> sh]
        else:
            if bits11 >= 0x100: v = ((7 - (bits11 >> 8)) << 4) + 3
            else:
                v = 0
                for k, thr in enumerate((0x80, 0x40, 0x20, 0x10, 0x8, 0x4, 0x2, 0x1)):
                    if bits11 >= thr:
                        v = ((7 + k) << 4) | (4 + k)
                        break
            if (v >> 4) > zl: v = 0
        return (v & 0xF, v >> 4) if v else None

    for zl in range(1, 15):
        for b11 in range(2048):
            mine = L.b200_cavlc_probe(3, zl, b11 << 5)
            got = (mine & 15, mine >> 4) if mine & 15 else None
            if got is not None and got[1] > zl:
                got = None   # the decoder rejects run_before > zerosLeft, as the reference's INFO(value) > zerosLeft test does
            assert got == ref_run(b11, zl), (zl, b11)


@pytest.mark.skipif(not os.path.exists(REF_VLC), reason="reference sources not mounted")
def test_coded_block_pattern_mapping_matches_reference():
    A = _ref_arrays(REF_VLC)
    L = _lib.load()
    L.b200_cbp_probe.restype = C.c_uint32
    L.b200_cbp_probe.argtypes = [C.c_uint32, C.c_int]
    intra = [v for k, v in A.items() if "ntra" in k and len(v) == 48]
    inter = [v for k, v in A.items() if "nter" in k and len(v) == 48]
    assert intra and inter
    for code in range(48):
        assert L.b200_cbp_probe(code, 1) == intra[0][code]
        assert L.b200_cbp_probe(code, 0) == inter[0][code]


# ---- multi-GPU host logic on CPU (gloo, world size 2) ----------------------------------------------------------------
def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = bench.shard_streams(1000, world, rank)
    # what bench.py reduces across ranks: slowest rank's time, summed work
    t = torch.tensor([10.0 + rank]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(count)], dtype=torch.float64); dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, first, count, float(t.item()), float(n.item())))
    dist.destroy_process_group()


def test_shard_streams_gloo():
    import bench
    # static contiguous shards cover every stream exactly once
    for total, world in ((4096, 8), (1000, 3), (5, 8)):
        seen = []
        for r in range(world):
            first, count = bench.shard_streams(total, world, r)
            seen += list(range(first, first + count))
        assert seen == list(range(total))
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(60)
    assert [r[1:3] for r in res] == [(0, 500), (500, 500)]
    assert all(r[3] == 11.0 and r[4] == 1000.0 for r in res)


def _build_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "decode_file")
    libdir = os.path.join(root, "h264bsd_b200")
    subprocess.check_call(["gcc", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "decode_file.c"), "-L" + libdir, "-lh264bsd_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_example_links_against_public_headers(tmp_path):
    """a C program written like posix/test_h264bsd.c:130-179 compiles against include/ alone and links against the
    shared library (the drop-in boundary); without a GPU it must fail loudly, not decode on the CPU"""
    import subprocess
    exe = _build_example(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, os.path.join(_oracle.GOLDEN, "test_640x360.h264")], capture_output=True, text=True)
        assert r.returncode != 0 and "no CUDA device" in r.stderr


def test_reparse_many_matches_single_parse():
    """h264bsdB200ReparseStreams: n streams on native threads give the tapes a single parse gives (and tapes are re-usable)"""
    data = _oracle.stream_bytes("test_640x360.h264")
    bits = (C.c_uint8 * len(data)).from_buffer_copy(data)
    ref = ParsedStream(data)
    many = [ParsedStream() for _ in range(5)]
    for _ in range(2):
        ParsedStream.reparse_many(many, bits, 3)
        for p in many:
            assert p.status == 0 and p.num_pics == ref.num_pics and p.rec_bytes == ref.rec_bytes and p.coef_bytes == ref.coef_bytes
            a = np.ctypeslib.as_array(C.cast(p.ptr.contents.mbRecs, C.POINTER(C.c_uint8)), shape=(ref.rec_bytes,))
            b = np.ctypeslib.as_array(C.cast(ref.ptr.contents.mbRecs, C.POINTER(C.c_uint8)), shape=(ref.rec_bytes,))
            assert (a == b).all()
    for p in many:
        p.close()
    ref.close()


@pytest.mark.parametrize("name", ["test_640x360.h264", "test_1920x1080.h264"])
def test_picture_headers_count_the_passes(name):
    """the records lie in raster order and the GPU sorts them itself (pass A: inter / I_PCM, pass B: intra-predicted); the
    picture header's counts say what it will find, and an intra macroblock only waits for neighbours that are intra-predicted
    themselves"""
    ps = ParsedStream(_oracle.stream_bytes(name))
    t = ps.ptr.contents
    nmb = ps.width_mbs * ps.height_mbs
    W = ps.width_mbs
    recs = np.ctypeslib.as_array(C.cast(t.mbRecs, C.POINTER(C.c_uint8)), shape=(ps.num_pics * nmb, 96))
    assert t.numOrder == 0 and all(h.numConceal == 0 for h in ps.pics), "nothing is concealed in a valid stream"
    for k, h in enumerate(ps.pics):
        r = recs[k * nmb:(k + 1) * nmb]
        types = r[:, 0]
        mask = np.ascontiguousarray(r[:, 4:8]).view("<u4")[:, 0]
        mv0 = np.ascontiguousarray(r[:, 32:36]).view("<i2")
        intra = (types > 5) & (types != 31)
        copy = (types <= 1) & (mask == 0) & (mv0[:, 0] == 0) & (mv0[:, 1] == 0)
        assert h.numPassB == int(intra.sum()) and h.numPassA == nmb - h.numPassB and h.numCopy == int(copy.sum())
        wm = r[:, 28]
        assert (wm[~intra] == 0).all()
        g = intra.reshape(-1, W)
        left = np.pad(g, ((0, 0), (1, 0)))[:, :-1]
        up = np.pad(g, ((1, 0), (0, 0)))[:-1]
        upright = np.pad(g, ((1, 0), (0, 1)))[:-1, 1:]
        upleft = np.pad(g, ((1, 0), (1, 0)))[:-1, :-1]
        flags = r[:, 3].reshape(-1, W)
        want = ((flags & 1) * left) | (((flags >> 1) & 1) * up << 1) | (((flags >> 2) & 1) * upright << 2) | (((flags >> 3) & 1) * upleft << 3)
        assert (wm.reshape(-1, W)[g] == want[g]).all()
    ps.close()


POSIX_B200 = os.path.join(ROOT, "oracle", "_ref", "test_h264bsd_b200")


@pytest.mark.skipif(not os.path.exists(POSIX_B200), reason="oracle/_ref/test_h264bsd_b200 not built (make -C oracle ref, needs the reference sources)")
def test_reference_posix_front_end_links_unchanged_and_fails_loudly_without_gpu():
    """the reference's own posix/test_h264bsd.c, unmodified, compiled against include/ and linked with the shared library
    (oracle/Makefile: posix).  Without a GPU it must say so and stop -- no silent CPU path."""
    r = subprocess.run([POSIX_B200, os.path.join(ROOT, "tests", "golden", "test_640x360.h264")], capture_output=True, text=True)
    if _lib.load().h264bsdB200DeviceCount() <= 0:
        assert r.returncode != 0 and "no CUDA device" in r.stderr
    else:
        assert "73 pictures decoded" in r.stdout


def test_tape_reuse_across_different_streams():
    """a tape re-used for another stream (other size, other slot count, damaged or not) holds exactly what a fresh parse gives:
    the records are built in place in memory nobody cleared, so every byte of a record must be written"""
    import hashlib
    import synth_h264

    def signature(ps):
        t = ps.ptr.contents
        return (ps.status, ps.num_pics, ps.width_mbs, ps.height_mbs, ps.num_slots, tuple(ps.outputs),
                hashlib.md5(C.string_at(t.mbRecs, t.mbRecBytes)).hexdigest(), hashlib.md5(C.string_at(t.coefs, t.coefBytes)).hexdigest(),
                hashlib.md5(C.string_at(t.mbOrder, t.numOrder * 2)).hexdigest() if t.numOrder else None)

    reused = ParsedStream(synth_h264.make_stream(0))
    for seed in range(1, 60):
        damaged = seed % 3 == 0
        data = synth_h264.make_damaged_stream(seed) if damaged else synth_h264.make_stream(seed)
        fresh = ParsedStream(data, resilient=damaged)
        reused.reparse(data, resilient=damaged)
        assert signature(fresh) == signature(reused), f"seed {seed}"
        fresh.close()
    reused.close()
