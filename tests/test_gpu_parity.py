"""GPU suite: the CUDA engine, called through the C-ABI, against the golden vectors of the reference and
against the CPU oracle.  Bit-exact or fail."""
import ctypes as C
import hashlib
import json
import os
import numpy as np
import pytest
import _oracle
from h264bsd_b200 import _lib
from h264bsd_b200.batch import Batch, ParsedStream
from h264bsd_b200.decoder import H264bsdDecoder, decode_stream, PIC_RDY, HDRS_RDY

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(_oracle.GOLDEN, "md5.json")))
STREAMS = list(GOLD.keys())


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def parsed():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = ParsedStream(_oracle.stream_bytes(name))
        return cache[name]
    return get


def test_gpu_present_and_native_library_loaded():
    L = _lib.load()
    assert L.h264bsdB200DeviceCount() >= 1
    assert os.path.basename(_lib.LIB_PATH) == "libh264bsd_b200.so"


@pytest.mark.parametrize("name", STREAMS)
def test_legacy_api_bit_exact(name):
    """config 2 of BASELINE.json: the stream through h264bsdInit/Decode/NextOutputPicture/Shutdown, every frame
    at full coded size compared with the reference decoder's (posix/test_h264bsd.c:146-177 loop)"""
    g = GOLD[name]
    frames = decode_stream(_oracle.stream_bytes(name))
    assert len(frames) == g["pictures"]
    for k, f in enumerate(frames):
        assert md5(f) == g["post_frame_md5"][k], f"{name}: output picture {k} differs from the reference"
    h = hashlib.md5()
    for f in frames:
        h.update(f.tobytes())
    assert h.hexdigest() == g["post_md5"]


def test_legacy_api_getters_and_headers_ready():
    d = H264bsdDecoder()
    d.queueInput(_oracle.stream_bytes("test_640x360.h264"))
    seen_hdrs = 0
    pics = 0
    while d.inputBytesRemaining() > 0:
        r = d.decode()
        if r == HDRS_RDY:
            seen_hdrs += 1
            assert (d.outputPictureWidth(), d.outputPictureHeight()) == (640, 368)
            assert d.croppingParams() == {"left": 0, "width": 640, "top": 0, "height": 360}
        elif r == PIC_RDY:
            assert d.nextOutputPicture() is not None
            assert d.nextOutputPicture() is None  # exactly one picture per PIC_RDY when nothing is reordered
            pics += 1
        assert r in (0, 1, 2)
    assert seen_hdrs == 1 and pics == 73
    assert d.videoRange() == 0
    d.release()
    d2 = H264bsdDecoder()
    d2.queueInput(_oracle.stream_bytes("test_1920x1080_fullRange.h264")[:200000])
    while d2.inputBytesRemaining() > 0 and d2.decode() != PIC_RDY:
        pass
    assert d2.videoRange() == 1
    d2.release()


@pytest.mark.parametrize("name,pics", [("test_640x360.h264", list(range(0, 10)) + [39, 40, 41, 72]),
                                       ("test_1920x1080.h264", list(range(0, 42)))])
def test_stages_in_isolation_vs_oracle(parsed, name, pics):
    """reconstruction and in-loop filter each checked alone: the GPU gets the oracle's frames as input state (1080p: every
    picture of the first IDR period and the IDR picture that follows)"""
    ps = parsed(name)
    orc = _oracle.OracleDecoder(ps)
    b = Batch(1, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    for k in range(max(pics) + 1):
        h = ps.pics[k]
        if k in pics:
            for s in range(ps.num_slots):
                b.write_frame(0, s, orc.frame(s))
        orc.recon(k)
        if k in pics:
            pre = orc.frame(h.curSlot).copy()
            b.debug_stage(k, True, False)
            assert np.array_equal(b.read_frame(0, h.curSlot), pre), f"{name}: reconstruction of picture {k}"
            assert md5(pre) == GOLD[name]["pre_frame_md5"][k]
        orc.deblock(k)
        if k in pics:
            b.write_frame(0, h.curSlot, pre)
            b.debug_stage(k, False, True)
            assert np.array_equal(b.read_frame(0, h.curSlot), orc.frame(h.curSlot)), f"{name}: deblocking of picture {k}"
    assert b.idct_errors() == 0 and b.watchdog() == (0, 0)
    b.close()


@pytest.mark.parametrize("name,n_streams", [("test_640x360.h264", 16), ("test_1920x1080.h264", 8)])
def test_batch_of_replicated_streams(parsed, name, n_streams):
    """many independent instances of one stream, each with its own HBM work-list and frame slots: every picture
    of the first instance matches the reference, every other instance matches the first"""
    ps = parsed(name)
    g = GOLD[name]
    b = Batch(n_streams, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    b.replicate(0)
    for k in range(ps.num_pics):
        b.decode_picture(k)
        slot = ps.pics[k].curSlot
        if k % 6 == 0 or k == ps.num_pics - 1:
            assert md5(b.read_frame(0, slot)) == g["post_frame_md5"][k], f"picture {k}"
            assert md5(b.read_frame(n_streams - 1, slot)) == g["post_frame_md5"][k], f"picture {k}, last instance"
            assert b.compare_streams([slot] * n_streams) == 0
    assert b.idct_errors() == 0 and b.watchdog() == (0, 0)
    b.close()


def test_looped_stream_is_idempotent(parsed):
    """looping the tape (what the throughput bench does) leaves every slot as after the first pass"""
    ps = parsed("test_640x360.h264")
    b = Batch(4, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    b.replicate(0)
    b.run(0, ps.num_pics)
    first = [b.read_frame(1, s).copy() for s in range(ps.num_slots)]
    for _ in range(2):
        b.run(0, ps.num_pics)
    for s in range(ps.num_slots):
        assert np.array_equal(b.read_frame(2, s), first[s])
    assert md5(first[ps.pics[-1].curSlot]) == GOLD["test_640x360.h264"]["post_frame_md5"][-1]
    b.close()


@pytest.mark.parametrize("name", [STREAMS[0], STREAMS[2]])
def test_colour_conversion_kernel(parsed, name):
    """config 5: YUV -> 32-bit pixels, against h264bsdConvertToRGBA/BGRA of the reference (golden md5) and the oracle"""
    ps = parsed(name)
    g = GOLD[name]
    W, H = ps.width_mbs * 16, ps.height_mbs * 16
    b = Batch(1, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    h = hashlib.md5()
    for k in range(2):
        b.decode_picture(k)
        yuv = b.read_frame(0, ps.pics[k].curSlot)
        for mode in (0, 1):
            px = b.convert_frame(0, ps.pics[k].curSlot, mode)
            assert np.array_equal(px, _oracle.oracle_convert(mode, W, H, yuv))
            h.update(px.tobytes())
        assert np.array_equal(b.convert_frame(0, ps.pics[k].curSlot, 2), _oracle.oracle_convert(2, W, H, yuv))
    assert h.hexdigest() == g["rgba_bgra_first2_md5"]
    b.close()
    # the same through the preserved entry points
    L = _lib.load()
    out = np.empty(W * H, np.uint32)
    yuv = np.ascontiguousarray(yuv)
    for mode, fn in ((0, L.h264bsdConvertToRGBA), (1, L.h264bsdConvertToBGRA), (2, L.h264bsdConvertToYCbCrA)):
        fn(W, H, yuv.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, _oracle.oracle_convert(mode, W, H, yuv))
    d = H264bsdDecoder()
    d.queueInput(_oracle.stream_bytes(name))
    while d.decode() != PIC_RDY:
        pass
    rgba = d.nextOutputPictureRGBA()
    d.release()
    frames0 = decode_stream(_oracle.stream_bytes(name)[:400000])[0] if name == STREAMS[0] else None
    if frames0 is not None:
        assert np.array_equal(rgba, _oracle.oracle_convert(0, W, H, frames0))


def test_example_decoder_matches_oracle(tmp_path):
    """the C driver of examples/ (the posix/test_h264bsd.c loop) linked against the shared library writes the same
    I420 file as the reference decoder"""
    import subprocess
    from test_cpu_host import _build_example
    exe = _build_example(tmp_path)
    name = "test_640x360.h264"
    out = str(tmp_path / "out.yuv")
    r = subprocess.run([exe, "-o", out, os.path.join(_oracle.GOLDEN, name)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"{GOLD[name]['pictures']} pictures decoded" in r.stdout
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == GOLD[name]["post_md5"]


def test_streamed_upload_matches_golden(parsed):
    """work-lists uploaded in groups of pictures on the copy stream while earlier pictures decode (the end-to-end path of
    the bench): same pictures as the reference, for every stream, twice in a row (device arrays re-used)"""
    name = "test_640x360.h264"
    ps = parsed(name)
    g = GOLD[name]
    n = 6
    b = Batch(n, ps.width_mbs, ps.height_mbs, ps.num_slots)
    for rep in range(2):
        for g0 in range(0, ps.num_pics, 4):
            gn = min(4, ps.num_pics - g0)
            for s in range(n):
                b.upload_range(s, ps, g0, gn)
            b.upload_fence(g0 + gn)
        for k in range(ps.num_pics):
            b.decode_picture(k)
            if k % 7 == 0 or k == ps.num_pics - 1:
                slot = ps.pics[k].curSlot
                assert md5(b.read_frame(0, slot)) == g["post_frame_md5"][k], f"pass {rep} picture {k}"
                assert b.compare_streams([slot] * n) == 0
        b.sync()
    assert b.idct_errors() == 0 and b.watchdog() == (0, 0)
    b.close()


@pytest.mark.parametrize("name,n_streams", [("test_640x360.h264", 6), ("test_1920x1080.h264", 4)])
def test_end_to_end_path_every_stream_every_picture(name, n_streams):
    """the data path of the bench's end-to-end leg -- bitstream bytes parsed and uploaded by the host threads
    (h264bsdB200BatchParseUploadBegin / Wait), decode_picture, read_picture_all into page-locked memory -- two passes over two
    alternating batches: every picture of every stream against the reference's md5"""
    L = _lib.load()
    data = _oracle.stream_bytes(name)
    g = GOLD[name]
    bits = (C.c_uint8 * len(data)).from_buffer_copy(data)
    ps = ParsedStream(data)
    fb, npics = ps.frame_bytes, ps.num_pics
    batches = [Batch(n_streams, ps.width_mbs, ps.height_mbs, ps.num_slots) for _ in range(2)]
    pool = L.h264bsdB200ParseUploadPoolCreate(3)
    host = L.h264bsdB200HostAlloc(fb * n_streams)
    assert pool and host
    view = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_uint8)), shape=(fb * n_streams,))
    tok = batches[0].parse_upload_begin(pool, [bits] * n_streams)
    for i in range(3):
        cur = batches[i & 1]
        cur.parse_upload_wait(tok)
        if i + 1 < 3:
            tok = batches[(i + 1) & 1].parse_upload_begin(pool, [bits] * n_streams)   # overlaps this pass's GPU work
        for k in range(npics):
            cur.decode_picture(k)
            cur.read_picture_all(k, host, fb)
            cur.sync()
            for s in range(n_streams):
                assert md5(view[s * fb:(s + 1) * fb]) == g["post_frame_md5"][k], f"{name}: pass {i}, stream {s}, picture {k}"
        assert cur.idct_errors() == 0 and cur.watchdog() == (0, 0)
    L.h264bsdB200ParseUploadPoolDestroy(pool)
    L.h264bsdB200HostFree(host)
    for b in batches:
        b.close()
    ps.close()


def test_64_stream_1080p_batch_every_picture(parsed):
    """the throughput configuration in small: 64 instances of the 1080p stream, EVERY picture: stream 0 against the reference's
    md5 and every other stream against stream 0 (device-side compare)"""
    ps = parsed("test_1920x1080.h264")
    g = GOLD["test_1920x1080.h264"]
    n = 64
    b = Batch(n, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    b.replicate(0)
    for k in range(ps.num_pics):
        b.decode_picture(k)
        slot = ps.pics[k].curSlot
        assert md5(b.read_frame(0, slot)) == g["post_frame_md5"][k], f"picture {k}"
        assert b.compare_streams([slot] * n) == 0, f"picture {k}: the instances differ"
    assert md5(b.read_frame(n - 1, ps.pics[-1].curSlot)) == g["post_frame_md5"][-1]
    assert b.idct_errors() == 0 and b.watchdog() == (0, 0)
    b.close()


def test_cropped_and_nv12_output(parsed):
    """the output kernel: picture k of every stream de-stripped, cropped to the display rectangle of
    h264bsdCroppingParams and delivered as I420 or NV12 in one transfer -- against the coded-size picture cropped on the host"""
    L = _lib.load()
    name = "test_640x360.h264"
    ps = parsed(name)
    t = ps.ptr.contents
    assert t.cropFlag and (t.cropWidth, t.cropHeight) == (640, 360)
    W, H = ps.width_mbs * 16, ps.height_mbs * 16
    n = 3
    b = Batch(n, ps.width_mbs, ps.height_mbs, ps.num_slots)
    b.upload(0, ps)
    b.replicate(0)
    for crop in ((t.cropLeft, t.cropTop, t.cropWidth, t.cropHeight), (16, 2, 600, 300), (34, 20, 90, 66)):
        x0, y0, cw, ch = crop
        ob = cw * ch * 3 // 2
        host = L.h264bsdB200HostAlloc(ob * n)
        view = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_uint8)), shape=(ob * n,))
        for k in range(3):
            b.decode_picture(k) if crop[2] == t.cropWidth else None
            full = b.read_frame(1, ps.pics[k].curSlot)
            Y = full[:W * H].reshape(H, W)[y0:y0 + ch, x0:x0 + cw]
            Cb = full[W * H:W * H * 5 // 4].reshape(H // 2, W // 2)[y0 // 2:(y0 + ch) // 2, x0 // 2:(x0 + cw) // 2]
            Cr = full[W * H * 5 // 4:].reshape(H // 2, W // 2)[y0 // 2:(y0 + ch) // 2, x0 // 2:(x0 + cw) // 2]
            for nv12 in (False, True):
                view[:] = 0xEE
                b.read_picture_all_ex(k, host, ob, crop, nv12)
                b.sync()
                chroma = np.stack([Cb, Cr], axis=2).reshape(-1) if nv12 else np.concatenate([Cb.reshape(-1), Cr.reshape(-1)])
                want = np.concatenate([Y.reshape(-1), chroma])
                for s in range(n):
                    assert np.array_equal(view[s * ob:(s + 1) * ob], want), f"crop {crop}, nv12 {nv12}, picture {k}, stream {s}"
        L.h264bsdB200HostFree(host)
    b.close()


def test_legacy_api_sequence_parameter_set_change():
    """a stream that activates another sequence parameter set (another picture size) in the middle: h264bsdDecode re-creates
    the engine on H264BSD_HDRS_RDY like the reference re-allocates (h264bsd_decoder.c:343-389, h264bsd_storage.c:297-420) --
    same, smaller and larger size, the RGBA output path included"""
    small = _oracle.stream_bytes("test_640x360.h264")
    big = _oracle.stream_bytes("test_1920x1080.h264")
    big = big[:big.rindex(b"\x00\x00\x00\x01", 0, 330000)]      # the first pictures, cut at a NAL unit boundary
    for first, second in ((small, small), (small, big), (big, small)):
        want = decode_stream(first) + decode_stream(second)
        d = H264bsdDecoder()
        got = []
        for part in (first, second):
            d.queueInput(part)
            while d.inputBytesRemaining() > 0:
                if d.decode() == PIC_RDY:
                    while (f := d.nextOutputPicture()) is not None:
                        got.append(f.copy())
            d.flush()
            while (f := d.nextOutputPicture()) is not None:
                got.append(f.copy())
        assert len(got) == len(want)
        for k, (a, w) in enumerate(zip(got, want)):
            assert a.shape == w.shape and np.array_equal(a, w), f"picture {k}"
        # the converted output after the switch
        d.queueInput(first)
        while d.decode() != PIC_RDY:
            pass
        rgba = d.nextOutputPictureRGBA()
        ref = decode_stream(first)[0]
        Wd, Hd = d.outputPictureWidth(), d.outputPictureHeight()
        assert np.array_equal(rgba, _oracle.oracle_convert(0, Wd, Hd, ref))
        d.release()


POSIX_B200 = os.path.join(_oracle.ROOT, "oracle", "_ref", "test_h264bsd_b200")


@pytest.mark.skipif(not os.path.exists(POSIX_B200), reason="oracle/_ref/test_h264bsd_b200 not built (make -C oracle ref, needs the reference sources)")
def test_reference_posix_front_end_unchanged(tmp_path):
    """BASELINE.json configs[0], literally: the reference's own posix/test_h264bsd.c -- compiled unmodified against include/ and
    linked with libh264bsd_b200.so (oracle/Makefile: posix) -- decodes test_640x360.h264, prints what the reference build
    prints, and its "-o" dump has the md5 of the reference build's dump"""
    import subprocess
    gold = json.load(open(os.path.join(_oracle.GOLDEN, "posix_front_end.json")))["test_640x360.h264"]
    out = tmp_path / "dump.yuv"
    r = subprocess.run([POSIX_B200, "-o", str(out), os.path.join(_oracle.GOLDEN, "test_640x360.h264")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().splitlines()[-1] == gold["stdout_tail"]
    data = out.read_bytes()
    assert len(data) == gold["o_dump_bytes"] and hashlib.md5(data).hexdigest() == gold["o_dump_md5"]
