"""GPU suite, part two: the CUDA engine on the synthetic streams of tests/synth_h264.py (every macroblock type and
partition shape, several reference frames, vectors far outside the picture, I_PCM, FMO / ASO slice orders, pictures
down to one macroblock), through the C-ABI, against the committed md5s of the reference decoder and -- picture by
picture, before and after the in-loop filter -- against the CPU oracle.  Bit-exact or fail."""
import hashlib
import json
import os
import numpy as np
import pytest
import _oracle
import synth_h264
from h264bsd_b200.batch import Batch, ParsedStream
from h264bsd_b200.decoder import decode_stream

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(_oracle.GOLDEN, "synth_md5.json")))
SEEDS = sorted(int(s) for s in GOLD if not s.startswith("L"))
LARGE = sorted(s for s in GOLD if s.startswith("L"))


def _stream(seed):
    data = synth_h264.make_stream(seed)
    if hashlib.md5(data).hexdigest() != GOLD[str(seed)]["stream_md5"]:
        pytest.skip("generator drifted from tests/golden/synth_md5.json: re-run tests/make_synth_golden.py")
    return data


@pytest.mark.parametrize("chunk", range(8))
def test_batched_engine_matches_oracle_picture_by_picture(chunk):
    """two instances of every stream; each picture after reconstruction and after the filter equals the oracle's"""
    for seed in SEEDS[chunk::8]:
        ps = ParsedStream(_stream(seed))
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
            assert np.array_equal(b.read_frame(1, slot), orc.frame(slot)), f"seed {seed}: in-loop filter of picture {k}"
            assert b.compare_streams([slot, slot]) == 0, f"seed {seed}: the two instances differ at picture {k}"
        assert b.idct_errors() == 0 and b.watchdog() == (0, 0), f"seed {seed}"
        b.close()
        orc.close()
        ps.close()


@pytest.mark.parametrize("chunk", range(4))
def test_legacy_api_matches_reference_golden(chunk):
    """h264bsdInit/Decode/NextOutputPicture over the synthetic streams: output pictures, output order"""
    for seed in SEEDS[chunk::4]:
        g = GOLD[str(seed)]
        frames = decode_stream(_stream(seed))
        assert len(frames) == g["outputs"], f"seed {seed}"
        h = hashlib.md5()
        for f in frames:
            h.update(np.ascontiguousarray(f).tobytes())
        assert h.hexdigest() == g["post_md5"], f"seed {seed}: output pictures differ from the reference"


def test_large_still_streams_match_oracle_and_golden():
    """rows wider than a copy run, runs cut by slice / slice-group borders, several reference slots: four instances per stream"""
    for key in LARGE:
        g = GOLD[key]
        data = synth_h264.make_stream(g["seed"], **g["knobs"])
        if hashlib.md5(data).hexdigest() != g["stream_md5"]:
            pytest.skip("generator drifted from tests/golden/synth_md5.json: re-run tests/make_synth_golden.py")
        ps = ParsedStream(data)
        orc = _oracle.OracleDecoder(ps)
        b = Batch(4, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        b.replicate(0)
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            b.decode_picture(k)
            orc.recon(k)
            orc.deblock(k)
            assert np.array_equal(b.read_frame(3, slot), orc.frame(slot)), f"{key}: picture {k}"
            assert b.compare_streams([slot] * 4) == 0, f"{key}: instances differ at picture {k}"
        assert b.idct_errors() == 0 and b.watchdog() == (0, 0), key
        b.close()
        orc.close()
        ps.close()
        frames = decode_stream(data)
        h = hashlib.md5()
        for f in frames:
            h.update(np.ascontiguousarray(f).tobytes())
        assert len(frames) == g["outputs"] and h.hexdigest() == g["post_md5"], f"{key}: legacy API output differs from the reference"


DAMAGED = json.load(open(os.path.join(_oracle.GOLDEN, "synth_damaged_md5.json")))


@pytest.mark.parametrize("chunk", range(4))
def test_legacy_api_matches_reference_golden(chunk):
    """h264bsdInit/Decode/NextOutputPicture over the synthetic streams: output pictures, output order"""
    for seed in SEEDS[chunk::4]:
        g = GOLD[str(seed)]
        frames = decode_stream(_stream(seed))
        assert len(frames) == g["outputs"], f"seed {seed}"
        h = hashlib.md5()
        for f in frames:
            h.update(np.ascontiguousarray(f).tobytes())
        assert h.hexdigest() == g["post_md5"], f"seed {seed}: output pictures differ from the reference"


def test_large_still_streams_match_oracle_and_golden():
    """rows wider than a copy run, runs cut by slice / slice-group borders, several reference slots: four instances per stream"""
    for key in LARGE:
        g = GOLD[key]
        data = synth_h264.make_stream(g["seed"], **g["knobs"])
        if hashlib.md5(data).hexdigest() != g["stream_md5"]:
            pytest.skip("generator drifted from tests/golden/synth_md5.json: re-run tests/make_synth_golden.py")
        ps = ParsedStream(data)
        orc = _oracle.OracleDecoder(ps)
        b = Batch(4, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        b.replicate(0)
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            b.decode_picture(k)
            orc.recon(k)
            orc.deblock(k)
            assert np.array_equal(b.read_frame(3, slot), orc.frame(slot)), f"{key}: picture {k}"
            assert b.compare_streams([slot] * 4) == 0, f"{key}: instances differ at picture {k}"
        assert b.idct_errors() == 0 and b.watchdog() == (0, 0), key
        b.close()
        orc.close()
        ps.close()
        frames = decode_stream(data)
        h = hashlib.md5()
        for f in frames:
            h.update(np.ascontiguousarray(f).tobytes())
        assert len(frames) == g["outputs"] and h.hexdigest() == g["post_md5"], f"{key}: legacy API output differs from the reference"


DAMAGED = json.load(open(os.path.join(_oracle.GOLDEN, "synth_damaged_md5.json")))


@pytest.mark.xfail(strict=False, reason="error-concealment path (concealKernel, concealed-copy records) was written after round 1's GPU "
                                        "budget was spent: checked on the host by emulation (tests/test_cpu_kernel_emu.py), not yet run on "
                                        "hardware -- the first run decides (DESIGN.md, known gaps)")
@pytest.mark.parametrize("chunk", range(4))
def test_damaged_streams_concealment_matches_oracle(chunk):
    """damaged streams in resilient mode: lost macroblocks copied from the reference picture or estimated from their
    neighbours (concealKernel), then filtered as intra / QP 40 -- every picture against the CPU oracle, the output against the
    reference decoder's md5"""
    seeds = sorted(int(s) for s in DAMAGED)[chunk::4]
    concealed = 0
    for seed in seeds:
        g = DAMAGED[str(seed)]
        data = synth_h264.make_damaged_stream(seed)
        if hashlib.md5(data).hexdigest() != g["stream_md5"]:
            pytest.skip("generator drifted from tests/golden/synth_damaged_md5.json")
        ps = ParsedStream(data, resilient=True)
        if ps.status != 0 or ps.num_pics == 0:
            ps.close()
            continue
        orc = _oracle.OracleDecoder(ps)
        b = Batch(1, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            b.debug_stage(k, True, False)
            orc.recon(k)
            assert np.array_equal(b.read_frame(0, slot), orc.frame(slot)), f"seed {seed}: reconstruction / concealment of picture {k}"
            b.debug_stage(k, False, True)
            orc.deblock(k)
            assert np.array_equal(b.read_frame(0, slot), orc.frame(slot)), f"seed {seed}: in-loop filter of picture {k}"
            concealed += ps.pics[k].numErrMbs > 0
        assert b.watchdog() == (0, 0), f"seed {seed}"
        b.close()
        orc.close()
        ps.close()
    assert concealed > 0


def _bulk_copy_body():
    """runs in a process of its own (see the test below): B200_COPY_BULK=1 is read when a Batch is created"""
    md5s = json.load(open(os.path.join(_oracle.GOLDEN, "md5.json")))
    checked = 0
    # (1) still scenes with long runs that start at odd and even columns, three instances, every picture against the oracle
    for seed, w, hh in ((3, 11, 4), (4, 40, 3), (6, 7, 6), (9, 37, 2), (11, 33, 5)):
        ps = ParsedStream(synth_h264.make_stream(seed, still=True, W=w, H=hh, pictures=4))
        assert ps.status == 0
        orc = _oracle.OracleDecoder(ps)
        b = Batch(3, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        b.replicate(0)
        for k in range(ps.num_pics):
            slot = ps.pics[k].curSlot
            b.decode_picture(k)
            orc.recon(k)
            orc.deblock(k)
            for st in range(3):
                assert np.array_equal(b.read_frame(st, slot), orc.frame(slot)), f"still seed {seed}, picture {k}, instance {st}"
            checked += ps.pics[k].numRun > 0
        assert b.watchdog() == (0, 0)
        b.close(); orc.close(); ps.close()
    assert checked >= 8, checked
    # (2) a real stream, four instances, every picture against the reference's md5; one launch more per picture that has both
    # runs and single copies is how the variant shows it was the one that ran
    ps = ParsedStream(_oracle.stream_bytes("test_640x360.h264"))
    g = md5s["test_640x360.h264"]
    both = sum(1 for k in range(ps.num_pics) if ps.pics[k].numRun and ps.pics[k].numCopy)
    counts = {}
    for bulk, variant in (("0", "0"), ("1", "0"), ("0", "1"), ("0", "2")):   # B200_COPY_VARIANT: reconCopyKernelOcc4 / ...Deep
        os.environ["B200_COPY_BULK"] = bulk
        os.environ["B200_COPY_VARIANT"] = variant
        b = Batch(4, ps.width_mbs, ps.height_mbs, ps.num_slots)
        b.upload(0, ps)
        b.replicate(0)
        for k in range(ps.num_pics):
            b.decode_picture(k)
            slot = ps.pics[k].curSlot
            assert hashlib.md5(b.read_frame(3, slot).tobytes()).hexdigest() == g["post_frame_md5"][k], f"bulk={bulk}, variant={variant}, picture {k}"
        assert b.idct_errors() == 0 and b.watchdog() == (0, 0)
        counts[bulk + variant] = b.launches()
        b.close()
    assert counts["10"] == counts["00"] + both and both > 0, (counts, both)
    assert counts["01"] == counts["00"] and counts["02"] == counts["00"], counts
    print(f"bulk copy ok: {checked} still pictures, {ps.num_pics} pictures of test_640x360.h264, launches {counts}")


@pytest.mark.xfail(strict=False, reason="reconCopyBulkKernel (B200_COPY_BULK=1) and the two B200_COPY_VARIANT kernels, all off by default, were written after round 1's GPU budget was spent: "
                                        "checked on the host by emulation (tests/test_cpu_kernel_emu.py), never run on hardware -- the first run "
                                        "decides; it runs in a process of its own so that a fault cannot take the suite's CUDA context with it")
def test_experimental_copy_variants_match_oracle():
    """the copy pass with the zero-motion runs moved by cp.async.bulk (copy_bulk_kernel.cuh) instead of through registers, and
    the two A/B variants of reconCopyKernel (four CTAs per SM; four steps in flight)"""
    import subprocess
    import sys
    env = dict(os.environ, B200_COPY_BULK="1")
    r = subprocess.run([sys.executable, "-c", "import conftest, test_gpu_synth as t; t._bulk_copy_body()"], cwd=os.path.dirname(os.path.abspath(__file__)),
                       env=env, capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:], r.stderr[-4000:])
    assert r.returncode == 0, r.stderr[-2000:]
