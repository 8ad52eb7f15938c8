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


def test_still_scenes_copies_match_oracle():
    """still scenes: whole columns of zero-motion copies (pass A moves them as contiguous bursts of the strip layout) next to a
    few coded macroblocks, odd and even picture widths, three instances, every picture against the oracle"""
    checked = 0
    for seed, w, hh in ((3, 11, 4), (4, 40, 3), (6, 7, 6), (9, 37, 2), (11, 33, 5), (12, 3, 35)):
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
            checked += ps.pics[k].numCopy > 0
        assert b.watchdog() == (0, 0)
        b.close(); orc.close(); ps.close()
    assert checked >= 8, checked
