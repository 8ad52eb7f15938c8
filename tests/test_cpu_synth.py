"""Host syntax decoder + CPU oracle against the reference decoder on synthetic streams that walk through the
Baseline syntax the three encoder-made fixtures never touch (tests/synth_h264.py: FMO, ASO, several slices and
reference frames, list reordering, I_PCM, every partition shape and intra mode, vectors far outside the picture,
level escape codes, long-term IDR, frame_num wrap, POC types 0-2 ...).  No GPU: this pins the *record* half of
record-then-replay and the oracle the GPU suite is checked against."""
import hashlib
import json
import os
import numpy as np
import pytest
import _oracle
import synth_h264
from h264bsd_b200.batch import ParsedStream

GOLD = json.load(open(os.path.join(_oracle.GOLDEN, "synth_md5.json")))
SEEDS = sorted(int(s) for s in GOLD if not s.startswith("L"))
LARGE = sorted(s for s in GOLD if s.startswith("L"))


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def decode_with_oracle(data):
    ps = ParsedStream(data)
    try:
        assert ps.status == 0, f"host parser reports status {ps.status} after {ps.num_pics} pictures"
        post, pre, errs = _oracle.oracle_run_tape(ps, want_pre=True)
        assert errs == 0
        return len(ps.outputs), ps.num_pics, (ps.width_mbs, ps.height_mbs), post, pre
    finally:
        ps.close()


def stream_info(data):
    """the tape's copy of what h264bsdCroppingParams / h264bsdVideoRange / h264bsdMatrixCoefficients return"""
    ps = ParsedStream(data)
    t = ps.ptr.contents
    if t.cropFlag:
        crop = [1, t.cropLeft, t.cropWidth, t.cropTop, t.cropHeight]
    else:
        crop = [0, 0, 0, 0, 0]
    info = {"crop": crop, "video_range": int(t.videoRange), "matrix_coefficients": int(t.matrixCoefficients)}
    ps.close()
    return info


@pytest.mark.parametrize("chunk", range(8))
def test_synthetic_streams_match_reference_golden(chunk):
    """committed md5s of the reference decoder's output (tests/make_synth_golden.py); needs no reference build"""
    for seed in SEEDS[chunk::8]:
        g = GOLD[str(seed)]
        data = synth_h264.make_stream(seed)
        if hashlib.md5(data).hexdigest() != g["stream_md5"]:
            pytest.skip("tests/synth_h264.py (or Python's random) no longer writes the streams the golden file was made from: "
                        "re-run tests/make_synth_golden.py")
        n_out, n_dec, dims, post, pre = decode_with_oracle(data)
        assert dims == (g["width_mbs"], g["height_mbs"]), f"seed {seed}"
        assert stream_info(data) == g["info"], f"seed {seed}: cropping / video range / matrix coefficients"
        assert (n_out, n_dec) == (g["outputs"], g["decoded"]), f"seed {seed}: picture counts"
        assert md5(pre) == g["pre_md5"], f"seed {seed}: pictures before the in-loop filter differ from the reference"
        assert md5(post) == g["post_md5"], f"seed {seed}: output pictures differ from the reference"


def test_large_still_streams_match_reference_golden():
    """pictures up to 120 macroblocks wide, mostly zero-vector copies"""
    runs = 0
    for key in LARGE:
        g = GOLD[key]
        data = synth_h264.make_stream(g["seed"], **g["knobs"])
        if hashlib.md5(data).hexdigest() != g["stream_md5"]:
            pytest.skip("generator drifted from the golden file: re-run tests/make_synth_golden.py")
        ps = ParsedStream(data)
        runs += sum(p.numCopy for p in ps.pics)
        ps.close()
        n_out, n_dec, dims, post, pre = decode_with_oracle(data)
        assert dims == (g["width_mbs"], g["height_mbs"]) and (n_out, n_dec) == (g["outputs"], g["decoded"]), key
        assert md5(pre) == g["pre_md5"] and md5(post) == g["post_md5"], f"{key}: pictures differ from the reference"
    assert runs > 5000


def test_golden_streams_cover_the_syntax():
    """the point of the synthetic set is coverage; fail if the generator stops producing it"""
    import ctypes as C
    rec = np.dtype([("mbType", "u1"), ("pad0", "u1", 14), ("sub", "u1"), ("refSlot", "u1", 4), ("icm", "u1"), ("idc", "u1"),
                    ("sliceId", "<u2"), ("refIdx", "u1", 4), ("pad1", "u1", 4), ("mv", "<i2", (16, 2))])
    assert rec.itemsize == 96
    types, subs, modes, idcs = set(), set(), set(), set()
    multi_ref = multi_slice = far_mv = reordered = 0
    for seed in SEEDS[:60]:
        ps = ParsedStream(synth_h264.make_stream(seed))
        t = ps.ptr.contents
        n = ps.mbs_per_pic
        area = C.string_at(t.mbRecs, t.mbRecBytes)      # (a picture's records may be followed by records for the filter alone)
        a = np.concatenate([np.frombuffer(area, rec, count=n, offset=p.mbRecOffset) for p in ps.pics])
        types |= set(np.unique(a["mbType"]).tolist())
        p8 = a[(a["mbType"] == 4) | (a["mbType"] == 5)]
        for q in range(4):
            subs |= set(np.unique((p8["sub"] >> (2 * q)) & 3).tolist())
        i4 = a[a["mbType"] == 6]
        modes |= set(np.unique(i4["mv"].view("u1").reshape(len(i4), 64)[:, :16]).tolist())
        idcs |= set(np.unique(a["idc"]).tolist())
        inter = a[a["mbType"] <= 5]
        multi_ref += int((inter["refIdx"] > 0).any(axis=1).sum())
        far_mv += int((np.abs(inter["mv"][:, :, 0]) > 1000).any(axis=1).sum())
        multi_slice += sum(1 for k in range(ps.num_pics) if len(np.unique(a["sliceId"][k * n:(k + 1) * n])) > 1)
        reordered += int(ps.outputs != sorted(ps.outputs))
        ps.close()
    assert types == set(range(32)), sorted(set(range(32)) - types)
    assert subs == {0, 1, 2, 3} and modes == set(range(9)) and idcs == {0, 1, 2}
    assert multi_ref > 50 and multi_slice > 50 and far_mv > 50 and reordered > 0


@pytest.mark.skipif(_oracle.reference() is None, reason="oracle/_ref not built (needs the reference sources)")
def test_fresh_seeds_against_the_compiled_reference():
    """seeds outside the golden file, decoded by the reference here and now"""
    from make_synth_golden import reference_decode
    for seed in range(10_000, 10_120):
        data = synth_h264.make_stream(seed)
        n, fb, rpost, rpre, ndec, dims = reference_decode(data)
        assert n >= 0, f"seed {seed}: reference decode error (generator bug)"
        n_out, n_dec, odims, post, pre = decode_with_oracle(data)
        assert (n_out, n_dec, odims) == (n, ndec, dims), f"seed {seed}"
        assert np.array_equal(pre, rpre), f"seed {seed}: pre-filter pictures"
        assert np.array_equal(post, rpost), f"seed {seed}: output pictures"


def test_sequence_change_same_size_and_other_size():
    """two coded video sequences back to back: with the same picture size the tape simply goes on (and matches the reference,
    new IDR, new parameter sets); a change of picture size ends the tape (B200_TAPE_SIZE_CHANGE) with the first sequence intact"""
    a, b = synth_h264.make_stream(3, W=4, H=3), synth_h264.make_stream(7, W=4, H=3)
    n_a = decode_with_oracle(a)[1]
    n_out, n_dec, dims, post, pre = decode_with_oracle(a + b)
    assert dims == (4, 3) and n_dec == n_a + decode_with_oracle(b)[1]
    if _oracle.reference() is not None:
        from make_synth_golden import reference_decode
        n, fb, rpost, rpre, ndec, rdims = reference_decode(a + b)
        assert (n, ndec) == (n_out, n_dec) and np.array_equal(post, rpost) and np.array_equal(pre, rpre)
    ps = ParsedStream(a + synth_h264.make_stream(7, W=6, H=2))
    assert ps.status == 100 and (ps.width_mbs, ps.height_mbs) == (4, 3) and ps.num_pics == n_a
    ps.close()


@pytest.mark.skipif(_oracle.reference() is None, reason="oracle/_ref not built (needs the reference sources)")
def test_concatenated_sequences_against_the_compiled_reference():
    """two or three coded video sequences of one picture size back to back: parameter sets replaced under the same ids, a new
    IDR with other POC / frame_num / DPB parameters, pictures of the old sequence still waiting for output"""
    import random
    from make_synth_golden import reference_decode
    r = random.Random(5)
    for i in range(40):
        W, H = r.choice([1, 2, 3, 4, 6]), r.choice([1, 2, 3, 5])
        seeds = [r.randrange(10 ** 6) for _ in range(r.choice([2, 2, 3]))]
        data = b"".join(synth_h264.make_stream(s, W=W, H=H) for s in seeds)
        n, fb, rpost, rpre, ndec, dims = reference_decode(data)
        assert n >= 0, f"{seeds}: reference decode error"
        n_out, n_dec, odims, post, pre = decode_with_oracle(data)
        assert (n_out, n_dec, odims) == (n, ndec, dims), f"{seeds}"
        assert np.array_equal(pre, rpre) and np.array_equal(post, rpost), f"{seeds}: pictures differ from the reference"


@pytest.mark.skipif(_oracle.reference() is None, reason="oracle/_ref not built (needs the reference sources)")
def test_no_output_reordering_mode_against_the_compiled_reference():
    """h264bsdInit(storage, noOutputReordering = 1): pictures leave in decoding order, the DPB keeps reference frames only"""
    import ctypes as C
    from make_synth_golden import reference_decode
    L = _oracle.reference()
    L.ref_set_no_reordering.argtypes = [C.c_int]
    for seed in range(40_000, 40_080):
        data = synth_h264.make_stream(seed)
        L.ref_set_no_reordering(1)
        try:
            n, fb, rpost, rpre, ndec, dims = reference_decode(data)
        finally:
            L.ref_set_no_reordering(0)
        assert n >= 0, f"seed {seed}: reference decode error"
        ps = ParsedStream(data, no_output_reordering=True)
        try:
            assert ps.status == 0 and (len(ps.outputs), ps.num_pics) == (n, ndec), f"seed {seed}"
            post, pre, _ = _oracle.oracle_run_tape(ps, want_pre=True)
        finally:
            ps.close()
        assert np.array_equal(pre, rpre) and np.array_equal(post, rpost), f"seed {seed}: pictures differ from the reference"
