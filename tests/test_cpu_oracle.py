"""CPU suite, part 1: the oracle is pinned against the reference's golden vectors.

tests/golden/md5.json was produced by the UNMODIFIED reference decoder (oracle/make_golden.py via
oracle/_ref).  Here the host syntax decoder (product code) parses each stream into a tape and the CPU
oracle (oracle/px_oracle.c) replays it; the result must reproduce, byte for byte, both the frames the
reference hands to its in-loop filter and its output frames."""
import ctypes as C
import hashlib
import json
import os
import numpy as np
import pytest
import _oracle
from h264bsd_b200.batch import ParsedStream

GOLD = json.load(open(os.path.join(_oracle.GOLDEN, "md5.json")))
STREAMS = list(GOLD.keys())


@pytest.mark.parametrize("name", STREAMS)
def test_input_fixture_is_the_reference_stream(name):
    assert hashlib.md5(_oracle.stream_bytes(name)).hexdigest() == GOLD[name]["input_md5"]


@pytest.mark.parametrize("name", STREAMS)
def test_oracle_reproduces_reference_frames(name):
    g = GOLD[name]
    ps = ParsedStream(_oracle.stream_bytes(name))
    assert ps.status == 0
    assert (ps.num_pics, ps.width_mbs, ps.height_mbs) == (g["pictures"], g["width_mbs"], g["height_mbs"])
    assert ps.outputs == list(range(g["pictures"]))  # POC type 2: no reordering (storage.c:363-368)
    post, pre, errs = _oracle.oracle_run_tape(ps, want_pre=True)
    assert errs == 0
    fb = ps.frame_bytes
    for k in range(ps.num_pics):
        assert hashlib.md5(pre[k * fb:(k + 1) * fb].tobytes()).hexdigest() == g["pre_frame_md5"][k], f"pre-deblock picture {k}"
        assert hashlib.md5(post[k * fb:(k + 1) * fb].tobytes()).hexdigest() == g["post_frame_md5"][k], f"output picture {k}"
    assert hashlib.md5(post.tobytes()).hexdigest() == g["post_md5"]
    assert hashlib.md5(pre.tobytes()).hexdigest() == g["pre_md5"]


@pytest.mark.parametrize("name", STREAMS[:1] + STREAMS[2:])
def test_oracle_colour_conversion(name):
    """h264bsdConvertToRGBA then BGRA of the first two pictures (decoder.c:1163-1298)."""
    g = GOLD[name]
    ps = ParsedStream(_oracle.stream_bytes(name))
    post, _, _ = _oracle.oracle_run_tape(ps)
    fb, W, H = ps.frame_bytes, ps.width_mbs * 16, ps.height_mbs * 16
    h = hashlib.md5()
    for k in range(2):
        for mode in (0, 1):
            h.update(_oracle.oracle_convert(mode, W, H, post[k * fb:(k + 1) * fb]).tobytes())
    assert h.hexdigest() == g["rgba_bgra_first2_md5"]


@pytest.mark.skipif(_oracle.reference() is None, reason="oracle/_ref not built (needs the reference sources)")
def test_compiled_reference_still_matches_golden():
    ref = _oracle.reference()
    name = STREAMS[0]
    data = _oracle.stream_bytes(name)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    info = (C.c_uint32 * 8)()
    cap = GOLD[name]["pictures"] * GOLD[name]["width_mbs"] * GOLD[name]["height_mbs"] * 384
    post = np.zeros(cap, np.uint8)
    n = ref.ref_decode_stream(buf, len(data), post.ctypes.data, cap, None, 0, None, 0, info)
    assert n == GOLD[name]["pictures"]
    assert hashlib.md5(post.tobytes()).hexdigest() == GOLD[name]["post_md5"]
    assert ref.ref_sizeof_storage() == 4648  # the ABI size include/h264bsd_storage.h preserves
