"""TEST INFRASTRUCTURE: the engine's real host code -- Batch in engine.cu (pool, tensor maps, job tables, the per-picture launch
sequence, staging, read-back) and the C-ABI in api.cpp -- built for the host: tests/emu/stubs_rt/cuda_runtime.h is a CUDA runtime
whose device memory is host memory and whose kernel launches run the kernels' sources under tests/emu/warp_emu.hpp.  The library
that comes out exports the product's C-ABI, so the Python bindings drive it like the real one (B200_LIB, in a process of its
own) and the same checks as the GPU suite run against the CPU oracle and the reference's md5s -- on small synthetic streams, the
emulation is slow.  What it cannot show: anything about asynchrony (every call completes before
it returns) or the hardware.  The product never loads this library."""
import os
import re
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
BUILD = os.path.join(EMU_DIR, "_build")
SO = os.path.join(BUILD, "libh264bsd_b200_hostemu.so")
CSRC = os.path.join(ROOT, "h264bsd_b200", "csrc")


def split_top(s):
    """split at commas that are not inside parentheses"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def prepare_engine_source():
    """engine.cu with every kernel<<<grid, block, smem, stream>>>(args); spelled EMU_LAUNCH(kernel, grid, block, args); and the
    reconstruction kernels' header with pass A's dynamic shared memory declared as a plain array"""
    src = open(os.path.join(CSRC, "engine", "engine.cu")).read()
    launch = re.compile(r"([A-Za-z_]\w*(?:<\w+>)?)<<<(.*?)>>>\((.*)\);")
    n = 0
    lines = []
    for line in src.split("\n"):
        m = launch.search(line)
        if m:
            cfg = split_top(m.group(2))
            assert len(cfg) == 4, line
            line = line[:m.start()] + f"EMU_LAUNCH({m.group(1)}, {cfg[0]}, {cfg[1]}, {m.group(3)});" + line[m.end():]
            n += 1
        lines.append(line)
    assert n == src.count("<<<") and n >= 12, n
    src = "\n".join(lines)
    assert src.count('#include "recon_kernel.cuh"') == 1
    src = src.replace('#include "recon_kernel.cuh"', '#include "recon_kernel_emu.cuh"')
    src += "\nnamespace b200 {\nalignas(128) uint8_t interSmemRaw[(sizeof(PassAWarpSmem) > sizeof(MultiWarpSmem) ? sizeof(PassAWarpSmem) : sizeof(MultiWarpSmem)) * kPassAWarps];\n}\n"
    with open(os.path.join(BUILD, "engine_hostemu.cpp"), "w") as f:
        f.write(src)


@pytest.fixture(scope="module")
def hostemu_lib():
    os.makedirs(BUILD, exist_ok=True)
    rk = open(os.path.join(CSRC, "engine", "recon_kernel.cuh")).read()
    line = "extern __shared__ __align__(128) uint8_t interSmemRaw[];"
    assert rk.count(line) == 2   # (the two pass-A kernels)
    with open(os.path.join(BUILD, "recon_kernel_emu.cuh"), "w") as f:
        f.write(rk.replace(line, "extern uint8_t interSmemRaw[];"))
    prepare_engine_source()
    host = [os.path.join(CSRC, "host", n) for n in ("cavlc.cpp", "params.cpp", "dpb.cpp", "picture.cpp", "stream_decoder.cpp", "tape_builder.cpp")]
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function",
                           "-I" + os.path.join(EMU_DIR, "stubs_rt"), "-I" + EMU_DIR, "-I" + BUILD,
                           "-I" + os.path.join(CSRC, "engine"), "-I" + os.path.join(CSRC, "host"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(BUILD, "engine_hostemu.cpp"), os.path.join(CSRC, "api", "api.cpp"), *host, "-o", SO, "-lpthread"])
    return SO


def run_body(lib, body, timeout=900):
    env = dict(os.environ, B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", f"import conftest, _hostemu_bodies as t; t.{body}"], cwd=os.path.join(ROOT, "tests"), env=env,
                       capture_output=True, text=True, timeout=timeout)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout


def test_batched_api_on_the_host_matches_oracle(hostemu_lib):
    """upload + replicate + decode_picture + read_frame / read_picture_all / compare_streams of the real Batch, two streams, every
    picture of a few synthetic streams (all four copy-pass settings among them) against the oracle"""
    assert "batched ok" in run_body(hostemu_lib, "batched()")


def test_legacy_api_on_the_host_matches_reference_md5(hostemu_lib):
    """h264bsdInit / Decode / NextOutputPicture / Shutdown of api.cpp (submitHostPicture, staging, output mirror) on synthetic
    streams -- redundant slices with filter-only records and damaged streams with concealment among them -- against the md5s of
    the reference decoder"""
    assert "legacy ok" in run_body(hostemu_lib, "legacy()")


@pytest.mark.parametrize("kind", ["valid", "damaged", "large", "reference"])
def test_kernels_on_the_host_match_oracle(hostemu_lib, kind):
    """the shipped kernel sources under the warp emulation on wider ground: synthetic streams with every macroblock type (stage by
    stage), damaged streams (concealment), still scenes larger than a pass-A chunk / a filter stretch, and the first pictures of
    the reference's own test_640x360.h264 with three instances"""
    assert "kernels ok" in run_body(hostemu_lib, f"kernels('{kind}')", timeout=1500)
