import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The native library is a build artefact (git-ignored, shipped in-tree): if a checkout arrives without it, build it
    once with the repo's own script rather than fail every test with the same ImportError.  No fallback: if nvcc is not
    there either, the tests fail loudly in _lib.load()."""
    lib = os.path.join(ROOT, "h264bsd_b200", "libh264bsd_b200.so")
    if not os.path.exists(lib):
        import subprocess
        try:
            subprocess.check_call(["bash", os.path.join(ROOT, "h264bsd_b200", "build.sh")])
        except Exception as e:  # noqa: BLE001
            print(f"conftest: could not build {lib}: {e}", file=sys.stderr)
