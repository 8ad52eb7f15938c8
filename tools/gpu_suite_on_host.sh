#!/bin/bash
# TEST INFRASTRUCTURE: parts of the GPU suite (tests/test_gpu_*.py, unchanged) against the HOST build of the engine -- engine.cu and
# api.cpp compiled with tests/emu/stubs_rt/cuda_runtime.h, kernels run by tests/emu/warp_emu.hpp (see tests/test_cpu_engine_hostemu.py,
# which builds the library).  For a machine without a GPU: it checks the host code and the kernels' arithmetic on the real
# 640x360 fixture; it says nothing about the hardware.  Takes 5 - 15 minutes (256 host threads per emulated block).
#   bash tools/gpu_suite_on_host.sh ['pytest -k expression']
cd "$(dirname "$0")/.." || exit 1
python -m pytest tests/test_cpu_engine_hostemu.py -q -k batched || exit 1      # builds tests/emu/_build/libh264bsd_b200_hostemu.so
sel="${1:-(legacy_api_bit_exact and 640x360) or getters or (stages_in_isolation and 640x360) or streamed_upload or (colour and 640x360)}"
B200_LIB="$PWD/tests/emu/_build/libh264bsd_b200_hostemu.so" python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py -m gpu -q -x --durations=10 -k "$sel"
B200_LIB="$PWD/tests/emu/_build/libh264bsd_b200_hostemu.so" python -c "import __graft_entry__ as g; g.smoke()"
