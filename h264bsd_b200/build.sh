#!/bin/bash
# Build libh264bsd_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/csrc"
out="${B200_OUT:-$here/libh264bsd_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared \
  -I"$here/../include" -I"$src/engine" -I"$src/host" \
  "$src/engine/engine.cu" "$src/api/api.cpp" \
  "$src/host/cavlc.cpp" "$src/host/params.cpp" "$src/host/dpb.cpp" "$src/host/picture.cpp" \
  "$src/host/stream_decoder.cpp" "$src/host/tape_builder.cpp" \
  -o "$out" -lcudart_static -ldl -lrt -lpthread "$@"
echo "built $out"
