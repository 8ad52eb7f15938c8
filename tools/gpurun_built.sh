#!/bin/bash
# build the library here (nvcc cross-compiles), then hand the command to gpurun: a stale .so never travels
cd "$(dirname "$0")/.." || exit 1
bash h264bsd_b200/build.sh > /dev/null || exit 1
exec /usr/local/graft/bin/gpurun "$@"
