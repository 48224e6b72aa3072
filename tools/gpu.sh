#!/bin/bash
# Build in-tree, then run a command on a B200 box:  tools/gpu.sh [--timeout S] '<command>'
set -e
cd "$(dirname "$0")/.."
python distributedconvrl-pde-control_b200/build.py >/dev/null
make -C oracle -s
T=900
if [ "$1" == "--timeout" ]; then T=$2; shift 2; fi
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
