#!/bin/bash
# Build in-tree, then run a command on a B200 box:  tools/gpu.sh [--timeout S] [--gpus N] '<command>'
set -e
cd "$(dirname "$0")/.."
python distributedconvrl-pde-control_b200/build.py >/dev/null
make -C oracle -s
T=900
G=()
while [[ "$1" == --* ]]; do
  case "$1" in
    --timeout) T=$2; shift 2;;
    --gpus) G=(--gpus "$2"); shift 2;;
    *) break;;
  esac
done
exec /usr/local/graft/bin/gpurun --timeout "$T" "${G[@]}" -- "$@"
