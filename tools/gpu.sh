#!/bin/bash
# Rebuild the in-tree library, then run a command on the GPU box: tools/gpu.sh [--gpus N] <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
./mrhash_b200/build.sh | grep -E "error|built"
make -s -C oracle oracle
T=$1; shift
exec /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "mkdir -p gpurun_out; $*"
