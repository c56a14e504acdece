#!/bin/bash
# Build the in-tree library, then run a command on the B200 box:  scripts/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -m cds_mvsnet_b200.build
T=${1:-600}; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
