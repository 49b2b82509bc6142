#!/bin/bash
# compute-sanitizer targets for the hand-written kernels (the reference has no race detection at all).
#   bash scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]     (run on a GPU box, e.g. under gpurun)
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout ${SAN_TMO:-1500} compute-sanitizer --tool $TOOL --error-exitcode 1 \
  python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm_nt or rmsnorm or qkln or embedding or adamw or attention" -p no:cacheprovider \
  > gpurun_out/sanitize_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL exit $?"
tail -n 15 gpurun_out/sanitize_$TOOL.log
