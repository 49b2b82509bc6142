#!/bin/bash
# compute-sanitizer over every hand-written kernel family (the reference has no race detection at all).
#   bash scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck] ...     (run on a GPU box, e.g. under gpurun)
# Targets: tests/sanitize_targets.py (tiny shapes: pair / single-CTA / fp8 / MX GEMMs with every epilogue incl. split-K wgrad,
# attention forward + backward v3 + v2-with-dropout for both head dims, element-wise, cross-entropy, optimizer).
mkdir -p gpurun_out
for TOOL in "${@:-memcheck}"; do
  # synccheck flags every mbarrier phase that completes without a waiter; the forward attention awaits pv_done lazily by design
  # (only when a row maximum forces a rescale), so that tool runs the kernel with eager waits (MIDGPT_ATTN_DEBUG bit 8)
  [ "$TOOL" = synccheck ] && export MIDGPT_ATTN_DEBUG=8 || unset MIDGPT_ATTN_DEBUG
  timeout ${SAN_TMO:-900} compute-sanitizer --tool $TOOL --error-exitcode 1 \
    python -m pytest tests/sanitize_targets.py -q -m gpu -p no:cacheprovider > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "compute-sanitizer $TOOL exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|passed|failed|Error" gpurun_out/sanitize_$TOOL.log | tail -n 8
done
