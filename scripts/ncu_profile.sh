#!/bin/bash
# Full ncu capture of the hot kernels (one GPU; ~40 replays per launch) → gpurun_out/prof_<name>.ncu-rep
mkdir -p gpurun_out
NAME=${1:-hot}; shift
REGEX=${1:-"gemm_bf16_tcgen05|attn_fwd_kernel|attn_bwd_kernel"}; shift
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s ${SKIPK:-3} -c ${COUNT:-5} -f -o gpurun_out/prof_$NAME \
   python scripts/prof_targets.py "$@" > gpurun_out/ncu_$NAME.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$NAME.log
