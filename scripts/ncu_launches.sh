#!/bin/bash
# Launch list with device time per kernel for ONE optimizer step of the 1.5B config (32 sequences), serialised by ncu.
# Shares are what matters (cold cache, serialised) — see B200_PROFILING.md.
mkdir -p gpurun_out
SKIP=${SKIP:-1900}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 2600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --seqs-per-gpu 32 --micro-bs 32 --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt | head -40
