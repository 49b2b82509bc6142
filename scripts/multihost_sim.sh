#!/bin/bash
# Multi-host rehearsal on ONE box: two torchrun agents ("nodes", --nnodes=2) with disjoint CUDA_VISIBLE_DEVICES rendezvous over
# 127.0.0.1, so launch.py --multihost runs exactly the code path of a real 2-node job (reference launch.py:22-23,69-70):
# shard group = the GPUs of one agent (in-kernel NVLink collectives inside it), replica group across the agents (NCCL
# all-reduce of shard-sized gradients, clip norm taken AFTER it).  Then: checkpoint -> kill -> relaunch with the same --rundir
# (restart-from-checkpoint recovery, reference README.md:29-33).
#   scripts/multihost_sim.sh <gpus per node> <rundir> [max_steps]
NPER=${1:-2}; RUNDIR=${2:-/tmp/mh_run}; STEPS=${3:-12}; PORT=${MH_PORT:-29611}
launch() {   # $1 = max_steps
  for node in 0 1; do
    first=$((node * NPER)); devs=$(seq -s, $first $((first + NPER - 1)))
    CUDA_VISIBLE_DEVICES=$devs python -m torch.distributed.run --nnodes=2 --node-rank=$node --nproc-per-node=$NPER \
      --master-addr=127.0.0.1 --master-port=$PORT launch.py --config=shakespeare_char --rundir=$RUNDIR --multihost --synthetic \
      --shard_model=1 --max_steps=$1 --set=eval_interval=5 --set=batch_size=16 --set=model_config.dropout=0.0 \
      --set=engine.log_every=1 > $RUNDIR.node$node.log 2>&1 &
    pids[$node]=$!
  done
  rc=0; for p in ${pids[@]}; do wait $p || rc=$?; done; return $rc
}
rm -rf $RUNDIR; mkdir -p $(dirname $RUNDIR)
launch $((STEPS / 2 + 1)) && echo "[multihost_sim] first run ok" || { echo "[multihost_sim] first run FAILED"; tail -20 $RUNDIR.node0.log; exit 1; }
grep -E "engine\]|fused-comm" $RUNDIR.node0.log | head -3
launch $STEPS && echo "[multihost_sim] resumed run ok" || { echo "[multihost_sim] resume FAILED"; tail -20 $RUNDIR.node0.log; exit 1; }
grep -E "resumed from step" $RUNDIR.node0.log
python - <<PY
import json
rows = [json.loads(l) for l in open("$RUNDIR/metrics.jsonl")]
opt = [(r["step"], round(r["loss/optimized"], 4), round(r.get("grad_norm", 0), 4)) for r in rows if "loss/optimized" in r]
print("[multihost_sim] steps logged:", [s for s, _, _ in opt])
print("[multihost_sim] loss first/last:", opt[0][1], opt[-1][1])
assert opt[-1][1] < opt[0][1]
PY
