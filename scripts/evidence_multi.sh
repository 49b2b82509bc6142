#!/bin/bash
# One multi-GPU evidence pass (run with `gpurun --gpus N -- bash scripts/evidence_multi.sh N "<stages>"`): every stage has its own
# timeout; logs land in gpurun_out/ (copy what matters into profiles/).
#   stages: tests comm bench layer seq4096 fp8 timeline
N=${1:-2}; STAGES="${2:-tests comm bench layer seq4096 timeline}"
mkdir -p gpurun_out; rm -f gpurun_out/stages_multi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; shift; timeout ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "stage $name exit $?" >> gpurun_out/stages_multi.txt; }
short() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); s = d.get("secondary", {}).get("openwebtext_124M")
        print(sys.argv[1].split("/")[-1], d["config"]["model"].split()[0], d["config"]["gather_mode"], d["config"]["comm"], "remat", d["config"]["remat"],
              "tok/s", round(d["value"]), "ms", round(d["ms_per_step"], 1), "mfu", round(d["mfu_nominal_2.25PF"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]),
              "exposed_ms", round(d["comm_exposed_ms"], 2), "peakGB", round(d["peak_mem_GB"], 1), d["comm_kernel_launches"], d["clocks"])
        if s: print("   124M tok/s", round(s["value"]), "ms", round(s["ms_per_step"], 1), "mfu", round(s["mfu_nominal_2.25PF"], 3), "e2e", s["e2e"] and round(s["e2e"]["value"]), "exposed_ms", round(s["comm_exposed_ms"], 2))
PY
}
for s in $STAGES; do case $s in
  tests)    TMO=600 run m_tests python -m pytest tests/test_fused_multigpu.py -q -m multigpu -p no:cacheprovider; tail -n 3 gpurun_out/m_tests.log ;;
  comm)     run comm_bench_n${N} $TR --master-port 29544 bench/comm_bench.py; grep "^{" gpurun_out/comm_bench_n${N}.log > gpurun_out/comm_bench_n${N}_r2.jsonl; cut -c1-330 gpurun_out/comm_bench_n${N}_r2.jsonl ;;
  bench)    TMO=500 run bench_n${N} $TR --master-port 29545 bench.py --gpus $N --steps ${STEPS:-6} --warmup 3; short gpurun_out/bench_n${N}.log ;;
  layer)    for gm in step layer; do run bench_n${N}_${gm}_remat $TR --master-port 29546 bench.py --gpus $N --steps 4 --warmup 3 --no-secondary --no-e2e --remat block --gather-mode $gm; short gpurun_out/bench_n${N}_${gm}_remat.log; done
            run bench_n${N}_zero3 $TR --master-port 29547 bench.py --gpus $N --steps 4 --warmup 3 --no-secondary --no-e2e --remat block --gather-mode layer --set engine.grad_shard=true; short gpurun_out/bench_n${N}_zero3.log ;;
  seq4096)  run bench_n${N}_seq4096 $TR --master-port 29548 bench.py --gpus $N --steps 4 --warmup 3 --no-secondary --set model_config.block_size=4096 --seqs-per-gpu 32 --remat block; short gpurun_out/bench_n${N}_seq4096.log ;;
  fp8)      run bench_n${N}_fp8 $TR --master-port 29549 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --fp8; short gpurun_out/bench_n${N}_fp8.log ;;
  timeline) run timeline_n${N} $TR --master-port 29550 scripts/timeline_multi.py; grep "^{" gpurun_out/timeline_n${N}.log | cut -c1-1500 ;;
esac; done
cat gpurun_out/stages_multi.txt
