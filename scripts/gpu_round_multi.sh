#!/bin/bash
# Multi-GPU round (run with `gpurun --gpus N`): fused-comm tests, distributed smoke script, bench at N GPUs for both comm paths.
N=${1:-2}
STAGES="${2:-dist tests nccl fused comm}"
mkdir -p gpurun_out
rm -f gpurun_out/stages_multi.txt
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "stage $name exit $?" >> gpurun_out/stages_multi.txt; tail -n 3 gpurun_out/$name.log | cut -c1-400; }
[[ $STAGES == *dist* ]] && run m_dist python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/test_dist.py /tmp/ckpt_test
[[ $STAGES == *tests* ]] && run m_tests python -m pytest tests/test_fused_multigpu.py -q -m gpu -p no:cacheprovider
[[ $STAGES == *nccl* ]] && TMO=900 run m_bench_nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 --comm nccl --no-e2e
[[ $STAGES == *fused* ]] && TMO=900 run m_bench_fused python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 3 --warmup 3 --comm fused
[[ $STAGES == *comm* ]] && TMO=600 run m_commbench python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench/comm_bench.py
cat gpurun_out/stages_multi.txt
