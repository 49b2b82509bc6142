#!/bin/bash
# One gpurun call = one pass over: kernel numerics, micro-benchmarks, model test, bench.  Every stage has its own timeout
# so a hung kernel cannot take the box (and a strike) with it.  Logs land in gpurun_out/.
mkdir -p gpurun_out
STAGES="${@:-gemm elem attn model kbench bench}"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for s in $STAGES; do
  case $s in
    gemm)   timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1 ;;
    elem)   timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not gemm and not attention" -p no:cacheprovider > gpurun_out/t_elem.log 2>&1 ;;
    attn)   timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" -p no:cacheprovider > gpurun_out/t_attn.log 2>&1 ;;
    model)  timeout 400 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_model.log 2>&1 ;;
    kbench) timeout 400 python bench/kernel_bench.py > gpurun_out/kbench.log 2>&1 ;;
    kquick) timeout 300 python bench/kernel_bench.py --quick > gpurun_out/kbench.log 2>&1 ;;
    bench)  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1 ;;
    bench_torch) timeout 900 python bench.py --steps 3 --warmup 3 --backend torch --no-e2e > gpurun_out/bench_torch.log 2>&1 ;;
  esac
  echo "stage $s exit $?" >> gpurun_out/stages.txt
  tail -n 4 gpurun_out/*.log 2>/dev/null | tail -n 12
done
cat gpurun_out/stages.txt
