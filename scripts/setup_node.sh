#!/bin/bash
# Per-node setup — the counterpart of the reference's scripts/setup.sh (install the accelerator stack + requirements, mount the
# read-only dataset disk at /mnt/disks/persist).  On a B200 node nothing is installed from an index: the image carries torch +
# CUDA; this script verifies the stack, builds the in-tree sm_100a extensions, wires up the dataset directory and runs the smoke.
#   bash scripts/setup_node.sh [dataset_dir]        (run on every node: `gpu setup` fans it out)
set -u
DATA_SRC=${1:-${MIDGPT_DATA:-}}
cd "$(dirname "$0")/.."
echo "== host $(hostname): driver / GPUs"
nvidia-smi --query-gpu=index,name,driver_version,memory.total --format=csv,noheader || { echo "no NVIDIA driver"; exit 1; }
python - <<'PY' || exit 1
import torch, sys
cap = torch.cuda.get_device_capability(0) if torch.cuda.is_available() else None
print(f"torch {torch.__version__}, cuda {torch.version.cuda}, nccl {'.'.join(map(str, torch.cuda.nccl.version()))}, capability {cap}, gpus {torch.cuda.device_count()}")
if cap is None or cap[0] != 10:
    sys.exit("this engine is sm_100a-only: no compute-capability-10 device found")
PY
echo "== NVLink / NVSwitch"
nvidia-smi topo -m 2>/dev/null | head -12
nvidia-smi nvlink -s 2>/dev/null | grep -c "GB/s" | xargs -I{} echo "{} active NVLink lanes reported"
echo "== build (nvcc -gencode arch=compute_100a,code=sm_100a, in-tree)"
python -m midgpt_b200.ops.build || exit 1
if [ -n "$DATA_SRC" ]; then
  # the reference mounts a persistent disk at /mnt/disks/persist and points data_dir there (openwebtext_mh / _xl presets)
  sudo mkdir -p /mnt/disks 2>/dev/null; sudo ln -sfn "$DATA_SRC" /mnt/disks/persist 2>/dev/null || ln -sfn "$DATA_SRC" ./persist
  for f in openwebtext/train.bin openwebtext/val.bin; do [ -f "$DATA_SRC/$f" ] && echo "dataset ok: $DATA_SRC/$f ($(du -h "$DATA_SRC/$f" | cut -f1))" || echo "WARNING: $DATA_SRC/$f missing (launch.py falls back to a synthetic corpus)"; done
fi
echo "== smoke"
python __graft_entry__.py --smoke
