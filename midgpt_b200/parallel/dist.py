"""Process-group bootstrap and thin collective wrappers (X10 of SURVEY.md §2.5).

One process per GPU; ``torch.distributed`` with NCCL on GPUs and gloo on CPU (the plumbing config
runs world_size=2 on gloo).  The reference's analogue is ``jax.distributed.initialize()``
(launch.py:22-23) plus the ('replica','data') mesh (src/train.py:128-130); here the two mesh axes
become two process groups: ``shard`` (FSDP group, ≤ 8 ranks = one NVSwitch domain) and
``replica`` (ranks holding the same shard on different nodes).
"""
from __future__ import annotations

import os
import datetime
import typing as tp
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class Mesh:
    rank: int
    world: int
    local_rank: int
    device: torch.device
    shard_size: int                    # ranks per FSDP group
    shard_rank: int
    replica_size: int
    replica_rank: int
    shard_group: tp.Any = None         # None => trivial (size 1)
    replica_group: tp.Any = None
    backend: str = "none"

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def init_distributed(device_type: str | None = None) -> tp.Tuple[int, int, int, torch.device, str]:
    """Initialise torch.distributed from torchrun-style env vars; returns (rank, world, local_rank, device, backend)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if device_type is None:
        device_type = "cuda" if torch.cuda.is_available() else "cpu"
    if device_type == "cuda":
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        backend = "nccl"
    else:
        device = torch.device("cpu")
        backend = "gloo"
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(minutes=10), **kwargs)
    return rank, world, local_rank, device, (backend if world > 1 else "none")


def make_mesh(shard_model: bool, device_type: str | None = None, max_shard: int = 8) -> Mesh:
    """shard group = the ranks of ONE NODE (``LOCAL_WORLD_SIZE`` consecutive ranks, at most 8) when ``shard_model`` else 1
    (pure DP); replica groups run across nodes (HSDP).

    The reference hard-codes 8 (``n_devices // 8``, src/train.py:129-130) and therefore cannot run on 1/2/4 devices; here the
    group is the NVLink domain the launcher reports: parameters are sharded where peer memory / NVSwitch multicast exist (the
    in-kernel collectives need it) and replicated across nodes, where only NCCL all-reduces of shard-sized gradients travel.
    """
    rank, world, local_rank, device, backend = init_distributed(device_type)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if local_world < 1 or world % local_world:
        local_world = world
    S = min(world, max_shard, local_world) if shard_model else 1
    while world % S:
        S -= 1
    R = world // S
    shard_group = replica_group = None
    if world > 1:
        if S == world:
            shard_group = dist.group.WORLD
        elif S > 1:
            for r in range(R):
                g = dist.new_group(list(range(r * S, (r + 1) * S)))
                if rank // S == r:
                    shard_group = g
        if R == world:
            replica_group = dist.group.WORLD
        elif R > 1:
            for s in range(S):
                g = dist.new_group(list(range(s, world, S)))
                if rank % S == s:
                    replica_group = g
    return Mesh(rank, world, local_rank, device, S, rank % S, R, rank // S, shard_group, replica_group, backend)


# ------------------------------------------------------------------------------------------------
# Collective wrappers: NCCL fast paths, gloo fall-backs (gloo has no reduce_scatter)
# ------------------------------------------------------------------------------------------------

COLLECTIVE_CALLS = {"n": 0}      # library collectives issued through this module (bench.py reports the count)


def all_gather_flat(out_full: torch.Tensor, in_shard: torch.Tensor, group, backend: str) -> None:
    if group is None:
        if out_full.data_ptr() != in_shard.data_ptr():
            out_full[: in_shard.numel()].copy_(in_shard)
        return
    COLLECTIVE_CALLS["n"] += 1
    if backend == "nccl":
        dist.all_gather_into_tensor(out_full, in_shard, group=group)
    else:
        n = dist.get_world_size(group)
        chunks = list(out_full.view(n, -1).unbind(0))
        dist.all_gather(chunks, in_shard.clone(), group=group)


def reduce_scatter_avg(out_shard: torch.Tensor, in_full: torch.Tensor, group, backend: str) -> None:
    """out_shard = mean over the group of in_full[rank-th chunk]."""
    if group is None:
        if out_shard.data_ptr() != in_full.data_ptr():
            out_shard.copy_(in_full[: out_shard.numel()])
        return
    n = dist.get_world_size(group)
    COLLECTIVE_CALLS["n"] += 1
    if backend == "nccl":
        dist.reduce_scatter_tensor(out_shard, in_full, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(in_full, group=group)
        r = dist.get_rank(group)
        out_shard.copy_(in_full.view(n, -1)[r]).div_(n)


def all_reduce_avg(t: torch.Tensor, group, backend: str) -> None:
    if group is None:
        return
    COLLECTIVE_CALLS["n"] += 1
    if backend == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, group=group)
        t.div_(dist.get_world_size(group))


def all_reduce_sum(t: torch.Tensor, group) -> None:
    if group is not None:
        dist.all_reduce(t, group=group)


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
