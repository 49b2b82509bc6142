"""Multi-GPU: the in-kernel NVLink path (comm="fused": rs_sumsq pull, adamw_allgather push, device barriers over
symmetric memory) must train exactly like the NCCL collective path, and FSDP must match the single-GPU run."""
import pytest
import torch

from tests._mp import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _cfg(comm, gather_mode="step", shard=True, grad_shard=False):
    from midgpt_b200 import load_config
    from midgpt_b200.config import GPTConfig
    cfg = load_config("shakespeare_char")
    cfg.model_config = GPTConfig(block_size=256, vocab_size=65, n_layer=3, n_head=2, n_embd=256, dropout=0.0)
    cfg.batch_size, cfg.g_accum_iters, cfg.shard_model = 8, 1, shard
    cfg.warmup_steps, cfg.lr_decay_steps, cfg.max_steps = 2, 50, 50
    cfg.engine.backend, cfg.engine.comm, cfg.engine.micro_batch_size = "b200", comm, 2
    cfg.engine.gather_mode = gather_mode
    cfg.engine.grad_shard = grad_shard
    return cfg


def _run(rank, world, comm, gather_mode, steps, grad_shard=False, local_world=0):
    import os
    from midgpt_b200.train import Trainer
    torch.cuda.set_device(rank)
    if local_world:                      # pretend the ranks live on world / local_world nodes: shard inside a "node", replicate across
        os.environ["LOCAL_WORLD_SIZE"] = str(local_world)
    tr = Trainer(_cfg(comm, gather_mode, grad_shard=grad_shard), device_type="cuda")
    g = torch.Generator().manual_seed(7)
    losses = []
    for _ in range(steps):
        x = torch.randint(0, 65, (8, 256), generator=g, dtype=torch.int32)
        y = torch.randint(0, 65, (8, 256), generator=g, dtype=torch.int32)
        n = 8 // world
        losses.append(tr.train_step(x[rank * n:(rank + 1) * n].cuda(), y[rank * n:(rank + 1) * n].cuda()).item())
    state = tr.engine.full_state("master")
    return losses, {k: v.float().cpu() for k, v in state.items()}, tr.engine.last_grad_norm.item()


@pytest.mark.parametrize("gather_mode,grad_shard", [("step", False), ("layer", False), ("layer", True)])
def test_fused_comm_matches_nccl(gather_mode, grad_shard):
    """step: rs_sumsq pull + adamw_allgather push;  layer: rs_sumsq pull + all-gather ⊕ GEMM (csrc/gemm.cu gather path)."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("not enough GPUs")
    world = 2 if world < 4 else world
    steps = 5
    nccl = run_distributed(_run, world, "nccl", "step", steps, timeout=600)
    fused = run_distributed(_run, world, "fused", gather_mode, steps, grad_shard, timeout=600)
    for r in range(world):
        ln, sn, gn = nccl[r]
        lf, sf, gf = fused[r]
        assert max(abs(a - b) for a, b in zip(ln, lf)) < 2e-3, (ln, lf)
        assert abs(gn - gf) < 1e-3 * max(1.0, gn)
        for k in sn:
            err = (torch.as_tensor(sf[k]) - torch.as_tensor(sn[k])).norm() / (torch.as_tensor(sn[k]).norm() + 1e-9)
            assert err < 2e-3, (k, err.item())
    assert nccl[0][0][-1] < nccl[0][0][0]


def test_hsdp_fused_matches_nccl():
    """4 GPUs as 2 "nodes" of 2: with the in-kernel path the clip norm must be taken AFTER the replica
    all-reduce, otherwise every replica group clips with its own norm and the replicas' weights drift apart silently."""
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    steps = 5
    nccl = run_distributed(_run, 4, "nccl", "step", steps, False, 2, timeout=600)
    fused = run_distributed(_run, 4, "fused", "step", steps, False, 2, timeout=600)
    for r in range(4):
        ln, sn, gn = nccl[r]
        lf, sf, gf = fused[r]
        assert max(abs(a - b) for a, b in zip(ln, lf)) < 2e-3, (ln, lf)
        assert abs(gn - gf) < 1e-3 * max(1.0, gn), (gn, gf)
        for k in sn:
            err = (torch.as_tensor(sf[k]) - torch.as_tensor(sn[k])).norm() / (torch.as_tensor(sn[k]).norm() + 1e-9)
            assert err < 2e-3, (k, err.item())
    # replicas hold identical weights (rank 0 is in replica group 0, rank 2 in replica group 1, same shard rank)
    for name, runs in (("nccl", nccl), ("fused", fused)):
        for k in runs[0][1]:
            a, b = torch.as_tensor(runs[0][1][k]), torch.as_tensor(runs[2][1][k])
            assert torch.equal(a, b), (name, k, (a - b).abs().max().item())
    assert nccl[0][2] > 1.0, "the test must exercise an active clip"


def _flag_stress(rank, world, iters):
    """Every iteration: stamp the local symmetric buffer, device barrier, read EVERY peer's buffer over NVLink and check its stamp,
    device barrier again (so nobody overwrites a buffer that is still being read).  Random per-rank spin kernels in between
    shake the relative timing of the ranks; a lost release/acquire or an epoch mix-up shows up as a stale stamp."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm
    from midgpt_b200.ops.b200_backend import B200Backend
    from midgpt_b200.parallel.dist import init_distributed
    rank, world, local, dev, _ = init_distributed("cuda")
    be = B200Backend(dev)
    n = 1 << 16
    buf = symm.empty(n, dtype=torch.int32, device=dev)
    pad = symm.empty(1024, dtype=torch.int32, device=dev)
    buf.zero_(); pad.zero_()
    h_buf, h_pad = symm.rendezvous(buf, dist.group.WORLD), symm.rendezvous(pad, dist.group.WORLD)
    peers = [h_buf.get_buffer(p, (n,), torch.int32) for p in range(world)]
    pads = [int(p) for p in h_pad.buffer_ptrs]
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.Generator().manual_seed(100 + rank)
    bad = torch.zeros((), dtype=torch.int64, device=dev)
    epoch = 0
    for it in range(iters):
        if torch.rand((), generator=g).item() < 0.5:
            torch.cuda._sleep(int(torch.randint(1, 200_000, (), generator=g)))
        buf.fill_(it * 16 + rank)
        epoch += 1
        be.ext.xrank_barrier(pads, rank, 1, epoch)
        for p in range(world):
            bad += (peers[p] != it * 16 + p).sum()
        epoch += 1
        be.ext.xrank_barrier(pads, rank, 1, epoch)
    torch.cuda.synchronize()
    return int(bad.item())


def test_device_barrier_flag_protocol_stress():
    """signal/flag-protocol stress test for the in-kernel P2P paths (csrc/comm.cu xrank_barrier)."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("not enough GPUs")
    out = run_distributed(_flag_stress, world, 300, timeout=300)
    assert out == [0] * world, out
