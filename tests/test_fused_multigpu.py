"""Multi-GPU: the in-kernel NVLink path (comm="fused": rs_sumsq pull, adamw_allgather push, device barriers over
symmetric memory) must train exactly like the NCCL collective path, and FSDP must match the single-GPU run."""
import pytest
import torch

from tests._mp import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _cfg(comm, gather_mode="step", shard=True):
    from midgpt_b200 import load_config
    from midgpt_b200.config import GPTConfig
    cfg = load_config("shakespeare_char")
    cfg.model_config = GPTConfig(block_size=256, vocab_size=65, n_layer=3, n_head=2, n_embd=256, dropout=0.0)
    cfg.batch_size, cfg.g_accum_iters, cfg.shard_model = 8, 1, shard
    cfg.warmup_steps, cfg.lr_decay_steps, cfg.max_steps = 2, 50, 50
    cfg.engine.backend, cfg.engine.comm, cfg.engine.micro_batch_size = "b200", comm, 2
    cfg.engine.gather_mode = gather_mode
    return cfg


def _run(rank, world, comm, gather_mode, steps):
    from midgpt_b200.train import Trainer
    torch.cuda.set_device(rank)
    tr = Trainer(_cfg(comm, gather_mode), device_type="cuda")
    g = torch.Generator().manual_seed(7)
    losses = []
    for _ in range(steps):
        x = torch.randint(0, 65, (8, 256), generator=g, dtype=torch.int32)
        y = torch.randint(0, 65, (8, 256), generator=g, dtype=torch.int32)
        n = 8 // world
        losses.append(tr.train_step(x[rank * n:(rank + 1) * n].cuda(), y[rank * n:(rank + 1) * n].cuda()).item())
    state = tr.engine.full_state("master")
    return losses, {k: v.float().cpu() for k, v in state.items()}, tr.engine.last_grad_norm.item()


@pytest.mark.parametrize("gather_mode", ["step", "layer"])
def test_fused_comm_matches_nccl(gather_mode):
    """step: rs_sumsq pull + adamw_allgather push;  layer: rs_sumsq pull + all-gather ⊕ GEMM (csrc/gemm.cu gather path)."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("not enough GPUs")
    world = 2 if world < 4 else world
    steps = 5
    nccl = run_distributed(_run, world, "nccl", "step", steps, timeout=600)
    fused = run_distributed(_run, world, "fused", gather_mode, steps, timeout=600)
    for r in range(world):
        ln, sn, gn = nccl[r]
        lf, sf, gf = fused[r]
        assert max(abs(a - b) for a, b in zip(ln, lf)) < 2e-3, (ln, lf)
        assert abs(gn - gf) < 1e-3 * max(1.0, gn)
        for k in sn:
            err = (torch.as_tensor(sf[k]) - torch.as_tensor(sn[k])).norm() / (torch.as_tensor(sn[k]).norm() + 1e-9)
            assert err < 2e-3, (k, err.item())
    assert nccl[0][0][-1] < nccl[0][0][0]
