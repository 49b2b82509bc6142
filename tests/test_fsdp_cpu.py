"""FSDP (gloo, world_size=2, CPU) ≡ single process: losses and post-step parameters (BASELINE config 1 plumbing)."""

import pytest
import torch

from midgpt_b200 import load_config
from midgpt_b200.config import GPTConfig
from tests._mp import run_distributed


def _tiny_cfg(shard_model=True, gather_mode="step", g=2, dropout=0.0):
    cfg = load_config("shakespeare_char")
    cfg.model_config = GPTConfig(block_size=16, vocab_size=65, n_layer=3, n_head=2, n_embd=32, dropout=dropout)
    cfg.batch_size, cfg.g_accum_iters, cfg.shard_model = 4, g, shard_model
    cfg.warmup_steps, cfg.lr_decay_steps, cfg.max_steps = 2, 20, 20
    cfg.compute_dtype = "float32"
    cfg.engine.gather_mode = gather_mode
    cfg.engine.micro_batch_size = 1
    return cfg


def _batches(cfg, n_steps):
    g = torch.Generator().manual_seed(123)
    T, N = cfg.model_config.block_size, cfg.batch_size * cfg.g_accum_iters
    return [(torch.randint(0, 65, (N, T), generator=g, dtype=torch.int32),
             torch.randint(0, 65, (N, T), generator=g, dtype=torch.int32)) for _ in range(n_steps)]


def _run(rank, world, shard_model, gather_mode, n_steps, grad_shard=False):
    from midgpt_b200.train import Trainer
    torch.manual_seed(0)
    cfg = _tiny_cfg(shard_model, gather_mode)
    cfg.engine.grad_shard = grad_shard
    tr = Trainer(cfg, device_type="cpu")
    losses = []
    for x, y in _batches(cfg, n_steps):
        # global batch → this rank's slice of every accumulation step
        G, B, T = cfg.g_accum_iters, cfg.batch_size, cfg.model_config.block_size
        xs = x.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        ys = y.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        losses.append(tr.train_step(xs, ys).item())
    state = {k: v.clone() for k, v in tr.engine.full_state("master").items()}
    if grad_shard and world > 1:       # ZeRO-2/3 gradient memory: no full-size accumulator exists
        assert all(u.grad_full is None for u in tr.engine.units)
        assert tr.engine.memory_report()["grad_full_GB"] < 4 * sum(u.padded for u in tr.engine.units) * 4 / 2 ** 30
    return losses, state, tr.engine.last_grad_norm.item()


@pytest.mark.parametrize("shard_model,gather_mode,grad_shard", [(True, "step", False), (True, "layer", False), (False, "step", False),
                                                              (True, "layer", True), (True, "step", True)])
def test_fsdp_ws2_equals_single(shard_model, gather_mode, grad_shard):
    """grad_shard=True: gradients are reduce-scattered every micro-step (g_accum = 2 x 2 micro-batches here) into a shard-sized
    accumulator -- the reference's per-microstep sharding constraint (src/train.py:87)."""
    n_steps = 4
    ref_losses, ref_state, ref_gn = _run(0, 1, False, "step", n_steps)
    outs = run_distributed(_run, 2, shard_model, gather_mode, n_steps, grad_shard)
    for losses, state, gn in outs:
        assert torch.allclose(torch.tensor(losses), torch.tensor(ref_losses), rtol=1e-5, atol=1e-6)
        assert abs(gn - ref_gn) < 1e-4 * max(1.0, ref_gn)
        for k in ref_state:
            assert torch.allclose(torch.as_tensor(state[k]), ref_state[k], rtol=1e-4, atol=1e-6), k
    assert ref_losses[-1] < ref_losses[0] + 0.5   # sanity: not diverging


def test_loss_decreases_single_process():
    from midgpt_b200.train import Trainer
    cfg = _tiny_cfg(False)
    cfg.learning_rate, cfg.warmup_steps = 3e-3, 2
    tr = Trainer(cfg, device_type="cpu")
    x, y = _batches(cfg, 1)[0]
    losses = [tr.train_step(x, y).item() for _ in range(12)]
    assert losses[-1] < losses[1] - 0.3, losses


def _run_hsdp(rank, world, n_steps):
    """world=4 as 2 shard groups of 2 (HSDP: reduce-scatter inside a group, all-reduce across replicas — X5)."""
    from midgpt_b200.train import Trainer
    from midgpt_b200.parallel import dist as D
    torch.manual_seed(0)
    cfg = _tiny_cfg(True, "step")
    mesh = D.make_mesh(True, "cpu", max_shard=2)
    assert mesh.shard_size == 2 and mesh.replica_size == 2
    tr = Trainer(cfg, mesh=mesh)
    losses = []
    for x, y in _batches(cfg, n_steps):
        G, B, T = cfg.g_accum_iters, cfg.batch_size, cfg.model_config.block_size
        xs = x.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        ys = y.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        losses.append(tr.train_step(xs, ys).item())
    return losses, {k: v.clone() for k, v in tr.engine.full_state("master").items()}


def test_hsdp_world4_equals_single():
    n_steps = 3
    ref_losses, ref_state, _ = _run(0, 1, False, "step", n_steps)
    outs = run_distributed(_run_hsdp, 4, n_steps)
    for losses, state in outs:
        assert torch.allclose(torch.tensor(losses), torch.tensor(ref_losses), rtol=1e-5, atol=1e-6)
        for k in ref_state:
            assert torch.allclose(torch.as_tensor(state[k]), ref_state[k], rtol=1e-4, atol=1e-6), k


def _run_mesh_from_launcher_env(rank, world):
    """torchrun --nnodes=2 --nproc-per-node=2 sets LOCAL_WORLD_SIZE=2: the shard group must be one node, replicas across nodes."""
    import os
    from midgpt_b200.parallel import dist as D
    os.environ["LOCAL_WORLD_SIZE"] = "2"
    mesh = D.make_mesh(True, "cpu")
    dp = D.make_mesh(False, "cpu")
    return (mesh.shard_size, mesh.replica_size, mesh.shard_rank, mesh.replica_rank, dp.shard_size, dp.replica_size)


def test_shard_group_is_the_node_reported_by_the_launcher():
    outs = run_distributed(_run_mesh_from_launcher_env, 4)
    assert [o[:4] for o in outs] == [(2, 2, 0, 0), (2, 2, 1, 0), (2, 2, 0, 1), (2, 2, 1, 1)]
    assert all(o[4:] == (1, 4) for o in outs)          # shard_model=False: pure DP over all ranks
