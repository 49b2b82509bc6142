"""The ZeRO-3 "carried gather" protocol of the FSDP engine (gather_mode="layer" on the in-kernel path), exercised on CPU.

On a GPU the gather of every unit is carried by a GEMM kernel (csrc/gemm.cu); which GEMM carries which unit, when pool slots are
handed out and returned, and what happens at the forward/backward turn-around or in evaluation is host-side logic in
``FSDPEngine._acquire_fused_layer`` / ``_open_gather`` / ``release``.  Here a fake ``fused`` object performs the collectives with
gloo and a wrapper backend "carries" a gather by running the all-gather at the moment the carrier GEMM is called -- so a unit
whose gather nobody carried, a stale pool slot or a leaked slot shows up as wrong numbers or as the engine's own errors.
world_size = 2 must reproduce the single-process run exactly (same test as tests/test_fsdp_cpu.py, different machinery)."""
import torch

from tests._mp import run_distributed
from tests.test_fsdp_cpu import _tiny_cfg, _batches


class _Spec:
    def __init__(self, unit):
        self.unit, self.done, self.carried_by = unit, False, None


class FakeFused:
    """Stands in for parallel.fused.FusedComm: same interface towards the engine, gloo underneath."""

    def __init__(self, engine):
        self.e = engine
        self.layer_mode = True
        self.carried, self.pulled = [], []
        self.stats = {}

    def gather_spec(self, u):
        return _Spec(u)

    def all_gather(self, u):                       # stand-alone pull (the embedding)
        from midgpt_b200.parallel import dist as D
        D.all_gather_flat(u.lp_full, u.lp_shard, self.e.mesh.shard_group, self.e.mesh.backend)
        self.pulled.append(u.spec.name)

    def carry(self, spec, by):                     # what the carrier GEMM's copy warps do
        from midgpt_b200.parallel import dist as D
        assert not spec.done, "a gather must be carried exactly once"
        u = spec.unit
        D.all_gather_flat(u.lp_full, u.lp_shard, self.e.mesh.shard_group, self.e.mesh.backend)
        spec.done, spec.carried_by = True, by
        self.carried.append((u.spec.name, by))

    def reduce_scatter(self, u):
        from midgpt_b200.parallel import dist as D
        D.reduce_scatter_avg(u.grad_shard, u.grad_full, self.e.mesh.shard_group, self.e.mesh.backend)

    def optimizer_step(self):
        from midgpt_b200.optim import step_scalars, clip_factor_from_sumsq
        from midgpt_b200.parallel import dist as D
        e = self.e
        sumsq = torch.zeros((), dtype=torch.float32)
        for u in e.units:
            sumsq = sumsq + e.be.sumsq(u.grad_shard).float()
        D.all_reduce_sum(sumsq, e.mesh.shard_group)
        clip, gnorm = clip_factor_from_sumsq(sumsq)
        sc = step_scalars(e.count, e.cfg)
        for u in e.units:
            e.be.adamw_step(u.master, u.exp_avg, u.exp_avg_sq, u.grad_shard, u.lp_shard, lr=sc.lr, b1=sc.b1, b2=sc.b2, eps=sc.eps,
                            wd_over_lr_peak=sc.wd_over_lr_peak, bc1=sc.bc1, bc2=sc.bc2, grad_scale=clip)
            u.grad_full.zero_()
        return gnorm


def _carrying_backend(base, fused):
    """TorchBackend whose GEMM entry points honour ``gather=`` like B200Backend._gemm: the first call that receives a pending
    spec carries it."""
    class Carrying(type(base)):
        def _take(self, spec, by):
            if spec is not None and not spec.done:
                fused.carry(spec, by)

        def linear_fwd(self, x, w, res=None, act=None, gather=None):
            self._take(gather, "gelu_fwd" if act == "gelu" else "fwd")
            return super().linear_fwd(x, w, res=res, act=act)

        def linear_wgrad(self, dy, x, dw_acc, gather=None):
            self._take(gather, "wgrad")
            return super().linear_wgrad(dy, x, dw_acc)

        def lmhead_ce(self, h, w, targets, dw_acc=None, grad_scale=0.0, chunk=8192, gather=None, prefetch=None):
            self._take(gather, "head_fwd")
            self._take(prefetch, "head_fwd_prefetch")
            return super().lmhead_ce(h, w, targets, dw_acc, grad_scale, chunk)

    base.__class__ = Carrying
    return base


def _run(rank, world, n_steps, remat):
    from midgpt_b200.train import Trainer
    torch.manual_seed(0)
    cfg = _tiny_cfg(True, "layer")
    cfg.engine.remat = remat
    tr = Trainer(cfg, device_type="cpu")
    eng = tr.engine
    assert eng.gather_mode == "layer"
    fused = FakeFused(eng)
    eng.fused = fused
    _carrying_backend(tr.backend, fused)
    losses, evals = [], []
    n_units = len(eng.units)
    for i, (x, y) in enumerate(_batches(cfg, n_steps)):
        G, B, T = cfg.g_accum_iters, cfg.batch_size, cfg.model_config.block_size
        xs = x.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        ys = y.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        if i == 1:                                   # an evaluation between two training steps (no backward follows the head)
            evals.append(tr.eval_step(xs[:1], ys[:1]).item())
            assert len(eng._pool_free) == len(eng._pool) and not eng._open_specs, "evaluation leaked a pool slot / an open gather"
        fused.carried.clear()
        losses.append(tr.train_step(xs, ys).item())
        assert len(eng._pool_free) == len(eng._pool) and not eng._open_specs, "a training step leaked a pool slot / an open gather"
        # per micro-step: blocks + head gathered in forward, blocks again in backward; every gather carried exactly once
        n_micro = tr.n_micro
        assert len(fused.carried) == n_micro * ((n_units - 1) + (n_units - 2)), fused.carried
    kinds = {by for _, by in fused.carried}
    # forward: first block self-carried by its QKV GEMM, the others prefetched by the previous block's MLP-in GEMM, the last block
    # (for the backward) by the LM-head GEMM; backward: prefetched by the recomputed MLP-in GEMM (remat) or the first wgrad GEMM
    assert {"fwd", "gelu_fwd", "head_fwd_prefetch"} <= kinds and (("wgrad" in kinds) == (remat == "none")), kinds
    state = {k: v.clone() for k, v in eng.full_state("master").items()}
    return losses, state, evals


import pytest


@pytest.mark.parametrize("remat", ["block", "none"])
def test_carried_gather_protocol_equals_single_process(remat):
    from tests.test_fsdp_cpu import _run as ref_run
    n_steps = 3
    ref_losses, ref_state, _ = ref_run(0, 1, False, "step", n_steps)
    outs = run_distributed(_run, 2, n_steps, remat)
    for losses, state, evals in outs:
        assert torch.allclose(torch.tensor(losses), torch.tensor(ref_losses), rtol=1e-5, atol=1e-6), (losses, ref_losses)
        for k in ref_state:
            assert torch.allclose(torch.as_tensor(state[k]), ref_state[k], rtol=1e-4, atol=1e-6), k
        assert all(e == e for e in evals)
