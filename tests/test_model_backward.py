"""Explicit forward/backward engine (models/gpt.py + ops/torch_backend.py) vs autograd of the oracle."""
import pytest
import torch

from midgpt_b200.config import GPTConfig
from midgpt_b200.models import oracle, params as P
from midgpt_b200.models.gpt import GPTModel, DictProvider
from midgpt_b200.ops.torch_backend import TorchBackend


def _setup(dropout=0.0, dtype=torch.float64):
    torch.manual_seed(0)
    cfg = GPTConfig(block_size=16, vocab_size=37, n_layer=3, n_head=2, n_embd=16, dropout=dropout)
    prm = P.init_full_params(cfg, seed=1, dtype=dtype)
    # make the LN gains non-trivial so their gradient is exercised
    for l in range(cfg.n_layer):
        prm[f"block.{l}.q_ln"] = prm[f"block.{l}.q_ln"] + 0.1 * torch.randn(cfg.head_dim, dtype=dtype)
        prm[f"block.{l}.k_ln"] = prm[f"block.{l}.k_ln"] + 0.1 * torch.randn(cfg.head_dim, dtype=dtype)
    idx = torch.randint(0, cfg.vocab_size, (2, cfg.block_size))
    tgt = torch.randint(0, cfg.vocab_size, (2, cfg.block_size))
    return cfg, prm, idx, tgt


@pytest.mark.parametrize("remat", ["block", "none"])
def test_manual_backward_matches_autograd(remat):
    cfg, prm, idx, tgt = _setup()
    leaf = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    loss = oracle.loss_fn(leaf, idx, tgt, cfg)
    loss.backward()

    grads = {k: torch.zeros_like(v) for k, v in prm.items()}
    wp = DictProvider(cfg, prm, grads)
    model = GPTModel(cfg, TorchBackend(), remat=remat)
    loss_sum = model.forward_backward(wp, idx, tgt, grad_scale=1.0 / idx.numel())
    assert torch.allclose(loss_sum / idx.numel(), loss.detach(), rtol=1e-9, atol=1e-10)
    for k in prm:
        assert torch.allclose(grads[k], leaf[k].grad, rtol=1e-7, atol=1e-10), k
    # grads become ready head-first, embedding last
    assert wp.ready_order[0] == cfg.n_layer + 1 and wp.ready_order[-1] == 0


def test_eval_loss_and_logits_match_oracle():
    cfg, prm, idx, tgt = _setup()
    model = GPTModel(cfg, TorchBackend())
    wp = DictProvider(cfg, prm)
    tol = dict(rtol=1e-9, atol=1e-10)
    assert torch.allclose(model.forward_loss(wp, idx, tgt), oracle.loss_fn(prm, idx, tgt, cfg), **tol)
    assert torch.allclose(model.forward_logits(wp, idx), oracle.forward_logits(prm, idx, cfg), **tol)
    pos = torch.tensor([3, 7])
    lg = model.forward_logits(wp, idx, positions=pos)
    ref = oracle.forward_logits(prm, idx, cfg)[torch.arange(2), pos]
    assert torch.allclose(lg, ref, **tol)


def test_dropout_backward_is_consistent_with_finite_differences():
    """With dropout the masks are functions of (seed, site); check d loss/d θ numerically."""
    cfg, prm, idx, tgt = _setup(dropout=0.25)
    model = GPTModel(cfg, TorchBackend(), remat="block")
    grads = {k: torch.zeros_like(v) for k, v in prm.items()}
    model.forward_backward(DictProvider(cfg, prm, grads), idx, tgt, 1.0, step_seed=5, micro=1)

    def f(p):
        g0 = {k: torch.zeros_like(v) for k, v in p.items()}
        return model.forward_backward(DictProvider(cfg, p, g0), idx, tgt, 1.0, step_seed=5, micro=1)

    torch.manual_seed(1)
    for name in ["block.1.c_fc", "block.0.c_attn", "embed.wte", "head.lm_head", "block.2.q_ln"]:
        d = torch.randn_like(prm[name])
        eps = 1e-6
        pp = dict(prm); pp[name] = prm[name] + eps * d
        pm = dict(prm); pm[name] = prm[name] - eps * d
        fd = (f(pp) - f(pm)) / (2 * eps)
        an = (grads[name] * d).sum()
        assert torch.allclose(fd, an, rtol=1e-5, atol=1e-7), (name, fd.item(), an.item())


def test_count_params_matches_reference_numbers():
    from midgpt_b200 import load_config
    assert P.count_params(load_config("shakespeare_char").model_config) == 10_642_560
    assert P.count_params(load_config("openwebtext").model_config) == 123_569_664
    xl = load_config("openwebtext_xl").model_config
    assert P.count_params(xl) == 1_310_988_288
    assert P.stored_params(xl) == 1_414_010_880
    assert P.flops_per_token(xl) == 8_469_909_504
