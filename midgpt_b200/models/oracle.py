"""Pure-PyTorch oracle of the model semantics (numerical ground truth, CPU-capable).

Every hand-written kernel and the explicit forward/backward engine in ``gpt.py`` is tested
against this file.  Semantics follow SURVEY.md §2.4 with citations into the reference:

* batch-major ``[B, T]`` tokens instead of the reference's per-example ``vmap`` (src/train.py:76)
* weightless RMSNorm, eps 1e-6 in blocks, 1e-5 for ln_f      (src/layers.py:70-75, src/model.py:94-95,133)
* QK-LayerNorm: mean-subtracting LN over head dim, gain only   (src/model.py:52-53,64-65)
* interleaved (GPT-J) RoPE, after QK-LN, on Q and K            (src/layers.py:79-99, src/model.py:67-69)
* fp32 softmax of ``QKᵀ/√C`` with a causal ``-inf`` mask          (src/model.py:71-77)
* tanh-GELU 4× MLP, no biases anywhere                          (src/model.py:24-31, src/layers.py:37-57)
* logits → fp32 → mean cross-entropy                            (src/train.py:76-77)
* dropout at embedding / attention probs / both residual branches (src/model.py:31,78,80,147)

Norm statistics are always computed in fp32 (the reference computes RMSNorm in the activation
dtype; fp32 statistics are a strict accuracy improvement and what the kernels do).
"""
from __future__ import annotations

import math
import typing as tp

import torch
import torch.nn.functional as F

from ..config import GPTConfig

Params = tp.Dict[str, torch.Tensor]


def _f(x: torch.Tensor) -> torch.Tensor:
    """fp32 statistics; fp64 passes through so gradient checks can run in double."""
    return x if x.dtype == torch.float64 else x.float()


def rmsnorm(x: torch.Tensor, eps: float) -> torch.Tensor:
    xf = _f(x)
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)


def qk_layernorm(x: torch.Tensor, gain: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm over the last (head) dim with gain, no bias: (x-mean)*rsqrt(var+eps)*g."""
    xf = _f(x)
    mu = xf.mean(-1, keepdim=True)
    var = (xf - mu).pow(2).mean(-1, keepdim=True)
    return ((xf - mu) * torch.rsqrt(var + eps) * _f(gain)).to(x.dtype)


def rope_tables(C: int, T: int, device=None, base: float = 10000.0, offset: int = 0):
    """sin/cos of shape [T, C/2] (fp64; cast by the user) — inv_freq_i = base^(-2i/C), positions offset..offset+T-1."""
    inv_freq = 1.0 / (base ** (torch.arange(0, C, 2, dtype=torch.float64, device=device) / C))
    ang = torch.outer(torch.arange(offset, offset + T, dtype=torch.float64, device=device), inv_freq)
    return ang.sin(), ang.cos()


def apply_rope(x: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor) -> torch.Tensor:
    """Interleaved-pair rotation on the last dim of x[..., T, C]: (a,b) -> (a·cos − b·sin, b·cos + a·sin).

    The reference casts sin/cos to the activation dtype before multiplying (src/layers.py:93-94);
    here the rotation is done in fp32 and rounded once.
    """
    xf = _f(x)
    sin, cos = sin.to(xf.dtype), cos.to(xf.dtype)
    a, b = xf[..., 0::2], xf[..., 1::2]
    out = torch.stack((a * cos - b * sin, b * cos + a * sin), dim=-1)
    return out.flatten(-2).to(x.dtype)


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return F.gelu(x, approximate="tanh")


def _dropout(x, p, training, gen):
    if not training or p == 0.0:
        return x
    keep = torch.rand(x.shape, device=x.device, generator=gen) >= p
    return torch.where(keep, x / (1.0 - p), torch.zeros_like(x))


def attention(q, k, v, dropout_p=0.0, training=False, gen=None):
    """q,k,v: [B,H,T,C] → [B,H,T,C]; fp32 softmax, causal."""
    C, T = q.shape[-1], q.shape[-2]
    s = _f(q @ k.transpose(-1, -2)) / math.sqrt(C)
    mask = torch.ones(T, T, dtype=torch.bool, device=q.device).tril_()
    s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1).to(q.dtype)
    p = _dropout(p, dropout_p, training, gen)
    return p @ v


def block_forward(x, w: Params, prefix: str, cfg: GPTConfig, sin, cos, training=False, gen=None):
    B, T, D = x.shape
    H, C = cfg.n_head, cfg.head_dim
    h = rmsnorm(x, 1e-6)
    qkv = h @ w[prefix + "c_attn"].t()
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, T, H, C).transpose(1, 2)
    k = k.view(B, T, H, C).transpose(1, 2)
    v = v.view(B, T, H, C).transpose(1, 2)
    q = apply_rope(qk_layernorm(q, w[prefix + "q_ln"]), sin, cos)
    k = apply_rope(qk_layernorm(k, w[prefix + "k_ln"]), sin, cos)
    o = attention(q, k, v, cfg.dropout, training, gen).transpose(1, 2).reshape(B, T, D)
    x = x + _dropout(o @ w[prefix + "attn_proj"].t(), cfg.dropout, training, gen)
    h = rmsnorm(x, 1e-6)
    a = gelu_tanh(h @ w[prefix + "c_fc"].t())
    return x + _dropout(a @ w[prefix + "mlp_proj"].t(), cfg.dropout, training, gen)


def forward_hidden(params: Params, idx: torch.Tensor, cfg: GPTConfig, training=False, gen=None):
    """idx [B,T] int → final normed hidden [B,T,D]."""
    B, T = idx.shape
    x = params["embed.wte"][idx]
    x = _dropout(x, cfg.dropout, training, gen)
    sin, cos = rope_tables(cfg.head_dim, T, device=idx.device)
    for l in range(cfg.n_layer):
        x = block_forward(x, params, f"block.{l}.", cfg, sin, cos, training, gen)
    return rmsnorm(x, 1e-5)


def forward_logits(params: Params, idx: torch.Tensor, cfg: GPTConfig, training=False, gen=None):
    return forward_hidden(params, idx, cfg, training, gen) @ params["head.lm_head"].t()


def loss_fn(params: Params, idx: torch.Tensor, targets: torch.Tensor, cfg: GPTConfig,
            training=False, gen=None) -> torch.Tensor:
    logits = _f(forward_logits(params, idx, cfg, training, gen))
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), targets.reshape(-1).long())


def cast_params(params: Params, dtype: torch.dtype) -> Params:
    """reference ``cast_pytree`` (src/train.py:47-53)."""
    return {k: v.to(dtype) for k, v in params.items()}
