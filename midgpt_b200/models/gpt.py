"""Explicit forward / backward GPT over a primitive-op backend.

There is no autograd here: ``forward_backward`` walks the layers forward, then backward, calling
the backend's primitives (``ops/torch_backend.py`` on CPU / library baseline,
``ops/b200_backend.py`` for the hand-written sm_100a kernels).  This is the B200-first
replacement for the reference's ``jax.value_and_grad(loss_fn)`` over a ``lax.scan`` of
``jax.checkpoint``-ed blocks (src/train.py:86, src/model.py:149-155):

* per-block activation recompute is explicit (``remat="block"`` keeps only each block's input;
  ``remat="none"`` keeps all intermediates — affordable with 180 GB of HBM);
* weights are obtained from a *weight provider* (the FSDP engine) unit by unit so that
  all-gathers can be prefetched, and weight gradients are accumulated straight into the
  engine's fp32 accumulators, which is where the reduce-scatter ⊕ AdamW path picks them up;
* the LM head and the cross-entropy are one fused primitive, so ``[M, V]`` logits never exist.

Tensor convention: activations are ``[M, ·]`` with ``M = B·T`` (token-major).
"""
from __future__ import annotations

import typing as tp
from contextlib import nullcontext

import torch

from ..config import GPTConfig

W = tp.Dict[str, torch.Tensor]


class WeightProvider(tp.Protocol):
    """What the model needs from the parallel engine.  Unit ids: 0 = embed, 1..L = blocks, L+1 = head."""

    def acquire(self, unit: int) -> W: ...          # compute-dtype weights of a unit (gathered)
    def release(self, unit: int) -> None: ...       # the caller is done with them (for now)
    def grads(self, unit: int) -> W: ...            # fp32 gradient accumulators (same names/shapes)
    def grads_ready(self, unit: int) -> None: ...   # all contributions of this micro-step are in


class DictProvider:
    """Trivial provider over un-sharded ``{"embed.wte", "block.l.x", "head.lm_head"}`` dicts."""

    def __init__(self, cfg: GPTConfig, params: W, grads: W | None = None):
        self.cfg, self.params, self._grads = cfg, params, grads
        self.ready_order: tp.List[int] = []

    def _prefix(self, unit: int) -> str:
        L = self.cfg.n_layer
        return "embed." if unit == 0 else ("head." if unit == L + 1 else f"block.{unit - 1}.")

    def _sub(self, d: W, unit: int) -> W:
        p = self._prefix(unit)
        return {k[len(p):]: v for k, v in d.items() if k.startswith(p)}

    def acquire(self, unit): return self._sub(self.params, unit)
    def release(self, unit): pass
    def grads(self, unit): return self._sub(self._grads, unit)
    def grads_ready(self, unit): self.ready_order.append(unit)


def _carry(w: W):
    """The all-gather a GEMM of this unit should carry next (ZeRO-3 layer mode, all-gather ⊕ GEMM): the unit's own gather if no
    kernel has taken it yet, else the prefetch of the next unit of the schedule; None when there is nothing (left) to carry."""
    for key in ("__gather__", "__prefetch__"):
        spec = w.get(key)
        if spec is not None and not spec.done:
            return spec
    return None


def _site_seed(step_seed: int, micro: int, layer: int, site: int) -> int:
    return (step_seed * 1000003 + micro * 7919 + layer * 31 + site) & 0x7FFFFFFFFFFFFFFF


class GPTModel:
    def __init__(self, cfg: GPTConfig, backend, remat: str = "block", nvtx: bool = False):
        assert remat in ("block", "none")
        self.cfg, self.be, self.remat, self.nvtx = cfg, backend, remat, nvtx

    # NVTX ranges use the reference's jax.named_scope names (src/model.py:28,55,97,140).
    def _range(self, name):
        if self.nvtx and torch.cuda.is_available():
            return torch.autograd.profiler.record_function(name)
        return nullcontext()

    # ------------------------------------------------------------------ block forward
    def _block_fwd(self, x, w: W, B, T, layer, step_seed, micro, training, save: dict | None):
        cfg, be = self.cfg, self.be
        H, C = cfg.n_head, cfg.head_dim
        p = cfg.dropout if training else 0.0
        with self._range("block"):
            h1, rstd1 = be.rmsnorm_fwd(x, 1e-6)
            with self._range("causal_sa"):
                qkv_raw = be.linear_fwd(h1, w["c_attn"], gather=w.get("__gather__"))
                qkv, stats = be.qkln_rope_fwd(qkv_raw, w["q_ln"], w["k_ln"], B, T, H, C)
                o, lse = be.attention_fwd(qkv, B, T, H, C, p, _site_seed(step_seed, micro, layer, 1))
                if p == 0.0:
                    x2 = be.linear_fwd(o, w["attn_proj"], res=x)
                else:
                    x2 = x + be.dropout(be.linear_fwd(o, w["attn_proj"]), p, _site_seed(step_seed, micro, layer, 2))
            h2, rstd2 = be.rmsnorm_fwd(x2, 1e-6)
            with self._range("mlp"):
                pre, a = be.linear_fwd(h2, w["c_fc"], act="gelu", gather=_carry(w))     # MLP-in GEMM: carries the NEXT unit's gather
                if p == 0.0:
                    y = be.linear_fwd(a, w["mlp_proj"], res=x2)
                else:
                    y = x2 + be.dropout(be.linear_fwd(a, w["mlp_proj"]), p, _site_seed(step_seed, micro, layer, 3))
        if save is not None:
            save.update(x=x, h1=h1, rstd1=rstd1, qkv_raw=qkv_raw, qkv=qkv, stats=stats, o=o, lse=lse,
                        x2=x2, h2=h2, rstd2=rstd2, pre=pre, a=a)
        return y

    # ------------------------------------------------------------------ block backward
    def _block_bwd(self, dy, s: dict, w: W, g: W, B, T, layer, step_seed, micro, training):
        cfg, be = self.cfg, self.be
        H, C = cfg.n_head, cfg.head_dim
        p = cfg.dropout if training else 0.0
        dmlp = dy if p == 0.0 else be.dropout(dy, p, _site_seed(step_seed, micro, layer, 3))
        # (ZeRO-3 / layer gathers without remat: this first GEMM of the block's backward carries the unit's all-gather)
        be.linear_wgrad(dmlp, s["a"], g["mlp_proj"], gather=_carry(w))
        dpre = be.linear_dgrad(dmlp, w["mlp_proj"], gelu_pre=s["pre"])
        be.linear_wgrad(dpre, s["h2"], g["c_fc"], gather=_carry(w))
        dh2 = be.linear_dgrad(dpre, w["c_fc"])
        dx2 = be.rmsnorm_bwd(dh2, s["x2"], s["rstd2"], dres=dy)
        dproj = dx2 if p == 0.0 else be.dropout(dx2, p, _site_seed(step_seed, micro, layer, 2))
        be.linear_wgrad(dproj, s["o"], g["attn_proj"])
        do = be.linear_dgrad(dproj, w["attn_proj"])
        # fuse_dq: dq stays in the kernel's fp32 accumulator and is consumed (and cleared) by qkln_rope_bwd right below
        dqkv = be.attention_bwd(do, s["qkv"], s["o"], s["lse"], B, T, H, C, p, _site_seed(step_seed, micro, layer, 1), fuse_dq=True)
        dqkv_raw = be.qkln_rope_bwd(dqkv, s["qkv_raw"], s["stats"], w["q_ln"], w["k_ln"],
                                    g["q_ln"], g["k_ln"], B, T, H, C)
        be.linear_wgrad(dqkv_raw, s["h1"], g["c_attn"])
        dh1 = be.linear_dgrad(dqkv_raw, w["c_attn"])
        return be.rmsnorm_bwd(dh1, s["x"], s["rstd1"], dres=dx2)

    # ------------------------------------------------------------------ public API
    def forward_backward(self, wp: WeightProvider, idx: torch.Tensor, targets: torch.Tensor,
                         grad_scale: float, step_seed: int = 0, micro: int = 0) -> torch.Tensor:
        """One micro-step: accumulates ``grad_scale · ∂(Σ_token CE)/∂θ`` into ``wp.grads(·)``.

        Returns the *sum* of token losses (fp32 scalar tensor on device; no host sync).
        """
        cfg, be = self.cfg, self.be
        B, T = idx.shape
        L = cfg.n_layer
        p = cfg.dropout
        with self._range("gpt"):
            w0 = wp.acquire(0)
            x = be.embed_fwd(w0["wte"], idx)
            if p > 0.0:
                x = be.dropout(x, p, _site_seed(step_seed, micro, 0, 0))
            wp.release(0)
            saved: tp.List[tp.Any] = []
            for l in range(L):
                w = wp.acquire(l + 1)
                if self.remat == "block":
                    saved.append(x)
                    x = self._block_fwd(x, w, B, T, l + 1, step_seed, micro, True, None)
                else:
                    s: dict = {}
                    x = self._block_fwd(x, w, B, T, l + 1, step_seed, micro, True, s)
                    saved.append(s)
                wp.release(l + 1)
            # final norm + fused LM head / CE (forward and backward in one pass)
            wh = wp.acquire(L + 1)
            hf, rstdf = be.rmsnorm_fwd(x, 1e-5)
            loss_sum, dhf = be.lmhead_ce(hf, wh["lm_head"], targets, wp.grads(L + 1)["lm_head"], grad_scale, gather=wh.get("__gather__"),
                                         prefetch=wh.get("__prefetch__"))
            wp.release(L + 1)
            wp.grads_ready(L + 1)
            dx = be.rmsnorm_bwd(dhf, x, rstdf)
            del hf, dhf, x
            for l in reversed(range(L)):
                w = wp.acquire(l + 1)
                if self.remat == "block":
                    s = {}
                    self._block_fwd(saved[l], w, B, T, l + 1, step_seed, micro, True, s)
                else:
                    s = saved[l]
                saved[l] = None
                dx = self._block_bwd(dx, s, w, wp.grads(l + 1), B, T, l + 1, step_seed, micro, True)
                del s
                wp.release(l + 1)
                wp.grads_ready(l + 1)
            if p > 0.0:
                dx = be.dropout(dx, p, _site_seed(step_seed, micro, 0, 0))
            be.embed_bwd(dx, idx, wp.grads(0)["wte"])
            wp.grads_ready(0)
        return loss_sum

    def forward_hidden(self, wp: WeightProvider, idx: torch.Tensor) -> torch.Tensor:
        """Inference-mode forward to the final normed hidden state [M, D] (K19/K20)."""
        cfg, be = self.cfg, self.be
        B, T = idx.shape
        w0 = wp.acquire(0)
        x = be.embed_fwd(w0["wte"], idx)
        wp.release(0)
        for l in range(cfg.n_layer):
            w = wp.acquire(l + 1)
            x = self._block_fwd(x, w, B, T, l + 1, 0, 0, False, None)
            wp.release(l + 1)
        hf, _ = be.rmsnorm_fwd(x, 1e-5)
        return hf

    def forward_loss(self, wp: WeightProvider, idx: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """Mean token CE in eval mode (reference ``simple_loss``, src/train.py:99-103)."""
        hf = self.forward_hidden(wp, idx)
        L = self.cfg.n_layer
        wh = wp.acquire(L + 1)
        loss_sum, _ = self.be.lmhead_ce(hf, wh["lm_head"], targets, None, 0.0, gather=wh.get("__gather__"))
        wp.release(L + 1)
        return loss_sum / idx.numel()

    def forward_logits(self, wp: WeightProvider, idx: torch.Tensor, positions: torch.Tensor | None = None):
        """Full logits [B, T, V] (or only at ``positions`` [B]) — used by sample.py."""
        hf = self.forward_hidden(wp, idx)
        B, T = idx.shape
        L = self.cfg.n_layer
        wh = wp.acquire(L + 1)
        if positions is not None:
            rows = torch.arange(B, device=idx.device) * T + positions.to(idx.device)
            hf = hf[rows]
        logits = self.be.lmhead_logits(hf, wh["lm_head"], gather=wh.get("__gather__"))
        wp.release(L + 1)
        return logits if positions is not None else logits.view(B, T, -1)
