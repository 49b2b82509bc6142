"""Library/PyTorch implementation of the primitive-op set (forward *and* hand-derived backward).

This backend serves three purposes:
  1. the CPU path (tests, gloo world_size=2 plumbing config — BASELINE.json config 1);
  2. the "library baseline" on GPU (cuBLAS GEMMs + aten flash attention) that the hand-written
     sm_100a backend in ``b200_backend.py`` must beat;
  3. an executable specification of each primitive's backward, validated against autograd of
     ``models/oracle.py`` in ``tests/test_model_backward.py``.

All activations are token-major 2-D tensors ``[M = B·T, features]``.
Op ↔ SURVEY.md §2.4 mapping: embed K1, rmsnorm K2, linear K3/K7/K8/K9, qkln_rope K4+K5,
attention K6, lmhead_ce K10+K11, dropout K12.
"""
from __future__ import annotations

import math
import typing as tp

import torch
import torch.nn.functional as F

_SQRT_2_OVER_PI = math.sqrt(2.0 / math.pi)


def _f(x: torch.Tensor) -> torch.Tensor:
    """Accumulation dtype: fp32, or fp64 when the caller already runs in fp64 (gradient-check tests)."""
    return x if x.dtype == torch.float64 else x.float()


def _acc(x: torch.Tensor) -> torch.dtype:
    return torch.float64 if x.dtype == torch.float64 else torch.float32


def _gelu_grad(pre: torch.Tensor) -> torch.Tensor:
    """d/dx tanh-GELU, fp32."""
    x = _f(pre)
    u = _SQRT_2_OVER_PI * (x + 0.044715 * x * x * x)
    t = torch.tanh(u)
    du = _SQRT_2_OVER_PI * (1.0 + 3 * 0.044715 * x * x)
    return 0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * du


class TorchBackend:
    name = "torch"

    def __init__(self, use_aten_flash: bool = True):
        self.use_aten_flash = use_aten_flash
        self._rope_cache: tp.Dict[tp.Tuple, tp.Tuple[torch.Tensor, torch.Tensor]] = {}

    # ---------------------------------------------------------------- embedding (K1)
    def embed_fwd(self, wte: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        return wte[idx.reshape(-1).long()]

    def embed_bwd(self, dx: torch.Tensor, idx: torch.Tensor, dwte_acc: torch.Tensor) -> None:
        dwte_acc.index_add_(0, idx.reshape(-1).long(), dx.to(dwte_acc.dtype))

    # ---------------------------------------------------------------- dropout (K12)
    def dropout_mask(self, shape, p: float, seed: int, device) -> torch.Tensor:
        g = torch.Generator(device=device)
        g.manual_seed(seed & 0x7FFFFFFFFFFFFFFF)
        return torch.rand(shape, device=device, generator=g) >= p

    def dropout(self, x: torch.Tensor, p: float, seed: int) -> torch.Tensor:
        """Forward and backward are the same function of (tensor, p, seed): y = x·mask/(1-p)."""
        if p == 0.0:
            return x
        keep = self.dropout_mask(x.shape, p, seed, x.device)
        return torch.where(keep, x / (1.0 - p), torch.zeros_like(x))

    # ---------------------------------------------------------------- RMSNorm (K2)
    def rmsnorm_fwd(self, x: torch.Tensor, eps: float):
        xf = _f(x)
        rstd = torch.rsqrt(xf.pow(2).mean(-1) + eps)
        return (xf * rstd[:, None]).to(x.dtype), rstd

    def rmsnorm_bwd(self, dy, x, rstd, dres=None):
        """dx = rstd·(dy − ŷ·mean(dy·ŷ)) [+ dres] with ŷ = x·rstd."""
        yf = _f(x) * rstd[:, None]
        dyf = _f(dy)
        dx = rstd[:, None] * (dyf - yf * (dyf * yf).mean(-1, keepdim=True))
        if dres is not None:
            dx = dx + _f(dres)
        return dx.to(x.dtype)

    # ---------------------------------------------------------------- linear (K3,K7,K8,K9)
    def linear_fwd(self, x, w, res=None, act: str | None = None, gather=None):
        """y = x @ wᵀ (+res).  act == "gelu" → returns (pre, gelu(pre))."""
        y = x @ w.t()
        if act == "gelu":
            return y, F.gelu(_f(y), approximate="tanh").to(y.dtype)
        if res is not None:
            y = y + res
        return y

    def linear_dgrad(self, dy, w, gelu_pre=None, dres=None, gather=None):
        """dx = dy @ w, optionally ·gelu'(pre) (K8 backward) and/or + dres."""
        dx = dy @ w
        if gelu_pre is not None:
            dx = (_f(dx) * _gelu_grad(gelu_pre)).to(dy.dtype)
        if dres is not None:
            dx = dx + dres
        return dx

    def linear_wgrad(self, dy, x, dw_acc: torch.Tensor, gather=None) -> None:
        """dw_acc (fp32 or wider) += dyᵀ @ x."""
        if dw_acc.dtype == dy.dtype:
            dw_acc.addmm_(dy.t(), x)
        else:
            dw_acc.add_((dy.t() @ x).to(dw_acc.dtype))

    # ---------------------------------------------------------------- QK-LN + RoPE (K4,K5)
    def _rope(self, C, T, device):
        key = (C, T, str(device))
        if key not in self._rope_cache:
            inv = 1.0 / (10000.0 ** (torch.arange(0, C, 2, dtype=torch.float64, device=device) / C))
            ang = torch.outer(torch.arange(T, dtype=torch.float64, device=device), inv)
            self._rope_cache[key] = (ang.sin(), ang.cos())
        return self._rope_cache[key]

    def qkln_rope_fwd(self, qkv, q_gain, k_gain, B, T, H, C, eps=1e-6):
        """qkv [M,3D] → (qkv' [M,3D] with LN+RoPE applied to the q and k thirds, stats [M,2,H,2] fp32)."""
        M, D = B * T, H * C
        x = qkv.view(B, T, 3, H, C)
        qk = _f(x[:, :, :2])                                   # [B,T,2,H,C]
        mu = qk.mean(-1, keepdim=True)
        rstd = torch.rsqrt((qk - mu).pow(2).mean(-1, keepdim=True) + eps)
        gains = torch.stack((_f(q_gain), _f(k_gain)), 0).view(1, 1, 2, 1, C)
        n = (qk - mu) * rstd * gains
        sin, cos = (t.to(qk.dtype) for t in self._rope(C, T, qkv.device))
        sin, cos = sin.view(1, T, 1, 1, C // 2), cos.view(1, T, 1, 1, C // 2)
        a, b = n[..., 0::2], n[..., 1::2]
        r = torch.stack((a * cos - b * sin, b * cos + a * sin), -1).flatten(-2)
        out = torch.empty_like(qkv).view(B, T, 3, H, C)
        out[:, :, :2] = r.to(qkv.dtype)
        out[:, :, 2] = x[:, :, 2]
        stats = torch.cat((mu, rstd), -1).reshape(M, 2, H, 2)
        return out.view(M, 3 * D), stats

    def qkln_rope_bwd(self, dqkv, qkv_raw, stats, q_gain, k_gain, dq_gain_acc, dk_gain_acc, B, T, H, C):
        """Gradient w.r.t. the raw projection output; accumulates the two gain grads (fp32)."""
        M, D = B * T, H * C
        d = dqkv.view(B, T, 3, H, C)
        x = qkv_raw.view(B, T, 3, H, C)
        st = stats.view(B, T, 2, H, 2)
        mu, rstd = st[..., 0:1], st[..., 1:2]
        sin, cos = (t.to(_acc(dqkv)) for t in self._rope(C, T, dqkv.device))
        sin, cos = sin.view(1, T, 1, 1, C // 2), cos.view(1, T, 1, 1, C // 2)
        g = _f(d[:, :, :2])
        da, db = g[..., 0::2], g[..., 1::2]
        dn = torch.stack((da * cos + db * sin, db * cos - da * sin), -1).flatten(-2)   # inverse rotation
        xhat = (_f(x[:, :, :2]) - mu) * rstd
        gains = torch.stack((_f(q_gain), _f(k_gain)), 0).view(1, 1, 2, 1, C)
        dgain = (dn * xhat).sum((0, 1, 3))                          # [2, C]
        dq_gain_acc.add_(dgain[0].to(dq_gain_acc.dtype))
        dk_gain_acc.add_(dgain[1].to(dk_gain_acc.dtype))
        dxhat = dn * gains
        dx = rstd * (dxhat - dxhat.mean(-1, keepdim=True) - xhat * (dxhat * xhat).mean(-1, keepdim=True))
        out = torch.empty_like(dqkv).view(B, T, 3, H, C)
        out[:, :, :2] = dx.to(dqkv.dtype)
        out[:, :, 2] = d[:, :, 2]
        return out.view(M, 3 * D)

    # ---------------------------------------------------------------- attention (K6)
    def attention_fwd(self, qkv, B, T, H, C, dropout_p=0.0, seed=0):
        """causal softmax(QKᵀ/√C)·V on the packed qkv [M,3D] → (o [M,D], lse [B,H,T] fp32 natural-log)."""
        x = qkv.view(B, T, 3, H, C)
        if self.use_aten_flash and qkv.is_cuda and dropout_p == 0.0 and qkv.dtype in (torch.bfloat16, torch.float16):
            q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
            out = torch.ops.aten._scaled_dot_product_flash_attention(q, k, v, 0.0, True, False)
            o, lse = out[0], out[1]
            self._flash_ctx = out  # philox seed/offset etc. for the aten backward
            return o.transpose(1, 2).reshape(B * T, H * C), lse
        q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        s = (_f(q) @ _f(k).transpose(-1, -2)) / math.sqrt(C)
        causal = torch.ones(T, T, dtype=torch.bool, device=qkv.device).tril_()
        s = s.masked_fill(~causal, float("-inf"))
        lse = torch.logsumexp(s, -1)
        p = torch.exp(s - lse[..., None]).to(qkv.dtype)
        if dropout_p > 0.0:
            p = self.dropout(p, dropout_p, seed)
        o = (p @ v).permute(0, 2, 1, 3).reshape(B * T, H * C)
        return o, lse

    def attention_bwd(self, do, qkv, o, lse, B, T, H, C, dropout_p=0.0, seed=0, fuse_dq: bool = False):
        x = qkv.view(B, T, 3, H, C)
        if self.use_aten_flash and qkv.is_cuda and dropout_p == 0.0 and qkv.dtype in (torch.bfloat16, torch.float16):
            q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
            ctx = self._flash_ctx
            o4 = o.view(B, T, H, C).transpose(1, 2)
            do4 = do.view(B, T, H, C).transpose(1, 2)
            dq, dk, dv = torch.ops.aten._scaled_dot_product_flash_attention_backward(
                do4, q, k, v, o4, lse, ctx[2], ctx[3], ctx[4], ctx[5], 0.0, True, ctx[6], ctx[7])
            out = torch.empty_like(qkv).view(B, T, 3, H, C)
            out[:, :, 0], out[:, :, 1], out[:, :, 2] = dq.transpose(1, 2), dk.transpose(1, 2), dv.transpose(1, 2)
            return out.view(B * T, 3 * H * C)
        q, k, v = (_f(x[:, :, i].permute(0, 2, 1, 3)) for i in range(3))
        scale = 1.0 / math.sqrt(C)
        s = (q @ k.transpose(-1, -2)) * scale
        causal = torch.ones(T, T, dtype=torch.bool, device=qkv.device).tril_()
        s = s.masked_fill(~causal, float("-inf"))
        p = torch.exp(s - lse[..., None])
        do4 = _f(do.view(B, T, H, C).permute(0, 2, 1, 3))
        o4 = _f(o.view(B, T, H, C).permute(0, 2, 1, 3))
        pd = _f(p.to(qkv.dtype))
        if dropout_p > 0.0:
            keep = self.dropout_mask(p.shape, dropout_p, seed, p.device).to(p.dtype) / (1.0 - dropout_p)
            pd = pd * keep
        dv = pd.transpose(-1, -2) @ do4
        dp = do4 @ v.transpose(-1, -2)
        if dropout_p > 0.0:
            dp = dp * keep
        delta = (do4 * o4).sum(-1, keepdim=True)
        ds = p * (dp - delta) * scale
        dq = ds @ k
        dk = ds.transpose(-1, -2) @ q
        out = torch.empty_like(qkv).view(B, T, 3, H, C)
        out[:, :, 0] = dq.permute(0, 2, 1, 3).to(qkv.dtype)
        out[:, :, 1] = dk.permute(0, 2, 1, 3).to(qkv.dtype)
        out[:, :, 2] = dv.permute(0, 2, 1, 3).to(qkv.dtype)
        return out.view(B * T, 3 * H * C)

    # ---------------------------------------------------------------- LM head + cross-entropy (K10,K11)
    def lmhead_logits(self, h, w, gather=None):
        return h @ w.t()

    def lmhead_ce(self, h, w, targets, dw_acc=None, grad_scale: float = 0.0, chunk: int = 8192, gather=None, prefetch=None):
        """Σ_tokens CE(h·wᵀ, targets) in fp32 without materialising [M,V] at once.

        If ``dw_acc`` is given the backward is fused: returns (loss_sum, dh) and accumulates
        ``dw_acc += grad_scale · dlogitsᵀ h``; ``dh = grad_scale · dlogits · w``.
        """
        M = h.shape[0]
        targets = targets.reshape(-1).long()
        loss_sum = torch.zeros((), dtype=_acc(h), device=h.device)
        dh = torch.empty_like(h) if dw_acc is not None else None
        for s in range(0, M, chunk):
            e = min(M, s + chunk)
            logits = _f(h[s:e] @ w.t())
            lse = torch.logsumexp(logits, -1)
            tgt = targets[s:e]
            loss_sum += (lse - logits.gather(1, tgt[:, None]).squeeze(1)).sum()
            if dw_acc is not None:
                dl = torch.exp(logits - lse[:, None])
                dl[torch.arange(e - s, device=h.device), tgt] -= 1.0
                dl = (dl * grad_scale).to(h.dtype)
                dh[s:e] = dl @ w
                self.linear_wgrad(dl, h[s:e], dw_acc)
        return loss_sum, dh

    # ---------------------------------------------------------------- optimizer primitives (K13-K18)
    def sumsq(self, g: torch.Tensor) -> torch.Tensor:
        return _f(g).pow(2).sum()

    def adamw_step(self, p, m, v, g, p_lp, *, lr, b1, b2, eps, wd_over_lr_peak, bc1, bc2, grad_scale):
        """Fused K14-K18 on one fp32 shard (see optim.py for the algebra).  In-place.

        g_eff = g · grad_scale ; m,v Adam moments ; u = m̂/(√v̂+eps) + (wd/lr_peak)·p ; p -= lr·u ;
        p_lp (bf16 compute copy, may be None) ← p.
        """
        gf = _f(g) * grad_scale
        m.mul_(b1).add_(gf, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(gf, gf, value=1.0 - b2)
        u = (m / bc1) / ((v / bc2).sqrt() + eps) + wd_over_lr_peak * p
        p.add_(u, alpha=-lr)
        if p_lp is not None:
            p_lp.copy_(p)
