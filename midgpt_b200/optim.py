"""Learning-rate schedule and optimizer algebra (K15–K18 of SURVEY.md §2.4).

The reference builds ``optax.chain(clip_by_global_norm(1.0), scale_by_adam(b2=β₂),
add_decayed_weights(wd / lr_peak), scale_by_schedule(warmup_cosine), scale(-1))``
(src/train.py:147-159).  Collapsed into one per-element update that the fused kernels implement:

    g'   = g · min(1, 1/‖g‖₂)                       (global norm over *all* parameters)
    m    ← β₁ m + (1-β₁) g' ;  v ← β₂ v + (1-β₂) g'²
    u    = (m / (1-β₁ᵗ)) / (√(v / (1-β₂ᵗ)) + ε) + (wd / lr_peak) · p          t = count + 1
    p    ← p − lr(count) · u                                                  count = steps done so far

Notes that are easy to get wrong (all verified in tests/test_optim.py):
  * the schedule is evaluated at the *pre-increment* count ⇒ the very first step uses lr = 0;
  * weight decay is "independent": after LR scaling the decay is ``lr_t/lr_peak · wd · p``;
  * weight decay applies to every tensor including embeddings and LayerNorm gains (no mask);
  * ε is added outside the square root, bias correction uses the incremented count.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

BETA1 = 0.9
ADAM_EPS = 1e-8
CLIP_NORM = 1.0


def warmup_cosine_lr(count: int, peak: float, warmup: int, decay_steps: int, end: float) -> float:
    """optax.warmup_cosine_decay_schedule(0, peak, warmup, decay_steps, end) evaluated at ``count``."""
    if count < warmup:
        return peak * count / max(warmup, 1)
    span = max(decay_steps - warmup, 1)
    c = min(count - warmup, span)
    cos = 0.5 * (1.0 + math.cos(math.pi * c / span))
    alpha = end / peak
    return peak * ((1.0 - alpha) * cos + alpha)


@dataclass
class StepScalars:
    """Host-side scalars for one optimizer step (everything except the clip factor, which stays on device)."""
    lr: float
    bc1: float       # 1 - β₁ᵗ
    bc2: float       # 1 - β₂ᵗ
    wd_over_lr_peak: float
    b1: float
    b2: float
    eps: float


def step_scalars(count: int, cfg) -> StepScalars:
    t = count + 1
    return StepScalars(
        lr=warmup_cosine_lr(count, cfg.learning_rate, cfg.warmup_steps, cfg.lr_decay_steps, cfg.min_lr),
        bc1=1.0 - BETA1 ** t, bc2=1.0 - cfg.beta2 ** t,
        wd_over_lr_peak=cfg.weight_decay / cfg.learning_rate,
        b1=BETA1, b2=cfg.beta2, eps=ADAM_EPS,
    )


def clip_factor_from_sumsq(total_sumsq, max_norm: float = CLIP_NORM):
    """Device-side: min(1, max_norm/‖g‖) with optax's ``g_norm < max_norm`` trigger; no host sync."""
    import torch
    norm = total_sumsq.sqrt()
    return torch.where(norm < max_norm, torch.ones_like(norm), max_norm / norm), norm
