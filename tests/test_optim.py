"""Schedule + optimizer algebra vs closed forms (SURVEY.md §2.4 K15–K18, §7.4 item 7)."""
import math

import torch

from midgpt_b200 import load_config
from midgpt_b200.optim import warmup_cosine_lr, step_scalars, clip_factor_from_sumsq
from midgpt_b200.ops.torch_backend import TorchBackend


def test_schedule_shape():
    lr = lambda c: warmup_cosine_lr(c, 1e-3, 100, 5000, 1e-4)
    assert lr(0) == 0.0                                   # first step uses lr = 0
    assert math.isclose(lr(50), 0.5e-3)
    assert math.isclose(lr(100), 1e-3)
    mid = 100 + (5000 - 100) // 2
    assert math.isclose(lr(mid), 1e-4 + (1e-3 - 1e-4) * 0.5, rel_tol=1e-9)
    assert math.isclose(lr(5000), 1e-4)
    assert math.isclose(lr(10_000), 1e-4)                 # flat after decay
    assert all(lr(c + 1) <= lr(c) + 1e-15 for c in range(100, 5000, 37))


def test_clip_factor():
    f, n = clip_factor_from_sumsq(torch.tensor(0.25))
    assert f.item() == 1.0 and math.isclose(n.item(), 0.5)
    f, n = clip_factor_from_sumsq(torch.tensor(16.0))
    assert math.isclose(f.item(), 0.25)


def test_adamw_step_matches_optax_chain_algebra():
    cfg = load_config("shakespeare_char")
    torch.manual_seed(0)
    p = torch.randn(1000, dtype=torch.float64)
    g = torch.randn(1000, dtype=torch.float64)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    be = TorchBackend()
    p0 = p.clone()
    ref_m, ref_v, ref_p = torch.zeros_like(p), torch.zeros_like(p), p.clone()
    for count in range(0, 4):
        sc = step_scalars(count, cfg)
        clip = 0.7
        be.adamw_step(p, m, v, g, None, lr=sc.lr, b1=sc.b1, b2=sc.b2, eps=sc.eps,
                      wd_over_lr_peak=sc.wd_over_lr_peak, bc1=sc.bc1, bc2=sc.bc2, grad_scale=clip)
        # literal optax chain
        gg = g * clip
        ref_m = 0.9 * ref_m + 0.1 * gg
        ref_v = cfg.beta2 * ref_v + (1 - cfg.beta2) * gg * gg
        t = count + 1
        u = (ref_m / (1 - 0.9 ** t)) / (torch.sqrt(ref_v / (1 - cfg.beta2 ** t)) + 1e-8)
        u = u + (cfg.weight_decay / cfg.learning_rate) * ref_p
        ref_p = ref_p - warmup_cosine_lr(count, cfg.learning_rate, cfg.warmup_steps, cfg.lr_decay_steps, cfg.min_lr) * u
        assert torch.allclose(p, ref_p, rtol=1e-12, atol=1e-14)
        if count == 0:
            assert torch.equal(p, p0)                     # lr(0) == 0 ⇒ parameters unchanged by step 0
    # independent weight decay: with zero grads the decay per step is lr_t/lr_peak * wd * p
    p = torch.ones(4, dtype=torch.float64)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    sc = step_scalars(cfg.warmup_steps, cfg)              # lr == peak
    be.adamw_step(p, m, v, torch.zeros_like(p), None, lr=sc.lr, b1=sc.b1, b2=sc.b2, eps=sc.eps,
                  wd_over_lr_peak=sc.wd_over_lr_peak, bc1=sc.bc1, bc2=sc.bc2, grad_scale=1.0)
    assert torch.allclose(p, torch.full_like(p, 1 - cfg.weight_decay))


def test_step_watchdog_fires_on_a_hung_step_and_not_on_a_paused_one(tmp_path):
    """utils/obs.StepWatchdog: a step that does not complete within the limit dumps stacks and calls the timeout action (exit 3
    in production); evaluation / checkpoint pauses do not count."""
    import time
    from midgpt_b200.utils.obs import StepWatchdog
    fired = []
    wd = StepWatchdog(str(tmp_path), rank=0, factor=5.0, min_timeout_s=0.3, on_timeout=fired.append, poll_s=0.05)
    for _ in range(5):
        time.sleep(0.02)
        wd.beat()
    wd.beat(pause=True)
    time.sleep(0.6)                      # a long evaluation: disarmed
    assert not fired
    wd.beat()
    time.sleep(0.02)
    wd.beat()
    time.sleep(0.8)                      # a hung step
    assert fired and "no optimizer step completed" in fired[0]
    assert (tmp_path / "watchdog_rank0.txt").exists() and "Thread" in (tmp_path / "watchdog_rank0.txt").read_text()
    wd.close()
