"""Small-shape launches of EVERY hand-written kernel family, for compute-sanitizer (scripts/sanitize.sh runs this file under
memcheck / racecheck / synccheck / initcheck; shapes are tiny because the tools slow tcgen05/TMA kernels down by 100-1000x).
Not collected by the default test run (file name); every case still checks its numerics so a sanitizer run is also a test."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from midgpt_b200.ops.b200_backend import B200Backend
    return B200Backend(torch.device("cuda", 0))


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_gemm_pair_all_epilogues(be):
    """CTA-pair kernels (M >= 256): store, residual, gelu (register path, K > 1024 -> 6 stages; and TMA-box path, K <= 1024),
    dgrad, dgelu, wgrad with split-K."""
    for K in (256, 1152):
        M, N = 512, 512
        x, w, res = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(M, N)
        base = x.float() @ w.float().t()
        assert rel(be.linear_fwd(x, w), base) < 6e-3
        assert rel(be.linear_fwd(x, w, res=res), base + res.float()) < 6e-3
        pre, act = be.linear_fwd(x, w, act="gelu")
        assert rel(pre, base) < 6e-3 and rel(act, torch.nn.functional.gelu(pre.float(), approximate="tanh")) < 6e-3
    M, N, K = 512, 256, 384
    dy, w, x = rnd(M, N), rnd(N, K, scale=N ** -0.5), rnd(M, K)
    assert rel(be.linear_dgrad(dy, w), dy.float() @ w.float()) < 6e-3
    pre = rnd(M, K)
    u = 0.7978845608028654 * (pre.float() + 0.044715 * pre.float() ** 3)
    g = 0.5 * (1 + torch.tanh(u)) + 0.5 * pre.float() * (1 - torch.tanh(u) ** 2) * 0.7978845608028654 * (1 + 3 * 0.044715 * pre.float() ** 2)
    assert rel(be.linear_dgrad(dy, w, gelu_pre=pre), (dy.float() @ w.float()) * g) < 1e-2
    acc = torch.zeros(N, K, device="cuda")
    be.linear_wgrad(dy, x, acc)
    assert rel(acc, dy.float().t() @ x.float()) < 3e-3
    # long K, few tiles: hybrid split-K schedule
    M2 = 4096
    dy2, x2 = rnd(M2, 256), rnd(M2, 256)
    acc2 = torch.zeros(256, 256, device="cuda")
    be.linear_wgrad(dy2, x2, acc2)
    assert rel(acc2, dy2.float().t() @ x2.float()) < 3e-3


def test_gemm_single_cta_and_fp8(be):
    x, w = rnd(128, 128), rnd(256, 128, scale=128 ** -0.5)          # M < 256: 1-CTA kernel
    assert rel(be.linear_fwd(x, w), x.float() @ w.float().t()) < 6e-3
    from midgpt_b200.ops.b200_backend import B200Backend
    for fmt in ("row", "mx"):
        b8 = B200Backend(torch.device("cuda", 0), fp8=True, fp8_format=fmt)
        x, w = rnd(256, 256), rnd(256, 256, scale=1 / 16)
        assert rel(b8.linear_fwd(x, w), x.float() @ w.float().t()) < 6e-2


@pytest.mark.parametrize("C", [64, 128])
def test_attention_all_kernels(be, C):
    """forward, backward v3 (persistent, several items per CTA would need > #SMs items: covered by timing-free logic here),
    backward v2 with in-kernel dropout."""
    from tests.test_attention_gpu import make_inputs, ref_attention, row_err
    B, T, H = 1, 256, 2
    h, q, k, v = make_inputs(B, T, H, C, 3)
    o, lse = be.attention_fwd(h, B, T, H, C)
    o_r, _ = ref_attention(q, k, v)
    assert row_err(o, o_r) < 2e-2
    do = rnd(B * T, H * C)
    dqkv = be.attention_bwd(do, h, o, lse, B, T, H, C)
    assert torch.isfinite(dqkv.float()).all()
    od, lsed = be.attention_fwd(h, B, T, H, C, 0.2, 7)
    dd = be.attention_bwd(do, h, od, lsed, B, T, H, C, 0.2, 7)
    assert torch.isfinite(dd.float()).all() and torch.isfinite(od.float()).all()


def test_elementwise_and_optimizer(be):
    from midgpt_b200.ops.torch_backend import TorchBackend
    ref = TorchBackend(use_aten_flash=False)
    x = rnd(256, 256)
    y, rstd = be.rmsnorm_fwd(x, 1e-6)
    yr, _ = ref.rmsnorm_fwd(x.float(), 1e-6)
    assert rel(y, yr) < 6e-3
    dx = be.rmsnorm_bwd(rnd(256, 256), x, rstd, x)
    assert torch.isfinite(dx.float()).all()
    B, T, H, C = 1, 128, 2, 64
    qkv = rnd(B * T, 3 * H * C)
    gq, gk = torch.ones(C, device="cuda"), torch.ones(C, device="cuda")
    hq, stats = be.qkln_rope_fwd(qkv, gq, gk, B, T, H, C)
    dgq, dgk = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    be.qkln_rope_bwd(rnd(B * T, 3 * H * C), qkv, stats, gq, gk, dgq, dgk, B, T, H, C)
    wte = rnd(65, 128)
    idx = torch.randint(0, 65, (2, 64), device="cuda", dtype=torch.int32)
    e = be.embed_fwd(wte, idx)
    acc = torch.zeros(65, 128, device="cuda")
    be.embed_bwd(e, idx, acc)
    hf, w = rnd(256, 128), rnd(65, 128, scale=0.1)
    dw = torch.zeros(65, 128, device="cuda")
    loss, dh = be.lmhead_ce(hf, w, torch.randint(0, 65, (256,), device="cuda"), dw, 1.0 / 256)
    assert math.isfinite(loss.item())
    n = 1 << 14
    p, m, v, g = (torch.randn(n, device="cuda") for _ in range(4))
    v.abs_()
    lp = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    be.adamw_step(p, m, v, g, lp, lr=1e-3, b1=0.9, b2=0.95, eps=1e-8, wd_over_lr_peak=0.1, bc1=0.5, bc2=0.5, grad_scale=torch.ones(1, device="cuda"))
    assert math.isfinite(be.sumsq(g).item())
    assert be.dropout(x, 0.2, 1).shape == x.shape
