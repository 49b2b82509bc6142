"""Numerics of every hand-written sm_100a kernel against a plain PyTorch fp32 reference of the same op."""

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from midgpt_b200.ops.b200_backend import B200Backend
    return B200Backend(torch.device("cuda", 0))


@pytest.fixture(scope="module")
def ref():
    from midgpt_b200.ops.torch_backend import TorchBackend
    return TorchBackend(use_aten_flash=False)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def rnd(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (384, 1152, 384), (1000, 776, 200), (4096, 2048, 2048)])
def test_gemm_nt(be, M, N, K):
    torch.manual_seed(0)
    x, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    y = be.linear_fwd(x, w)
    assert rel_err(y, x.float() @ w.float().t()) < 6e-3


def test_gemm_epilogues(be):
    torch.manual_seed(1)
    M, N, K = 512, 768, 256
    x, w, res = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(M, N)
    y = be.linear_fwd(x, w, res=res)
    base = x.float() @ w.float().t()
    assert rel_err(y, base + res.float()) < 6e-3
    pre, act = be.linear_fwd(x, w, act="gelu")
    assert rel_err(pre, base) < 6e-3
    assert rel_err(act, torch.nn.functional.gelu(pre.float(), approximate="tanh")) < 6e-3


@pytest.mark.parametrize("M,N,K", [(256, 512, 128), (640, 384, 1152), (1000, 200, 776)])
def test_gemm_dgrad_and_wgrad(be, ref, M, N, K):
    torch.manual_seed(2)
    dy, w, x = rnd(M, N), rnd(N, K, scale=N ** -0.5), rnd(M, K)
    dx = be.linear_dgrad(dy, w)
    assert rel_err(dx, dy.float() @ w.float()) < 6e-3
    if K % 64 == 0:
        pre = rnd(M, K)
        dxg = be.linear_dgrad(dy, w, gelu_pre=pre)
        assert rel_err(dxg, ref.linear_dgrad(dy.float(), w.float(), gelu_pre=pre.float())) < 8e-3
    acc = torch.randn(N, K, device="cuda")
    want = acc + dy.float().t() @ x.float()
    be.linear_wgrad(dy, x, acc)
    assert rel_err(acc, want) < 3e-3
    be.linear_wgrad(dy, x, acc)     # accumulates
    assert rel_err(acc, want + dy.float().t() @ x.float()) < 3e-3


@pytest.mark.parametrize("M,N,K", [(8192, 1024, 768), (16384, 2560, 2048)])
def test_gemm_wgrad_hybrid_split_k(be, M, N, K):
    """Long-K weight gradients take the hybrid split-K schedule (csrc/gemm.cu WorkIter): fewer tiles than workers (12 on 74
    CTA pairs) and one full wave + 6 leftover tiles cut into K-slices; every piece must add into the fp32 accumulator exactly once."""
    torch.manual_seed(21)
    dy, x = rnd(M, N), rnd(M, K)
    acc = torch.randn(N, K, device="cuda")
    want = acc.double() + dy.double().t() @ x.double()
    be.linear_wgrad(dy, x, acc)
    assert rel_err(acc, want) < 2e-3
    be.linear_wgrad(dy, x, acc)
    assert rel_err(acc, want + dy.double().t() @ x.double()) < 2e-3


# ----------------------------------------------------------------------------------------------- element-wise
@pytest.mark.parametrize("D", [384, 768, 2048])
def test_rmsnorm(be, ref, D):
    torch.manual_seed(3)
    x, dy, dres = rnd(300, D), rnd(300, D), rnd(300, D)
    y, rstd = be.rmsnorm_fwd(x, 1e-6)
    y_r, rstd_r = ref.rmsnorm_fwd(x.float(), 1e-6)
    assert rel_err(y, y_r) < 4e-3 and rel_err(rstd, rstd_r) < 1e-5
    dx = be.rmsnorm_bwd(dy, x, rstd, dres)
    assert rel_err(dx, ref.rmsnorm_bwd(dy.float(), x.float(), rstd_r, dres.float())) < 4e-3


@pytest.mark.parametrize("H,C", [(6, 64), (4, 128)])
def test_qkln_rope(be, ref, H, C):
    torch.manual_seed(4)
    B, T = 2, 96
    D = H * C
    qkv = rnd(B * T, 3 * D)
    qg, kg = (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16(), (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16()
    h, stats = be.qkln_rope_fwd(qkv, qg, kg, B, T, H, C)
    out_r, stats_r = ref.qkln_rope_fwd(qkv.float(), qg.float(), kg.float(), B, T, H, C)
    assert rel_err(h.qk, out_r[:, : 2 * D]) < 4e-3
    assert rel_err(stats, stats_r) < 1e-4
    dqkv = rnd(B * T, 3 * D)
    gq, gk = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    gq_r, gk_r = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    want = ref.qkln_rope_bwd(dqkv.float(), qkv.float(), stats_r, qg.float(), kg.float(), gq_r, gk_r, B, T, H, C)
    got = be.qkln_rope_bwd(dqkv.clone(), qkv, stats, qg, kg, gq, gk, B, T, H, C)
    assert rel_err(got, want) < 5e-3
    assert rel_err(gq, gq_r) < 5e-3 and rel_err(gk, gk_r) < 5e-3


def test_embedding(be):
    torch.manual_seed(5)
    V, D, M = 1000, 256, 777
    wte = rnd(V, D)
    idx = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32)
    assert torch.equal(be.embed_fwd(wte, idx), wte[idx.long()])
    dx = rnd(M, D)
    acc = torch.zeros(V, D, device="cuda")
    be.embed_bwd(dx, idx, acc)
    want = torch.zeros(V, D, device="cuda").index_add_(0, idx.long(), dx.float())
    assert rel_err(acc, want) < 1e-5


@pytest.mark.parametrize("V", [65, 50304])
def test_lmhead_ce(be, ref, V):
    torch.manual_seed(6)
    M, D = 640, 256
    h, w = rnd(M, D), rnd(V, D, scale=D ** -0.5)
    tgt = torch.randint(0, V, (M,), device="cuda")
    acc, acc_r = torch.zeros(V, D, device="cuda"), torch.zeros(V, D, device="cuda")
    loss, dh = be.lmhead_ce(h, w, tgt, acc, 1.0 / M, chunk=256)
    loss_r, dh_r = ref.lmhead_ce(h.float(), w.float(), tgt, acc_r, 1.0 / M)
    assert abs(loss.item() - loss_r.item()) / loss_r.item() < 3e-3
    assert rel_err(dh, dh_r) < 2e-2 and rel_err(acc, acc_r) < 2e-2
    loss_e, none = be.lmhead_ce(h, w, tgt, None, 0.0)
    assert none is None and abs(loss_e.item() - loss_r.item()) / loss_r.item() < 3e-3


def test_adamw_and_sumsq(be, ref):
    torch.manual_seed(7)
    n = 1 << 20
    p, g = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    m, v = torch.rand(n, device="cuda") * 0.1, torch.rand(n, device="cuda") * 0.1
    pr, mr, vr = p.double(), m.double(), v.double()
    lp = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    kw = dict(lr=3e-4, b1=0.9, b2=0.95, eps=1e-8, wd_over_lr_peak=0.1, bc1=0.19, bc2=0.0975)
    clip = torch.tensor(0.37, device="cuda")
    assert abs(be.sumsq(g).item() - g.double().pow(2).sum().item()) / n < 1e-4
    be.adamw_step(p, m, v, g, lp, grad_scale=clip, **kw)
    ref.adamw_step(pr, mr, vr, g.double(), None, grad_scale=0.37, **kw)
    assert rel_err(p, pr) < 1e-6 and rel_err(m, mr) < 1e-6 and rel_err(v, vr) < 1e-6
    assert torch.equal(lp, p.bfloat16())


def test_dropout_statistics(be):
    x = torch.ones(1 << 20, device="cuda", dtype=torch.bfloat16)
    y = be.dropout(x, 0.2, 1234)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.8) < 5e-3
    assert torch.allclose(y[y != 0].float(), torch.tensor(1.25, device="cuda"), atol=1e-2)
    assert torch.equal(y, be.dropout(x, 0.2, 1234)) and not torch.equal(y, be.dropout(x, 0.2, 1235))


# ----------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,T,H,C", [(2, 256, 3, 64), (1, 512, 2, 128), (2, 200, 2, 64), (1, 1024, 2, 128), (2, 192, 2, 128), (1, 200, 2, 128)])
def test_flash_attention_fwd_bwd(be, ref, B, T, H, C):
    torch.manual_seed(8)
    D = H * C
    raw = rnd(B * T, 3 * D)
    qk = rnd(B * T, 2 * D)
    from midgpt_b200.ops.b200_backend import QKV
    h = QKV(qk, raw)
    o, lse = be.attention_fwd(h, B, T, H, C)
    packed = torch.cat((qk, raw[:, 2 * D:]), 1).float()
    o_r, lse_r = ref.attention_fwd(packed, B, T, H, C)
    assert rel_err(o, o_r) < 8e-3
    assert (lse - lse_r).abs().max().item() < 2e-2
    do = rnd(B * T, D)
    dqkv = be.attention_bwd(do, h, o, lse, B, T, H, C)
    dqkv_r = ref.attention_bwd(do.float(), packed, o_r, lse_r, B, T, H, C)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        assert rel_err(dqkv[:, sl], dqkv_r[:, sl]) < 1.5e-2, name


# ----------------------------------------------------------------------------------------------- FP8
def test_fp8_quant_and_gemm(be):
    torch.manual_seed(9)
    M, N, K = 512, 768, 1024
    x, w, res = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(M, N)
    q, s = be.ext.quant_rows_e4m3(x)
    deq = q.view(torch.float8_e4m3fn).float() * s[:, None]
    assert rel_err(deq, x) < 4e-2                       # e4m3 has 3 mantissa bits
    assert torch.allclose(s, x.float().abs().amax(1) / 448, rtol=1e-3)
    from midgpt_b200.ops.b200_backend import B200Backend
    be8 = B200Backend(torch.device("cuda", 0), fp8=True, fp8_format="row")
    wq, sw = be8.ext.quant_rows_e4m3(w)
    ref = (q.view(torch.float8_e4m3fn).float() * s[:, None]) @ (wq.view(torch.float8_e4m3fn).float() * sw[:, None]).t()
    y = be8.linear_fwd(x, w)
    assert rel_err(y, ref) < 6e-3                       # exact up to bf16 output rounding given the quantised operands
    assert rel_err(y, x.float() @ w.float().t()) < 6e-2
    y2 = be8.linear_fwd(x, w, res=res)
    assert rel_err(y2, ref + res.float()) < 6e-3
    pre, act = be8.linear_fwd(x, w, act="gelu")
    assert rel_err(pre, ref) < 6e-3 and rel_err(act, torch.nn.functional.gelu(pre.float(), approximate="tanh")) < 6e-3


def test_mx_quantisation_layout_and_values(be):
    """MX e4m3: per-32-element UE8M0 scales in the tcgen05 scale-factor atom order; dequantised values match."""
    torch.manual_seed(10)
    R, C = 300, 512
    x = rnd(R, C) * torch.exp(torch.randn(R, C // 32, device="cuda")).repeat_interleave(32, 1).bfloat16()
    q, sf = be.ext.quant_mx_e4m3(x)
    rows = torch.arange(R, device="cuda")
    kb = torch.arange(C // 32, device="cuda")
    rb, m = rows // 128, rows % 128
    off = (m % 32) * 16 + (m // 32) * 4
    idx = ((rb[:, None] * (C // 128) + (kb // 4)[None, :]) * 512 + off[:, None] + (kb % 4)[None, :])
    e = sf.view(-1)[idx].float() - 127.0                       # [R, C/32] exponents
    amax = x.float().abs().view(R, C // 32, 32).amax(-1)
    want = torch.ceil(torch.log2(amax.clamp_min(1e-30) / 448.0)).clamp(-127, 127)
    assert (e != want).float().mean().item() < 1e-3 and (e - want).abs().max().item() <= 1      # ties at exact powers of two aside
    deq = q.view(torch.float8_e4m3fn).float().view(R, C // 32, 32) * torch.exp2(e)[..., None]
    assert rel_err(deq.view(R, C), x) < 4e-2


def _mx_dequant(q, sf, R, C):
    rows = torch.arange(R, device="cuda")
    kb = torch.arange(C // 32, device="cuda")
    rb, m = rows // 128, rows % 128
    off = (m % 32) * 16 + (m // 32) * 4
    idx = ((rb[:, None] * (C // 128) + (kb // 4)[None, :]) * 512 + off[:, None] + (kb % 4)[None, :])
    e = sf.view(-1)[idx].float() - 127.0
    return (q.view(torch.float8_e4m3fn).float().view(R, C // 32, 32) * torch.exp2(e)[..., None]).view(R, C)


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (512, 768, 1024), (300, 2048, 2048)])
def test_mxfp8_block_scaled_gemm(be, M, N, K):
    """tcgen05 kind::mxf8f6f4.block_scale GEMM == fp32 matmul of the de-quantised MX operands."""
    torch.manual_seed(11)
    from midgpt_b200.ops.b200_backend import B200Backend
    be8 = B200Backend(torch.device("cuda", 0), fp8=True, fp8_format="mx")
    x = rnd(M, K) * torch.exp(0.5 * torch.randn(M, K // 32, device="cuda")).repeat_interleave(32, 1).bfloat16()
    w, res = rnd(N, K, scale=K ** -0.5), rnd(M, N)
    xq, sx = be8.ext.quant_mx_e4m3(x)
    wq, sw = be8.ext.quant_mx_e4m3(w)
    ref = _mx_dequant(xq, sx, M, K) @ _mx_dequant(wq, sw, N, K).t()
    y = be8.linear_fwd(x, w)
    assert rel_err(y, ref) < 6e-3
    assert rel_err(y, x.float() @ w.float().t()) < 6e-2
    assert rel_err(be8.linear_fwd(x, w, res=res), ref + res.float()) < 6e-3
    pre, act = be8.linear_fwd(x, w, act="gelu")
    assert rel_err(pre, ref) < 6e-3 and rel_err(act, torch.nn.functional.gelu(pre.float(), approximate="tanh")) < 6e-3


# ----------------------------------------------------------------------------------------------- GEMM at the models' real shapes
@pytest.mark.parametrize("name,N,K", [("1.5B c_attn", 6144, 2048), ("1.5B mlp_proj", 2048, 8192), ("124M c_attn", 2304, 768),
                                      ("124M mlp_proj", 768, 3072), ("124M c_fc", 3072, 768)])
def test_gemm_model_shapes_rowwise(be, ref, name, N, K):
    """Forward (+gelu / +residual), dgrad (+gelu') and wgrad on the layer shapes of both headline models at M = 8192 tokens, checked
    per output row (max over rows of the relative row error) so that a mis-routed tile / box cannot hide in a Frobenius norm.
    K = 768 exercises the TMA-box epilogue (5 operand stages), K >= 2048 the register epilogue (6 stages)."""
    torch.manual_seed(5)
    M = 8192
    x, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)

    def rows(a, b):
        a, b = a.float(), b.float()
        return ((a - b).norm(dim=1) / (b.norm(dim=1) + 1e-6)).max().item()

    base = x.float() @ w.float().t()
    assert rows(be.linear_fwd(x, w), base) < 8e-3
    res = rnd(M, N)
    assert rows(be.linear_fwd(x, w, res=res), base + res.float()) < 8e-3
    pre, act = be.linear_fwd(x, w, act="gelu")
    assert rows(pre, base) < 8e-3 and rows(act, torch.nn.functional.gelu(pre.float(), approximate="tanh")) < 8e-3
    dy = rnd(M, N)
    want = dy.float() @ w.float()
    assert rows(be.linear_dgrad(dy, w), want) < 8e-3
    gp = rnd(M, K)
    assert rows(be.linear_dgrad(dy, w, gelu_pre=gp), ref.linear_dgrad(dy.float(), w.float(), gelu_pre=gp.float())) < 1.2e-2
    acc = torch.zeros(N, K, device="cuda")
    be.linear_wgrad(dy, x, acc)
    assert rows(acc, dy.float().t() @ x.float()) < 4e-3
