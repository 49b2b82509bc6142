"""Flash-attention kernels where the code is riskiest: enough work items that the persistent / dynamic
schedules of the backward kernels engage (items >= 3 x #SMs), T = 4096, head_dim 64 on the transposed (v3) backward, and the
in-kernel attention-probability dropout (forward/backward mask consistency against an fp32 PyTorch reference that uses the
mask extracted from the kernel itself).  Errors are checked per row (max over rows of the relative row error) and as max-abs,
not only as one Frobenius norm, so a few wrong rows or columns cannot hide."""
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


def row_err(a, b):
    """Worst row: max_r ||a_r - b_r|| / max(||b_r||, rms_r ||b_r||).  Short causal rows are ill-conditioned in any flash
    backward (dS = P ∘ (dP − δ) cancels, and δ is taken from the bf16-rounded O: row 0 of dq is exactly 0 in exact arithmetic),
    so rows below the typical norm are measured against the typical norm; a row written by the wrong lane / tile is off by
    about a typical norm and scores ~1."""
    a, b = a.float(), b.float()
    nb = b.norm(dim=-1)
    floor = nb.pow(2).mean().sqrt()
    return ((a - b).norm(dim=-1) / torch.maximum(nb, floor)).max().item()


def ref_attention(q, k, v, mask=None, inv_keep=1.0):
    """fp32 causal attention on [B,T,H,C] tensors; ``mask`` [B,H,T,T] (1 = keep) applies dropout to the probabilities."""
    B, T, H, C = q.shape
    q, k, v = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(C)
    causal = torch.ones(T, T, dtype=torch.bool, device=q.device).tril_()
    s = s.masked_fill(~causal, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.softmax(s, -1)
    if mask is not None:
        p = p * mask * inv_keep
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * T, H * C)
    return o, lse


def make_inputs(B, T, H, C, seed):
    from midgpt_b200.ops.b200_backend import QKV
    torch.manual_seed(seed)
    D = H * C
    raw, qk = rnd(B * T, 3 * D), rnd(B * T, 2 * D)
    q = qk.view(B, T, 2, H, C)[:, :, 0]
    k = qk.view(B, T, 2, H, C)[:, :, 1]
    v = raw.view(B, T, 3, H, C)[:, :, 2]
    return QKV(qk, raw), q, k, v


@pytest.mark.parametrize("B,T,H,C", [(8, 1024, 16, 128), (2, 4096, 16, 128), (16, 1024, 12, 64), (3, 320, 5, 64)])
def test_attention_many_items_rowwise(be, B, T, H, C):
    """(8,1024,16,128) / (2,4096,16,128) / (16,1024,12,64): 1024-1536 (batch, head, key-tile) items on 148 persistent CTAs —
    the cross-item pipeline of attn_bwd3_kernel that the benchmark runs; (3,320,5,64): T % 128 == 64 on the head_dim-64 v3 path."""
    h, q, k, v = make_inputs(B, T, H, C, 11)
    D = H * C
    o, lse = be.attention_fwd(h, B, T, H, C)
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    o_r, lse_r = ref_attention(qr, kr, vr)
    assert row_err(o, o_r) < 1.5e-2
    assert (o.float() - o_r).abs().max().item() < 4e-2
    assert (lse - lse_r).abs().max().item() < 2e-2
    do = rnd(B * T, D)
    dqkv = be.attention_bwd(do, h, o, lse, B, T, H, C)
    o_r.backward(do.float())
    for name, got, want in (("dq", dqkv[:, :D], qr.grad), ("dk", dqkv[:, D:2 * D], kr.grad), ("dv", dqkv[:, 2 * D:], vr.grad)):
        want = want.reshape(B * T, D)
        # per (token, head) vector: the unit a wrong tile / lane mapping would corrupt
        e = row_err(got.reshape(B * T, H, C), want.reshape(B * T, H, C))
        assert e < 6e-2, (name, e)
        assert (got.float() - want).abs().max().item() < 0.05 * want.abs().max().item() + 2e-2, name


def extract_keep_mask(be, B, T, H, C, p, seed):
    """The kernel's own dropout mask: with q = 0 every causal probability of row i is 1/(i+1), and with V = one-hot of the key
    index inside a C-wide block the output columns are exactly the (dropped, rescaled) probabilities of that block."""
    from midgpt_b200.ops.b200_backend import QKV
    D = H * C
    qk = torch.zeros(B * T, 2 * D, device="cuda", dtype=torch.bfloat16)
    mask = torch.zeros(B, H, T, T, device="cuda")
    for j in range((T + C - 1) // C):
        raw = torch.zeros(B, T, 3, H, C, device="cuda", dtype=torch.bfloat16)
        n = min(C, T - j * C)
        raw[:, j * C:j * C + n, 2, :, :n] = torch.eye(n, device="cuda", dtype=torch.bfloat16)[None, :, None, :]
        o, _ = be.attention_fwd(QKV(qk, raw.view(B * T, 3 * D)), B, T, H, C, p, seed)
        mask[:, :, :, j * C:j * C + n] = (o.view(B, T, H, C)[..., :n] != 0).permute(0, 2, 1, 3).float()
    return mask


@pytest.mark.parametrize("B,T,H,C", [(2, 256, 3, 64), (1, 384, 2, 128)])
def test_attention_dropout_in_kernel(be, B, T, H, C):
    p, seed = 0.2, 4242
    thr = round(p * 65536)
    inv_keep = 1.0 / (1.0 - thr / 65536.0)
    mask = extract_keep_mask(be, B, T, H, C, p, seed)
    causal = torch.ones(T, T, dtype=torch.bool, device="cuda").tril_()
    keep = mask[:, :, causal].mean().item()
    assert abs(keep - (1 - p)) < 6e-3, keep                      # ~100k-200k Bernoulli draws: sigma ~ 1e-3
    assert mask[:, :, ~causal].sum().item() == 0
    # rows and columns are not correlated with each other (a counter bug would repeat the mask along an axis)
    m = mask[0, 0]
    assert not torch.equal(m[T // 2, : T // 4], m[T // 2 + 1, : T // 4])
    assert not torch.equal(m[T - 8:, 8:16], m[T - 8:, 16:24])
    # forward / backward consistency with random inputs and the SAME seed (the mask does not depend on the data)
    h, q, k, v = make_inputs(B, T, H, C, 5)
    D = H * C
    o, lse = be.attention_fwd(h, B, T, H, C, p, seed)
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    o_r, lse_r = ref_attention(qr, kr, vr, mask, inv_keep)
    assert row_err(o, o_r) < 2e-2
    assert (lse - lse_r).abs().max().item() < 2e-2
    do = rnd(B * T, D)
    dqkv = be.attention_bwd(do, h, o, lse, B, T, H, C, p, seed)
    o_r.backward(do.float())
    for name, got, want in (("dq", dqkv[:, :D], qr.grad), ("dk", dqkv[:, D:2 * D], kr.grad), ("dv", dqkv[:, 2 * D:], vr.grad)):
        e = row_err(got.reshape(B * T, H, C), want.reshape(B * T, H, C))
        assert e < 6e-2, (name, e)
    # determinism in the seed
    o2, _ = be.attention_fwd(h, B, T, H, C, p, seed)
    o3, _ = be.attention_fwd(h, B, T, H, C, p, seed + 1)
    assert torch.equal(o, o2) and not torch.equal(o, o3)
    # p = 0 is the plain kernel
    o0, _ = be.attention_fwd(h, B, T, H, C, 0.0, seed)
    o0_r, _ = ref_attention(q, k, v)
    assert row_err(o0, o0_r) < 1.5e-2


@pytest.mark.parametrize("B,T,H,C", [(2, 512, 3, 64), (2, 256, 2, 128)])
def test_fused_dq_path_equals_unfused(be, B, T, H, C):
    """Model path: dq stays in the backward's fp32 accumulator, qkln_rope_bwd reads it there and clears it -- no zero-fill, no
    fp32->bf16 conversion kernel.  Must equal the unfused pipeline bit for bit and leave the accumulator all-zero."""
    h, q, k, v = make_inputs(B, T, H, C, 21)
    D = H * C
    o, lse = be.attention_fwd(h, B, T, H, C)
    do = rnd(B * T, D)
    gq, gk = torch.rand(C, device="cuda") + 0.5, torch.rand(C, device="cuda") + 0.5
    _, stats = be.qkln_rope_fwd(h.raw, gq, gk, B, T, H, C)
    outs = []
    for fuse in (False, True, True):            # twice fused: the accumulator must be reusable
        dgq, dgk = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        dqkv = be.attention_bwd(do, h, o, lse, B, T, H, C, fuse_dq=fuse)
        out = be.qkln_rope_bwd(dqkv, h.raw, stats, gq, gk, dgq, dgk, B, T, H, C)
        outs.append((out.clone(), dgq, dgk))
    for got in outs[1:]:
        assert torch.equal(got[0], outs[0][0])
        assert torch.allclose(got[1], outs[0][1], rtol=1e-4, atol=1e-4) and torch.allclose(got[2], outs[0][2], rtol=1e-4, atol=1e-4)
    assert be._dq_accumulator(B * T, D).abs().max().item() == 0.0
