"""Error map of the attention backward: which (batch, head, 64-row query tile) blocks of dq/dk/dv are wrong."""
import sys, math, torch
sys.path.insert(0, ".")
from midgpt_b200.ops.b200_backend import B200Backend, QKV
from tests.test_attention_gpu import ref_attention, make_inputs, rnd

be = B200Backend(torch.device("cuda", 0))
for (B, T, H, C) in [(1, 1024, 16, 128), (2, 1024, 16, 128), (8, 1024, 16, 128), (16, 1024, 12, 64), (2, 1024, 2, 64)]:
    h, q, k, v = make_inputs(B, T, H, C, 11)
    D = H * C
    o, lse = be.attention_fwd(h, B, T, H, C)
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    o_r, lse_r = ref_attention(qr, kr, vr)
    do = rnd(B * T, D)
    o_r.backward(do.float())
    for rep in range(2):
        dqkv = be.attention_bwd(do, h, o, lse, B, T, H, C)
        for name, got, want in (("dq", dqkv[:, :D], qr.grad), ("dk", dqkv[:, D:2 * D], kr.grad), ("dv", dqkv[:, 2 * D:], vr.grad)):
            g = got.float().reshape(B, T // 64, 64, H, C)
            w = want.reshape(B, T // 64, 64, H, C)
            e = ((g - w).pow(2).sum((2, 4)) / (w.pow(2).sum((2, 4)) + 1e-9)).sqrt()      # [B, T/64, H]
            bad = (e > 0.05).nonzero()
            print(f"{(B,T,H,C)} rep{rep} {name}: max blk err {e.max().item():.3f}, bad blocks {bad.shape[0]}/{e.numel()}", bad[:12].tolist() if bad.shape[0] else "")
