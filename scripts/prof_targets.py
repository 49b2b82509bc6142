"""Launch each hot kernel a few times on representative 1.5B shapes (for `ncu -k regex:… --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midgpt_b200.ops.b200_backend import B200Backend, QKV  # noqa: E402

dev = torch.device("cuda", 0)
be = B200Backend(dev)
r = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16)
which = sys.argv[1:] or ["gemm", "attn"]
M = 16384
if "gemm" in which:
    x, w = r(M, 2048), r(8192, 2048) * 0.02
    dy = r(M, 8192)
    acc = torch.zeros(8192, 2048, device=dev)
    for _ in range(2):
        be.linear_fwd(x, w, act="gelu")      # <0,0,2>
        be.linear_dgrad(dy, w)                # <0,1,0>
        be.linear_wgrad(dy, x, acc)           # <1,1,4>
if "attn" in which:
    B, T, H, C = 16, 1024, 16, 128
    D = H * C
    h = QKV(r(B * T, 2 * D), r(B * T, 3 * D))
    do = r(B * T, D)
    for _ in range(2):
        o, lse = be.attention_fwd(h, B, T, H, C)
        be.attention_bwd(do, h, o, lse, B, T, H, C)
torch.cuda.synchronize()
print("done")
# ---- round 2: the 124M (K = 768 / head_dim 64) shapes
if "gemm124" in which:
    M2 = 32768
    x, x4 = r(M2, 768), r(M2, 3072)
    w_fc, w_pr = r(3072, 768) * 0.03, r(768, 3072) * 0.02
    for _ in range(2):
        be.linear_fwd(x, w_fc, act="gelu")            # <0,0,2,...,AUXB> TMA-box epilogue
        be.linear_fwd(x4, w_pr, res=x)                # <0,0,1> K = 3072 register path
        be.linear_dgrad(x, w_pr, gelu_pre=x4)         # <0,1,3,...,AUXB>
        be.linear_dgrad(x4, w_fc)                     # <0,1,0> N = 768: 5.19 waves
if "attn64" in which:
    B, T, H, C = 32, 1024, 12, 64
    D = H * C
    h = QKV(r(B * T, 2 * D), r(B * T, 3 * D))
    do = r(B * T, D)
    for _ in range(2):
        o, lse = be.attention_fwd(h, B, T, H, C)
        be.attention_bwd(do, h, o, lse, B, T, H, C)
torch.cuda.synchronize()
print("done r2 targets")
