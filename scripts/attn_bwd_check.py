"""Attention backward: per-tensor error against the fp32 reference and timing, for the kernel version selected by
MIDGPT_ATTN_BWD (unset = v3 where it applies, "v2" = previous kernel).  Debug aid; the real test is tests/test_kernels_gpu.py.
    python scripts/attn_bwd_check.py B T H C
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from midgpt_b200.ops.b200_backend import B200Backend, QKV  # noqa: E402
from midgpt_b200.ops.torch_backend import TorchBackend  # noqa: E402


def main():
    B, T, H, C = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (1, 512, 2, 128)
    dev = torch.device("cuda", 0)
    be, ref = B200Backend(dev), TorchBackend(use_aten_flash=False)
    torch.manual_seed(8)
    D = H * C
    rnd = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
    raw, qk, do = rnd(B * T, 3 * D), rnd(B * T, 2 * D), rnd(B * T, D)
    h = QKV(qk, raw)
    o, lse = be.attention_fwd(h, B, T, H, C)
    packed = torch.cat((qk, raw[:, 2 * D:]), 1).float()
    o_r, lse_r = ref.attention_fwd(packed, B, T, H, C)
    g = be.attention_bwd(do, h, o, lse, B, T, H, C)
    torch.cuda.synchronize()
    g_r = ref.attention_bwd(do.float(), packed, o_r, lse_r, B, T, H, C)
    out = {"ver": os.environ.get("MIDGPT_ATTN_BWD", "v3"), "shape": [B, T, H, C]}
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        a, b = g[:, sl].float(), g_r[:, sl].float()
        out[name] = round(((a - b).norm() / b.norm()).item(), 5)
        out[name + "_nan"] = bool(torch.isnan(a).any())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        be.attention_bwd(do, h, o, lse, B, T, H, C)
    e0.record()
    for _ in range(10):
        be.attention_bwd(do, h, o, lse, B, T, H, C)
    e1.record()
    torch.cuda.synchronize()
    out["bwd_ms"] = round(e0.elapsed_time(e1) / 10, 4)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
