"""Attention-backward what-if timings (profiling aid).  Runs the bwd kernel with MIDGPT_ATTN_DEBUG experiment bits
(see csrc/attention.cu; results are numerically wrong when a bit is set) to attribute its time:
    python scripts/attn_bwd_probe.py            # spawns one process per variant, prints one JSON line each
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(B=32, T=1024, H=16, C=128):
    import torch
    from midgpt_b200.ops.b200_backend import B200Backend, QKV
    dev = torch.device("cuda", 0)
    be = B200Backend(dev)
    D = H * C
    g = torch.Generator(device=dev).manual_seed(0)
    raw = torch.randn(B * T, 3 * D, device=dev, generator=g).to(torch.bfloat16)
    h = QKV(torch.randn(B * T, 2 * D, device=dev, generator=g).to(torch.bfloat16), raw)
    do = torch.randn(B * T, D, device=dev, generator=g).to(torch.bfloat16)
    o, lse = be.attention_fwd(h, B, T, H, C)
    for _ in range(3):
        be.attention_bwd(do, h, o, lse, B, T, H, C)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            be.attention_bwd(do, h, o, lse, B, T, H, C)
        torch.cuda.synchronize()
    out = {}
    for ev in prof.key_averages():
        if "attn_bwd" in ev.key or "dq_convert" in ev.key:
            out[ev.key.split("<")[0].split("::")[-1]] = round(ev.device_time_total / ev.count / 1e3, 4)
    print(json.dumps({"dbg": int(os.environ.get("MIDGPT_ATTN_DEBUG", "0")), "B": B, "T": T, "ms": out}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for dbg in (0, 1, 2, 3, 5, 7):
            env = dict(os.environ, MIDGPT_ATTN_DEBUG=str(dbg))
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, timeout=120)
