"""Single-GPU stand-in for the all-gather ⊕ GEMM copy engines: the "peer" shard is another local buffer, so the copy-engine code
path (segment loop, loads, stores, fence, flag) runs unchanged and can be timed / profiled with ncu on one GPU."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midgpt_b200.ops.b200_backend import B200Backend  # noqa: E402
from midgpt_b200.parallel.fused import GatherSpec  # noqa: E402

dev = torch.device("cuda", 0)
be = B200Backend(dev)
N, K = 6144, 2048
unit = 50_331_648
world = 2
shard = unit // world
peer = torch.randn(shard, device=dev).bfloat16()            # pretend rank 1's shard
mine = torch.randn(shard, device=dev).bfloat16()
local = torch.zeros(unit, dtype=torch.bfloat16, device=dev)
local[:shard].copy_(mine)
flags = torch.zeros(4096, dtype=torch.int32, device=dev)
ptrs = [mine.data_ptr(), peer.data_ptr()]
calls = 0
for M in (128, 16384):
    x = torch.randn(M, K, device=dev).bfloat16()
    for n_ctas in (8, 16, 32, 64):
        ts = []
        for it in range(6):
            calls += 1
            spec = GatherSpec(ptrs, local, flags, shard, calls, 0, n_ctas)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = be.linear_fwd(x, local[: N * K].view(N, K), gather=spec)
            e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        ok = torch.equal(local[shard:], peer)
        print(f"M={M} n_ctas={n_ctas}: {t:.4f} ms  copy {shard * 2 / t / 1e6:.0f} GB/s (local stand-in)  gathered_ok={ok}", flush=True)
    t0 = []
    for it in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); be.linear_fwd(x, local[: N * K].view(N, K)); e1.record(); e1.synchronize()
        t0.append(e0.elapsed_time(e1))
    print(f"M={M} plain GEMM: {sorted(t0)[3]:.4f} ms", flush=True)
