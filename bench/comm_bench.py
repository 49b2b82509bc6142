"""Roofline check of the fused compute+collective kernels (run under torchrun, one rank per GPU).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 bench/comm_bench.py

For each kernel: device time (CUDA events, max over ranks), bytes that must cross NVLink per GPU, achieved GB/s and the
fraction of BOTH the measured per-direction peer bandwidth (770 GB/s, B200_PROFILING.md) and the 900 GB/s link rate
BASELINE.json names; for all-gather ⊕ GEMM the target is max(GEMM alone on the same box, NVLink bytes / 900 GB/s).
NCCL equivalents are timed on the same buffers for reference.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midgpt_b200.parallel.dist import init_distributed  # noqa: E402
from midgpt_b200.utils.obs import peak_flops_per_gpu  # noqa: E402

NVLINK = 770e9
LINK_RATE = 900e9


def timeit(fn, pre=None, iters=8, warmup=3):
    ts = []
    for i in range(warmup + iters):
        if pre is not None:
            pre()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world, local, dev, _ = init_distributed("cuda")
    import torch.distributed._symmetric_memory as symm
    from midgpt_b200.ops.b200_backend import B200Backend
    from midgpt_b200.parallel.fused import GatherSpec
    be = B200Backend(dev)
    ext = be.ext
    peak, how = peak_flops_per_gpu()
    unit = 50_331_648                      # one 1.5B transformer block (elements), padded to world*1024
    unit = (unit + world * 1024 - 1) // (world * 1024) * (world * 1024)
    shard = unit // world
    grad = symm.empty(unit, dtype=torch.float32, device=dev); grad.normal_()
    lp = symm.empty(unit, dtype=torch.bfloat16, device=dev); lp.zero_()
    lps = symm.empty(shard, dtype=torch.bfloat16, device=dev); lps.normal_()
    pad = symm.empty(1024, dtype=torch.int32, device=dev); pad.zero_()
    torch.cuda.synchronize()
    hg, hl, hs, hp = (symm.rendezvous(t, dist.group.WORLD) for t in (grad, lp, lps, pad))
    gp, lpp, sp, pp = ([int(p) for p in h.buffer_ptrs] for h in (hg, hl, hs, hp))
    mc_g = int(hg.multicast_ptr or 0)
    mc_l = int(hl.multicast_ptr or 0) if mc_g else 0
    out = torch.empty(shard, device=dev); ss = torch.zeros(1, device=dev)
    p32, m, v = (torch.randn(shard, device=dev) for _ in range(3)); v.abs_()
    clip = torch.ones(1, device=dev)
    res = []
    ep = [0]

    def barrier():
        ep[0] += 1
        ext.xrank_barrier(pp, rank, 0, ep[0])

    # reduce-scatter ⊕ mean ⊕ Σg²  (fp32): each GPU must receive (world-1)/world * unit * 4 bytes
    by = (world - 1) * shard * 4
    for name, mc in (("rs_sumsq_p2p", 0), ("rs_sumsq_multimem", mc_g)):
        if name.endswith("multimem") and not mc:
            continue
        t = timeit(lambda: ext.rs_sumsq(gp, mc, out, ss, rank * shard, rank, False, True), pre=barrier)
        res.append({"kernel": name, "ms": t, "nvlink_GBs_per_gpu": by / t / 1e6, "frac_of_770GBs": by / (t * 1e-3) / NVLINK, "frac_of_900GBs": by / (t * 1e-3) / LINK_RATE})
    t = timeit(lambda: dist.reduce_scatter_tensor(out, grad, op=dist.ReduceOp.AVG))
    res.append({"kernel": "nccl_reduce_scatter_fp32", "ms": t, "nvlink_GBs_per_gpu": by / t / 1e6, "frac_of_770GBs": by / (t * 1e-3) / NVLINK, "frac_of_900GBs": by / (t * 1e-3) / LINK_RATE})
    # AdamW ⊕ all-gather push (bf16): each GPU must send (world-1) * shard * 2 bytes
    by = (world - 1) * shard * 2
    for name, mc in (("adamw_allgather_p2p", 0), ("adamw_allgather_multimem", mc_l)):
        if name.endswith("multimem") and not mc:
            continue
        t = timeit(lambda: ext.adamw_allgather(p32, m, v, out, lpp, mc, rank * shard, clip, 1e-4, 0.9, 0.95, 1e-8, 0.1, 0.5, 0.5), pre=barrier)
        res.append({"kernel": name, "ms": t, "nvlink_GBs_per_gpu": by / t / 1e6, "frac_of_770GBs": by / (t * 1e-3) / NVLINK, "frac_of_900GBs": by / (t * 1e-3) / LINK_RATE})
    t = timeit(lambda: dist.all_gather_into_tensor(lp, lp[rank * shard:(rank + 1) * shard]))
    res.append({"kernel": "nccl_all_gather_bf16", "ms": t, "nvlink_GBs_per_gpu": by / t / 1e6, "frac_of_770GBs": by / (t * 1e-3) / NVLINK, "frac_of_900GBs": by / (t * 1e-3) / LINK_RATE})
    # all-gather ⊕ GEMM: a GEMM of the block carries the gather of the WHOLE unit (88 MB remote at N = 8) in its own CTAs.
    #   fwd   QKV projection (N=6144, K=2048): consumes the c_attn rows as they land (X1)
    #   dgrad mlp_proj dgrad ⊕ gelu' (the first dgrad of a block's backward): W_mlp_proj k-rows as they land (X2)
    #   wgrad the first GEMM of a block's backward: reads no weight, simply hides the pull (X2)
    M, D = 16384, 2048
    x, x4 = torch.randn(M, D, device=dev).bfloat16(), torch.randn(M, 4 * D, device=dev).bfloat16()
    local = torch.zeros(unit, dtype=torch.bfloat16, device=dev)
    local[rank * shard:(rank + 1) * shard].copy_(lps)
    flags = torch.zeros(4096, dtype=torch.int32, device=dev)
    calls = [0]
    o_attn, o_proj, o_fc = 0, 3 * D * D, 4 * D * D                        # unit layout: c_attn | attn_proj | c_fc | mlp_proj | gains
    w_attn = local[o_attn:o_attn + 3 * D * D].view(3 * D, D)
    w_mlp_proj = local[o_fc + 4 * D * D:o_fc + 8 * D * D].view(D, 4 * D)
    acc = torch.zeros(D, 4 * D, device=dev)

    def spec():
        calls[0] += 1
        return GatherSpec(sp, local, flags, shard, calls[0], rank)

    cases = {
        "qkv_fwd": (lambda g: be.linear_fwd(x, w_attn, gather=g), 2.0 * M * 3 * D * D),
        "mlp_proj_dgrad_dgelu": (lambda g: be.linear_dgrad(x, w_mlp_proj, gelu_pre=x4, gather=g), 2.0 * M * 4 * D * D),
        "mlp_proj_wgrad": (lambda g: be.linear_wgrad(x, x4, acc, gather=g), 2.0 * M * 4 * D * D),
    }
    by = (world - 1) * shard * 2
    for name, (fn, fl) in cases.items():
        t_plain = timeit(lambda: fn(None))
        t = timeit(lambda: fn(spec()), pre=barrier)
        def nccl_then(fn=fn):
            dist.all_gather_into_tensor(local, lps)
            fn(None)
        t_nccl = timeit(nccl_then)
        target = max(t_plain, by / LINK_RATE * 1e3)
        res.append({"kernel": f"allgather_gemm_{name}", "ms": t, "gemm_only_ms": t_plain, "nccl_all_gather_then_gemm_ms": t_nccl,
                    "target_ms_max(gemm_only, bytes/900GBs)": target, "ratio_to_target": t / target, "speedup_vs_nccl": t_nccl / t,
                    "tflops": fl / t / 1e9, "nvlink_GBs_per_gpu_if_pull_bound": by / t / 1e6})
    # correctness of the carried gather: the local buffer must now equal the all-gathered shards
    want = torch.empty(unit, dtype=torch.bfloat16, device=dev)
    dist.all_gather_into_tensor(want, lps)
    local[: (rank) * shard].zero_(); local[(rank + 1) * shard:].zero_()
    barrier(); be.linear_dgrad(x, w_mlp_proj, gelu_pre=x4, gather=spec()); torch.cuda.synchronize()
    res.append({"kernel": "allgather_gemm_dgrad_gathered_buffer_equal", "ok": bool(torch.equal(local, want))})
    # stand-alone NVLink pull kernel (units whose first consumer is not a GEMM)
    for n_ctas in (16, 32, 64):
        sflags = torch.zeros(8, dtype=torch.int32, device=dev)
        tt = timeit(lambda n_ctas=n_ctas: ext.ag_pull(sp, local, shard, rank, sflags, n_ctas), pre=barrier)
        res.append({"kernel": f"ag_pull_lsu_{n_ctas}ctas", "ms": tt, "nvlink_GBs_per_gpu": by / tt / 1e6, "frac_of_770GBs": by / (tt * 1e-3) / NVLINK, "frac_of_900GBs": by / (tt * 1e-3) / LINK_RATE})
    if rank == 0:
        for r in res:
            print(json.dumps({"n_gpus": world, **{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
