"""Per-kernel micro-benchmarks (CUDA events, warm-up, L2-sized rotation) vs the library equivalents.

    python bench/kernel_bench.py [--quick]    → one JSON line per kernel/shape on stdout
Roofline denominators come from MEASURED_PEAKS.json (cuBLAS bf16 burst, STREAM copy).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from midgpt_b200.ops.b200_backend import B200Backend, QKV  # noqa: E402
from midgpt_b200.utils.obs import peak_flops_per_gpu  # noqa: E402


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()          # write a buffer larger than L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_epilogue_bench(be, dev, flush, peak):
    """Every epilogue variant of the bf16 GEMM on the block shapes of both headline models (M = 32768 tokens)."""
    r = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16)
    M = 32768
    for model, D in (("124M", 768), ("1.5B", 2048)):
        x, x4 = r(M, D), r(M, 4 * D)
        w_fc, w_pr = r(4 * D, D) * D ** -0.5, r(D, 4 * D) * (4 * D) ** -0.5
        w_qkv, w_o = r(3 * D, D) * D ** -0.5, r(D, D) * D ** -0.5
        acc_fc = torch.zeros(4 * D, D, device=dev)
        cases = {
            "store  c_attn   fwd": (lambda: be.linear_fwd(x, w_qkv), 2.0 * M * 3 * D * D),
            "gelu   c_fc     fwd": (lambda: be.linear_fwd(x, w_fc, act="gelu"), 2.0 * M * 4 * D * D),
            "resid  mlp_proj fwd": (lambda: be.linear_fwd(x4, w_pr, res=x), 2.0 * M * 4 * D * D),
            "resid  attn_proj fwd": (lambda: be.linear_fwd(x, w_o, res=x), 2.0 * M * D * D),
            "store  c_fc     dgrad": (lambda: be.linear_dgrad(x4, w_fc), 2.0 * M * 4 * D * D),
            "dgelu  mlp_proj dgrad": (lambda: be.linear_dgrad(x, w_pr, gelu_pre=x4), 2.0 * M * 4 * D * D),
            "f32acc c_fc     wgrad": (lambda: be.linear_wgrad(x4, x, acc_fc), 2.0 * M * 4 * D * D),
        }
        for name, (fn, fl) in cases.items():
            t = timeit(fn, flush=flush)
            print(json.dumps({"kernel": "gemm_epilogue", "model": model, "case": name, "ms": round(t, 4), "tflops": round(fl / t / 1e9, 1),
                              "frac_of_cublas_peak": round(fl / (t * 1e-3) / peak, 3)}), flush=True)


def dq_fuse_bench(be, dev, flush):
    """attention backward + QK-LN/RoPE backward: dq through bf16 (zero-fill + kernel + convert + qkln) vs dq left in the fp32
    accumulator and consumed by qkln_rope_bwd."""
    r = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16)
    for (B, T, H, C) in [(32, 1024, 16, 128), (32, 1024, 12, 64)]:
        D = H * C
        h = QKV(r(B * T, 2 * D), r(B * T, 3 * D))
        o, lse = be.attention_fwd(h, B, T, H, C)
        do = r(B * T, D)
        gq, gk = torch.ones(C, device=dev), torch.ones(C, device=dev)
        _, stats = be.qkln_rope_fwd(h.raw, gq, gk, B, T, H, C)
        dgq, dgk = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        res = {}
        for fuse in (False, True):
            def fn(fuse=fuse):
                d = be.attention_bwd(do, h, o, lse, B, T, H, C, fuse_dq=fuse)
                be.qkln_rope_bwd(d, h.raw, stats, gq, gk, dgq, dgk, B, T, H, C)
            res["fused_ms" if fuse else "unfused_ms"] = round(timeit(fn, flush=flush), 4)
        print(json.dumps({"kernel": "attn_bwd+qkln_bwd", "B": B, "T": T, "H": H, "C": C, **res}), flush=True)


def main():
    quick = "--quick" in sys.argv
    if "--dq-fuse" in sys.argv:
        dev = torch.device("cuda", 0)
        return dq_fuse_bench(B200Backend(dev), dev, torch.empty(256 << 20, dtype=torch.uint8, device=dev))
    if "--gemm-epi" in sys.argv:
        dev = torch.device("cuda", 0)
        peak, _ = peak_flops_per_gpu()
        return gemm_epilogue_bench(B200Backend(dev), dev, torch.empty(256 << 20, dtype=torch.uint8, device=dev), peak)
    dev = torch.device("cuda", 0)
    be = B200Backend(dev)
    peak, how = peak_flops_per_gpu()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    r = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16)
    M = 8192 if quick else 32768
    shapes = [("c_attn", M, 6144, 2048), ("attn_proj", M, 2048, 2048), ("c_fc", M, 8192, 2048), ("mlp_proj", M, 2048, 8192)]
    for name, m, n, k in shapes:
        x, w = r(m, k), r(n, k) * k ** -0.5
        dy = r(m, n)
        acc = torch.zeros(n, k, device=dev)
        fl = 2.0 * m * n * k
        rows = {}
        rows["fwd"] = timeit(lambda: be.linear_fwd(x, w), flush=flush)
        rows["fwd_cublas"] = timeit(lambda: x @ w.t(), flush=flush)
        rows["dgrad"] = timeit(lambda: be.linear_dgrad(dy, w), flush=flush)
        rows["dgrad_cublas"] = timeit(lambda: dy @ w, flush=flush)
        rows["wgrad"] = timeit(lambda: be.linear_wgrad(dy, x, acc), flush=flush)
        rows["wgrad_cublas"] = timeit(lambda: dy.t() @ x, flush=flush)
        print(json.dumps({"kernel": "gemm", "layer": name, "M": m, "N": n, "K": k,
                          **{f"{a}_ms": round(b, 4) for a, b in rows.items()},
                          **{f"{a}_tflops": round(fl / b / 1e9, 1) for a, b in rows.items()},
                          "frac_of_%s_cublas_peak" % how: {a: round(fl / (b * 1e-3) / peak, 3) for a, b in rows.items() if "cublas" not in a}}), flush=True)
    # FP8 forward GEMMs (quantisation of the activation included; weights are quantised once per optimizer step)
    be_row, be_mx = B200Backend(dev, fp8=True, fp8_format="row"), B200Backend(dev, fp8=True, fp8_format="mx")
    for name, m, n, k in shapes:
        x, w = r(m, k), r(n, k) * k ** -0.5
        fl = 2.0 * m * n * k
        t_row = timeit(lambda: be_row.linear_fwd(x, w), flush=flush)
        t_mx = timeit(lambda: be_mx.linear_fwd(x, w), flush=flush)
        xq, sx = be_mx.ext.quant_mx_e4m3(x)
        wq, sw = be_mx.ext.quant_mx_e4m3(w)
        y = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        t_mx_gemm = timeit(lambda: be_mx.ext.gemm_mx(xq, sx, wq, sw, y, 0, None, None), flush=flush)
        print(json.dumps({"kernel": "gemm_fp8", "layer": name, "M": m, "N": n, "K": k, "row_scaled_fwd_ms": round(t_row, 4),
                          "mx_fwd_ms": round(t_mx, 4), "mx_gemm_only_ms": round(t_mx_gemm, 4),
                          "row_scaled_tflops": round(fl / t_row / 1e9, 1), "mx_tflops": round(fl / t_mx / 1e9, 1),
                          "mx_gemm_only_tflops": round(fl / t_mx_gemm / 1e9, 1)}), flush=True)
    for (B, T, H, C) in ([(4, 1024, 16, 128)] if quick else [(16, 1024, 16, 128), (16, 1024, 12, 64), (4, 4096, 16, 128)]):
        D = H * C
        raw, qk = r(B * T, 3 * D), r(B * T, 2 * D)
        h = QKV(qk, raw)
        o, lse = be.attention_fwd(h, B, T, H, C)
        do = r(B * T, D)
        fl_f = 4.0 * B * H * T * T * C / 2
        t_f = timeit(lambda: be.attention_fwd(h, B, T, H, C), flush=flush)
        t_b = timeit(lambda: be.attention_bwd(do, h, o, lse, B, T, H, C), flush=flush)
        q, k, v = (t.transpose(1, 2) for t in B200Backend._views(h, B, T, H, C))
        t_lib = timeit(lambda: torch.ops.aten._scaled_dot_product_flash_attention(q, k, v, 0.0, True, False), flush=flush)
        print(json.dumps({"kernel": "attention", "B": B, "T": T, "H": H, "C": C, "fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4),
                          "aten_flash_fwd_ms": round(t_lib, 4), "fwd_tflops": round(fl_f / t_f / 1e9, 1),
                          "bwd_tflops": round(2.5 * fl_f / t_b / 1e9, 1), "aten_fwd_tflops": round(fl_f / t_lib / 1e9, 1)}), flush=True)
    # memory-bound kernels vs the measured copy bandwidth
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
    except Exception:
        hbm = 6.65e12
    Mtok, D = (8192 if quick else 32768), 2048
    x, dy = r(Mtok, D), r(Mtok, D)
    t = timeit(lambda: be.rmsnorm_fwd(x, 1e-6), flush=flush)
    print(json.dumps({"kernel": "rmsnorm_fwd", "ms": round(t, 4), "GBs": round(2 * x.numel() * 2 / t / 1e6, 1), "frac_of_copy_bw": round(2 * x.numel() * 2 / (t * 1e-3) / hbm, 3)}), flush=True)
    y, rstd = be.rmsnorm_fwd(x, 1e-6)
    t = timeit(lambda: be.rmsnorm_bwd(dy, x, rstd, dy), flush=flush)
    print(json.dumps({"kernel": "rmsnorm_bwd", "ms": round(t, 4), "GBs": round(4 * x.numel() * 2 / t / 1e6, 1), "frac_of_copy_bw": round(4 * x.numel() * 2 / (t * 1e-3) / hbm, 3)}), flush=True)
    n = 1 << 27
    p, mm, vv, g = (torch.randn(n, device=dev) for _ in range(4))
    vv.abs_()
    lp = torch.empty(n, device=dev, dtype=torch.bfloat16)
    clip = torch.ones(1, device=dev)
    t = timeit(lambda: be.adamw_step(p, mm, vv, g, lp, lr=1e-3, b1=0.9, b2=0.95, eps=1e-8, wd_over_lr_peak=0.1, bc1=0.5, bc2=0.5, grad_scale=clip))
    by = n * (4 * 4 + 3 * 4 + 2)
    print(json.dumps({"kernel": "adamw", "numel": n, "ms": round(t, 4), "GBs": round(by / t / 1e6, 1), "frac_of_copy_bw": round(by / (t * 1e-3) / hbm, 3)}), flush=True)


if __name__ == "__main__":
    main()
