"""GPU-busy probe: profile ONE training step of the bench configuration with the torch profiler (CUDA activity only)
and report how much of the step the device was idle, plus the largest gaps and which kernel preceded them.
    python scripts/busy_probe.py [--config openwebtext_xl] [--seqs 128] [--micro-bs 32]
Timing under a profiler is not a bench number; the idle FRACTION and the gap list are what this is for.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="openwebtext_xl")
    ap.add_argument("--seqs", type=int, default=128)
    ap.add_argument("--micro-bs", type=int, default=32)
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    from midgpt_b200 import load_config
    from midgpt_b200.train import Trainer

    cfg = load_config(args.config)
    cfg.batch_size = args.seqs
    cfg.g_accum_iters = 1
    cfg.engine.backend = "b200"
    cfg.engine.micro_batch_size = args.micro_bs
    cfg.engine.remat = "none"
    tr = Trainer(cfg, device_type="cuda")
    dev = tr.mesh.device
    T, V = cfg.model_config.block_size, cfg.model_config.vocab_size
    x = torch.randint(0, V, (tr.local_seqs, T), dtype=torch.int32, device=dev)
    y = torch.randint(0, V, (tr.local_seqs, T), dtype=torch.int32, device=dev)
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.train_step(x, y)
        torch.cuda.synchronize()
    evs = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                  if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    busy, cur_end, gaps = 0.0, t0, []
    prev = "<start>"
    for s, e, name in evs:
        if s > cur_end:
            gaps.append((s - cur_end, prev, name))
            busy += e - s
            cur_end = e
        elif e > cur_end:
            busy += e - cur_end
            cur_end = e
        prev = name
    gaps.sort(reverse=True)
    out = {"span_ms": (t1 - t0) / 1e3, "busy_ms": busy / 1e3, "idle_frac": 1 - busy / (t1 - t0), "kernels": len(evs),
           "gap_total_ms": sum(g[0] for g in gaps) / 1e3, "n_gaps_over_20us": sum(1 for g in gaps if g[0] > 20),
           "top_gaps": [{"us": round(g[0], 1), "after": g[1][:60], "before": g[2][:60]} for g in gaps[:12]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
