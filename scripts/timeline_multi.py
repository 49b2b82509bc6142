"""Per-rank timeline of steady-state training steps (torch profiler, CUDA activities) + exposed-communication attribution.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/timeline_multi.py [--config openwebtext_xl] [--steps 2]

Writes, per rank, gpurun_out/timeline_rank<r>.json.gz (chrome trace) and prints (rank 0) / saves one summary JSON per rank:
device time per kernel family, and for the communication kernels (rs_sumsq, adamw_allgather, xrank_barrier, NCCL) how much of
their time was EXPOSED, i.e. not overlapped by any compute kernel on another stream of the same GPU."""
import argparse
import gzip
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="openwebtext_xl")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--seqs-per-gpu", type=int, default=128)
    ap.add_argument("--gather-mode", default="auto")
    ap.add_argument("--remat", default="none")
    args = ap.parse_args()
    from midgpt_b200 import load_config
    from midgpt_b200.train import Trainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = load_config(args.config)
    cfg.batch_size, cfg.g_accum_iters, cfg.shard_model = args.seqs_per_gpu * world, 1, True
    cfg.engine.backend, cfg.engine.gather_mode, cfg.engine.remat = "b200", args.gather_mode, args.remat
    tr = Trainer(cfg, device_type="cuda")
    m = tr.mesh
    g = torch.Generator().manual_seed(m.rank)
    V, T = cfg.model_config.vocab_size, cfg.model_config.block_size
    x = torch.randint(0, V, (tr.local_seqs, T), generator=g, dtype=torch.int32).to(m.device)
    y = torch.randint(0, V, (tr.local_seqs, T), generator=g, dtype=torch.int32).to(m.device)
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    os.makedirs("gpurun_out", exist_ok=True)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            tr.train_step(x, y)
        torch.cuda.synchronize()
    path = f"gpurun_out/timeline_rank{m.rank}.json"
    prof.export_chrome_trace(path)
    trace = json.load(open(path))
    ev = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    with gzip.open(path + ".gz", "wt") as f:
        json.dump({"traceEvents": ev}, f)
    os.remove(path)

    def is_comm(n):
        return any(k in n for k in ("rs_sumsq", "adamw_allgather", "xrank_barrier", "xrank_signal", "ag_pull", "nccl", "ncclDevKernel"))

    fam = {}
    for e in ev:
        n = e["name"].split("<")[0].replace("void ", "")
        fam[n] = fam.get(n, 0.0) + e["dur"] / 1e3
    # exposed communication: comm-kernel time not covered by the union of non-comm kernel intervals
    comp = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev if not is_comm(e["name"]))
    merged = []
    for a, b in comp:
        if merged and a <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], b)
        else:
            merged.append([a, b])
    import bisect
    starts = [a for a, _ in merged]

    def covered(a, b):
        tot, i = 0.0, max(0, bisect.bisect_right(starts, a) - 1)
        while i < len(merged) and merged[i][0] < b:
            tot += max(0.0, min(b, merged[i][1]) - max(a, merged[i][0]))
            i += 1
        return tot

    comm = {}
    for e in ev:
        if is_comm(e["name"]):
            n = e["name"].split("<")[0].replace("void ", "")
            a, b = e["ts"], e["ts"] + e["dur"]
            c = comm.setdefault(n, {"launches": 0, "ms": 0.0, "exposed_ms": 0.0})
            c["launches"] += 1
            c["ms"] += e["dur"] / 1e3
            c["exposed_ms"] += (e["dur"] - covered(a, b)) / 1e3
    t0, t1 = min(e["ts"] for e in ev), max(e["ts"] + e["dur"] for e in ev)
    busy = sum(b - a for a, b in merged) / 1e3
    out = {"rank": m.rank, "world": world, "config": args.config, "steps": args.steps, "gather_mode": tr.engine.gather_mode, "comm_path": tr.engine.comm,
           "wall_ms_per_step": (t1 - t0) / 1e3 / args.steps, "compute_busy_ms_per_step": busy / args.steps,
           "comm_kernels_per_step": {k: {kk: (vv / args.steps if kk != "launches" else vv // args.steps) for kk, vv in v.items()} for k, v in comm.items()},
           "top_kernels_ms_per_step": {k: round(v / args.steps, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:14]}}
    json.dump(out, open(f"gpurun_out/timeline_summary_rank{m.rank}.json", "w"), indent=1)
    if m.is_main:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
