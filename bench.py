#!/usr/bin/env python
"""Headline benchmark: training throughput (tokens/s) of the reference's 1.5B config (``openwebtext_xl``:
L24 H16 D2048 T1024 V50304, bf16 compute / fp32 master, FSDP) — BASELINE.md, reference README.md:55.

    python bench.py                                   # 1 GPU, defaults that finish in minutes
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 \
        bench.py --gpus 8 --steps 10 --warmup 3       # one rank per GPU (what the driver runs)
    python bench.py --impl reference                  # the unmodified reference (unavailable here: no jax)

Weak scaling: every GPU processes ``--seqs-per-gpu`` (default 128) sequences of 1024 tokens per optimizer
step, so N = 8 is exactly the reference's global batch of 1024 sequences.  Each step is a full training
step: forward, backward, gradient reduction, global-norm clip, AdamW, bf16 re-gather.  Synthetic tokens,
random-init weights (no network for datasets/checkpoints).

The same invocation also measures the other BASELINE.json configs under ``"secondary"``: the 124M ``openwebtext`` config
(reference src/configs/openwebtext.py:4-22) run as FSDP with the same weak-scaling layout — the second half of the metric —,
the 1.5B model at T = 4096 with per-block recompute, and the 1.5B model with MX-fp8 forward GEMMs (``--secondary`` selects).

Two numbers:
  * ``value``  — device-timed (CUDA events, max over ranks) with batches already resident on the GPU;
  * ``e2e``    — the same K steps through the public API (``PinnedBatchLoader`` → ``Trainer.train_step`` →
                 ``loss.item()``): H2D copy of every batch from pinned memory and a D2H read of the loss
                 inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PUBLISHED_TOKENS_PER_S = 444_000.0     # reference README.md:55 (TPU v3-128), BASELINE.md


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=4)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="openwebtext_xl")
    p.add_argument("--backend", default="b200", choices=["b200", "torch"])
    p.add_argument("--comm", default="auto")
    p.add_argument("--seqs-per-gpu", type=int, default=128)
    p.add_argument("--micro-bs", type=int, default=0)
    p.add_argument("--remat", default="auto", choices=["auto", "block", "none"])
    p.add_argument("--gather-mode", default="auto")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-secondary", action="store_true", help="skip the secondary configs")
    p.add_argument("--secondary", default="124M,seq4096,fp8", help="secondary BASELINE.json configs measured after the headline")
    p.add_argument("--fp8", action="store_true", help="e4m3 forward GEMMs (separate config; the headline number is bf16)")
    p.add_argument("--fp8-format", default="mx", choices=["mx", "row"])
    p.add_argument("--set", action="append", default=[])
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2]); power.append(float(r[3]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def reference_arm():
    """The reference needs jax/jaxlib/equinox/optax/orbax; none is installed or present in /opt/wheelhouse and the
    repo has no setup.py/pyproject (pip: 'not installable'), see DESIGN.md §Reference arm."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    why = "reference is a Jax/Equinox script tree (no setup.py/pyproject; pip --no-index install fails) and jax/jaxlib/equinox/optax/orbax are not in the image or /opt/wheelhouse"
    try:
        import jax  # noqa: F401
        why = "jax present but reference not installed under baseline/_ref" if not os.path.isdir(ref) else why
    except Exception:
        pass
    if int(os.environ.get("RANK", "0")) == 0:          # one line per job, also under torchrun
        print(json.dumps({"impl": "reference", "unavailable": why}), flush=True)
    return 0


def run_config(args, name: str, steps: int, warmup: int, want_e2e: bool, clocks_on: bool, variant: dict | None = None):
    """Train ``name`` for ``warmup`` + ``steps`` optimizer steps (weak scaling: --seqs-per-gpu sequences per GPU, FSDP over
    all GPUs of the node) and return the measurements of one config as a dict (rank 0) — the headline config and the
    secondary configs go through exactly the same code.  ``variant``: block_size / seqs_per_gpu / remat / fp8 of a BASELINE.json
    config that differs from the preset (seq-4096 stress, block-scaled fp8)."""
    variant = variant or {}
    import torch
    import torch.distributed as dist
    from midgpt_b200 import load_config
    from midgpt_b200.data import BatchSampler, PinnedBatchLoader, synthetic_tokens
    from midgpt_b200.models.params import flops_per_token
    from midgpt_b200.train import Trainer
    from midgpt_b200.utils import obs
    from launch import apply_overrides

    cfg = load_config(name)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    seqs_per_gpu = variant.get("seqs_per_gpu", args.seqs_per_gpu)
    cfg.batch_size = seqs_per_gpu * world
    cfg.g_accum_iters = 1
    cfg.shard_model = True            # BASELINE.json runs every GPU config as FSDP (the 124M preset itself is pure DP)
    cfg.engine.backend = args.backend
    cfg.engine.comm = args.comm
    cfg.engine.gather_mode = args.gather_mode
    cfg.engine.micro_batch_size = args.micro_bs
    cfg.engine.fp8 = args.fp8 or bool(variant.get("fp8"))
    cfg.engine.fp8_format = args.fp8_format
    apply_overrides(cfg, args.set)
    if "block_size" in variant:
        cfg.model_config.block_size = variant["block_size"]
    T = cfg.model_config.block_size
    # full activation stash needs ~70 KB/token for the 1.5 B model; keep it when it fits comfortably in 180 GB
    cfg.engine.remat = variant.get("remat") or ("none" if args.remat == "auto" else args.remat)

    torch.cuda.reset_peak_memory_stats()
    tr = Trainer(cfg, device_type="cuda")
    m = tr.mesh
    dev = m.device
    n_resident = 4
    gen = torch.Generator().manual_seed(1234 + m.rank)
    V = cfg.model_config.vocab_size
    resident = [(torch.randint(0, V, (tr.local_seqs, T), generator=gen, dtype=torch.int32).to(dev),
                 torch.randint(0, V, (tr.local_seqs, T), generator=gen, dtype=torch.int32).to(dev)) for _ in range(n_resident)]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ------------------------------------------------------------------ warm-up (also compiles nothing: no JIT)
    loss = None
    for i in range(warmup):
        x, y = resident[i % n_resident]
        loss = tr.train_step(x, y)
    sync()
    loss_w = loss.item() if warmup else float("nan")
    tr.engine.measure_comm = True
    tr.engine.exposed_comm_ms()                         # reset the accounting: only the timed steps count

    # ------------------------------------------------------------------ device-timed region
    clocks = ClockSampler(m.local_rank)
    if m.is_main and clocks_on:
        clocks.start()
    launches0 = getattr(tr.backend, "launches", 0)
    comm0 = tr.engine.comm_stats()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        x, y = resident[(warmup + i) % n_resident]
        loss = tr.train_step(x, y)
    e1.record()
    sync()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = getattr(tr.backend, "launches", 0) - launches0
    comm1 = tr.engine.comm_stats()
    clk = clocks.stop() if (m.is_main and clocks_on) else {}
    ms_per_step = ms_total / steps
    tokens_per_step = tr.tokens_per_step
    value = tokens_per_step / (ms_per_step * 1e-3)
    final_loss = loss.item()

    # exposed communication: time the compute stream spent waiting for the comm stream / running the un-overlappable
    # tail (Σg² all-reduce, AdamW ⊕ all-gather push, barriers), measured with events around those regions, max over ranks
    comm = {k: comm1[k] - comm0[k] for k in comm1 if isinstance(comm1[k], (int, float))}
    comm_exposed_ms = max_over_ranks(tr.engine.exposed_comm_ms()) / max(steps, 1) if world > 1 else 0.0
    tr.engine.measure_comm = False

    # ------------------------------------------------------------------ end-to-end region (public API, H2D + D2H every step)
    e2e = None
    if want_e2e:
        corpus = synthetic_tokens(max(8 * T * tr.local_seqs, 1 << 22), V, seed=7 + m.rank)
        loader = PinnedBatchLoader(BatchSampler(corpus, T, seed=11 + m.rank), tr.local_seqs, dev)
        x, y = loader.next()
        tr.train_step(x, y).item()                      # one untimed step to prime the loader pipeline
        sync()
        t0 = time.perf_counter()
        lv = float("nan")
        for i in range(steps):
            x, y = loader.next()                        # pinned host -> device copy of this step's batch
            lv = tr.train_step(x, y).item()             # device -> host read of this step's loss
        sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        loader.close()
        e2e = {"value": tokens_per_step * steps / dt, "unit": "tokens/s",
               "h2d_bytes_per_step": loader.h2d_bytes_per_batch * world, "d2h_bytes_per_step": 4 * world,
               "ms_per_step": dt / steps * 1e3, "last_loss": lv}

    peak, how = obs.peak_flops_per_gpu()
    fpt = flops_per_token(cfg.model_config)
    mc = cfg.model_config
    out = {
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "mfu_nominal_2.25PF": value * fpt / (2.25e15 * world),
        f"mfu_{how}_cublas_peak": value * fpt / (peak * world),
        "config": {"model": f"{name} L{mc.n_layer} H{mc.n_head} D{mc.n_embd} T{T} V{V}",
                   "global_batch": cfg.batch_size, "seq_len": T,
                   "parallelism": (f"fsdp{m.shard_size}" if m.shard_size > 1 else "single") + (f"xdp{m.replica_size}" if m.replica_size > 1 else ""),
                   "micro_batch": tr.micro_bs, "n_micro": tr.n_micro, "remat": cfg.engine.remat, "backend": tr.backend.name,
                   "comm": tr.engine.comm, "gather_mode": tr.engine.gather_mode,
                   "l2": "no L2 flush needed: per-step working set (GBs of bf16 weights + activations, 4 rotating batches) >> 126 MB L2"},
        "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons"),
                   "power_w_max": clk.get("power_w_max")},
        "e2e": e2e, "gpu_launches": launches, "loss_after_warmup": loss_w, "final_loss": final_loss,
        "comm_exposed_ms": comm_exposed_ms, "comm_kernel_launches": comm,
        "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30,
    }
    # free everything before the next config: symmetric arenas, pinned buffers, activations
    del tr, resident
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm()
    if args.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    rank = int(os.environ.get("RANK", "0"))

    head = run_config(args, args.config, args.steps, args.warmup, not args.no_e2e, True)
    # The other BASELINE.json configs, measured by the same invocation so that every BENCH / SCALE record carries them:
    #   124M   reference src/configs/openwebtext.py:4-22, run as FSDP (the other half of "tokens/sec and MFU at 124M & 1.5B")
    #   seq4096  1.5B, T = 4096, activation-checkpointed (long-context stress; same tokens per GPU: 32 sequences)
    #   fp8    1.5B with block-scaled (MX e4m3) forward GEMMs -- as far as fp8 exists here (backward GEMMs stay bf16)
    secondary = {}
    keep = ("value", "unit", "ms_per_step", "mfu_nominal_2.25PF", "e2e", "config", "steps", "warmup", "gpu_launches", "comm_exposed_ms",
            "peak_mem_GB", "final_loss")
    todo = [] if (args.config != "openwebtext_xl" or args.no_secondary or args.fp8) else [k for k in args.secondary.split(",") if k]
    specs = {"124M": ("openwebtext_124M", "openwebtext", None, max(args.steps, 6), not args.no_e2e),
             "seq4096": ("openwebtext_xl_seq4096_remat", "openwebtext_xl", {"block_size": 4096, "seqs_per_gpu": max(1, args.seqs_per_gpu // 4), "remat": "block"}, min(args.steps, 4), False),
             "fp8": ("openwebtext_xl_fp8_mx_forward", "openwebtext_xl", {"fp8": True}, min(args.steps, 4), False)}
    for key in todo:
        label, cfg_name, variant, k_steps, e2e_on = specs[key]
        try:
            sec = run_config(args, cfg_name, k_steps, max(args.warmup, 3), e2e_on, False, variant)
            secondary[label] = {k: sec[k] for k in keep}
            if key == "fp8":
                secondary[label]["dtype"] = "fp8-e4m3 MX block-scaled forward GEMMs / bf16 backward"
                secondary[label]["vs_bf16_headline"] = sec["value"] / head["value"]
        except Exception as e:      # (deterministic failures hit every rank alike; the headline line is printed regardless)
            secondary[label] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        fp8 = args.fp8
        out = {"metric": "training throughput, 1.5B openwebtext_xl (tokens/s)" if args.config == "openwebtext_xl" else f"training throughput, {args.config} (tokens/s)"}
        out.update(head)
        out.update({
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": head["value"] / PUBLISHED_TOKENS_PER_S if args.config == "openwebtext_xl" else None,
            "dtype": "bf16" if not fp8 else f"fp8-e4m3 GEMMs ({'MX block-scaled' if args.fp8_format == 'mx' else 'row-scaled'})",
            "data": "synthetic tokens, random-init weights", "impl": "ours", "secondary": secondary,
        })
        print(json.dumps(out))
    if world > 1 and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
