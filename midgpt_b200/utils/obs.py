"""Metrics / logging / profiling.

Reference: wandb project "midgpt" with keys ``loss/train``, ``loss/val`` (eval steps) and
``loss/optimized`` (every 20 steps) (src/train.py:200-201,212-213), a tqdm bar with postfix
``train_loss,val_loss,loss,lr,thpt`` (src/train.py:198-220), ``jax.profiler`` trace of step 0 under
``--debug`` (src/train.py:205-211).

Here: the same keys go to ``<rundir>/metrics.jsonl`` always and to wandb when it is importable and
not disabled (benchmark hosts have no network ⇒ ``WANDB_MODE=offline``/``disabled`` honoured);
extra keys the reference lacks: tokens/s, MFU (BASELINE.md formula), grad-norm, step ms, peak HBM.
``--debug`` captures a ``torch.profiler`` chrome trace of the first steps with NVTX-style ranges
named like the reference's ``jax.named_scope``s.
"""
from __future__ import annotations

import json
import os
import time
import typing as tp

import torch

from ..models.params import flops_per_token


def peak_flops_per_gpu() -> tp.Tuple[float, str]:
    """bf16 dense peak used as MFU denominator: measured (MEASURED_PEAKS.json) else the recipe's fallback."""
    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    path = os.path.join(here, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops"]) * 1e12, "measured"
    except Exception:
        return 1.59e15, "fallback"


def mfu(tokens_per_s: float, model_cfg, n_gpus: int, peak: float | None = None) -> float:
    peak = peak or 2.25e15   # nominal dense bf16 (comparable with the reference's nominal-peak MFU)
    return tokens_per_s * flops_per_token(model_cfg) / (peak * n_gpus)


class MetricLogger:
    def __init__(self, rundir: str | None, config_dict: dict | None = None, enabled: bool = True,
                 wandb_id: str | None = None, use_wandb: bool | None = None):
        self.enabled = enabled
        self.path = os.path.join(rundir, "metrics.jsonl") if (rundir and enabled) else None
        self._wandb = None
        if use_wandb is None:
            use_wandb = os.environ.get("WANDB_MODE", "") not in ("disabled",) and bool(os.environ.get("WANDB_API_KEY") or os.environ.get("WANDB_MODE") == "offline")
        if enabled and use_wandb:
            try:
                import wandb
                wandb.init(project="midgpt", id=wandb_id, resume="allow", config=config_dict,
                           dir=rundir or None)
                self._wandb = wandb
            except Exception as e:  # no network, no key, …
                print(f"[obs] wandb disabled: {e}")

    def log(self, metrics: dict, step: int) -> None:
        if not self.enabled:
            return
        if self.path:
            with open(self.path, "a") as f:
                f.write(json.dumps({"step": step, "time": time.time(), **metrics}) + "\n")
        if self._wandb is not None:
            self._wandb.log(metrics, step=step)

    def finish(self) -> None:
        if self._wandb is not None:
            self._wandb.finish()


class StepTimer:
    """CUDA-event step timer (device time on the launching stream); wall clock on CPU."""

    def __init__(self, device: torch.device):
        self.cuda = device.type == "cuda"
        self._t0 = None

    def start(self):
        if self.cuda:
            self._e0 = torch.cuda.Event(enable_timing=True)
            self._e0.record()
        else:
            self._t0 = time.perf_counter()

    def stop_ms(self) -> float:
        if self.cuda:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            e1.synchronize()
            return self._e0.elapsed_time(e1)
        return (time.perf_counter() - self._t0) * 1e3


class DebugProfiler:
    """``--debug``: trace the first ``n_steps`` steps to ``<rundir>/trace_rank<r>.json`` (reference traces step 0)."""

    def __init__(self, rundir: str | None, rank: int, n_steps: int = 2):
        self.dir, self.rank, self.n, self._prof, self._seen = rundir or ".", rank, n_steps, None, 0

    def step_begin(self):
        if self._seen == 0:
            acts = [torch.profiler.ProfilerActivity.CPU]
            if torch.cuda.is_available():
                acts.append(torch.profiler.ProfilerActivity.CUDA)
            self._prof = torch.profiler.profile(activities=acts, record_shapes=False)
            self._prof.__enter__()

    def step_end(self):
        self._seen += 1
        if self._prof is not None and self._seen == self.n:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self._prof.__exit__(None, None, None)
            os.makedirs(self.dir, exist_ok=True)
            out = os.path.join(self.dir, f"trace_rank{self.rank}.json")
            self._prof.export_chrome_trace(out)
            print(f"[debug] wrote profiler trace {out}")
            self._prof = None


class StepWatchdog:
    """Hang detection (the reference has none: a stuck collective on one host blocks a pod forever).

    ``beat()`` is called once per completed optimizer step.  A daemon thread checks that the next beat arrives within
    ``max(min_timeout_s, factor x median step time)``; if not it dumps every Python thread's stack (``faulthandler``), writes
    ``<rundir>/watchdog_rank<r>.txt`` and terminates the process with exit code 3, so that the launcher's retry loop
    (``gpu retry_launch``, ``scripts/cluster.sh``) relaunches and the run resumes from the last checkpoint.  NCCL's own
    timeout covers library collectives; this also covers a spinning in-kernel barrier (a peer that died mid-step) and a
    stalled data loader."""

    def __init__(self, rundir: str | None, rank: int, factor: float = 20.0, min_timeout_s: float = 600.0,
                 on_timeout: tp.Callable[[str], None] | None = None, poll_s: float = 1.0):
        import threading
        self.rundir, self.rank, self.factor, self.min_timeout_s = rundir, rank, factor, min_timeout_s
        self._on_timeout = on_timeout
        self._last = time.monotonic()
        self._durs: tp.List[float] = []
        self._stop = False
        self._poll = poll_s
        self._armed = False                      # armed after the first beat: start-up (init, first eval) is not a step
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def timeout_s(self) -> float:
        if not self._durs:
            return self.min_timeout_s
        med = sorted(self._durs)[len(self._durs) // 2]
        return max(self.min_timeout_s, self.factor * med)

    def beat(self, pause: bool = False) -> None:
        """``pause=True``: something long and legitimate is about to happen (evaluation, checkpoint wait): disarm until the next beat."""
        now = time.monotonic()
        if self._armed and not pause:
            self._durs = (self._durs + [now - self._last])[-50:]
        self._last = now
        self._armed = not pause

    def close(self) -> None:
        self._stop = True

    def _run(self) -> None:
        while not self._stop:
            time.sleep(self._poll)
            if not self._armed:
                continue
            waited = time.monotonic() - self._last
            if waited > self.timeout_s():
                msg = (f"[watchdog] rank {self.rank}: no optimizer step completed for {waited:.0f} s "
                       f"(limit {self.timeout_s():.0f} s) -- dumping stacks and exiting with code 3")
                try:
                    import faulthandler
                    path = os.path.join(self.rundir or ".", f"watchdog_rank{self.rank}.txt")
                    with open(path, "w") as f:
                        f.write(msg + "\n")
                        faulthandler.dump_traceback(file=f, all_threads=True)
                except Exception:
                    pass
                print(msg, flush=True)
                if self._on_timeout is not None:
                    self._on_timeout(msg)
                    self._stop = True
                    return
                os._exit(3)
