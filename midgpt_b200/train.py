"""Training runtime: ``Trainer`` (step / evaluate / fit) — the counterpart of reference src/train.py.

Reference call stack: ``train(config)`` builds mesh → loads data → checkpoint
manager → optimizer → sharded init → resume → loop {eval every ``eval_interval``; sample batch;
``step``; log; save}.  ``Trainer.fit`` keeps that shape; what differs is underneath:

* ``Trainer.train_step`` = explicit micro-batch loop over ``GPTModel.forward_backward`` with the
  FSDP engine providing gathered weights and consuming gradients, then one sharded optimizer step;
  nothing synchronises with the host inside a step (loss / grad-norm stay on the device);
* ``batch_size`` is GLOBAL: each of the ``world`` ranks takes ``batch_size / world`` sequences per
  accumulation step (SURVEY.md §5.6 caveat), split further into engine micro-batches.
"""
from __future__ import annotations

import math
import typing as tp

import torch

from .config import ExperimentConfig
from .data import BatchSampler, PinnedBatchLoader, load_tokens, split_array_by_idx, synthetic_tokens, dataset_available
from .models.gpt import GPTModel
from .models.params import count_params
from .ops import get_backend
from .parallel import dist as D
from .parallel.fsdp import FSDPEngine
from .utils import obs


def auto_micro_batch(cfg: ExperimentConfig, local_seqs: int, device: torch.device) -> int:
    """Largest divisor of ``local_seqs`` whose token count stays ≤ a target that keeps GEMMs large
    but activations modest (32 K tokens on GPU, everything on CPU)."""
    if cfg.engine.micro_batch_size > 0:
        mb = min(cfg.engine.micro_batch_size, local_seqs)
        while local_seqs % mb:          # a remainder would silently drop sequences and mis-scale loss / gradients
            mb -= 1
        return mb
    if device.type != "cuda":
        return local_seqs
    # 32 K tokens per micro-batch for the 1.5B model (≈ 55 GB of stashed activations without recompute); activation bytes per
    # token scale with n_layer x n_embd, so smaller models take proportionally more tokens per micro-step (124M: the whole 128 K
    # tokens of a GPU in ONE micro-step): larger GEMMs waste less of their last wave (N = 768: 384 tiles on 74 CTA pairs = 5.19 -> 6
    # waves at 32 K tokens, 20.8 -> 21 at 128 K), weight gradients are reduce-added once instead of four times, 4x fewer launches.
    mc = cfg.model_config
    scale = max(1.0, min(8.0, (24 * 2048) / float(mc.n_layer * mc.n_embd)))
    target = max(1, int(32768 * scale) // mc.block_size)
    mb = min(local_seqs, target)
    while local_seqs % mb:
        mb -= 1
    return mb


class Trainer:
    def __init__(self, cfg: ExperimentConfig, mesh: D.Mesh | None = None, device_type: str | None = None,
                 shard_model: bool | None = None):
        self.cfg = cfg
        shard = cfg.shard_model if shard_model is None else shard_model
        self.mesh = mesh or D.make_mesh(shard, device_type)
        m = self.mesh
        if cfg.batch_size % m.world:
            raise ValueError(f"global batch_size {cfg.batch_size} not divisible by world size {m.world}")
        self.backend = get_backend(cfg.engine.backend, m.device, fp8=cfg.engine.fp8, fp8_format=cfg.engine.fp8_format)
        self.engine = FSDPEngine(cfg, m, self.backend, seed=cfg.engine.seed)
        self.model = GPTModel(cfg.model_config, self.backend, remat=cfg.engine.remat, nvtx=cfg.debug)
        self.T = cfg.model_config.block_size
        self.seqs_per_g = cfg.batch_size // m.world                 # per rank, per accumulation step
        self.local_seqs = self.seqs_per_g * cfg.g_accum_iters        # per rank, per optimizer step
        self.micro_bs = auto_micro_batch(cfg, self.local_seqs, m.device)
        self.n_micro = self.local_seqs // self.micro_bs
        self.tokens_per_step = cfg.batch_size * cfg.g_accum_iters * self.T   # global
        self.launch_count = 0

    # ------------------------------------------------------------------ one optimizer step
    def train_step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """x, y: int32 device tensors [local_seqs, T].  Returns the global mean loss (device scalar)."""
        eng, m = self.engine, self.mesh
        eng.begin_step()
        grad_scale = 1.0 / (self.local_seqs * self.T)
        loss_sum = torch.zeros((), dtype=torch.float32, device=m.device)
        step_seed = self.cfg.engine.seed * 7_777_777 + eng.count * 97 + m.rank
        for k in range(self.n_micro):
            eng.set_micro(k == self.n_micro - 1)
            xs = x[k * self.micro_bs:(k + 1) * self.micro_bs]
            ys = y[k * self.micro_bs:(k + 1) * self.micro_bs]
            loss_sum = loss_sum + self.model.forward_backward(eng, xs, ys, grad_scale, step_seed, k).float()
        eng.optimizer_step()
        loss = loss_sum * grad_scale
        if m.world > 1:
            D.all_reduce_avg(loss, torch.distributed.group.WORLD, m.backend)
        return loss

    @torch.no_grad()
    def eval_step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        eng, m = self.engine, self.mesh
        eng.begin_step(backward=False)
        tot = torch.zeros((), dtype=torch.float32, device=m.device)
        n = x.shape[0]
        mb = min(self.micro_bs, n)
        for s in range(0, n, mb):
            xs, ys = x[s:s + mb], y[s:s + mb]
            tot = tot + self.model.forward_loss(eng, xs, ys).float() * xs.shape[0]
        loss = tot / n
        if m.world > 1:
            D.all_reduce_avg(loss, torch.distributed.group.WORLD, m.backend)
        return loss

    def evaluate(self, sampler: BatchSampler, num_batches: int | None = None) -> float:
        """Mean loss over 200 (1 in debug) random batches of the full batch size (src/train.py:107-117)."""
        n = num_batches if num_batches is not None else (1 if self.cfg.debug else 200)
        tot = 0.0
        for _ in range(n):
            x, y = sampler.sample(self.seqs_per_g)
            xd = torch.from_numpy(x).to(self.mesh.device)
            yd = torch.from_numpy(y).to(self.mesh.device)
            tot += self.eval_step(xd, yd).item()
        return tot / n

    # ------------------------------------------------------------------ data
    def make_samplers(self, synthetic: bool | None = None) -> tp.Tuple[BatchSampler, BatchSampler]:
        cfg, m = self.cfg, self.mesh
        use_syn = (not dataset_available(cfg.data_dir)) if synthetic is None else synthetic
        if use_syn:
            V = cfg.model_config.vocab_size
            n = max(4 * self.T * max(self.local_seqs, 8), 1 << 20)
            train = synthetic_tokens(n, V, seed=1)
            val = synthetic_tokens(max(n // 8, 4 * self.T), V, seed=2)
            if m.is_main:
                print(f"[data] {cfg.data_dir!r} not found — using a synthetic corpus ({n} tokens, vocab {V})")
        else:
            train, val = load_tokens(cfg.data_dir, "train"), load_tokens(cfg.data_dir, "val")
            if m.is_main:
                print("Raw shapes", train.shape, val.shape)
        train = split_array_by_idx(train, m.rank, m.world) if len(train) > 2 * m.world * (self.T + 1) else train
        val = split_array_by_idx(val, m.rank, m.world) if len(val) > 2 * m.world * (self.T + 1) else val
        seed = cfg.engine.seed
        # the training stream has its own generator; evaluate() on the training split uses a third sampler so that it
        # neither races with the prefetch thread nor perturbs the (checkpointed) training stream
        self.train_eval_sampler = BatchSampler(train, self.T, seed=seed * 1000 + 250 + m.rank)
        return (BatchSampler(train, self.T, seed=seed * 1000 + m.rank),
                BatchSampler(val, self.T, seed=seed * 1000 + 500 + m.rank))

    # ------------------------------------------------------------------ main loop
    def fit(self, synthetic: bool | None = None, max_steps: int | None = None) -> dict:
        from .checkpoint.manager import CheckpointManager
        cfg, m, eng = self.cfg, self.mesh, self.engine
        train_s, val_s = self.make_samplers(synthetic)
        if m.is_main:
            print(f"Model has {count_params(cfg.model_config)} parameters.")
            print(f"[engine] world={m.world} shard_group={m.shard_size} replicas={m.replica_size} backend={self.backend.name} "
                  f"comm={eng.comm} gather={eng.gather_mode} micro_bs={self.micro_bs}x{self.n_micro} remat={cfg.engine.remat}")
        mngr = None
        remote, local_rundir = None, cfg.rundir
        if not cfg.debug and cfg.rundir:
            from .utils.remote import RemoteDir, is_remote
            if is_remote(cfg.rundir):        # object-store run directory: work in the local mirror, push committed steps
                remote = RemoteDir(cfg.rundir)
                local_rundir = remote.local
            mngr = CheckpointManager(local_rundir, max_to_keep=1, save_interval_steps=cfg.eval_interval,
                                     async_save=cfg.engine.ckpt_async, remote=remote)
            mngr.sync_from_remote(m)
            mngr.prepare(eng)
        first_step = 0
        if mngr is not None and mngr.latest_step() is not None:
            extra = mngr.restore(mngr.latest_step(), eng)
            first_step = mngr.latest_step() + 1
            if extra and "sampler" in extra:
                train_s.load_state_dict(extra["sampler"])
                if extra.get("_from_other_rank") and m.rank > 0:
                    # world size changed since the save: only rank 0's stream is known; give every rank its own jump of it
                    train_s.rng = type(train_s.rng)(train_s.rng.bit_generator.jumped(m.rank))
            if m.is_main:
                print(f"[ckpt] resumed from step {first_step - 1}")
        logger = obs.MetricLogger(local_rundir or None, cfg.reference_dict(), enabled=m.is_main,
                                  wandb_id=getattr(self, "wandb_id", None))
        prof = obs.DebugProfiler(local_rundir or ".", m.rank) if cfg.debug else None
        loader = PinnedBatchLoader(train_s, self.local_seqs, m.device)
        timer = obs.StepTimer(m.device)
        watchdog = obs.StepWatchdog(local_rundir or None, m.rank, min_timeout_s=cfg.engine.watchdog_s) if cfg.engine.watchdog_s > 0 else None
        total = cfg.max_steps if max_steps is None else min(cfg.max_steps, max_steps)
        try:
            from tqdm import trange
            pbar = trange(first_step, total, initial=first_step, total=total, disable=not m.is_main)
        except Exception:
            pbar = range(first_step, total)
        postfix: dict = {}
        last = {}
        stop = _StopRequest(m)             # SIGTERM / SIGUSR1: checkpoint at the next step boundary and leave (preemption, `gpu stop`)
        for itr in pbar:
            if itr % cfg.eval_interval == 0:
                if watchdog:
                    watchdog.beat(pause=True)          # 2 x 200 evaluation batches are not a hung step
                tl, vl = self.evaluate(self.train_eval_sampler), self.evaluate(val_s)
                postfix.update(train_loss=tl, val_loss=vl)
                logger.log({"loss/train": tl, "loss/val": vl}, step=itr)
                last.update(train_loss=tl, val_loss=vl)
            x, y = loader.next()
            if prof:
                prof.step_begin()
            timer.start()
            lr = eng.current_lr()
            loss = self.train_step(x, y)
            ms = timer.stop_ms()
            if prof:
                prof.step_end()
            loss_v = loss.item()
            if watchdog:
                watchdog.beat()
            if not math.isfinite(loss_v):
                raise FloatingPointError(f"non-finite loss {loss_v} at step {itr}")
            tps = self.tokens_per_step / (ms * 1e-3)
            if itr % cfg.engine.log_every == 0:
                logger.log({"loss/optimized": loss_v, "lr": lr, "grad_norm": eng.last_grad_norm.item(),
                            "tokens_per_s": tps, "mfu": obs.mfu(tps, cfg.model_config, m.world),
                            "step_ms": ms,
                            "peak_mem_GB": (torch.cuda.max_memory_allocated() / 2 ** 30) if m.device.type == "cuda" else 0.0},
                           step=itr)
            if mngr is not None:
                if mngr.save(itr, eng, extra={"sampler": loader.state_dict()}):
                    logger.log({"ckpt_save_stall_ms": mngr.last_save_stall_ms}, step=itr)
            postfix.update(loss=loss_v, lr=lr, thpt=self.tokens_per_step / self.T / (ms * 1e-3))
            if hasattr(pbar, "set_postfix"):
                pbar.set_postfix(**postfix)
            last.update(loss=loss_v, step=itr, tokens_per_s=tps)
            if stop.agreed(itr):               # every rank sees the same answer at the same step
                if mngr is not None and not mngr.should_save(itr):
                    mngr.save(itr, eng, extra={"sampler": loader.state_dict()}, force=True)
                if m.is_main:
                    print(f"[signal] stop requested: checkpoint of step {itr} {'written' if mngr is not None else 'skipped (no rundir)'}; "
                          f"re-run with the same --rundir to resume at step {itr + 1}", flush=True)
                last["stopped_at"] = itr
                break
        if hasattr(pbar, "close"):
            pbar.close()
        loader.close()
        if watchdog:
            watchdog.close()
        logger.finish()
        if mngr is not None:
            mngr.wait_until_finished()
        if remote is not None and m.is_main:
            remote.push_file("metrics.jsonl")
        return last


class _StopRequest:
    """Cooperative preemption.  SIGTERM / SIGUSR1 on ANY rank sets a flag; every ``every`` steps the ranks agree on it with a
    one-element MAX all-reduce (a signal delivered to one rank only must not desynchronise the collectives), then the loop
    writes a forced checkpoint and returns.  The reference's pods are simply killed (``tpu stop`` = ``pkill -9``) and lose
    everything since the last interval checkpoint."""

    def __init__(self, mesh, every: int = 5):
        import signal
        import threading
        self.mesh, self.every, self.flag = mesh, every, False
        if threading.current_thread() is threading.main_thread():
            for sig in (signal.SIGTERM, signal.SIGUSR1):
                try:
                    signal.signal(sig, self._handler)
                except (ValueError, OSError):
                    pass

    def _handler(self, signum, frame):
        self.flag = True

    def agreed(self, itr: int) -> bool:
        if self.mesh.world == 1:
            return self.flag
        if itr % self.every:
            return False
        t = torch.tensor([1.0 if self.flag else 0.0], device=self.mesh.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return bool(t.item() > 0)


def train(config: ExperimentConfig, **kw) -> dict:
    """Reference-compatible entry point (src/train.py:127)."""
    return Trainer(config, **kw).fit()
