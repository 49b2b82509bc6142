"""Token data: uint16 ``.bin`` memmaps, random-crop batch sampling, synthetic corpora, pinned prefetch.

Reference behaviour (src/train.py:56-66,122-124,132-137): the whole ``train.bin``/``val.bin`` is
copied into host RAM, each *process* keeps a contiguous ``1/n_proc`` slice, and every batch is
``batch_size·G`` uniformly random crops of ``block_size+1`` tokens drawn with the global, unseeded
``np.random``.  Differences here, on purpose:

* one process per GPU ⇒ 8 processes per node; the file stays a shared read-only memmap (page
  cache is shared) instead of 8 private RAM copies;
* the sampler owns a seeded ``numpy.random.Generator``; the prefetching loader snapshots its state together with
  every batch it produces, and checkpoints store — per rank — the state that follows the batch the trainer actually
  *consumed*, so a resumed run continues the same sample stream exactly (the reference restarts it);
* batches are assembled straight into **pinned** host buffers by a background thread (or the
  native C++ gatherer in ``csrc/dataloader.cpp`` when built) and copied H2D on a side stream (X9).
* a synthetic corpus of the right vocabulary/length stands in when no dataset is on disk
  (there is no network on the benchmark hosts).
"""
from __future__ import annotations

import os
import threading
import queue
import typing as tp

import numpy as np
import torch


def split_array_by_idx(arr: np.ndarray, proc_idx: int, n_proc: int) -> np.ndarray:
    """Contiguous per-process slice, same arithmetic as the reference (src/train.py:122-124)."""
    n = int(arr.shape[0] / n_proc) + 1
    return arr[proc_idx * n:(proc_idx + 1) * n]


def load_tokens(data_dir: str, split: str, in_ram: bool = False) -> np.ndarray:
    path = os.path.join(data_dir, f"{split}.bin")
    arr = np.memmap(path, dtype=np.uint16, mode="r")
    return np.array(arr) if in_ram else arr


def synthetic_tokens(n_tokens: int, vocab_size: int, seed: int = 0) -> np.ndarray:
    """A Zipf-ish random token stream with a little local structure, so losses move during smoke runs."""
    rng = np.random.default_rng(seed)
    ranks = rng.zipf(1.3, size=n_tokens).astype(np.int64)
    toks = (ranks - 1) % max(vocab_size, 1)
    # inject a short-range copy pattern: every 4th token repeats the token 2 positions back
    toks[4::4] = toks[2:-2:4][: len(toks[4::4])]
    return toks.astype(np.uint16)


def dataset_available(data_dir: str) -> bool:
    return os.path.exists(os.path.join(data_dir, "train.bin")) and os.path.exists(os.path.join(data_dir, "val.bin"))


class BatchSampler:
    """Uniform random crops: x = data[i : i+T], y = data[i+1 : i+T+1]  (reference ``get_batch``)."""

    def __init__(self, data: np.ndarray, block_size: int, seed: int = 0):
        self.data, self.T = data, block_size
        self.rng = np.random.default_rng(seed)
        self._native = None
        try:
            from .ops import native_loader
            self._native = native_loader.get()
        except Exception:
            self._native = None

    def sample_into(self, x_out: np.ndarray, y_out: np.ndarray) -> None:
        """Fill int32 arrays of shape [N, T] in place (N crops)."""
        n = x_out.shape[0]
        ix = self.rng.integers(0, len(self.data) - self.T, size=(n,))
        if self._native is not None and self.data.dtype == np.uint16:
            self._native.gather_crops(self.data, ix, self.T, x_out, y_out)
            return
        win = np.lib.stride_tricks.sliding_window_view(self.data, self.T + 1)
        rows = win[ix]                        # [n, T+1] uint16 (gather)
        x_out[...] = rows[:, :-1]
        y_out[...] = rows[:, 1:]

    def sample(self, n: int) -> tp.Tuple[np.ndarray, np.ndarray]:
        x = np.empty((n, self.T), dtype=np.int32)
        y = np.empty((n, self.T), dtype=np.int32)
        self.sample_into(x, y)
        return x, y

    def state_dict(self) -> dict:
        return {"bit_generator": self.rng.bit_generator.state}

    def load_state_dict(self, st: dict) -> None:
        self.rng.bit_generator.state = st["bit_generator"]


def get_batch(data: np.ndarray, block_size: int, batch_size: int, g_accum_iters: int | None = None,
              rng: np.random.Generator | None = None) -> tp.Tuple[np.ndarray, np.ndarray]:
    """Functional form with the reference's signature/shape contract: (G, B, T) when G is given else (B, T)."""
    s = BatchSampler(data, block_size)
    if rng is not None:
        s.rng = rng
    x, y = s.sample(batch_size * (g_accum_iters or 1))
    if g_accum_iters is not None:
        x = x.reshape(g_accum_iters, batch_size, block_size)
        y = y.reshape(g_accum_iters, batch_size, block_size)
    return x, y


class PinnedBatchLoader:
    """Double-buffered pinned-host batches + async H2D on a side stream.

    ``next()`` returns device int32 tensors ``(x, y)`` of shape [N, T]; the copy of batch k+1
    overlaps the compute of batch k.  ``h2d_bytes_per_batch`` is what bench.py reports.
    """

    def __init__(self, sampler: BatchSampler, n_seqs: int, device: torch.device, prefetch: int = 2,
                 background: bool = True):
        self.sampler, self.n, self.device = sampler, n_seqs, device
        self.cuda = device.type == "cuda"
        T = sampler.T
        self._bufs = []
        for _ in range(prefetch + 1):
            x = torch.empty((n_seqs, T), dtype=torch.int32)
            y = torch.empty((n_seqs, T), dtype=torch.int32)
            if self.cuda:
                x, y = x.pin_memory(), y.pin_memory()
            self._bufs.append((x, y))
        self.h2d_bytes_per_batch = 2 * n_seqs * T * 4
        self._copy_stream = torch.cuda.Stream(device) if self.cuda else None
        self._free: "queue.Queue[int]" = queue.Queue()
        self._ready: "queue.Queue[int]" = queue.Queue()
        for i in range(len(self._bufs)):
            self._free.put(i)
        self._inflight: tp.List[tp.Tuple[int, tp.Any]] = []
        self._state_after: tp.List[dict | None] = [None] * len(self._bufs)
        self._consumed_state = sampler.state_dict()
        self._background = background
        self._stop = False
        if background:
            self._thread = threading.Thread(target=self._work, daemon=True)
            self._thread.start()

    def _fill(self, i: int) -> None:
        x, y = self._bufs[i]
        self.sampler.sample_into(x.numpy(), y.numpy())
        self._state_after[i] = self.sampler.state_dict()      # RNG state that follows this batch

    def state_dict(self) -> dict:
        """Sampler state to checkpoint: the one following the last batch handed out by ``next()`` — NOT the live state of
        the sampler, which the prefetch thread has already advanced by up to ``prefetch + 1`` batches."""
        return self._consumed_state

    def _work(self) -> None:
        while not self._stop:
            try:
                i = self._free.get(timeout=0.1)
            except queue.Empty:
                continue
            self._fill(i)
            self._ready.put(i)

    def next(self) -> tp.Tuple[torch.Tensor, torch.Tensor]:
        # recycle host buffers whose H2D copy has completed
        still = []
        for i, ev in self._inflight:
            if ev is None or ev.query():
                self._free.put(i)
            else:
                still.append((i, ev))
        self._inflight = still
        if self._background:
            i = self._ready.get()
        else:
            i = self._free.get()
            self._fill(i)
        hx, hy = self._bufs[i]
        self._consumed_state = self._state_after[i]
        if not self.cuda:
            out = (hx.clone(), hy.clone())
            self._free.put(i)
            return out
        with torch.cuda.stream(self._copy_stream):
            dx = hx.to(self.device, non_blocking=True)
            dy = hy.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        torch.cuda.current_stream().wait_event(ev)
        dx.record_stream(torch.cuda.current_stream())
        dy.record_stream(torch.cuda.current_stream())
        self._inflight.append((i, ev))
        return dx, dy

    def close(self) -> None:
        self._stop = True
