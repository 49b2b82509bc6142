"""Checkpoint manager with the reference's Orbax semantics, written by hand (no orbax/tensorstore here).

Reference (src/train.py:139-145,179-187,214-215,224-225): ``ocp.CheckpointManager(rundir,
AsyncCheckpointer(PyTreeCheckpointHandler()), max_to_keep=1, save_interval_steps=eval_interval)``;
``mngr.save(itr, (model_leaves, opt_leaves))`` is *called* every step and the manager skips steps off
the interval; resume = ``latest_step()`` → ``restore`` → ``first_step = latest + 1``.

On-disk layout (schema provenance in ``layout.py``):

    <rundir>/<step>/default/
        _METADATA            json: Orbax ``tree_metadata`` (one entry per leaf: key tuple, key types, value type)
        checkpoint           msgpack aggregate of the nested ``(model_leaves, opt_leaves)`` tree: scalar leaves inline,
                             array leaves as "PLACEHOLDER://<key>"
        0.0/ 0.5/ … 1.16/    one Zarr-v2 array per array leaf (see zarrio.py)
    <rundir>/<step>/extra.json, extra_rank<r>.json   (ours: per-rank sampler RNG state, world size — the reference loses these)

Distributed write — **no collective moves tensor data**: every rank copies only its own fp32 master / m / v shards to
pinned host memory on a side stream (the next optimizer step waits for that copy on the device; the host does not), and a
background thread writes exactly the byte ranges the rank owns into the (uncompressed, raw little-endian) Zarr chunk files
with ``pwrite`` — a chunk is one layer of one leaf, and a rank's flat shard covers a contiguous byte range of a few of
them.  The step directory is committed by an atomic rename once every rank's marker file exists.  Restore maps chunk files
with ``np.memmap`` and copies only the byte ranges of the local shard, so a run may resume on a different world size.
``save()`` returns after *enqueueing* the device→host copies; its stall is reported as ``last_save_stall_ms``.
"""
from __future__ import annotations

import json
import os
import shutil
import threading
import time
import typing as tp

import msgpack
import numpy as np
import torch

from ..models import params as P
from . import layout, zarrio

TMP_SUFFIX = ".orbax-checkpoint-tmp"


class CheckpointManager:
    def __init__(self, directory: str, max_to_keep: int | None = 1, save_interval_steps: int = 1,
                 async_save: bool = True, remote=None):
        """``remote``: a ``utils.remote.RemoteDir`` whose local mirror is ``directory`` — committed steps are pushed to it
        (rank 0, from the commit thread) and ``sync_from_remote`` pulls the newest one before a resume."""
        self.remote = remote
        self.dir = os.path.abspath(directory)
        self.max_to_keep, self.interval, self.async_save = max_to_keep, max(1, save_interval_steps), async_save
        self._thread: threading.Thread | None = None
        self._error: BaseException | None = None
        os.makedirs(self.dir, exist_ok=True)

    # ------------------------------------------------------------------ queries
    def all_steps(self) -> tp.List[int]:
        out = []
        for name in os.listdir(self.dir):
            if name.isdigit() and os.path.isdir(os.path.join(self.dir, name, "default")):
                out.append(int(name))
        return sorted(out)

    def latest_step(self) -> int | None:
        s = self.all_steps()
        return s[-1] if s else None

    def sync_from_remote(self, mesh) -> None:
        """Make the newest remote step available in the local mirror (one rank per node downloads; collective when world > 1)."""
        if self.remote is None:
            return
        if mesh.local_rank == 0:
            self.remote.pull_latest_step()
        if mesh.world > 1:
            torch.distributed.barrier()

    def should_save(self, step: int) -> bool:
        return step % self.interval == 0

    # ------------------------------------------------------------------ save
    def _host_buffer(self, n: int, cuda: bool) -> torch.Tensor:
        """One reusable (pinned, on CUDA) fp32 staging buffer; the previous save's writer thread has been joined."""
        buf = getattr(self, "_hbuf", None)
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=torch.float32, pin_memory=cuda)
            self._hbuf = buf
        return buf[:n]

    def prepare(self, engine) -> None:
        """Allocate the pinned staging buffer up front (page-locking 3 x the local fp32 shards takes seconds for a 1.5B model
        on one GPU): called once before the training loop so that no ``save()`` pays for it."""
        self._host_buffer(3 * sum(u.shard for u in engine.units), engine.cuda)

    def save(self, step: int, engine, extra: dict | None = None, force: bool = False) -> bool:
        if not (force or self.should_save(step)):
            return False
        t_begin = time.perf_counter()
        self.wait_until_finished()
        mesh = engine.mesh
        rank, world = mesh.rank, mesh.world
        final = os.path.join(self.dir, str(step))
        tmp = final + TMP_SUFFIX
        item = os.path.join(tmp, "default")
        cfg = engine.cfg
        mc = cfg.model_config
        L = mc.n_layer
        keys = layout.array_keys()
        shapes = {"wte": (mc.vocab_size, mc.n_embd), "lm_head": (mc.vocab_size, mc.n_embd)}
        shapes.update({k: (L,) + v for k, v in P.block_param_shapes(mc).items()})
        if rank == 0:
            if os.path.exists(tmp):
                shutil.rmtree(tmp)
            os.makedirs(item)
            scalars: tp.Dict[str, tp.Any] = {}
            for i, (kind, src) in enumerate(layout.MODEL_LEAVES):
                if kind == "scalar":
                    scalars[layout.model_key(i)] = src(cfg)
            for key, which, name in layout.opt_layout():
                if which == "count":
                    scalars[key] = int(engine.count)
            for (which, name), key in keys.items():
                shp = shapes[name]
                chunks = shp if name in ("wte", "lm_head") else (1,) + shp[1:]
                zarrio.write_zarray_meta(os.path.join(item, key), shp, chunks, "float32")
            with open(os.path.join(item, "_METADATA"), "w") as f:
                json.dump(layout.tree_metadata(set(keys.values()), scalars), f)
            with open(os.path.join(item, "checkpoint"), "wb") as f:
                f.write(msgpack.packb(layout.aggregate_tree(set(keys.values()), scalars)))
            with open(os.path.join(tmp, "extra.json"), "w") as f:
                json.dump({"extra": extra or {}, "world": world, "count": int(engine.count), "step": step}, f)
        if world > 1:
            torch.distributed.barrier()          # the directory skeleton exists before anybody writes into it
        with open(os.path.join(tmp, f"extra_rank{rank}.json"), "w") as f:
            json.dump({"extra": extra or {}, "world": world}, f)

        # ---- device -> pinned host, this rank's shards only, on a side stream; the optimizer waits for it, the host does not
        sr = engine.sr
        cuda = engine.cuda
        total = 3 * sum(u.shard for u in engine.units)
        host = self._host_buffer(total, cuda)
        jobs: tp.List[tp.Tuple[str, tp.Tuple[int, ...], int, int, int]] = []   # (leaf dir, chunk idx, byte offset in chunk, host off, n)
        if cuda:
            if getattr(self, "_stream", None) is None:
                self._stream = torch.cuda.Stream(engine.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            self._stream.wait_event(ready)
        off = 0
        for u in engine.units:
            lo, hi = sr * u.shard, (sr + 1) * u.shard
            layer = None if u.spec.name in ("embed", "head") else int(u.spec.name.split(".")[1])
            for which, src in (("master", u.master), ("exp_avg", u.exp_avg), ("exp_avg_sq", u.exp_avg_sq)):
                dst = host[off:off + u.shard]
                if cuda:
                    with torch.cuda.stream(self._stream):
                        dst.copy_(src if src.dtype == torch.float32 else src.float(), non_blocking=True)
                else:
                    dst.copy_(src.detach().float())
                for p in u.spec.params:
                    a, b = max(lo, p.offset), min(hi, p.offset + p.numel)
                    if a >= b:
                        continue
                    nd = len(shapes[p.name])
                    idx = (0,) * nd if layer is None else (layer,) + (0,) * (nd - 1)
                    jobs.append((os.path.join(item, keys[(which, p.name)]), idx, (a - p.offset) * 4, off + (a - lo), b - a))
                off += u.shard
        done = None
        if cuda:
            done = torch.cuda.Event()
            done.record(self._stream)
            engine.defer_state_update_until(done)       # the next AdamW sweep must not overwrite shards still being copied
        chunk_bytes = {name: 4 * int(np.prod(shapes[name] if name in ("wte", "lm_head") else shapes[name][1:])) for name in shapes}
        key_to_name = {key: name for (which, name), key in keys.items()}

        def commit():
            try:
                if done is not None:
                    done.synchronize()
                hv = host.numpy()
                for path, idx, boff, hoff, n in jobs:
                    zarrio.write_chunk_range(path, idx, boff, hv[hoff:hoff + n], chunk_bytes[key_to_name[os.path.basename(path)]])
                with open(os.path.join(tmp, f".done_rank{rank}"), "w") as f:
                    f.write("ok")
                if rank == 0:
                    deadline = time.time() + 3600
                    while not all(os.path.exists(os.path.join(tmp, f".done_rank{r}")) for r in range(world)):
                        if time.time() > deadline:
                            raise TimeoutError("checkpoint commit: ranks did not finish writing")
                        time.sleep(0.01)
                    for r in range(world):
                        os.remove(os.path.join(tmp, f".done_rank{r}"))
                    if os.path.exists(final):
                        shutil.rmtree(final)
                    os.replace(tmp, final)
                    self._prune()
                    if self.remote is not None:
                        self.remote.push_step(step, self.max_to_keep)
                        self.remote.push_file("metrics.jsonl")
            except BaseException as e:  # surfaced by wait_until_finished
                self._error = e

        if self.async_save:
            self._thread = threading.Thread(target=commit, daemon=False)
            self._thread.start()
        else:
            commit()
            self._raise()
        self.last_save_stall_ms = (time.perf_counter() - t_begin) * 1e3
        return True

    def _prune(self) -> None:
        if self.max_to_keep is None:
            return
        steps = self.all_steps()
        for s in steps[: max(0, len(steps) - self.max_to_keep)]:
            shutil.rmtree(os.path.join(self.dir, str(s)), ignore_errors=True)

    def _raise(self) -> None:
        if self._error is not None:
            e, self._error = self._error, None
            raise e

    def wait_until_finished(self) -> None:
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        self._raise()

    # ------------------------------------------------------------------ restore
    def restore(self, step: int, engine) -> dict:
        self.wait_until_finished()
        if engine.mesh.world > 1:
            torch.distributed.barrier()      # rank 0's commit thread must have renamed the directory
        item = os.path.join(self.dir, str(step), "default")
        keys = layout.array_keys()
        with open(os.path.join(item, "checkpoint"), "rb") as f:
            agg = msgpack.unpackb(f.read())
        metas: dict = {}

        def chunk(which: str, leaf: str, l: int | None) -> np.ndarray:
            path = os.path.join(item, keys[(which, leaf)])
            if path not in metas:
                metas[path] = zarrio.read_meta(path)
            m = metas[path]
            nd = len(m["shape"])
            if l is None:
                return zarrio.open_chunk(path, (0,) * nd, m)
            return zarrio.open_chunk(path, (l,) + (0,) * (nd - 1), m)[0]

        for i, u in enumerate(engine.units):
            l = None if u.spec.name in ("embed", "head") else int(u.spec.name.split(".")[1])
            for which in ("master", "exp_avg", "exp_avg_sq"):
                engine.load_unit_shard_np(i, which, lambda pname, which=which, l=l: chunk(which, pname, l))
        count_key = [k for k, w, n in layout.opt_layout() if w == "count"][0]
        engine.count = int(layout.aggregate_lookup(agg, count_key))
        # per-rank extra (sampler RNG) when the world size is unchanged, rank 0's otherwise (flagged, so the caller can
        # de-correlate the streams instead of replaying rank 0's on every rank)
        sdir = os.path.join(self.dir, str(step))
        mine = os.path.join(sdir, f"extra_rank{engine.mesh.rank}.json")
        if os.path.exists(mine):
            with open(mine) as f:
                d = json.load(f)
            if d.get("world") == engine.mesh.world:
                return d.get("extra", {})
        extra_path = os.path.join(sdir, "extra.json")
        if os.path.exists(extra_path):
            with open(extra_path) as f:
                ex = dict(json.load(f).get("extra", {}))
            ex["_from_other_rank"] = True
            return ex
        return {}

    def restore_full_params(self, step: int, cfg, which: str = "master") -> tp.Dict[str, torch.Tensor]:
        """Un-sharded ``{"embed.wte", "block.l.x", "head.lm_head"}`` (for sample.py; no engine needed)."""
        item = os.path.join(self.dir, str(step), "default")
        keys = layout.array_keys()
        leaves = {name: torch.from_numpy(zarrio.read_array(os.path.join(item, keys[(which, name)])))
                  for name in layout.ARRAY_ORDER}
        return P.from_reference_leaves(leaves, cfg.model_config)
