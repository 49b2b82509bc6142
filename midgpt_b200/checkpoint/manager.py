"""Checkpoint manager with the reference's Orbax semantics, written by hand (no orbax/tensorstore here).

Reference (src/train.py:139-145,179-187,214-215,224-225): ``ocp.CheckpointManager(rundir,
AsyncCheckpointer(PyTreeCheckpointHandler()), max_to_keep=1, save_interval_steps=eval_interval)``;
``mngr.save(itr, (model_leaves, opt_leaves))`` is *called* every step and the manager skips steps off
the interval; resume = ``latest_step()`` → ``restore`` → ``first_step = latest + 1``.

On-disk layout ([lib]-inferred):

    <rundir>/<step>/default/
        _METADATA            json: tree metadata (key → shape/dtype/kind)
        checkpoint           msgpack aggregate: scalar leaves inline, arrays as "PLACEHOLDER://<key>"
        0.0/ 0.5/ … 1.16/    one Zarr-v2 array per array leaf (see zarrio.py)
    <rundir>/<step>/extra.json   (ours: sampler RNG state, world size — the reference loses these)

Distributed write: for every FSDP unit the fp32 master / m / v shards are all-gathered (a
collective every rank joins), and rank ``unit % world`` writes that unit's chunks — one chunk per
layer for the stacked block leaves — so file IO is spread over all ranks.  Device→host copy is
synchronous, file writing happens on a background thread (async like Orbax's AsyncCheckpointer),
the step directory is committed by an atomic rename once every rank's marker file exists.
Restore maps chunk files with ``np.memmap`` and copies only the byte ranges of the local shard, so
a run may resume on a different world size.
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
    def save(self, step: int, engine, extra: dict | None = None, force: bool = False) -> bool:
        if not (force or self.should_save(step)):
            return False
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
            meta_tree, agg = {}, {}
            for (which, name), key in keys.items():
                shp = shapes[name]
                chunks = shp if name in ("wte", "lm_head") else (1,) + shp[1:]
                zarrio.write_zarray_meta(os.path.join(item, key), shp, chunks, "float32")
                meta_tree[key] = {"kind": "array", "shape": list(shp), "dtype": "float32", "which": which, "leaf": name}
                agg[key] = f"PLACEHOLDER://{key}"
            for i, (kind, src) in enumerate(layout.MODEL_LEAVES):
                if kind == "scalar":
                    agg[layout.model_key(i)] = src(cfg)
                    meta_tree[layout.model_key(i)] = {"kind": "scalar"}
            for key, which, name in layout.opt_layout():
                if which == "count":
                    agg[key] = int(engine.count)
                    meta_tree[key] = {"kind": "scalar", "leaf": f"{name}_count"}
            with open(os.path.join(item, "_METADATA"), "w") as f:
                json.dump({"tree_metadata": meta_tree, "format": "midgpt_b200-orbax-zarr2", "version": 1}, f)
            with open(os.path.join(item, "checkpoint"), "wb") as f:
                f.write(msgpack.packb(agg))
            with open(os.path.join(tmp, "extra.json"), "w") as f:
                json.dump({"extra": extra or {}, "world": world, "count": int(engine.count), "step": step}, f)
        if world > 1:
            torch.distributed.barrier()
        # synchronous part: gather + D2H
        jobs: tp.List[tp.Tuple[str, tp.Tuple[int, ...], np.ndarray]] = []
        for i, u in enumerate(engine.units):
            for which in ("master", "exp_avg", "exp_avg_sq"):
                full = engine.full_unit_state(i, which)            # collective
                if i % world != rank:
                    continue
                for pname, t in ((p.name, full[f"{u.spec.name}.{p.name}"]) for p in u.spec.params):
                    arr = t.detach().to("cpu", torch.float32).numpy().copy()
                    if u.spec.name in ("embed", "head"):
                        jobs.append((os.path.join(item, keys[(which, pname)]), (0,) * arr.ndim, arr))
                    else:
                        l = int(u.spec.name.split(".")[1])
                        jobs.append((os.path.join(item, keys[(which, pname)]), (l,) + (0,) * arr.ndim, arr[None]))
                del full

        def commit():
            try:
                for path, idx, arr in jobs:
                    zarrio.write_chunk(path, idx, arr)
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
        engine.count = int(agg[count_key])
        extra_path = os.path.join(self.dir, str(step), "extra.json")
        if os.path.exists(extra_path):
            with open(extra_path) as f:
                return json.load(f).get("extra", {})
        return {}

    def restore_full_params(self, step: int, cfg, which: str = "master") -> tp.Dict[str, torch.Tensor]:
        """Un-sharded ``{"embed.wte", "block.l.x", "head.lm_head"}`` (for sample.py; no engine needed)."""
        item = os.path.join(self.dir, str(step), "default")
        keys = layout.array_keys()
        leaves = {name: torch.from_numpy(zarrio.read_array(os.path.join(item, keys[(which, name)])))
                  for name in layout.ARRAY_ORDER}
        return P.from_reference_leaves(leaves, cfg.model_config)
