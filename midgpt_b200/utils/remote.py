"""Run directories on an object store (``gs://…``, ``s3://…``, anything fsspec can open).

The reference writes its run directory — ``config.json``, ``wandb_id.txt`` and the Orbax checkpoints — straight to GCS
through gcsfs when ``--rundir`` starts with ``gs://`` (launch.py:43-57, src/train.py:139-145).  Here every file is first
written to a LOCAL MIRROR (checkpoint chunks are memory-mapped and committed with an atomic rename, which object stores
cannot do) and then pushed; on start-up the small files and the latest committed checkpoint step are pulled back, so
``launch.py --rundir gs://bucket/run`` twice still means "resume".  A checkpoint step is uploaded under
``<step>.orbax-checkpoint-tmp`` and renamed once complete, so a reader never sees a partial step.

Plain paths are not remote; ``file://`` and ``memory://`` URLs are (that is how the CPU tests exercise this module
without a network).  gcsfs / s3fs are optional: without them fsspec raises ImportError for that scheme and the caller gets
a message saying which package is missing.
"""
from __future__ import annotations

import hashlib
import os
import tempfile
import typing as tp

TMP_SUFFIX = ".orbax-checkpoint-tmp"


def is_remote(path: str | None) -> bool:
    return bool(path) and "://" in path


def default_mirror(url: str) -> str:
    root = os.environ.get("MIDGPT_MIRROR_DIR") or os.path.join(tempfile.gettempdir(), "midgpt_b200_mirror")
    return os.path.join(root, hashlib.sha1(url.encode()).hexdigest()[:16])


class RemoteDir:
    def __init__(self, url: str, local: str | None = None):
        try:
            import fsspec
            self.fs, self.root = fsspec.core.url_to_fs(url)
        except ImportError as e:
            scheme = url.split("://", 1)[0]
            raise SystemExit(f"{scheme}:// run directories need the fsspec driver for that scheme "
                             f"(gs -> gcsfs, s3 -> s3fs), which is not installed here: {e}") from e
        self.url = url.rstrip("/")
        self.root = self.root.rstrip("/")
        self.local = os.path.abspath(local or default_mirror(self.url))
        os.makedirs(self.local, exist_ok=True)
        self.fs.makedirs(self.root, exist_ok=True)

    # ------------------------------------------------------------------ paths
    def _r(self, *parts: str) -> str:
        return "/".join((self.root,) + parts)

    def _l(self, *parts: str) -> str:
        return os.path.join(self.local, *parts)

    # ------------------------------------------------------------------ small files (config.json, wandb_id.txt, metrics.jsonl)
    def pull_file(self, name: str) -> bool:
        """Remote → mirror if the remote copy exists; returns whether the mirror now has the file."""
        if self.fs.exists(self._r(name)):
            self.fs.get_file(self._r(name), self._l(name))
        return os.path.exists(self._l(name))

    def push_file(self, name: str) -> None:
        if os.path.exists(self._l(name)):
            self.fs.put_file(self._l(name), self._r(name))

    # ------------------------------------------------------------------ checkpoint steps
    def steps(self) -> tp.List[int]:
        out = []
        for p in self.fs.ls(self.root, detail=False):
            name = p.rstrip("/").rsplit("/", 1)[-1]
            if name.isdigit() and self.fs.exists(self._r(name, "default", "checkpoint")):
                out.append(int(name))
        return sorted(out)

    def pull_step(self, step: int) -> None:
        """Download one committed step into the mirror (no-op when the mirror already has it)."""
        dst = self._l(str(step))
        if os.path.isdir(os.path.join(dst, "default")):
            return
        tmp = dst + ".pull-tmp"
        self.fs.get(self._r(str(step)) + "/", tmp + "/", recursive=True)
        os.replace(tmp, dst)

    def pull_latest_step(self) -> int | None:
        s = self.steps()
        if not s:
            return None
        self.pull_step(s[-1])
        return s[-1]

    def push_step(self, step: int, max_to_keep: int | None = 1) -> None:
        """Upload a committed mirror step; visible under its final name only when complete; then prune old remote steps."""
        src, final, tmp = self._l(str(step)), self._r(str(step)), self._r(str(step) + TMP_SUFFIX)
        if self.fs.exists(tmp):
            self.fs.rm(tmp, recursive=True)
        self.fs.put(src + "/", tmp + "/", recursive=True)
        if self.fs.exists(final):
            self.fs.rm(final, recursive=True)
        self.fs.mv(tmp, final, recursive=True)
        if max_to_keep is not None:
            for s in self.steps()[:-max_to_keep]:
                self.fs.rm(self._r(str(s)), recursive=True)
