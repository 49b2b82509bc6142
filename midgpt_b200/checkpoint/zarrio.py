"""Minimal Zarr-v2 array writer/reader (no compression), enough for Orbax's per-leaf array directories.

Orbax's ``PyTreeCheckpointHandler`` (non-OCDBT mode) stores every array leaf as a tensorstore
``zarr`` driver directory: a ``.zarray`` JSON + C-order chunk files named ``i.j.k``.  Neither
tensorstore nor zarr is installed here, so the (tiny) subset we need is written by hand.
Chunk files are raw little-endian buffers, which also makes partial reads (one rank's shard
range) a plain ``np.memmap`` slice.
"""
from __future__ import annotations

import json
import os
import typing as tp

import numpy as np

_DTYPES = {"float32": "<f4", "float64": "<f8", "int32": "<i4", "int64": "<i8", "uint16": "<u2", "bfloat16": "bfloat16"}


def write_zarray_meta(path: str, shape: tp.Sequence[int], chunks: tp.Sequence[int], dtype: str) -> None:
    os.makedirs(path, exist_ok=True)
    meta = {"zarr_format": 2, "shape": list(shape), "chunks": list(chunks), "dtype": _DTYPES[dtype],
            "compressor": None, "fill_value": 0, "filters": None, "order": "C", "dimension_separator": "."}
    with open(os.path.join(path, ".zarray"), "w") as f:
        json.dump(meta, f)


def chunk_name(idx: tp.Sequence[int]) -> str:
    return ".".join(str(i) for i in idx) if len(idx) else "0"


def write_chunk(path: str, idx: tp.Sequence[int], data: np.ndarray) -> None:
    os.makedirs(path, exist_ok=True)
    tmp = os.path.join(path, chunk_name(idx) + ".tmp")
    np.ascontiguousarray(data).tofile(tmp)
    os.replace(tmp, os.path.join(path, chunk_name(idx)))


def write_chunk_range(path: str, idx: tp.Sequence[int], byte_offset: int, data: np.ndarray, chunk_bytes: int) -> None:
    """Write ``data`` into the byte range [byte_offset, +nbytes) of one (raw, uncompressed) chunk file.  Several ranks write
    disjoint ranges of the same chunk concurrently: the file is created if missing and sized to ``chunk_bytes`` (ftruncate to
    the same size is idempotent and never discards bytes another rank already wrote)."""
    os.makedirs(path, exist_ok=True)
    fd = os.open(os.path.join(path, chunk_name(idx)), os.O_RDWR | os.O_CREAT, 0o644)
    try:
        if os.fstat(fd).st_size < chunk_bytes:
            os.ftruncate(fd, chunk_bytes)
        buf = memoryview(np.ascontiguousarray(data)).cast("B")
        done = 0
        while done < len(buf):
            done += os.pwrite(fd, buf[done:], byte_offset + done)
    finally:
        os.close(fd)


def read_meta(path: str) -> dict:
    with open(os.path.join(path, ".zarray")) as f:
        return json.load(f)


def open_chunk(path: str, idx: tp.Sequence[int], meta: dict | None = None) -> np.ndarray:
    """Memory-map one chunk (shape = meta['chunks'])."""
    meta = meta or read_meta(path)
    dt = np.dtype(meta["dtype"]) if meta["dtype"] != "bfloat16" else np.dtype("<u2")
    return np.memmap(os.path.join(path, chunk_name(idx)), dtype=dt, mode="r", shape=tuple(meta["chunks"]))


def read_array(path: str) -> np.ndarray:
    meta = read_meta(path)
    shape, chunks = meta["shape"], meta["chunks"]
    dt = np.dtype(meta["dtype"])
    out = np.zeros(shape, dtype=dt)
    if not shape:
        return np.fromfile(os.path.join(path, "0"), dtype=dt).reshape(())
    grid = [(-(-s // c)) for s, c in zip(shape, chunks)]
    for flat in range(int(np.prod(grid))):
        idx = np.unravel_index(flat, grid)
        ch = open_chunk(path, idx, meta)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = ch[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out
