"""Spawn helpers for multi-process (gloo / nccl) tests."""
import os
import socket
import traceback

import torch.multiprocessing as mp


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _plain(o):
    """Tensors → numpy so results pickle by value (fd-sharing breaks once the child exits)."""
    import torch
    if isinstance(o, torch.Tensor):
        return o.detach().cpu().float().numpy() if o.dtype == torch.bfloat16 else o.detach().cpu().numpy()
    if isinstance(o, dict):
        return {k: _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(_plain(v) for v in o)
    return o


def _entry(rank, world, port, fn, args, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    try:
        out = fn(rank, world, *args)
        q.put((rank, "ok", _plain(out)))
    except Exception:
        q.put((rank, "err", traceback.format_exc()))
    finally:
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


def run_distributed(fn, world: int, *args, timeout: float = 240.0, _retries: int = 2):
    """Run ``fn(rank, world, *args)`` in ``world`` processes; returns the list of return values by rank.  A rendezvous port
    that got taken between probing and binding (EADDRINUSE) is retried with a fresh port."""
    try:
        return _run_once(fn, world, args, timeout)
    except RuntimeError as e:
        if _retries > 0 and ("EADDRINUSE" in str(e) or "address already in use" in str(e)):
            return run_distributed(fn, world, *args, timeout=timeout, _retries=_retries - 1)
        raise


def _run_once(fn, world: int, args, timeout: float):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, status, out = q.get(timeout=timeout)
            if status != "ok":
                raise RuntimeError(f"rank {rank} failed:\n{out}")
            results[rank] = out
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    return [results[r] for r in range(world)]
