"""Primitive-op backends.

``get_backend("auto", device)`` returns the hand-written sm_100a backend on a CUDA device (and
raises loudly if its extension is not built — a silent PyTorch fallback on a GPU box would hide a
broken build) and the PyTorch backend on CPU.  ``"torch"`` on a GPU is the cuBLAS/aten "library
baseline" the kernels are measured against.
"""
from __future__ import annotations

import torch


def get_backend(name: str, device: torch.device, fp8: bool = False, fp8_format: str = "mx"):
    if name == "auto":
        name = "b200" if device.type == "cuda" else "torch"
    if name == "torch":
        from .torch_backend import TorchBackend
        return TorchBackend()
    if name == "b200":
        if device.type != "cuda":
            raise RuntimeError("backend 'b200' needs a CUDA (sm_100a) device")
        from .b200_backend import B200Backend
        return B200Backend(device, fp8=fp8, fp8_format=fp8_format)
    raise ValueError(f"unknown backend {name!r}")
