"""Loader for the C++ batch assembler (csrc/dataloader.cpp); returns None when it has not been built."""
from __future__ import annotations

import importlib.util
import os

_mod = None
_tried = False


def get():
    global _mod, _tried
    if _tried:
        return _mod
    _tried = True
    from .build import loader_path
    p = loader_path()
    if os.path.exists(p):
        spec = importlib.util.spec_from_file_location("_native_loader", p)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mod = m
    return _mod
