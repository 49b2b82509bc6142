"""Preset `openwebtext` (reference: src/configs/openwebtext.py:4-22); see _registry.py for the values."""
from ._registry import make

config = make("openwebtext")
