"""Preset `openwebtext_mh` (reference: src/configs/openwebtext_mh.py:4-22); see _registry.py for the values."""
from ._registry import make

config = make("openwebtext_mh")
