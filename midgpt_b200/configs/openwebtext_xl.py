"""Preset `openwebtext_xl` (reference: src/configs/openwebtext_xl.py:4-22); see _registry.py for the values."""
from ._registry import make

config = make("openwebtext_xl")
