"""Preset `shakespeare_char` (reference: src/configs/shakespeare_char.py:4-22); see _registry.py for the values."""
from ._registry import make

config = make("shakespeare_char")
