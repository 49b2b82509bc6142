"""Experiment configs — one module per experiment, each exporting ``config`` (reference: src/configs/*.py)."""
