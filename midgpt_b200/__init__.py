"""midgpt_b200 — a Blackwell (sm_100a) native FSDP GPT pre-training engine.

Capabilities mirror AllanYangZhou/midGPT (reference: /root/reference): decoder-only GPT with
RoPE, weightless RMSNorm, QK-LayerNorm, independent-weight-decay AdamW, FSDP sharding of
parameters / gradients / optimizer state, Orbax-layout checkpoints, ``launch.py``/``sample.py``.

The design is *not* a port: one process per GPU, explicit forward/backward over hand-written
sm_100a kernels (tcgen05 + TMEM + TMA), explicit per-unit collectives over NVLink.
"""
from .config import GPTConfig, ExperimentConfig, load_config, config_from_json, config_to_json  # noqa: F401

__version__ = "0.1.0"
