"""Configuration dataclasses and (de)serialisation.

Parity with the reference:
  * ``GPTConfig``          — reference ``src/model.py:108-115``
  * ``ExperimentConfig``   — reference ``src/train.py:26-44`` (same field names, same order, so a
    ``config.json`` written by either implementation can be read by the other)
  * ``load_config(name)``  — dynamic import of ``configs/<name>.py`` (reference ``launch.py:25-27``)
  * ``config_from_json``   — recursive dict→dataclass (reference ``sample.py:49-65``)

Extra, B200-specific knobs live in ``EngineConfig`` and are *not* part of ``config.json``'s
reference schema (they are stored under the ``"engine"`` key which the reference would ignore).
"""
from __future__ import annotations

import dataclasses
import importlib
import json
import typing as tp
from dataclasses import dataclass, field, asdict


@dataclass
class GPTConfig:
    block_size: int  # max sequence length T
    vocab_size: int  # V
    n_layer: int  # L
    n_head: int  # H
    n_embd: int  # D
    dropout: float

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head


@dataclass
class EngineConfig:
    """B200 engine knobs (no reference analogue; XLA decided these implicitly)."""
    backend: str = "auto"          # "auto" | "b200" (hand-written kernels) | "torch" (library / CPU oracle path)
    comm: str = "auto"             # "auto" | "fused" (in-kernel NVLink) | "nccl" | "gloo"
    remat: str = "block"           # "block" (reference behaviour: recompute every block) | "none"
    micro_batch_size: int = 0      # sequences per GPU per micro-step; 0 = choose automatically
    gather_mode: str = "auto"      # "step" (bf16 weights resident all step) | "layer" (gather per layer, free after use)
    prefetch_depth: int = 2        # units gathered ahead of the consumer when gather_mode == "layer"
    grad_shard: bool = False       # True: reduce-scatter every unit's gradients EVERY micro-step into a shard-sized fp32 accumulator
                                   # (ZeRO-2/3 gradient memory, the reference's behaviour, src/train.py:87); False: accumulate locally
                                   # in full size and reduce-scatter once per step (less traffic; 4 bytes/param extra per GPU)
    fp8: bool = False              # forward GEMMs of the blocks in e4m3 on the fp8 tensor cores; backward stays bf16
    fp8_format: str = "mx"         # "mx": OCP MX block scaling (UE8M0 per 32 elements, tcgen05 block_scale) | "row": one fp32 scale per row
    seed: int = 0
    watchdog_s: float = 1800.0     # hang detection: exit(3) if no optimizer step completes for max(this, 20 x median step) seconds; 0 = off
    log_every: int = 20
    ckpt_async: bool = True


@dataclass
class ExperimentConfig:
    rundir: str  # directory containing ckpts and logs
    data_dir: str  # dataset directory
    learning_rate: float
    batch_size: int  # GLOBAL across all devices
    warmup_steps: int
    min_lr: float
    lr_decay_steps: int
    max_steps: int
    beta2: float
    weight_decay: float
    eval_interval: int
    param_dtype: str  # "float32" | "bfloat16"
    compute_dtype: str
    g_accum_iters: int
    shard_model: bool
    model_config: GPTConfig
    debug: bool = False
    engine: EngineConfig = field(default_factory=EngineConfig)

    def reference_dict(self) -> dict:
        """The subset that the reference writes into config.json (launch.py:40,56-57)."""
        d = asdict(self)
        d.pop("engine", None)
        return d


_NESTED = {"model_config": GPTConfig, "engine": EngineConfig}


def _convert(obj: tp.Any, cls: type) -> tp.Any:
    if dataclasses.is_dataclass(cls) and isinstance(obj, dict):
        names = {f.name for f in dataclasses.fields(cls)}
        kwargs = {}
        for k, v in obj.items():
            if k not in names:
                continue  # forward compatibility: ignore unknown keys
            kwargs[k] = _convert(v, _NESTED[k]) if k in _NESTED else v
        return cls(**kwargs)
    return obj


def config_to_json(cfg: ExperimentConfig) -> str:
    return json.dumps(asdict(cfg))


def config_from_json(s: str) -> ExperimentConfig:
    return _convert(json.loads(s), ExperimentConfig)


def load_config(name: str) -> ExperimentConfig:
    """Import ``midgpt_b200.configs.<name>`` and return a *copy* of its ``config``."""
    mod = importlib.import_module(f"midgpt_b200.configs.{name}")
    return dataclasses.replace(
        mod.config,
        model_config=dataclasses.replace(mod.config.model_config),
        engine=dataclasses.replace(mod.config.engine),
    )


AVAILABLE_CONFIGS = ("shakespeare_char", "openwebtext", "openwebtext_mh", "openwebtext_xl")
