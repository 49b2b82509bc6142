"""Single table of experiment presets.

The reference keeps one hand-written dataclass literal per file (src/configs/*.py).  Here the
presets are rows of one table so that shared defaults (fp32 master / bf16 compute, the 1e-3 peak
LR and 1e-4 independent weight decay used by every reference run) are stated once, and the
per-preset modules are one-liners.  Values are taken from SURVEY.md §5.6.
"""
from __future__ import annotations

from midgpt_b200.config import ExperimentConfig, GPTConfig, EngineConfig

# name -> (model shape, optimisation schedule, batch layout, data dir, shard_model)
_MODELS = {
    "char6": dict(block_size=256, vocab_size=65, n_layer=6, n_head=6, n_embd=384, dropout=0.2),
    "gpt2_small": dict(block_size=1024, vocab_size=50304, n_layer=12, n_head=12, n_embd=768, dropout=0.0),
    "gpt2_xl_ish": dict(block_size=1024, vocab_size=50304, n_layer=24, n_head=16, n_embd=2048, dropout=0.0),
}

_PRESETS = {
    #                     model          data_dir                              B     G   warm   total  min_lr beta2 eval  shard
    "shakespeare_char": ("char6",       "data/shakespeare_char",              64,   1,  100,   5000,  1e-4,  0.99, 2000, False),
    "openwebtext":      ("gpt2_small",  "data/openwebtext",                   128,  16, 5000,  60000, 1e-5,  0.95, 1000, False),
    "openwebtext_mh":   ("gpt2_small",  "/mnt/disks/persist/openwebtext",     2048, 1,  5000,  60000, 1e-5,  0.95, 1000, False),
    "openwebtext_xl":   ("gpt2_xl_ish", "/mnt/disks/persist/openwebtext",     1024, 1,  2500,  25000, 1e-5,  0.95, 1000, True),
}


def make(name: str) -> ExperimentConfig:
    model, data_dir, B, G, warm, total, min_lr, beta2, eval_iv, shard = _PRESETS[name]
    return ExperimentConfig(
        rundir="", data_dir=data_dir,
        learning_rate=1e-3, min_lr=min_lr, warmup_steps=warm, lr_decay_steps=total, max_steps=total,
        batch_size=B, g_accum_iters=G, beta2=beta2, weight_decay=1e-4, eval_interval=eval_iv,
        param_dtype="float32", compute_dtype="bfloat16", shard_model=shard,
        model_config=GPTConfig(**_MODELS[model]), engine=EngineConfig(),
    )
