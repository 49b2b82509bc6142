"""The reference's checkpoint *tree*: ``(tree_leaves(model), tree_leaves(opt_state))`` (src/train.py:215).

Leaf order is the Jax pytree flattening of the Equinox ``GPT`` module and of the optax chain state
(inferred from dataclass field order — [lib], unverifiable without jax here, so the
mapping is confined to this file).  Array leaves of the blocks carry a leading ``L`` axis.
Keys on disk are ``"<tuple index>.<leaf index>"``.
"""
from __future__ import annotations

import typing as tp


# (key, kind, source) — kind "array": source is a reference-leaf name (models/params.py);
# kind "scalar": source is a callable(cfg) → python scalar.
MODEL_LEAVES: tp.List[tp.Tuple[str, tp.Any]] = [
    ("array", "wte"),
    ("scalar", lambda c: c.model_config.dropout),        # drop.p
    ("scalar", lambda c: False),                         # drop.inference
    ("scalar", lambda c: c.model_config.n_head),         # blocks.attn.n_head
    ("scalar", lambda c: c.model_config.n_embd),         # blocks.attn.n_embd
    ("array", "c_attn"),
    ("array", "attn_proj"),
    ("scalar", lambda c: c.model_config.dropout),        # attn_dropout.p
    ("scalar", lambda c: False),
    ("scalar", lambda c: c.model_config.dropout),        # resid_dropout.p
    ("scalar", lambda c: False),
    ("array", "q_ln"),
    ("array", "k_ln"),
    ("array", "c_fc"),
    ("array", "mlp_proj"),
    ("scalar", lambda c: c.model_config.dropout),        # mlp.dropout.p
    ("scalar", lambda c: False),
    ("scalar", lambda c: 1e-6),                          # ln1.eps
    ("scalar", lambda c: 1e-6),                          # ln2.eps
    ("scalar", lambda c: 1e-5),                          # ln_f.eps
    ("array", "lm_head"),
    ("scalar", lambda c: c.model_config.n_layer),
]

ARRAY_ORDER = [src for kind, src in MODEL_LEAVES if kind == "array"]   # 8 arrays


def model_key(i: int) -> str:
    return f"0.{i}"


def opt_layout() -> tp.List[tp.Tuple[str, str, str]]:
    """[(key, kind, source)]: Adam count, mu×8, nu×8, schedule count (optax chain of 5; EmptyStates have no leaves)."""
    out = [("1.0", "count", "adam")]
    k = 1
    for which in ("exp_avg", "exp_avg_sq"):
        for name in ARRAY_ORDER:
            out.append((f"1.{k}", which, name))
            k += 1
    out.append((f"1.{k}", "count", "schedule"))
    return out


def array_keys() -> tp.Dict[tp.Tuple[str, str], str]:
    """(which, reference-leaf name) → on-disk key, for which ∈ {master, exp_avg, exp_avg_sq}."""
    m = {}
    for i, (kind, src) in enumerate(MODEL_LEAVES):
        if kind == "array":
            m[("master", src)] = model_key(i)
    for key, which, name in opt_layout():
        if which != "count":
            m[(which, name)] = key
    return m
