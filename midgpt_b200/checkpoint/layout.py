"""The reference's checkpoint *tree*: ``(tree_leaves(model), tree_leaves(opt_state))`` (src/train.py:215).

Leaf order is the Jax pytree flattening of the Equinox ``GPT`` module and of the optax chain state
(inferred from dataclass field order — [lib], unverifiable without jax here, so the
mapping is confined to this file).  Array leaves of the blocks carry a leading ``L`` axis.
Keys on disk are ``"<tuple index>.<leaf index>"``.

Schema provenance (nothing here could be checked against a live Orbax: neither orbax nor tensorstore is installed and
there is no network; every item is marked with where it comes from):

* ``(list, list)`` tree, leaf order, 22 + 18 leaves ............ reference src/train.py:214-215 + dataclass field order
  (src/model.py, src/layers.py); optax chain of 5 → ``count, mu×8, nu×8, count`` (src/train.py:153-159).  [derived]
* param directory names ``"0.5"`` = tuple keys joined by ``.`` ... Orbax ``PyTreeCheckpointHandler`` flattens the tree with
  ``jax.tree_util`` key paths and joins sequence indices / dict keys with ``"."``.  [recalled, lib]
* ``_METADATA`` = ``{"tree_metadata": {"('0', '5')": {"key_metadata": [{"key": "0", "key_type": 1}, …],
  "value_metadata": {"value_type": "jax.Array", "skip_deserialize": false}}, …}, "use_zarr3": false}`` — the
  ``tree_metadata`` schema of orbax-checkpoint's ``tree_metadata.py`` (``KeyType.SEQUENCE = 1``, ``KeyType.DICT = 2``;
  ``value_type`` is the type-handler ``typestr()``: ``jax.Array`` / ``np.ndarray`` / ``scalar`` / ``string``).  The file was
  introduced after the 0.1.9 pin of the reference (requirements.txt:6) and newer readers use it, older ones ignore it.
  [recalled, lib]
* ``checkpoint`` = msgpack of the *nested* tree as flax ``to_state_dict`` renders it (sequences become dicts keyed by the
  decimal index as a string), array leaves replaced by ``"PLACEHOLDER://<param name>"``, python scalars inline — the
  0.1.x aggregate file.  [recalled, lib]
* per-leaf Zarr-v2 directories (``.zarray``, C-order raw chunks ``i.j.k``, no compressor) .......... zarr v2 spec.  [spec]
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


# ----------------------------------------------------------------------------------------------
# Orbax-shaped metadata (see the provenance notes in the module docstring)
# ----------------------------------------------------------------------------------------------
KEY_TYPE_SEQUENCE = 1


def all_keys() -> tp.List[str]:
    return [model_key(i) for i in range(len(MODEL_LEAVES))] + [k for k, _, _ in opt_layout()]


def tree_metadata(array_keys_: tp.Set[str], scalars: tp.Dict[str, tp.Any]) -> dict:
    tm = {}
    for key in all_keys():
        parts = key.split(".")
        if key in array_keys_:
            vt = "jax.Array"
        else:
            vt = "scalar"
        tm[str(tuple(parts))] = {
            "key_metadata": [{"key": p, "key_type": KEY_TYPE_SEQUENCE} for p in parts],
            "value_metadata": {"value_type": vt, "skip_deserialize": False},
        }
    return {"tree_metadata": tm, "use_zarr3": False}


def aggregate_tree(array_keys_: tp.Set[str], scalars: tp.Dict[str, tp.Any]) -> dict:
    out: tp.Dict[str, tp.Dict[str, tp.Any]] = {}
    for key in all_keys():
        a, b = key.split(".")
        out.setdefault(a, {})[b] = f"PLACEHOLDER://{key}" if key in array_keys_ else scalars[key]
    return out


def aggregate_lookup(agg: dict, key: str):
    """Value of ``key`` ("1.0") in a nested aggregate; also accepts the flat ``{key: value}`` form of older checkpoints."""
    if key in agg:
        return agg[key]
    a, b = key.split(".")
    return agg[a][b]
