"""Parameter inventory, initialisation and reference-compatible (stacked) views.

The reference model (src/model.py:118-138) has exactly eight array leaves, each block weight
stacked over a leading ``L`` axis:

    wte (V,D) · c_attn (L,3D,D) · attn.c_proj (L,D,D) · q_ln (L,C) · k_ln (L,C)
    c_fc (L,4D,D) · mlp.c_proj (L,D,4D) · lm_head (V,D)

This engine instead keeps one *FSDP unit* per transformer block (plus one for the embedding and
one for the LM head); inside a unit all tensors are concatenated into a single flat buffer.  This
module is the single source of truth for names, shapes, unit membership, init and for converting
between "per-unit" and "stacked reference leaf" representations (the latter is what checkpoints
store so that leaf order/shape match the reference's Orbax tree).
"""
from __future__ import annotations

import math
import typing as tp
from dataclasses import dataclass

import torch

from ..config import GPTConfig

# Order of tensors inside one block unit.  Big matrices first so they are 1024-element aligned.
BLOCK_PARAM_ORDER = ("c_attn", "attn_proj", "c_fc", "mlp_proj", "q_ln", "k_ln")

# Reference leaf order of the *model* part of the checkpoint (jax tree_leaves order of GPT with
# array leaves only; scalar leaves are handled in checkpoint/layout.py).
REFERENCE_ARRAY_LEAVES = ("wte", "c_attn", "attn_proj", "q_ln", "k_ln", "c_fc", "mlp_proj", "lm_head")


@dataclass(frozen=True)
class ParamSpec:
    name: str            # local name inside the unit, e.g. "c_attn"
    shape: tp.Tuple[int, ...]
    offset: int          # element offset inside the unit's flat buffer
    fan_in: int          # for init; 0 => constant-one init (LayerNorm gains)

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


@dataclass(frozen=True)
class UnitSpec:
    name: str                         # "embed" | "block.<l>" | "head"
    params: tp.Tuple[ParamSpec, ...]
    numel: int                        # un-padded

    def spec(self, pname: str) -> ParamSpec:
        for p in self.params:
            if p.name == pname:
                return p
        raise KeyError(pname)


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


def block_param_shapes(cfg: GPTConfig) -> tp.Dict[str, tp.Tuple[int, ...]]:
    D, C = cfg.n_embd, cfg.head_dim
    return {
        "c_attn": (3 * D, D),      # reference src/model.py:48  Linear(D, 3D) stores (out, in)
        "attn_proj": (D, D),       # src/model.py:49
        "c_fc": (4 * D, D),        # src/model.py:24
        "mlp_proj": (D, 4 * D),    # src/model.py:25
        "q_ln": (C,),              # src/model.py:52  LayerNorm gain, shared by all heads
        "k_ln": (C,),              # src/model.py:53
    }


def build_unit_specs(cfg: GPTConfig) -> tp.List[UnitSpec]:
    """embed, block.0 … block.L-1, head — the order in which forward consumes them."""
    V, D = cfg.vocab_size, cfg.n_embd
    units = [UnitSpec("embed", (ParamSpec("wte", (V, D), 0, -1),), V * D)]
    shapes = block_param_shapes(cfg)
    for l in range(cfg.n_layer):
        off, ps = 0, []
        for pname in BLOCK_PARAM_ORDER:
            shp = shapes[pname]
            fan_in = shp[1] if len(shp) == 2 else 0
            p = ParamSpec(pname, shp, off, fan_in)
            ps.append(p)
            off = _align(off + p.numel, 8)   # keep every tensor 16-byte aligned in bf16
        units.append(UnitSpec(f"block.{l}", tuple(ps), off))
    units.append(UnitSpec("head", (ParamSpec("lm_head", (V, D), 0, -1),), V * D))
    return units


def count_params(cfg: GPTConfig) -> int:
    """All array elements minus one (V,D) copy — reference src/model.py:161-164."""
    total = sum(u.numel_exact for u in _exact_units(cfg))
    return total - cfg.vocab_size * cfg.n_embd


@dataclass(frozen=True)
class _Exact:
    numel_exact: int


def _exact_units(cfg: GPTConfig):
    V, D = cfg.vocab_size, cfg.n_embd
    per_block = sum(math.prod(s) for s in block_param_shapes(cfg).values())
    return [_Exact(V * D)] + [_Exact(per_block)] * cfg.n_layer + [_Exact(V * D)]


def stored_params(cfg: GPTConfig) -> int:
    return sum(u.numel_exact for u in _exact_units(cfg))


def flops_per_token(cfg: GPTConfig) -> int:
    """Model FLOPs/token used for MFU: 6·N + 12·L·D·T (reproduces README's 47.8%)."""
    return 6 * count_params(cfg) + 12 * cfg.n_layer * cfg.n_embd * cfg.block_size


# ----------------------------------------------------------------------------------------------
# Initialisation (reference src/layers.py:49-50, src/model.py:134-138, K21 in SURVEY.md)
# ----------------------------------------------------------------------------------------------

def _seed_for(base_seed: int, unit: str, pname: str) -> int:
    # A stable 63-bit hash that does not depend on PYTHONHASHSEED, world size or device.
    h = 1469598103934665603
    for ch in f"{base_seed}/{unit}/{pname}".encode():
        h = ((h ^ ch) * 1099511628211) & 0x7FFFFFFFFFFFFFFF
    return h


def init_param(spec: ParamSpec, unit: str, cfg: GPTConfig, seed: int, device="cpu") -> torch.Tensor:
    """fp32 init of one tensor.  Generated on CPU so every world size / device sees identical bits."""
    g = torch.Generator(device="cpu")
    # wte and lm_head are initialised from the *same* values but are separate parameters
    # (reference src/model.py:135-138: "Share first and last layer parameters" — same array passed
    # to both, but they are independent pytree leaves and diverge after the first update).
    key_unit, key_name = ("embed", "wte") if spec.name in ("wte", "lm_head") else (unit, spec.name)
    g.manual_seed(_seed_for(seed, key_unit, key_name))
    if spec.fan_in == 0:
        w = torch.ones(spec.shape, dtype=torch.float32)
    elif spec.fan_in == -1:
        w = torch.randn(spec.shape, generator=g, dtype=torch.float32) / math.sqrt(cfg.n_embd)
    else:
        w = torch.empty(spec.shape, dtype=torch.float32)
        # truncated standard normal on [-2, 2], *not* re-normalised (std ≈ 0.88), times 1/sqrt(fan_in)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        w.mul_(1.0 / math.sqrt(spec.fan_in))
    return w.to(device)


def init_unit_flat(unit: UnitSpec, cfg: GPTConfig, seed: int, padded_numel: int | None = None,
                   device="cpu") -> torch.Tensor:
    flat = torch.zeros(padded_numel or unit.numel, dtype=torch.float32, device=device)
    for p in unit.params:
        flat[p.offset:p.offset + p.numel].copy_(init_param(p, unit.name, cfg, seed, device).reshape(-1))
    return flat


def unit_views(unit: UnitSpec, flat: torch.Tensor) -> tp.Dict[str, torch.Tensor]:
    """Named, shaped views into a (gathered) flat buffer."""
    return {p.name: flat[p.offset:p.offset + p.numel].view(p.shape) for p in unit.params}


def init_full_params(cfg: GPTConfig, seed: int = 0, device="cpu", dtype=torch.float32) -> tp.Dict[str, torch.Tensor]:
    """Un-sharded parameter dict ``{"embed.wte", "block.3.c_fc", …, "head.lm_head"}`` (oracle / tests / sampling)."""
    out = {}
    for u in build_unit_specs(cfg):
        for p in u.params:
            out[f"{u.name}.{p.name}"] = init_param(p, u.name, cfg, seed, device).to(dtype)
    return out


# ----------------------------------------------------------------------------------------------
# Stacked (reference-leaf) representation
# ----------------------------------------------------------------------------------------------

def to_reference_leaves(named: tp.Dict[str, torch.Tensor], cfg: GPTConfig) -> tp.Dict[str, torch.Tensor]:
    """{"embed.wte", "block.l.x", "head.lm_head"} → the 8 stacked arrays of the reference pytree."""
    out = {"wte": named["embed.wte"], "lm_head": named["head.lm_head"]}
    for pname in BLOCK_PARAM_ORDER:
        out[pname] = torch.stack([named[f"block.{l}.{pname}"] for l in range(cfg.n_layer)], 0)
    return out


def from_reference_leaves(leaves: tp.Dict[str, torch.Tensor], cfg: GPTConfig) -> tp.Dict[str, torch.Tensor]:
    out = {"embed.wte": leaves["wte"], "head.lm_head": leaves["lm_head"]}
    for pname in BLOCK_PARAM_ORDER:
        for l in range(cfg.n_layer):
            out[f"block.{l}.{pname}"] = leaves[pname][l]
    return out
