"""RoPE relative-position property — the reference's scripts/test_rotary.py:11-32, with an assert."""
import torch

from midgpt_b200.models.oracle import rope_tables, apply_rope


def test_rotary_scores_depend_only_on_relative_position():
    torch.manual_seed(0)
    T, C, shift = 32, 64, 5
    q = torch.randn(8, T, C, dtype=torch.float64)
    k = torch.randn(8, T, C, dtype=torch.float64)
    sin, cos = rope_tables(C, T)
    a = apply_rope(q, sin, cos) @ apply_rope(k, sin, cos).transpose(-1, -2)
    qs, ks = torch.roll(q, shift, 1), torch.roll(k, shift, 1)
    a_s = apply_rope(qs, sin, cos) @ apply_rope(ks, sin, cos).transpose(-1, -2)
    a_rolled = torch.roll(a, (shift, shift), (-2, -1))
    assert (a_s[:, shift:, shift:] - a_rolled[:, shift:, shift:]).abs().max() < 1e-10


def test_rope_is_interleaved_pairs():
    # position 1, pair (a, b) rotates by angle inv_freq_i: out = (a cos - b sin, b cos + a sin)
    C = 8
    sin, cos = rope_tables(C, 2)
    x = torch.zeros(2, C, dtype=torch.float64)
    x[1, 0], x[1, 1] = 1.0, 0.0
    y = apply_rope(x, sin, cos)
    assert torch.allclose(y[1, :2], torch.tensor([cos[1, 0], sin[1, 0]]))
    assert torch.all(y[1, 2:] == 0)
