"""Data pipeline: sampler contract, native gatherer ≡ numpy path, RNG state round-trip, per-process split."""
import numpy as np
import torch

from midgpt_b200.data import BatchSampler, PinnedBatchLoader, get_batch, split_array_by_idx, synthetic_tokens


def test_get_batch_contract():
    data = np.arange(10_000, dtype=np.uint16)
    x, y = get_batch(data, 16, 4, 3, rng=np.random.default_rng(0))
    assert x.shape == y.shape == (3, 4, 16) and x.dtype == np.int32
    assert np.array_equal(y, x + 1)                                   # y is x shifted by one token
    assert np.array_equal(x[..., 1:], x[..., :-1] + 1)                # contiguous crops
    x2, _ = get_batch(data, 16, 4, None, rng=np.random.default_rng(0))
    assert x2.shape == (4, 16)


def test_native_matches_numpy_and_state_roundtrip():
    data = synthetic_tokens(1 << 16, 50257, seed=3)
    a, b = BatchSampler(data, 64, seed=5), BatchSampler(data, 64, seed=5)
    b._native = None
    xa, ya = a.sample(32)
    xb, yb = b.sample(32)
    assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    st = a.state_dict()
    x1, _ = a.sample(8)
    a.load_state_dict(st)
    x2, _ = a.sample(8)
    assert np.array_equal(x1, x2)


def test_split_array_by_idx_matches_reference_arithmetic():
    arr = np.arange(103)
    parts = [split_array_by_idx(arr, i, 4) for i in range(4)]
    n = int(103 / 4) + 1
    assert all(len(p) <= n for p in parts) and np.array_equal(np.concatenate(parts), arr)


def test_pinned_loader_cpu():
    data = synthetic_tokens(1 << 14, 100, seed=1)
    ld = PinnedBatchLoader(BatchSampler(data, 32, seed=0), 4, torch.device("cpu"))
    x, y = ld.next()
    assert x.shape == (4, 32) and x.dtype == torch.int32 and torch.equal(x[:, 1:], y[:, :-1])
    ld.close()
