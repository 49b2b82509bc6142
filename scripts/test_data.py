"""Data-loader micro-benchmark — the reference's scripts/test_data.py:12-26 (time to load, batches/second), for the
numpy sampler and the native C++ gatherer.  Uses a synthetic corpus when no dataset directory is given."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midgpt_b200.data import BatchSampler, load_tokens, synthetic_tokens  # noqa: E402

data_dir = sys.argv[1] if len(sys.argv) > 1 else None
t0 = time.time()
data = load_tokens(data_dir, "train") if data_dir else synthetic_tokens(1 << 24, 50257)
print(f"Time to load train.bin: {time.time() - t0:.2f}s ({len(data):,} tokens)")
for native in (False, True):
    s = BatchSampler(data, 1024)
    if not native:
        s._native = None
    elif s._native is None:
        print("native loader not built (python -m midgpt_b200.ops.build)")
        continue
    x = np.empty((128 * 4, 1024), np.int32); y = np.empty_like(x)
    s.sample_into(x, y)
    t0 = time.time()
    for _ in range(100):
        s.sample_into(x, y)            # same shape as the reference's get_batch(…, 1024, 128, 4)
    dt = time.time() - t0
    print(f"{'native' if native else 'numpy '} sampler: {100 / dt:.1f} batches/s ({100 * x.size / dt / 1e6:.0f} M tokens/s)")
