"""Tokenise OpenWebText into ``train.bin`` / ``val.bin`` (flat uint16 GPT-2 BPE ids, EOT after each doc).

Output contract = reference data/openwebtext/prepare.py:41-73: ``encode_ordinary`` + EOT 50256 appended,
0.05 % validation split with seed 2357, uint16 memmaps (~9.04 B train / ~4.4 M val tokens).
Differences: documents can also come from local ``*.txt``/``*.jsonl`` shards (``--source``; no network on
the training hosts — HF ``datasets`` is used only when available), tokenisation is streamed through a
process pool straight into the output memmap (two passes: count, then write) instead of materialising a
tokenised HF dataset, and ``--synthetic N`` writes N Zipf-distributed tokens for benchmarking.
"""
import argparse
import glob
import json
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EOT = 50256


def iter_local_docs(source):
    for path in sorted(glob.glob(os.path.join(source, "*"))):
        if path.endswith(".jsonl"):
            with open(path, encoding="utf-8") as f:
                for line in f:
                    yield json.loads(line)["text"]
        elif path.endswith(".txt"):
            with open(path, encoding="utf-8") as f:
                yield f.read()


def _encode(text):
    import tiktoken
    enc = tiktoken.get_encoding("gpt2")
    ids = enc.encode_ordinary(text)
    ids.append(EOT)
    return np.asarray(ids, dtype=np.uint16)


def write_split(docs, path, workers):
    with ProcessPoolExecutor(workers) as ex:
        arrays = list(ex.map(_encode, docs, chunksize=64))
    total = int(sum(len(a) for a in arrays))
    out = np.memmap(path, dtype=np.uint16, mode="w+", shape=(total,))
    pos = 0
    for a in arrays:
        out[pos:pos + len(a)] = a
        pos += len(a)
    out.flush()
    print(f"{path}: {total:,} tokens")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default=None, help="directory of .txt/.jsonl shards (default: HF datasets 'openwebtext')")
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) // 2))
    ap.add_argument("--synthetic", type=int, default=0, help="write this many synthetic tokens instead")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    if args.synthetic:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from midgpt_b200.data import synthetic_tokens
        synthetic_tokens(args.synthetic, 50257, seed=1).tofile(os.path.join(args.out, "train.bin"))
        synthetic_tokens(max(args.synthetic // 2000, 4096), 50257, seed=2).tofile(os.path.join(args.out, "val.bin"))
        return
    if args.source:
        docs = list(iter_local_docs(args.source))
    else:
        from datasets import load_dataset
        docs = load_dataset("openwebtext", num_proc=args.workers)["train"]["text"]
    rng = np.random.default_rng(2357)
    perm = rng.permutation(len(docs))
    n_val = max(1, int(round(len(docs) * 0.0005)))
    val_idx = set(perm[:n_val].tolist())
    write_split([d for i, d in enumerate(docs) if i in val_idx], os.path.join(args.out, "val.bin"), args.workers)
    write_split([d for i, d in enumerate(docs) if i not in val_idx], os.path.join(args.out, "train.bin"), args.workers)


if __name__ == "__main__":
    main()
