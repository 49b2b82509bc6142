"""Tokenise OpenWebText into ``train.bin`` / ``val.bin`` (flat uint16 GPT-2 BPE ids, EOT after each doc).

Output contract = reference data/openwebtext/prepare.py:41-73: ``encode_ordinary`` + EOT 50256 appended,
0.05 % validation split with seed 2357, uint16 files (~9.04 B train / ~4.4 M val tokens).

* With HF ``datasets`` installed (the reference's path, prepare.py:21-25) the split is *the reference's split*:
  ``dataset["train"].train_test_split(test_size=0.0005, seed=2357, shuffle=True)`` — same documents in ``val.bin``.
* Without it (no network on the training hosts) documents come from local ``*.txt`` / ``*.jsonl`` shards (``--source``) and
  the split is a seeded (2357) permutation of the document indices — same fraction, not the same documents.
* Either way documents are **streamed**: a bounded window of documents is tokenised by a process pool and the ids are
  appended to the output file in document order, so peak RAM is O(window), not O(corpus).  (The reference sizes a memmap
  from a tokenised Arrow dataset and fills it in 1024 write-batches, prepare.py:58-73; appending produces the same bytes
  without knowing the total in advance, so no counting pass is needed.)
* ``--synthetic N`` writes N Zipf-distributed tokens for benchmarking.
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


def _encode(text, tokenizer="gpt2"):
    if tokenizer == "bytes":            # offline / test tokenizer: UTF-8 bytes (the GPT-2 BPE files need a download)
        ids = list(text.encode("utf-8"))
    else:
        import tiktoken
        enc = tiktoken.get_encoding("gpt2")
        ids = enc.encode_ordinary(text)
    ids.append(EOT)
    return np.asarray(ids, dtype=np.uint16)


def stream_tokenise(docs, files, route, workers, window=4096, tokenizer="gpt2"):
    """Tokenise ``docs`` (any iterable of strings, consumed lazily) and append every document's ids to
    ``files[route(i)]`` in document order.  At most ``window`` documents / token arrays are alive at any time."""
    from collections import deque
    totals = {k: 0 for k in files}
    pending = deque()

    def drain_one():
        k, fut = pending.popleft()
        a = fut.result()
        a.tofile(files[k])
        totals[k] += len(a)

    with ProcessPoolExecutor(workers) as ex:
        for i, text in enumerate(docs):
            pending.append((route(i), ex.submit(_encode, text, tokenizer)))
            if len(pending) >= window:
                drain_one()
        while pending:
            drain_one()
    return totals


def count_docs(source):
    return sum(1 for _ in iter_local_docs(source))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default=None, help="directory of .txt/.jsonl shards (default: HF datasets 'openwebtext')")
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) // 2))
    ap.add_argument("--synthetic", type=int, default=0, help="write this many synthetic tokens instead")
    ap.add_argument("--tokenizer", default="gpt2", choices=["gpt2", "bytes"])
    ap.add_argument("--window", type=int, default=4096, help="documents in flight (bounds peak RAM)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    if args.synthetic:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from midgpt_b200.data import synthetic_tokens
        synthetic_tokens(args.synthetic, 50257, seed=1).tofile(os.path.join(args.out, "train.bin"))
        synthetic_tokens(max(args.synthetic // 2000, 4096), 50257, seed=2).tofile(os.path.join(args.out, "val.bin"))
        return
    train_path, val_path = os.path.join(args.out, "train.bin"), os.path.join(args.out, "val.bin")
    if args.source:
        n_docs = count_docs(args.source)                      # cheap pass over the shards: no tokenisation, nothing kept
        rng = np.random.default_rng(2357)
        n_val = max(1, int(round(n_docs * 0.0005)))
        val_idx = set(rng.permutation(n_docs)[:n_val].tolist())
        with open(train_path, "wb") as ft, open(val_path, "wb") as fv:
            totals = stream_tokenise(iter_local_docs(args.source), {"train": ft, "val": fv},
                                     lambda i: "val" if i in val_idx else "train", args.workers, args.window, args.tokenizer)
    else:
        from datasets import load_dataset
        dataset = load_dataset("openwebtext", num_proc=args.workers)
        split = dataset["train"].train_test_split(test_size=0.0005, seed=2357, shuffle=True)   # the reference's split
        totals = {}
        for name, path, dset in (("train", train_path, split["train"]), ("val", val_path, split["test"])):
            with open(path, "wb") as f:
                totals.update(stream_tokenise((ex["text"] for ex in dset), {name: f}, lambda i, name=name: name,
                                              args.workers, args.window, args.tokenizer))
    for k, v in totals.items():
        print(f"{k}.bin: {v:,} tokens")


if __name__ == "__main__":
    main()
