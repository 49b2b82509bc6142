"""Build the character-level Shakespeare dataset: ``train.bin`` / ``val.bin`` (uint16) + ``meta.pkl``.

Output contract = reference data/shakespeare_char/prepare.py:49-61 (flat uint16 token files, 90/10
split, ``meta.pkl`` with ``vocab_size``/``itos``/``stoi`` over the sorted set of characters).
There is no network on the training hosts, so the corpus is looked up locally first
(``input.txt`` next to this file or ``$MIDGPT_SHAKESPEARE_TXT``); downloading is only attempted if
that fails; ``--synthetic`` writes a small pseudo-text with the same file formats for smoke runs.
"""
import argparse
import os
import pickle

import numpy as np

URL = "https://raw.githubusercontent.com/karpathy/char-rnn/master/data/tinyshakespeare/input.txt"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_corpus(synthetic: bool) -> str:
    if synthetic:
        rng = np.random.default_rng(0)
        words = ["the", "king", "and", "queen", "of", "night", "speak", "thou", "art", "love", "death",
                 "my", "lord", "good", "sir", "what", "is", "this", "that", "shall", "not", "be"]
        lines = []
        for i in range(20000):
            n = int(rng.integers(3, 9))
            lines.append(" ".join(words[int(j)] for j in rng.integers(0, len(words), n)).capitalize() + ".")
            if i % 7 == 0:
                lines.append("\nSPEAKER " + chr(65 + i % 26) + ":")
        return "\n".join(lines)
    for cand in (os.environ.get("MIDGPT_SHAKESPEARE_TXT"), os.path.join(HERE, "input.txt")):
        if cand and os.path.exists(cand):
            with open(cand, "r", encoding="utf-8") as f:
                return f.read()
    import requests  # network path, last resort
    text = requests.get(URL, timeout=30).text
    with open(os.path.join(HERE, "input.txt"), "w", encoding="utf-8") as f:
        f.write(text)
    return text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    text = read_corpus(args.synthetic)
    alphabet = sorted(set(text))
    stoi = {c: i for i, c in enumerate(alphabet)}
    itos = dict(enumerate(alphabet))
    print(f"{len(text):,} characters, vocab {len(alphabet)}: {''.join(alphabet)!r}")
    ids = np.fromiter((stoi[c] for c in text), dtype=np.uint16, count=len(text))
    cut = int(len(ids) * 0.9)
    os.makedirs(args.out, exist_ok=True)
    ids[:cut].tofile(os.path.join(args.out, "train.bin"))
    ids[cut:].tofile(os.path.join(args.out, "val.bin"))
    with open(os.path.join(args.out, "meta.pkl"), "wb") as f:
        pickle.dump({"vocab_size": len(alphabet), "itos": itos, "stoi": stoi}, f)
    print(f"train {cut:,} tokens, val {len(ids) - cut:,} tokens → {args.out}")


if __name__ == "__main__":
    main()
