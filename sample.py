#!/usr/bin/env python
"""Sample from a trained model — flags as in the reference ``sample.py`` (sample.py:29-37).

    python sample.py --ckpt_dir outputs/<run> --start "\\n" --num_samples 10 --max_new_tokens 500 --temperature 0.8

Reference behaviour kept: config comes from ``<ckpt_dir>/config.json`` (sample.py:98-99); the latest
checkpoint step is restored (sample.py:131-137); tokenizer = ``meta.pkl`` stoi/itos from the data dir if
present else GPT-2 BPE with ``<|endoftext|>`` allowed (sample.py:143-159); ``--start FILE:path`` reads the
prompt from a file (sample.py:167-170); the context is cropped to the last ``block_size`` tokens
(sample.py:74); logits at the last real position are divided by the temperature and sampled
categorically (sample.py:85-92); samples are separated by ``---------------`` (sample.py:194-197).

Difference: the reference right-pads to ``block_size`` to keep XLA shapes static (sample.py:76-81) and runs
a full 1024-token forward per new token; causality makes the padding irrelevant to the plucked logits, so
here the forward runs on the real context only and only the plucked position goes through the LM head.
"""
from __future__ import annotations

import argparse
import os
import pickle

import torch


def generate(model, provider, idx: torch.Tensor, max_new_tokens: int, block_size: int,
             temperature: float = 1.0, generator: torch.Generator | None = None) -> torch.Tensor:
    for _ in range(max_new_tokens):
        ctx = idx if idx.shape[1] <= block_size else idx[:, -block_size:]
        pos = torch.full((ctx.shape[0],), ctx.shape[1] - 1, dtype=torch.long)
        logits = model.forward_logits(provider, ctx.to(torch.int32), positions=pos).float() / temperature
        probs = torch.softmax(logits, dim=-1)
        nxt = torch.multinomial(probs, 1, generator=generator)
        idx = torch.cat([idx, nxt.to(idx.dtype)], dim=1)
    return idx


def load_tokenizer(data_dir: str):
    meta_path = os.path.join(data_dir, "meta.pkl")
    if os.path.exists(meta_path):
        print(f"Loading meta from LOCAL {meta_path}...")
        with open(meta_path, "rb") as f:
            meta = pickle.load(f)
        stoi, itos = meta["stoi"], meta["itos"]
        return (lambda s: [stoi[c] for c in s]), (lambda l: "".join(itos.get(i, "\ufffd") for i in l))
    print("No LOCAL meta.pkl found, assuming GPT-2 encodings...")
    try:
        import tiktoken
        enc = tiktoken.get_encoding("gpt2")
        return (lambda s: enc.encode(s, allowed_special={"<|endoftext|>"})), (lambda l: enc.decode(l))
    except ImportError:
        pass
    try:                                   # same vocabulary through a locally cached HF tokenizer (no network access attempted)
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained("gpt2", local_files_only=True)
        return (lambda s: tok.encode(s)), (lambda l: tok.decode(l))
    except Exception as e:
        raise SystemExit("GPT-2 BPE needs `tiktoken` (or a locally cached transformers 'gpt2' tokenizer); neither is available: "
                         f"{type(e).__name__}: {e}")


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--ckpt_dir", type=str, required=True)
    p.add_argument("--start", type=str, default="\n")
    p.add_argument("--num_samples", type=int, default=10)
    p.add_argument("--max_new_tokens", type=int, default=500)
    p.add_argument("--temperature", type=float, default=0.8)
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--seed", type=int, default=0)
    args = p.parse_args(argv)

    from midgpt_b200 import config_from_json
    from midgpt_b200.checkpoint.manager import CheckpointManager
    from midgpt_b200.models.gpt import GPTModel, DictProvider
    from midgpt_b200.ops import get_backend

    from midgpt_b200.utils.remote import RemoteDir, is_remote
    if is_remote(args.ckpt_dir):          # reference: gcsfs when the path starts with gs:// (sample.py:100-108)
        print(f"Using remote filesystem for checkpoint ({args.ckpt_dir.split('://', 1)[0]}://)")
        rd = RemoteDir(args.ckpt_dir)
        rd.pull_file("config.json")
        rd.pull_latest_step()
        args.ckpt_dir = rd.local
    else:
        print("Using local filesystem for checkpoint")
    with open(os.path.join(args.ckpt_dir, "config.json")) as f:
        config = config_from_json(f.read())
    print(config)
    mngr = CheckpointManager(args.ckpt_dir)
    step = mngr.latest_step()
    if step is None:
        raise SystemExit(f"no checkpoint found under {args.ckpt_dir}")
    device = torch.device(args.device or ("cuda" if torch.cuda.is_available() else "cpu"))
    compute = getattr(torch, config.compute_dtype) if device.type == "cuda" else torch.float32
    params = {k: v.to(device, compute) for k, v in mngr.restore_full_params(step, config).items()}
    mc = config.model_config
    model = GPTModel(mc, get_backend(config.engine.backend, device))
    provider = DictProvider(mc, params)
    encode, decode = load_tokenizer(config.data_dir)
    start = args.start
    if start.startswith("FILE:"):
        with open(start[5:], "r", encoding="utf-8") as f:
            start = f.read()
    ids = encode(start if start != "" else "\n")
    x = torch.tensor([ids for _ in range(args.num_samples)], dtype=torch.long, device=device)
    gen = torch.Generator(device=device).manual_seed(args.seed)
    print("generating samples...")
    with torch.no_grad():
        y = generate(model, provider, x, args.max_new_tokens, mc.block_size, args.temperature, gen)
    for i in range(args.num_samples):
        print(decode(y[i].tolist()))
        print("---------------")
    return y


if __name__ == "__main__":
    main()
