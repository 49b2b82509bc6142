#!/usr/bin/env python
"""Training entry point — same flags as the reference ``launch.py`` (launch.py:15-20), B200 process model.

    python launch.py --config=shakespeare_char                       # 1 GPU (or CPU)
    python launch.py --config=openwebtext_xl --nproc 8 --rundir=out/xl   # one process per GPU (self-spawning)
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 launch.py --config=openwebtext_xl --rundir=out/xl

Behaviour kept from the reference: default rundir ``outputs/<timestamp>`` (launch.py:32-34);
``--multihost`` requires ``--rundir`` (launch.py:31); ``--debug`` ⇒ profiler trace, one eval batch,
no checkpoints, no rundir files (launch.py:35-41); rank 0 writes ``config.json`` and reads-or-creates
``wandb_id.txt`` so a re-run with the same ``--rundir`` resumes the same run (launch.py:53-68);
re-running with an existing rundir resumes from ``latest_step + 1`` (src/train.py:179-187).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import uuid
from datetime import datetime


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--config", type=str, required=True)
    p.add_argument("--rundir", type=str)
    p.add_argument("--debug", action="store_true")
    p.add_argument("--multihost", action="store_true",
                   help="multi-node run: rendezvous comes from torchrun/SLURM env (RANK, WORLD_SIZE, MASTER_ADDR)")
    # additions (the reference edits config files instead)
    p.add_argument("--nproc", type=int, default=0, help="spawn this many local ranks (one per GPU); 0 = use env / single process")
    p.add_argument("--shard_model", type=int, default=None, help="override config.shard_model (1 = FSDP, 0 = pure DP)")
    p.add_argument("--backend", type=str, default=None, choices=["auto", "b200", "torch"])
    p.add_argument("--comm", type=str, default=None, choices=["auto", "fused", "nccl", "gloo"])
    p.add_argument("--max_steps", type=int, default=None)
    p.add_argument("--synthetic", action="store_true", help="train on a synthetic corpus (no dataset on disk)")
    p.add_argument("--device", type=str, default=None, choices=["cuda", "cpu"])
    p.add_argument("--set", action="append", default=[], metavar="FIELD=VALUE",
                   help="override any ExperimentConfig / model_config.* / engine.* field")
    return p.parse_args(argv)


def strip_nproc(argv):
    """Child argv for the self-spawn: drop ``--nproc N`` (two tokens) and ``--nproc=N`` (one token)."""
    out, skip = [], False
    for a in argv:
        if skip:
            skip = False
            continue
        if a == "--nproc":
            skip = True
            continue
        if a.startswith("--nproc="):
            continue
        out.append(a)
    return out


def apply_overrides(config, pairs):
    for kv in pairs:
        k, v = kv.split("=", 1)
        obj = config
        parts = k.split(".")
        for part in parts[:-1]:
            obj = getattr(obj, part)
        cur = getattr(obj, parts[-1])
        if isinstance(cur, bool):
            val = v.lower() in ("1", "true", "yes")
        elif isinstance(cur, int):
            val = int(v)
        elif isinstance(cur, float):
            val = float(v)
        else:
            val = v
        setattr(obj, parts[-1], val)
    return config


def main(argv=None):
    args = parse_args(argv)
    if args.nproc and "RANK" not in os.environ:
        # self-spawn: one process per GPU on this node (replaces the reference's pod fan-out, scripts/tpu_commands.sh)
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.nproc}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + strip_nproc(list(argv or sys.argv[1:]))
        if not args.rundir and not args.debug:
            cmd += ["--rundir", os.path.join("outputs", datetime.now().strftime("%Y-%m-%d-%H-%M-%S"))]
        raise SystemExit(subprocess.call(cmd))

    from midgpt_b200 import load_config, config_to_json
    from midgpt_b200.train import Trainer

    config = load_config(args.config)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.rundir is not None:
        config.rundir = args.rundir
    elif not args.debug:
        assert not args.multihost and world == 1, "Multi-process runs must pre-specify --rundir."
        config.rundir = os.path.join("outputs", datetime.now().strftime("%Y-%m-%d-%H-%M-%S"))
    if args.debug:
        config.debug = True
    if args.shard_model is not None:
        config.shard_model = bool(args.shard_model)
    if args.backend:
        config.engine.backend = args.backend
    if args.comm:
        config.engine.comm = args.comm
    apply_overrides(config, args.set)

    wandb_id = None
    from midgpt_b200.utils.remote import RemoteDir, is_remote
    remote = RemoteDir(config.rundir) if (not args.debug and is_remote(config.rundir)) else None
    if rank == 0 and not args.debug:
        print(f"Writing to {config.rundir}")
        if remote is not None:          # reference: gcsfs for gs:// paths (launch.py:43-50); here any fsspec scheme, via a local mirror
            print(f"Using remote filesystem ({config.rundir.split('://', 1)[0]}://), local mirror {remote.local}")
            local_dir = remote.local
            remote.pull_file("wandb_id.txt")
        else:
            print("Using local filesystem")
            config.rundir = os.path.abspath(config.rundir)
            local_dir = config.rundir
        os.makedirs(local_dir, exist_ok=True)
        with open(os.path.join(local_dir, "config.json"), "w") as f:
            f.write(config_to_json(config))
        wid_path = os.path.join(local_dir, "wandb_id.txt")
        if os.path.exists(wid_path):
            wandb_id = open(wid_path).read().strip()
        else:
            wandb_id = uuid.uuid4().hex[:8]
            with open(wid_path, "w") as f:
                f.write(wandb_id)
        if remote is not None:
            remote.push_file("config.json")
            remote.push_file("wandb_id.txt")
    elif not args.debug and remote is None:
        config.rundir = os.path.abspath(config.rundir)

    trainer = Trainer(config, device_type=args.device)
    trainer.wandb_id = wandb_id
    if trainer.mesh.is_main:
        print(json.dumps(config.reference_dict(), indent=1))
    out = trainer.fit(synthetic=True if args.synthetic else None, max_steps=args.max_steps)
    from midgpt_b200.parallel import dist as D
    D.barrier()
    D.shutdown()
    return out


if __name__ == "__main__":
    main()
