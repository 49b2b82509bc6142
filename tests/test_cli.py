"""End-to-end CLI: prepare (synthetic) → launch → resume with the same rundir → sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, **kw):
    env = dict(os.environ, WANDB_MODE="disabled", PYTHONPATH=ROOT)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600, **kw)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_prepare_launch_resume_sample(tmp_path):
    data = tmp_path / "data"
    _run(["data/shakespeare_char/prepare.py", "--synthetic", "--out", str(data)])
    assert (data / "train.bin").exists() and (data / "meta.pkl").exists()
    run = tmp_path / "run"
    common = ["launch.py", "--config=shakespeare_char", f"--rundir={run}", "--device=cpu",
              f"--set=data_dir={data}", "--set=model_config.n_layer=2", "--set=model_config.n_embd=32",
              "--set=model_config.n_head=2", "--set=model_config.block_size=32", "--set=batch_size=4",
              "--set=eval_interval=3", "--set=compute_dtype=float32", "--set=debug=false"]
    out = _run(common + ["--max_steps=4"])
    assert "Model has" in out
    cfgj = json.load(open(run / "config.json"))
    assert cfgj["model_config"]["n_layer"] == 2
    wid = open(run / "wandb_id.txt").read()
    assert (run / "3" / "default" / "checkpoint").exists()
    lines = [json.loads(l) for l in open(run / "metrics.jsonl")]
    assert any("loss/val" in l for l in lines) and any("loss/optimized" in l for l in lines)
    out2 = _run(common + ["--max_steps=7"])
    assert "resumed from step 3" in out2
    assert open(run / "wandb_id.txt").read() == wid
    assert (run / "6" / "default" / "checkpoint").exists() and not (run / "3").exists()   # max_to_keep=1
    out3 = _run(["sample.py", f"--ckpt_dir={run}", "--num_samples=2", "--max_new_tokens=8", "--device=cpu"])
    assert out3.count("---------------") == 2


def test_strip_nproc_both_spellings():
    sys.path.insert(0, ROOT)
    import launch
    assert launch.strip_nproc(["--config=a", "--nproc", "8", "--rundir", "r"]) == ["--config=a", "--rundir", "r"]
    assert launch.strip_nproc(["--nproc=8", "--config=a"]) == ["--config=a"]
    args = launch.parse_args(launch.strip_nproc(["--config", "shakespeare_char", "--nproc", "2", "--debug"]))
    assert args.config == "shakespeare_char" and args.debug and args.nproc == 0


def test_sigterm_checkpoints_and_resume_continues(tmp_path):
    """Preemption: SIGTERM makes a run write a checkpoint at the next step boundary and exit cleanly; the same command resumes."""
    import signal
    import time
    data = tmp_path / "data"
    _run(["data/shakespeare_char/prepare.py", "--synthetic", "--out", str(data)])
    run = tmp_path / "run"
    cmd = [sys.executable, "launch.py", "--config=shakespeare_char", f"--rundir={run}", "--device=cpu", f"--set=data_dir={data}",
           "--set=model_config.n_layer=2", "--set=model_config.n_embd=32", "--set=model_config.n_head=2", "--set=model_config.block_size=32",
           "--set=batch_size=4", "--set=eval_interval=100000", "--set=compute_dtype=float32", "--set=engine.log_every=1", "--max_steps=100000"]
    env = dict(os.environ, WANDB_MODE="disabled", PYTHONPATH=ROOT)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    p = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    metrics = run / "metrics.jsonl"
    deadline = time.time() + 300
    while time.time() < deadline and not (metrics.exists() and sum(1 for l in open(metrics) if "loss/optimized" in l) >= 5):
        time.sleep(0.2)
    p.send_signal(signal.SIGTERM)
    out, _ = p.communicate(timeout=120)
    assert p.returncode == 0, out[-3000:]
    assert "[signal] stop requested" in out
    steps = sorted(int(d) for d in os.listdir(run) if d.isdigit())
    assert len(steps) == 1 and steps[0] >= 4                      # a checkpoint off the (huge) interval: forced by the signal
    out2 = _run(cmd[1:-1] + [f"--max_steps={steps[0] + 3}"])
    assert f"resumed from step {steps[0]}" in out2


def test_resumed_run_continues_the_same_loss_sequence(tmp_path):
    """End to end: (train 10 steps) == (train 6 steps, stop, resume to 10) step for step -- model / optimizer state, schedule count
    AND the sample stream (the loader checkpoints the sampler state that follows the batch actually consumed)."""
    data = tmp_path / "data"
    _run(["data/shakespeare_char/prepare.py", "--synthetic", "--out", str(data)])

    def launch(run, max_steps):
        return _run(["launch.py", "--config=shakespeare_char", f"--rundir={run}", "--device=cpu", f"--set=data_dir={data}",
                     "--set=model_config.n_layer=2", "--set=model_config.n_embd=32", "--set=model_config.n_head=2",
                     "--set=model_config.block_size=32", "--set=model_config.dropout=0.0", "--set=batch_size=4", "--set=eval_interval=4",
                     "--set=compute_dtype=float32", "--set=engine.log_every=1", f"--max_steps={max_steps}"])

    def losses(run):
        out = {}
        for l in open(run / "metrics.jsonl"):
            d = json.loads(l)
            if "loss/optimized" in d:
                out[d["step"]] = d["loss/optimized"]          # a resumed run re-logs from its first step: the later entry wins
        return out

    a, b = tmp_path / "a", tmp_path / "b"
    launch(a, 10)
    launch(b, 6)                                               # checkpoint at step 4, step 5 is run and then lost
    out = launch(b, 10)
    assert "resumed from step 4" in out
    la, lb = losses(a), losses(b)
    assert sorted(lb) == list(range(10))
    for step in range(10):
        assert abs(la[step] - lb[step]) < 1e-6, (step, la[step], lb[step])


def test_two_rank_cli_run_resumes_exactly(tmp_path):
    """`launch.py --nproc 2` (self-spawned torchrun, gloo on CPU, FSDP over 2 ranks): per-rank shard writes into the same Zarr chunks,
    per-rank sampler state, resume == uninterrupted run."""
    data = tmp_path / "data"
    _run(["data/shakespeare_char/prepare.py", "--synthetic", "--out", str(data)])

    def launch(run, max_steps, port):
        env = dict(os.environ, WANDB_MODE="disabled", PYTHONPATH=ROOT, MASTER_PORT=str(port))
        env.pop("RANK", None); env.pop("WORLD_SIZE", None)
        r = subprocess.run([sys.executable, "launch.py", "--config=shakespeare_char", f"--rundir={run}", "--device=cpu", "--nproc", "2",
                            "--shard_model=1", f"--set=data_dir={data}", "--set=model_config.n_layer=2", "--set=model_config.n_embd=32",
                            "--set=model_config.n_head=2", "--set=model_config.block_size=32", "--set=model_config.dropout=0.0",
                            "--set=batch_size=4", "--set=eval_interval=4", "--set=compute_dtype=float32", "--set=engine.log_every=1",
                            f"--max_steps={max_steps}"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        return r.stdout + r.stderr

    def losses(run):
        out = {}
        for l in open(run / "metrics.jsonl"):
            d = json.loads(l)
            if "loss/optimized" in d:
                out[d["step"]] = d["loss/optimized"]
        return out

    from tests._mp import _free_port
    a, b = tmp_path / "a", tmp_path / "b"
    out_a = launch(a, 9, _free_port())
    assert "world=2 shard_group=2" in out_a
    launch(b, 6, _free_port())
    assert (b / "4" / "extra_rank1.json").exists()
    out_b = launch(b, 9, _free_port())
    assert "resumed from step 4" in out_b
    la, lb = losses(a), losses(b)
    for step in range(9):
        assert abs(la[step] - lb[step]) < 1e-6, (step, la[step], lb[step])
