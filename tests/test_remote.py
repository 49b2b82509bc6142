"""Object-store run directories (utils/remote.py): the reference writes config.json / wandb_id.txt / checkpoints to GCS when
--rundir starts with gs:// (launch.py:43-57).  Exercised without a network through fsspec's memory:// and file:// drivers."""
import json
import os
import shutil
import subprocess
import sys

import pytest

fsspec = pytest.importorskip("fsspec")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_remote_dir_files_and_steps(tmp_path):
    from midgpt_b200.utils.remote import RemoteDir, is_remote
    assert is_remote("gs://b/run") and is_remote("file:///x") and not is_remote("/data/run") and not is_remote(None)
    rd = RemoteDir("memory://bucket/run_a", local=str(tmp_path / "mirror"))
    assert rd.steps() == [] and rd.pull_latest_step() is None and not rd.pull_file("config.json")
    (tmp_path / "mirror" / "config.json").write_text("{}")
    rd.push_file("config.json")
    for step in (3, 6):
        d = tmp_path / "mirror" / str(step) / "default" / "0.0"
        d.mkdir(parents=True)
        (d.parent / "checkpoint").write_bytes(b"agg%d" % step)
        (d / "0.0").write_bytes(b"chunk%d" % step)
        rd.push_step(step, max_to_keep=1)
    assert rd.steps() == [6]                                   # step 3 pruned remotely, no .orbax-checkpoint-tmp left behind
    assert not any("tmp" in p for p in rd.fs.ls(rd.root, detail=False))
    # a fresh machine: empty mirror, same URL
    rd2 = RemoteDir("memory://bucket/run_a", local=str(tmp_path / "mirror2"))
    assert rd2.pull_file("config.json") and rd2.pull_latest_step() == 6
    assert (tmp_path / "mirror2" / "6" / "default" / "0.0" / "0.0").read_bytes() == b"chunk6"
    assert (tmp_path / "mirror2" / "6" / "default" / "checkpoint").read_bytes() == b"agg6"


def test_missing_driver_is_explained(tmp_path):
    from midgpt_b200.utils.remote import RemoteDir
    try:
        import gcsfs  # noqa: F401
        pytest.skip("gcsfs is installed here")
    except ImportError:
        pass
    with pytest.raises(SystemExit) as e:
        RemoteDir("gs://bucket/run", local=str(tmp_path / "m"))
    assert "gcsfs" in str(e.value)


def _run(args, env_extra):
    env = dict(os.environ, WANDB_MODE="disabled", PYTHONPATH=ROOT, **env_extra)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_launch_resume_sample_through_a_remote_rundir(tmp_path):
    """launch → (mirror wiped = another machine) → launch again resumes from the remote copy → sample reads the remote copy."""
    data, store, mirror = tmp_path / "data", tmp_path / "store" / "run", tmp_path / "mirrors"
    _run(["data/shakespeare_char/prepare.py", "--synthetic", "--out", str(data)], {})
    url = f"file://{store}"
    env = {"MIDGPT_MIRROR_DIR": str(mirror)}
    common = ["launch.py", "--config=shakespeare_char", f"--rundir={url}", "--device=cpu",
              f"--set=data_dir={data}", "--set=model_config.n_layer=2", "--set=model_config.n_embd=32",
              "--set=model_config.n_head=2", "--set=model_config.block_size=32", "--set=batch_size=4",
              "--set=eval_interval=3", "--set=compute_dtype=float32", "--set=debug=false"]
    out = _run(common + ["--max_steps=4"], env)
    assert "Using remote filesystem (file://)" in out
    assert json.load(open(store / "config.json"))["model_config"]["n_layer"] == 2
    wid = open(store / "wandb_id.txt").read()
    assert (store / "3" / "default" / "checkpoint").exists() and (store / "3" / "default" / "_METADATA").exists()
    assert any("loss/optimized" in l for l in open(store / "metrics.jsonl"))
    shutil.rmtree(mirror)
    out2 = _run(common + ["--max_steps=7"], env)
    assert "resumed from step 3" in out2
    assert open(store / "wandb_id.txt").read() == wid
    assert (store / "6" / "default" / "checkpoint").exists() and not (store / "3").exists()      # max_to_keep=1 holds remotely too
    shutil.rmtree(mirror)
    out3 = _run(["sample.py", f"--ckpt_dir={url}", "--num_samples=2", "--max_new_tokens=8", "--device=cpu"], env)
    assert out3.count("---------------") == 2
