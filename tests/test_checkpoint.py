"""Checkpoint round-trip, interval/max_to_keep semantics, resume equivalence, resharding on restore."""
import os

import numpy as np
import torch

from tests.test_fsdp_cpu import _tiny_cfg, _batches
from tests._mp import run_distributed
from midgpt_b200.checkpoint.manager import CheckpointManager
from midgpt_b200.checkpoint import layout


def _trainer(cfg):
    from midgpt_b200.train import Trainer
    return Trainer(cfg, device_type="cpu")


def test_layout_matches_reference_leaf_counts():
    assert len(layout.MODEL_LEAVES) == 22 and len(layout.ARRAY_ORDER) == 8
    opt = layout.opt_layout()
    assert len(opt) == 18 and opt[0][1] == "count" and opt[-1][1] == "count"    # count, mu×8, nu×8, count
    assert opt[-1][0] == "1.17"


def test_save_restore_roundtrip_and_interval(tmp_path):
    cfg = _tiny_cfg(False)
    tr = _trainer(cfg)
    batches = _batches(cfg, 3)
    for x, y in batches:
        tr.train_step(x, y)
    m = CheckpointManager(str(tmp_path), max_to_keep=1, save_interval_steps=2, async_save=True)
    assert not m.save(1, tr.engine)               # off-interval → skipped, like Orbax
    assert m.save(2, tr.engine, extra={"sampler": {"k": 1}})
    m.wait_until_finished()
    assert m.latest_step() == 2
    assert os.path.exists(tmp_path / "2" / "default" / "_METADATA")
    assert os.path.exists(tmp_path / "2" / "default" / "checkpoint")
    assert os.path.exists(tmp_path / "2" / "default" / "0.5" / ".zarray")       # c_attn stacked leaf
    assert m.save(4, tr.engine, force=True)
    m.wait_until_finished()
    assert m.all_steps() == [4]                   # max_to_keep=1
    ref = {w: tr.engine.full_state(w) for w in ("master", "exp_avg", "exp_avg_sq")}
    tr2 = _trainer(_tiny_cfg(False))
    tr2.engine.count = 0
    m.restore(4, tr2.engine)
    assert tr2.engine.count == tr.engine.count
    for w in ref:
        got = tr2.engine.full_state(w)
        for k in ref[w]:
            assert torch.equal(got[k], ref[w][k]), (w, k)
    full = m.restore_full_params(4, cfg)
    assert torch.equal(full["block.1.c_fc"], ref["master"]["block.1.c_fc"])


def test_resume_equivalence(tmp_path):
    """train 4 steps == train 2, checkpoint, restore into a fresh trainer, train 2 more."""
    cfg = _tiny_cfg(False)
    batches = _batches(cfg, 4)
    a = _trainer(cfg)
    la = [a.train_step(x, y).item() for x, y in batches]
    b = _trainer(_tiny_cfg(False))
    for x, y in batches[:2]:
        b.train_step(x, y)
    m = CheckpointManager(str(tmp_path), async_save=False)
    m.save(1, b.engine, force=True)
    c = _trainer(_tiny_cfg(False))
    m.restore(1, c.engine)
    lc = [c.train_step(x, y).item() for x, y in batches[2:]]
    assert np.allclose(lc, la[2:], rtol=1e-6, atol=1e-7)
    sa, sc = a.engine.full_state("master"), c.engine.full_state("master")
    for k in sa:
        assert torch.allclose(sa[k], sc[k], rtol=1e-6, atol=1e-8), k


def _save_ws2(rank, world, d):
    cfg = _tiny_cfg(True)
    tr = _trainer(cfg)
    for x, y in _batches(cfg, 2):
        G, B, T = cfg.g_accum_iters, cfg.batch_size, cfg.model_config.block_size
        xs = x.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        ys = y.view(G, B, T)[:, rank * B // world:(rank + 1) * B // world].reshape(-1, T)
        tr.train_step(xs, ys)
    m = CheckpointManager(d, async_save=True)
    m.save(1, tr.engine, force=True)
    m.wait_until_finished()
    torch.distributed.barrier()
    return tr.engine.full_state("master")


def test_sharded_save_then_restore_on_one_process(tmp_path):
    outs = run_distributed(_save_ws2, 2, str(tmp_path))
    cfg = _tiny_cfg(False)
    tr = _trainer(cfg)
    m = CheckpointManager(str(tmp_path))
    assert m.latest_step() == 1
    m.restore(1, tr.engine)
    got = tr.engine.full_state("master")
    for k, v in outs[0].items():
        assert torch.equal(got[k], torch.as_tensor(v)), k
    assert tr.engine.count == 2


def test_orbax_shaped_metadata_schema(tmp_path):
    """Pins the on-disk schema (layout.py documents where each field comes from)."""
    import json
    import msgpack
    cfg = _tiny_cfg(False)
    tr = _trainer(cfg)
    m = CheckpointManager(str(tmp_path), async_save=False)
    m.save(0, tr.engine, force=True)
    item = tmp_path / "0" / "default"
    meta = json.load(open(item / "_METADATA"))
    assert meta["use_zarr3"] is False and len(meta["tree_metadata"]) == 22 + 18
    e = meta["tree_metadata"]["('0', '5')"]                          # c_attn
    assert e["key_metadata"] == [{"key": "0", "key_type": 1}, {"key": "5", "key_type": 1}]
    assert e["value_metadata"] == {"value_type": "jax.Array", "skip_deserialize": False}
    assert meta["tree_metadata"]["('1', '0')"]["value_metadata"]["value_type"] == "scalar"     # Adam count
    agg = msgpack.unpackb(open(item / "checkpoint", "rb").read())
    assert set(agg) == {"0", "1"} and len(agg["0"]) == 22 and len(agg["1"]) == 18
    assert agg["0"]["5"] == "PLACEHOLDER://0.5" and agg["0"]["21"] == cfg.model_config.n_layer and agg["1"]["17"] == 0
    za = json.load(open(item / "0.5" / ".zarray"))
    L, D = cfg.model_config.n_layer, cfg.model_config.n_embd
    assert za["shape"] == [L, 3 * D, D] and za["chunks"] == [1, 3 * D, D] and za["compressor"] is None


def _save_no_collective(rank, world, d):
    """Saving must not move tensor data through collectives: every rank writes only the bytes it owns."""
    import midgpt_b200.parallel.dist as D
    cfg = _tiny_cfg(True)
    tr = _trainer(cfg)
    called = []
    orig = D.all_gather_flat
    D.all_gather_flat = lambda *a, **k: (called.append(1), orig(*a, **k))[1]
    m = CheckpointManager(d, async_save=True)
    m.save(0, tr.engine, force=True, extra={"sampler": {"rank": rank}})
    m.wait_until_finished()
    torch.distributed.barrier()
    D.all_gather_flat = orig
    ex = m.restore(0, tr.engine)
    return len(called), ex


def test_save_uses_no_all_gather_and_keeps_per_rank_extra(tmp_path):
    outs = run_distributed(_save_no_collective, 2, str(tmp_path))
    assert [o[0] for o in outs] == [0, 0]
    assert [o[1]["sampler"]["rank"] for o in outs] == [0, 1]
