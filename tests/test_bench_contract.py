"""bench.py / __graft_entry__ contracts that can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_unavailable():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line and isinstance(line["unavailable"], str)


def test_graft_entry_build_is_idempotent_and_products_exist():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from midgpt_b200.ops import build as b
    assert os.path.exists(b.ext_path()) and os.path.exists(b.loader_path())
    for name in ("gemm", "attn_fwd", "attn_bwd", "rmsnorm_fwd", "qkln_rope_fwd", "adamw_step", "rs_sumsq", "adamw_allgather",
                 "gemm_ag", "xrank_barrier", "ce_chunk", "embed_fwd", "dropout"):
        from midgpt_b200.ops.b200_backend import load_ext
        assert hasattr(load_ext(), name), name
