"""Host replay of the GEMM work schedule (csrc/gemm.cu: pick_split_k + WorkIter, the code the kernel's three roles run).
Every (tile, k-block) of the problem must be covered exactly once, whatever the tile count / worker count / K.  CPU tier: the
extension loads without a GPU; only the host-side planner is called."""
import itertools

import pytest

EPI_STORE, EPI_F32 = 0, 4


@pytest.fixture(scope="module")
def ext():
    try:
        from midgpt_b200.ops.b200_backend import load_ext
        return load_ext()
    except Exception as e:                     # not built in this checkout
        pytest.skip(f"extension not available: {e}")


def plan(ext, M, N, K, epi, acc, sms, pairs):
    flat = ext.gemm_work_plan(M, N, K, epi, acc, sms, pairs)
    return [tuple(flat[i:i + 4]) for i in range(0, len(flat), 4)]


@pytest.mark.parametrize("pairs", [True, False])
@pytest.mark.parametrize("M,N,K", [
    (6144, 2048, 32768), (2048, 2048, 32768), (8192, 2048, 32768), (2048, 8192, 32768),   # the 1.5B wgrad shapes
    (1024, 768, 8192), (2560, 2048, 16384), (256, 256, 64), (300, 200, 1000), (50304, 2048, 4096), (768, 768, 1024 * 33 + 7),
])
def test_wgrad_schedule_covers_every_kblock_once(ext, M, N, K, pairs):
    sms = 148
    tm = (M + 255) // 256 if pairs else (M + 127) // 128          # a CTA pair owns a 256-row tile
    tn = (N + 255) // 256
    num_k = (K + 63) // 64
    pieces = plan(ext, M, N, K, EPI_F32, 1, sms, pairs)
    seen = {}
    for w, tile, kb0, kb1 in pieces:
        assert 0 <= tile < tm * tn and 0 <= kb0 < kb1 <= num_k
        for kb in range(kb0, kb1):
            assert (tile, kb) not in seen, f"k-block {kb} of tile {tile} given to workers {seen[(tile, kb)]} and {w}"
            seen[(tile, kb)] = w
    assert len(seen) == tm * tn * num_k
    # balance: nobody does more than ~1.35x the mean work (epilogues not counted) once there is at least a wave of work
    per = {}
    for w, _, kb0, kb1 in pieces:
        per[w] = per.get(w, 0) + kb1 - kb0
    workers = 74 if pairs else 148
    if tm * tn * num_k >= 64 * workers:
        assert max(per.values()) <= 1.35 * (tm * tn * num_k / workers) + 16


def test_plain_schedule_is_whole_tiles(ext):
    for (M, N, K), pairs in itertools.product([(8192, 6144, 2048), (1000, 520, 384)], [True, False]):
        pieces = plan(ext, M, N, K, EPI_STORE, 0, 148, pairs)
        num_k = (K + 63) // 64
        tiles = sorted(t for _, t, _, _ in pieces)
        assert tiles == list(range(len(tiles))) and all(kb0 == 0 and kb1 == num_k for _, _, kb0, kb1 in pieces)
