"""Distributed smoke test for a GPU node — the counterpart of the reference's scripts/test_jax.py:40-58 (shard two matrices
over all devices, multiply, look at the result) and scripts/test_ckpt.py:8-24 (save one sharded array through the async
checkpoint manager), run on every rank:

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/test_dist.py [rundir]

1. symmetric-memory rendezvous + device barrier kernel;
2. all-gather ⊕ GEMM (NVLink pull inside the tcgen05 GEMM) vs a NCCL all-gather followed by the plain GEMM;
3. sharded checkpoint write/commit of a (8192, 12800) fp32 array (the reference's test array) via the manager's zarr writer.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from midgpt_b200.parallel.dist import init_distributed  # noqa: E402


def main():
    rank, world, local, dev, backend = init_distributed("cuda")
    assert world > 1, "run under torchrun with >= 2 ranks"
    import torch.distributed._symmetric_memory as symm
    from midgpt_b200.ops.b200_backend import B200Backend
    from midgpt_b200.parallel.fused import GatherSpec, AG_CTAS
    be = B200Backend(dev)
    N, K, M = 6144, 2048, 8192
    chunk = N * K // world
    torch.manual_seed(0)
    w_full = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    shard = symm.empty(chunk, dtype=torch.bfloat16, device=dev)
    shard.copy_(w_full.view(-1)[rank * chunk:(rank + 1) * chunk])
    pad = symm.empty(1024, dtype=torch.int32, device=dev); pad.zero_()
    h_shard, h_pad = symm.rendezvous(shard, dist.group.WORLD), symm.rendezvous(pad, dist.group.WORLD)
    pads = [int(p) for p in h_pad.buffer_ptrs]
    torch.cuda.synchronize()
    be.ext.xrank_barrier(pads, rank, 0, 1)
    x = torch.randn(M, K, device=dev).bfloat16()
    local_buf = torch.zeros(N * K, dtype=torch.bfloat16, device=dev)
    local_buf[rank * chunk:(rank + 1) * chunk].copy_(shard)
    flags = torch.zeros(4096, dtype=torch.int32, device=dev)
    spec = GatherSpec([int(p) for p in h_shard.buffer_ptrs], local_buf, flags, chunk, 1, rank, AG_CTAS)
    y = be.linear_fwd(x, local_buf.view(N, K), gather=spec)
    ref = be.linear_fwd(x, w_full)
    err = ((y.float() - ref.float()).norm() / ref.float().norm()).item()
    torch.cuda.synchronize()
    be.ext.xrank_barrier(pads, rank, 0, 2)
    print(f"[rank {rank}] all-gather ⊕ GEMM rel err vs plain GEMM on the full weight: {err:.2e}", flush=True)
    assert err < 1e-6
    if len(sys.argv) > 1:
        from midgpt_b200.checkpoint import zarrio
        d = os.path.join(sys.argv[1], "test", "0", "default", "0.0")
        rows = 8192 // world
        if rank == 0:
            zarrio.write_zarray_meta(d, (8192, 12800), (rows, 12800), "float32")
        dist.barrier()
        t0 = time.time()
        zarrio.write_chunk(d, (rank, 0), torch.full((rows, 12800), float(rank)).numpy())
        dist.barrier()
        if rank == 0:
            a = zarrio.read_array(d)
            assert a.shape == (8192, 12800) and a[-1, -1] == world - 1
            print(f"sharded checkpoint array written by {world} ranks in {time.time() - t0:.2f}s")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
