"""In-kernel NVLink communication for the FSDP engine (``comm="fused"``) — the product path; NCCL is the baseline.

Symmetric memory (``torch.distributed._symmetric_memory``: identical allocations on every rank, peer-mapped
through NVSwitch, optional NVLS multicast address) holds two arenas:

* ``grad_arena``  fp32 — every unit's full local gradient accumulator (what wgrad GEMMs add into);
* ``lp_arena``    bf16 — every unit's gathered compute weights (what forward/backward GEMMs read).

With both arenas peer-visible the two FSDP collectives disappear into kernels that were needed anyway
(csrc/comm.cu):

* ``rs_sumsq``         reduce-scatter ⊕ mean ⊕ Σg²: each rank *pulls* its shard of all peers' accumulators
                       (``multimem.ld_reduce`` = in-switch add when multicast is available, P2P loads otherwise);
* ``adamw_allgather``  clip·Adam·WD·LR on the fp32 shard, bf16 result *pushed* into every rank's ``lp_arena``
                       (``multimem.st`` or P2P stores) — the optimizer step is the all-gather.

Ordering uses a device-side all-ranks barrier kernel (system-scope release/acquire flags in a symmetric pad):
per unit before its pull (all ranks finished that unit's backward), once after all pulls (accumulators may
be zeroed), once after all pushes (weights visible everywhere).  No NCCL call moves tensor data; NCCL
carries only the 4-byte Σg² all-reduce (and, across nodes, the replica all-reduce of shard-sized gradients).

``gather_mode="layer"`` (ZeRO-3 weights): only the *shards* are symmetric; the gathered copy of a unit lives in a local pool
slot and is filled by a GEMM that **carries** the gather (all-gather ⊕ GEMM, csrc/gemm.cu: TMA bulk-copy rings in two extra
warps of every CTA pull the peers' shards; ``GatherSpec`` is the descriptor).  ``engine.grad_shard`` (ZeRO-2/3 gradients): the
symmetric gradient arena shrinks to three rotating full-size buffers and ``rs_sumsq`` runs per unit per micro-step in
accumulate mode (``reduce_scatter_accumulate``).
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist

from ..optim import step_scalars, clip_factor_from_sumsq


class GatherSpec:
    """Everything a GEMM needs to carry the all-gather of one unit while it multiplies (csrc/gemm.cu: copy warps in every CTA)."""

    def __init__(self, peer_ptrs, local, flags, chunk_elems, target, rank, n_ctas=0):
        self.peer_ptrs, self.local, self.flags = peer_ptrs, local, flags
        self.chunk_elems, self.target, self.rank = chunk_elems, target, rank
        self.done = False
        self.carried_by = None


AG_CTAS = 32     # CTAs of the stand-alone NVLink pull kernel (units whose first consumer is not a GEMM: the embedding)
SEG_ELEMS = 32768    # flag granularity (must match csrc/gemm.cu)


class FusedComm:
    def __init__(self, engine):
        import torch.distributed._symmetric_memory as symm_mem
        from ..ops.b200_backend import load_ext
        self.layer_mode = engine.gather_mode == "layer"
        self.ext = load_ext()
        self.e = engine
        mesh = engine.mesh
        self.rank, self.world = mesh.shard_rank, mesh.shard_size
        group = mesh.shard_group
        dev = engine.device
        total = sum(u.padded for u in engine.units)
        # gradient accumulators peers pull from: every unit in full size, or (sharded-gradient mode) the 3 rotating pool buffers
        self.grad_sharded = engine.grad_sharded
        gbig = max(u.padded for u in engine.units)
        self.grad_arena = symm_mem.empty(3 * gbig if self.grad_sharded else total, dtype=torch.float32, device=dev)
        # step mode: every rank holds the full gathered weights (peers push into it).  layer mode: only the local
        # shards are symmetric (peers pull from them inside the consuming GEMM) and gathered copies live in a local pool.
        lp_elems = total // self.world if self.layer_mode else total
        self.lp_arena = symm_mem.empty(lp_elems, dtype=engine.compute_dtype, device=dev)
        self.pad = symm_mem.empty(4096, dtype=torch.int32, device=dev)
        self.grad_arena.zero_()
        self.lp_arena.zero_()
        self.pad.zero_()
        torch.cuda.synchronize()
        self.h_grad = symm_mem.rendezvous(self.grad_arena, group)
        self.h_lp = symm_mem.rendezvous(self.lp_arena, group)
        self.h_pad = symm_mem.rendezvous(self.pad, group)
        self.grad_ptrs = [int(p) for p in self.h_grad.buffer_ptrs]
        self.lp_ptrs = [int(p) for p in self.h_lp.buffer_ptrs]
        self.pad_ptrs = [int(p) for p in self.h_pad.buffer_ptrs]
        use_mc = bool(int(__import__("os").environ.get("MIDGPT_MULTIMEM", "1")))
        self.mc_grad = int(self.h_grad.multicast_ptr or 0) if use_mc else 0          # 0 when NVLS multicast is unavailable
        self.mc_lp = int(self.h_lp.multicast_ptr or 0) if use_mc and self.mc_grad and not self.layer_mode else 0
        self.epoch = 0
        # re-point the engine's per-unit buffers into the arenas
        off = soff = 0
        self.offsets: tp.Dict[int, int] = {}
        self.shard_offsets: tp.Dict[int, int] = {}
        for i, u in enumerate(engine.units):
            self.offsets[i] = off
            self.shard_offsets[i] = soff
            if not self.grad_sharded:
                u.grad_full = self.grad_arena[off:off + u.padded]
            if self.layer_mode:
                u.lp_shard = self.lp_arena[soff:soff + u.shard]
            else:
                u.lp_full = self.lp_arena[off:off + u.padded]
                u.lp_shard = u.lp_full[self.rank * u.shard:(self.rank + 1) * u.shard]
            u.lp_shard.copy_(u.master)
            off += u.padded
            soff += u.shard
        if self.layer_mode:
            n_slots = len(engine._pool)
            n_seg = max(u.padded for u in engine.units) // SEG_ELEMS + 2 + 16
            self.slot_flags = [torch.zeros(n_seg, dtype=torch.int32, device=dev) for _ in range(n_slots)]
            self.slot_calls = [0] * n_slots
            self._scratch_flags = torch.zeros(8, dtype=torch.int32, device=dev)     # stand-alone pulls are ordered by the stream
        if self.grad_sharded:
            self.gbig = gbig
            engine._gpool = [self.grad_arena[k * gbig:(k + 1) * gbig] for k in range(3)]
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._pushed = False
        self.stats = {"xrank_barrier": 0, "rs_sumsq": 0, "adamw_allgather": 0, "ag_pull": 0, "gemm_ag": 0}
        self._barrier()          # every rank has zeroed / initialised its arenas
        if mesh.is_main:
            print(f"[fused-comm] symmetric arenas: grad {total * 4 / 2**30:.2f} GB fp32, weights {total * 2 / 2**30:.2f} GB bf16, "
                  f"multicast={'yes' if self.mc_grad else 'no (P2P loads/stores)'}")

    # ------------------------------------------------------------------
    def _barrier(self) -> None:
        self.epoch += 1
        self.ext.xrank_barrier(self.pad_ptrs, self.rank, 0, self.epoch)
        self.stats["xrank_barrier"] += 1

    def all_gather(self, u) -> None:
        """step mode: only the very first gather (initial weights) moves data; afterwards AdamW pushes keep lp_arena current.
        layer mode: stand-alone NVLink pull (used where no K-major GEMM is the first consumer, e.g. the embedding)."""
        if self.layer_mode:
            self.ext.ag_pull(self._peer_shard_ptrs(u), u.lp_full, u.shard, self.rank, self._scratch_flags, AG_CTAS)
            self.stats["ag_pull"] += 1
            return
        if self._pushed:
            return
        dist.all_gather_into_tensor(u.lp_full, u.lp_shard, group=self.e.mesh.shard_group)

    # ---- layer mode helpers ------------------------------------------------
    def _peer_shard_ptrs(self, u) -> tp.List[int]:
        i = self.e.units.index(u)
        return [p + self.shard_offsets[i] * 2 for p in self.lp_ptrs]

    def _next_flags(self, u) -> torch.Tensor:
        self.slot_calls[u.pool_slot] += 1
        return self.slot_flags[u.pool_slot]

    def gather_spec(self, u) -> GatherSpec:
        """Descriptor for the GEMM that will carry this unit's gather (FSDPEngine decides which one: normally a GEMM of the
        PREVIOUS unit of the schedule, so that the gather is hidden; the unit's own first GEMM for the first unit of a pass)."""
        flags = self._next_flags(u)
        return GatherSpec(self._peer_shard_ptrs(u), u.lp_full, flags, u.shard, self.slot_calls[u.pool_slot], self.rank, AG_CTAS)

    def reduce_scatter(self, u) -> None:
        """Called on the comm stream once this rank's backward of unit ``u`` is complete."""
        i = self.e.units.index(u)
        self._barrier()                                   # every rank finished accumulating this unit
        off = self.offsets[i] + self.rank * u.shard
        self.ext.rs_sumsq(self.grad_ptrs, self.mc_grad, u.grad_shard, self.sumsq, off, self.rank, False, True)
        self.stats["rs_sumsq"] += 1

    def reduce_scatter_accumulate(self, u, k: int, want_sumsq: bool) -> None:
        """Sharded-gradient mode, comm stream, every micro-step: pull this rank's shard of pool buffer k from every peer, ADD the
        mean to the fp32 shard accumulator (Σg² of the final value on the last micro-step), then zero the local buffer once every
        peer has pulled from it."""
        self._barrier()                                   # every rank finished this unit's backward of this micro-step
        off = k * self.gbig + self.rank * u.shard
        self.ext.rs_sumsq(self.grad_ptrs, self.mc_grad, u.grad_shard, self.sumsq, off, self.rank, True, want_sumsq)
        self.stats["rs_sumsq"] += 1
        self._barrier()                                   # every rank has pulled: the buffer may be cleared and reused
        self.grad_arena[k * self.gbig:k * self.gbig + u.padded].zero_()

    def optimizer_step(self) -> torch.Tensor:
        """Runs on the engine's compute stream after all reduce events have been waited on."""
        e = self.e
        if not self.grad_sharded:
            ev0 = e._mark()
            self._barrier()                               # all pulls everywhere are done -> accumulators reusable
            e._exposed(ev0)
            self.grad_arena.zero_()
        if e.mesh.replica_size > 1:
            # HSDP: rs_sumsq accumulated Σg² of the *intra-node* mean, before the replica all-reduce changed grad_shard;
            # every replica group would otherwise clip with its own norm and the replicas' weights would drift apart.
            total = torch.zeros(1, dtype=torch.float32, device=e.device)
            for u in e.units:
                self.ext.sumsq_accumulate(u.grad_shard, total)
        else:
            total = self.sumsq.clone()
        self.sumsq.zero_()
        ev0 = e._mark()
        dist.all_reduce(total, group=e.mesh.shard_group)   # 4 bytes
        if e.mesh.replica_size > 1:
            # Σg² is accumulated with floating-point atomics, whose order differs from GPU to GPU: without this (4-byte) average
            # the replicas would clip with factors one ulp apart and their weights would drift away from each other bit by bit
            dist.all_reduce(total, op=dist.ReduceOp.AVG, group=e.mesh.replica_group)
        e._exposed(ev0)
        clip, gnorm = clip_factor_from_sumsq(total[0])
        sc = step_scalars(e.count, e.cfg)
        clip1 = clip.reshape(1).float()
        for i, u in enumerate(e.units):
            if self.layer_mode:      # shards stay local (peers pull them later from inside their GEMMs)
                self.ext.adamw_step(u.master, u.exp_avg, u.exp_avg_sq, u.grad_shard, u.lp_shard, clip1, sc.lr, sc.b1, sc.b2, sc.eps,
                                    sc.wd_over_lr_peak, sc.bc1, sc.bc2, False)
                continue
            off = self.offsets[i] + self.rank * u.shard
            self.ext.adamw_allgather(u.master, u.exp_avg, u.exp_avg_sq, u.grad_shard, self.lp_ptrs, self.mc_lp, off, clip1,
                                     sc.lr, sc.b1, sc.b2, sc.eps, sc.wd_over_lr_peak, sc.bc1, sc.bc2)
            self.stats["adamw_allgather"] += 1
        if self.grad_sharded:
            for u in e.units:
                u.grad_shard.zero_()                      # shard-sized accumulators of the next step
        ev0 = e._mark()
        self._barrier()                                   # every rank's pushes have landed everywhere
        e._exposed(ev0)
        self._pushed = True
        return gnorm
