"""Hand-written sm_100a backend: every primitive is one of the kernels in ``csrc/`` (tcgen05/TMEM/TMA GEMM and
flash attention, fused element-wise kernels).  Same interface as ``TorchBackend`` (which is its numerical oracle
in ``tests/test_kernels_gpu.py``).  There is **no silent fallback**: if the extension is missing on a GPU box
the constructor raises.

GEMM operand conventions (csrc/gemm.cu): ``ext.gemm(a, a_mn, b, b_mn, d, epi, d2, aux, accumulate)`` computes
``d (+)= A·Bᵀ`` with ``A = a`` ([M,K]) or ``aᵀ`` (a stored [K,M], ``a_mn``), ``B = b`` ([N,K]) or ``bᵀ`` (b stored [K,N]).
"""
from __future__ import annotations

import importlib.util
import os
import typing as tp

import torch

from .torch_backend import TorchBackend

EPI_STORE, EPI_RESIDUAL, EPI_GELU, EPI_DGELU, EPI_F32 = 0, 1, 2, 3, 4

_ext = None


def load_ext():
    """Import the in-tree extension; build it first if it is missing or stale."""
    global _ext
    if _ext is not None:
        return _ext
    from . import build
    path = build.ext_path()
    if not os.path.exists(path) or os.environ.get("MIDGPT_REBUILD"):
        build.build_ext()
    spec = importlib.util.spec_from_file_location("_ext", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _ext = mod
    return mod


class QKV(tp.NamedTuple):
    """Opaque attention input: LN+RoPE'd q,k packed [M,2D] and the raw projection [M,3D] (v is read from it)."""
    qk: torch.Tensor
    raw: torch.Tensor


class B200Backend:
    name = "b200"

    def __init__(self, device: torch.device, fp8: bool = False, fp8_format: str = "mx"):
        assert fp8_format in ("mx", "row")
        self.fp8, self.fp8_format = fp8, fp8_format
        self._wq: tp.Dict[int, tp.Tuple[int, torch.Tensor, torch.Tensor]] = {}
        self._wver = 0
        if device.type != "cuda":
            raise RuntimeError("B200Backend needs a CUDA device")
        cap = torch.cuda.get_device_capability(device)
        if cap[0] != 10:
            raise RuntimeError(f"B200Backend is sm_100a-only; found compute capability {cap}")
        self.ext = load_ext()
        self.device = device
        self._torch = TorchBackend(use_aten_flash=False)
        self._rope: tp.Dict[tp.Tuple[int, int], tp.Tuple[torch.Tensor, torch.Tensor]] = {}
        self._logits_ws: torch.Tensor | None = None
        self._dq_acc: tp.Dict[tp.Tuple[int, int], torch.Tensor] = {}
        self.launches = 0          # kernels launched through this backend (bench.py reports it)
        self.carried = 0           # GEMM launches that carried an FSDP unit's all-gather (all-gather ⊕ GEMM)

    # ---------------------------------------------------------------- helpers
    def _gemm(self, a, a_mn, b, b_mn, d, epi=EPI_STORE, d2=None, aux=None, accumulate=False, gather=None):
        if gather is not None and not gather.done:
            # all-gather ⊕ GEMM: this GEMM carries the NVLink pull of the weight unit (copy warps in its own CTAs, csrc/gemm.cu);
            # if b is one of the unit's weights the kernel synchronises with the pull segment by segment
            self.ext.gemm_ag(a, a_mn, b, b_mn, d, epi, d2, aux, accumulate, gather.peer_ptrs, gather.local, gather.flags,
                             gather.chunk_elems, gather.target, gather.rank)
            gather.done = True
            self.carried += 1
            gather.carried_by = "wgrad" if epi == EPI_F32 else ("dgrad" if b_mn else "fwd")
        else:
            self.ext.gemm(a, a_mn, b, b_mn, d, epi, d2, aux, accumulate)
        self.launches += 1

    def _rope_tables(self, C: int, T: int):
        key = (C, T)
        if key not in self._rope:
            inv = 1.0 / (10000.0 ** (torch.arange(0, C, 2, dtype=torch.float64, device=self.device) / C))
            ang = torch.outer(torch.arange(T, dtype=torch.float64, device=self.device), inv)
            self._rope[key] = (ang.sin().float().contiguous(), ang.cos().float().contiguous())
        return self._rope[key]

    # ---------------------------------------------------------------- embedding
    def embed_fwd(self, wte, idx):
        self.launches += 1
        return self.ext.embed_fwd(wte, idx.reshape(-1).to(torch.int32).contiguous())

    def embed_bwd(self, dx, idx, dwte_acc):
        self.launches += 1
        self.ext.embed_bwd(dx.contiguous(), idx.reshape(-1).to(torch.int32).contiguous(), dwte_acc)

    # ---------------------------------------------------------------- dropout
    def dropout(self, x, p, seed):
        if p == 0.0:
            return x
        self.launches += 1
        return self.ext.dropout(x.contiguous(), float(p), int(seed) & 0x7FFFFFFFFFFFFFFF)

    # ---------------------------------------------------------------- RMSNorm
    def rmsnorm_fwd(self, x, eps):
        self.launches += 1
        y, rstd = self.ext.rmsnorm_fwd(x, float(eps))
        return y, rstd

    def rmsnorm_bwd(self, dy, x, rstd, dres=None):
        self.launches += 1
        return self.ext.rmsnorm_bwd(dy, x, rstd, dres)

    # ---------------------------------------------------------------- FP8 helpers
    def notify_weights_updated(self) -> None:
        """Called by the engine after every optimizer step: cached e4m3 weights are stale."""
        self._wver += 1

    def _quant_weight(self, w, mx: bool):
        key = (w.data_ptr(), tuple(w.shape), mx)
        hit = self._wq.get(key)
        if hit is None or hit[0] != self._wver:
            q, s = (self.ext.quant_mx_e4m3 if mx else self.ext.quant_rows_e4m3)(w.contiguous())
            self.launches += 1
            hit = (self._wver, q, s)
            self._wq[key] = hit
        return hit[1], hit[2]

    def _linear_fwd_fp8(self, x, w, y, epi, d2=None, aux=None):
        mx = self.fp8_format == "mx" and w.shape[1] % 128 == 0
        if mx:      # OCP MX: UE8M0 scale per 32 elements, applied inside tcgen05.mma.kind::mxf8f6f4.block_scale
            xq, sx = self.ext.quant_mx_e4m3(x)
            wq, sw = self._quant_weight(w, True)
            self.ext.gemm_mx(xq, sx, wq, sw, y, epi, d2, aux)
        else:       # one fp32 scale per row, applied in the epilogue of the kind::f8f6f4 GEMM
            xq, sx = self.ext.quant_rows_e4m3(x)
            wq, sw = self._quant_weight(w, False)
            self.ext.gemm_fp8(xq, sx, wq, sw, y, epi, d2, aux)
        self.launches += 2

    # ---------------------------------------------------------------- linear
    def linear_fwd(self, x, w, res=None, act=None, gather=None):
        M, N = x.shape[0], w.shape[0]
        y = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
        if gather is not None and gather.done:
            gather = None
        if self.fp8 and gather is None and x.is_contiguous() and w.shape[1] % 16 == 0 and N % 64 == 0:
            if act == "gelu":
                a = torch.empty_like(y)
                self._linear_fwd_fp8(x, w, y, EPI_GELU, d2=a)
                return y, a
            self._linear_fwd_fp8(x, w, y, EPI_RESIDUAL if res is not None else EPI_STORE, aux=res)
            return y
        if act == "gelu":
            a = torch.empty_like(y)
            self._gemm(x, False, w, False, y, EPI_GELU, d2=a, gather=gather)
            return y, a
        if res is not None:
            self._gemm(x, False, w, False, y, EPI_RESIDUAL, aux=res, gather=gather)
        else:
            self._gemm(x, False, w, False, y, EPI_STORE, gather=gather)
        return y

    def linear_dgrad(self, dy, w, gelu_pre=None, dres=None, gather=None):
        M, K = dy.shape[0], w.shape[1]
        dx = torch.empty((M, K), dtype=torch.bfloat16, device=dy.device)
        if gelu_pre is not None:
            self._gemm(dy, False, w, True, dx, EPI_DGELU, aux=gelu_pre, gather=gather)
        elif dres is not None:
            self._gemm(dy, False, w, True, dx, EPI_RESIDUAL, aux=dres, gather=gather)
        else:
            self._gemm(dy, False, w, True, dx, EPI_STORE, gather=gather)
        return dx

    def linear_wgrad(self, dy, x, dw_acc, gather=None):
        self._gemm(dy, True, x, True, dw_acc, EPI_F32, accumulate=True, gather=gather)

    # ---------------------------------------------------------------- QK-LN + RoPE
    def qkln_rope_fwd(self, qkv, q_gain, k_gain, B, T, H, C, eps=1e-6):
        sin, cos = self._rope_tables(C, T)
        self.launches += 1
        qk, stats = self.ext.qkln_rope_fwd(qkv, q_gain.float(), k_gain.float(), sin, cos, T, H, C, float(eps))
        return QKV(qk, qkv), stats

    def qkln_rope_bwd(self, dqkv, qkv_raw, stats, q_gain, k_gain, dq_gain_acc, dk_gain_acc, B, T, H, C):
        sin, cos = self._rope_tables(C, T)
        self.launches += 1
        # dq may still sit in the attention backward's fp32 accumulator (see attention_bwd): read it from there and clear it
        acc = getattr(dqkv, "_mg_dq_acc", None)
        self.ext.qkln_rope_bwd(dqkv, qkv_raw, stats, q_gain.float(), k_gain.float(), sin, cos, dq_gain_acc, dk_gain_acc, T, H, C, acc)
        if acc is not None:
            dqkv._mg_dq_acc = None
        return dqkv

    def _dq_accumulator(self, M: int, D: int) -> torch.Tensor:
        """A persistent, all-zero fp32 [M, D] buffer for the attention backward's dQ partial sums.  Invariant: zero between uses
        (qkln_rope_bwd clears what it reads), so no per-layer zero-fill kernel is launched."""
        key = (M, D)
        buf = self._dq_acc.get(key)
        if buf is None:
            buf = torch.zeros((M, D), dtype=torch.float32, device=self.device)
            self._dq_acc[key] = buf
        return buf

    # ---------------------------------------------------------------- attention
    @staticmethod
    def _views(h: QKV, B, T, H, C):
        qk = h.qk.view(B, T, 2, H, C)
        return qk[:, :, 0], qk[:, :, 1], h.raw.view(B, T, 3, H, C)[:, :, 2]

    def attention_fwd(self, h: QKV, B, T, H, C, dropout_p=0.0, seed=0):
        """Causal flash attention; ``dropout_p`` > 0 drops attention probabilities *inside* the kernel (reference
        src/model.py:78): a Philox keep-mask keyed by ``seed`` is applied to P in registers after the row sum."""
        q, k, v = self._views(h, B, T, H, C)
        self.launches += 1
        o, lse = self.ext.attn_fwd(q, k, v, float(dropout_p), int(seed) & 0x7FFFFFFFFFFFFFFF)
        return o, lse

    def attention_bwd(self, do, h: QKV, o, lse, B, T, H, C, dropout_p=0.0, seed=0, fuse_dq: bool = False):
        """Returns dqkv [M, 3D] (dq | dk | dv).  With ``fuse_dq`` the dq third is NOT materialised: dq stays in the fp32
        accumulator the kernel reduce-adds into, tagged onto the result for ``qkln_rope_bwd`` (its only consumer in the model),
        which saves the zero-fill and the fp32->bf16 conversion pass per layer."""
        q, k, v = self._views(h, B, T, H, C)
        # measured (bench/kernel_bench.py --dq-fuse, backward + QK-LN/RoPE backward): head_dim 64: 0.419 -> 0.398 ms fused;
        # head_dim 128: 0.809 -> 0.834 ms (the register-bound qkln kernel pays more for the fp32 read than the two small
        # streaming kernels cost), so the fused hand-over is used where it wins
        if not fuse_dq or C != 64:
            self.launches += 4          # prep, zero-fill, kernel, convert
            return self.ext.attn_bwd(q, k, v, o, do.contiguous(), lse, float(dropout_p), int(seed) & 0x7FFFFFFFFFFFFFFF, None)
        acc = self._dq_accumulator(B * T, H * C)
        self.launches += 2
        dqkv = self.ext.attn_bwd(q, k, v, o, do.contiguous(), lse, float(dropout_p), int(seed) & 0x7FFFFFFFFFFFFFFF, acc)
        dqkv._mg_dq_acc = acc
        return dqkv

    # ---------------------------------------------------------------- LM head + cross-entropy
    def _logits_buf(self, rows: int, vpad: int) -> torch.Tensor:
        need = rows * vpad
        if self._logits_ws is None or self._logits_ws.numel() < need:
            self._logits_ws = torch.empty(need, dtype=torch.bfloat16, device=self.device)
        return self._logits_ws[:need].view(rows, vpad)

    def lmhead_logits(self, h, w, gather=None):
        V = w.shape[0]
        vpad = (V + 7) // 8 * 8
        buf = torch.empty((h.shape[0], vpad), dtype=torch.bfloat16, device=h.device)
        self._gemm(h, False, w, False, buf[:, :V], EPI_STORE, gather=gather)
        return buf[:, :V]

    def lmhead_ce(self, h, w, targets, dw_acc=None, grad_scale: float = 0.0, chunk: int = 32768, gather=None, prefetch=None):
        """Chunk = 32768 tokens: one bf16 logits workspace of 32768 x V (3.3 GB for V = 50304 -- nothing on a 180 GB part).  With
        8192-token chunks (round 1) the dgrad GEMM of a chunk had 96 output tiles for 74 CTA pairs (1.3 waves -> 65 % of the
        machine) and the weight gradient was reduce-added into the fp32 accumulator four times as often."""
        M, V = h.shape[0], w.shape[0]
        vpad = (V + 7) // 8 * 8
        tg = targets.reshape(-1).to(torch.int32).contiguous()
        loss_sum = torch.zeros(1, dtype=torch.float32, device=h.device)
        dh = torch.empty_like(h) if dw_acc is not None else None
        for s in range(0, M, chunk):
            e = min(M, s + chunk)
            buf = self._logits_buf(e - s, vpad)
            logits = buf[:, :V]
            hc = h[s:e]
            carry = gather if (gather is not None and not gather.done) else prefetch     # own gather first, then the next unit's
            self._gemm(hc, False, w, False, logits, EPI_STORE, gather=carry)
            self.ext.ce_chunk(buf, tg[s:e], loss_sum, V, float(grad_scale), dw_acc is not None)
            self.launches += 1
            if dw_acc is not None:
                self._gemm(logits, False, w, True, dh[s:e], EPI_STORE)            # dh = dlogits · W
                self._gemm(logits, True, hc, True, dw_acc, EPI_F32, accumulate=True)   # dW += dlogitsᵀ · h
        return loss_sum[0], dh

    # ---------------------------------------------------------------- optimizer
    def sumsq(self, g):
        out = torch.zeros(1, dtype=torch.float32, device=g.device)
        self.ext.sumsq_accumulate(g, out)
        self.launches += 1
        return out[0]

    def adamw_step(self, p, m, v, g, p_lp, *, lr, b1, b2, eps, wd_over_lr_peak, bc1, bc2, grad_scale):
        if not torch.is_tensor(grad_scale):
            grad_scale = torch.full((1,), float(grad_scale), dtype=torch.float32, device=p.device)
        if p.dtype != torch.float32:
            return self._torch.adamw_step(p, m, v, g, p_lp, lr=lr, b1=b1, b2=b2, eps=eps, wd_over_lr_peak=wd_over_lr_peak,
                                          bc1=bc1, bc2=bc2, grad_scale=grad_scale)
        self.launches += 1
        self.ext.adamw_step(p, m, v, g, p_lp, grad_scale.reshape(1).float(), lr, b1, b2, eps, wd_over_lr_peak, bc1, bc2, False)
