"""Device-side step tail of the reference training loop: fused nRMSE loss,
and gradient-norm + adaptive clip + Adam + cosine LR in three launches with no host round trip.

    loss = nrmse_loss(model(xx, grid), yy)          # = nrmse(...).mean() of fno/train.py:34-40,:266
    opt = FusedClipAdam(model.parameters(), lr=1e-3, weight_decay=1e-4, t_max=T)
    opt.zero_grad(); loss.backward(); opt.step()     # fno/train.py:271-278

The arithmetic is torch.optim.Adam's (coupled L2 weight decay, bias correction, eps outside the
square root) and ``clip_grad_norm_``'s (coefficient ``clip / (norm + 1e-6)`` clamped to 1); complex
parameters are updated as interleaved real pairs exactly as torch does (``view_as_real``).  Because
nothing is evaluated on the host -- the reference computes ``max(5, 0.1 * total_norm)`` in Python --
the whole training step can be captured in one CUDA graph (fno_b200.train.GraphedTrainStep).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import numpy as np
import torch

from . import lib


class _NrmseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target):
        out, target = out.contiguous(), target.contiguous()
        lib._require(out, torch.float32, "output")
        lib._require(target, torch.float32, "target")
        if out.shape != target.shape:
            raise lib.FnoError(f"nrmse: output {tuple(out.shape)} vs target {tuple(target.shape)}")
        B, V = out.shape[0], out.shape[-1]
        P = out.numel() // (B * V)
        L = lib.load()
        work = torch.empty(L.fno_nrmse_workspace_bytes(B, V) // 4, dtype=torch.float32, device=out.device)
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        lib._check(L.fno_nrmse_fwd(out.data_ptr(), target.data_ptr(), loss.data_ptr(), work.data_ptr(), B, P, V,
                                   lib._stream()), "fno_nrmse_fwd")
        ctx.save_for_backward(out, target, work)
        ctx.dims = (B, P, V)
        return loss

    @staticmethod
    def backward(ctx, g):
        out, target, work = ctx.saved_tensors
        B, P, V = ctx.dims
        g = g.contiguous().float()
        dout = torch.empty_like(out)
        lib._check(lib.load().fno_nrmse_bwd(out.data_ptr(), target.data_ptr(), work.data_ptr(), g.data_ptr(),
                                            dout.data_ptr(), B, P, V, lib._stream()), "fno_nrmse_bwd")
        return dout, None


def nrmse_loss(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """``nrmse(output, target).mean()`` (fno/train.py:34-40, :266-267) for ``[B, *spatial, 1, V]`` tensors:
    per (sample, variable) MSE over the pixels normalised by the target's mean square, averaged."""
    if output.shape[-2] != 1 and output.dim() != 5:
        raise lib.FnoError("nrmse_loss: fused form needs one output step (or the 2-D layout [B, X, Y, T, V])")
    return _NrmseFn.apply(output, target)


class FusedClipAdam:
    """Adam(weight_decay) + ``clip_grad_norm_(params, max(clip_floor, clip_frac * total_norm))`` +
    per-iteration CosineAnnealingLR, as three sync-free launches over a chunk table.

    Gradients live in one flat fp32 buffer (``p.grad`` are views, created after the first backward
    so that parameters which never receive a gradient -- FNO3d's dead ``bn*`` -- are skipped exactly
    as the reference's norm code skips them).  With ``own_grads=False`` the existing ``p.grad``
    tensors are used in place (fno_b200.dp keeps them as views of its all-reduce buckets)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, t_max: float = 0.0, eta_min: float = 0.0, clip_floor: float = 5.0,
                 clip_frac: float = 0.1, own_grads: bool = True):
        self.params: List[torch.nn.Parameter] = []
        seen = set()
        for p in params:                      # de-duplicate (fno_aux shared_layers alias the trunk)
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("FusedClipAdam: no parameters")
        self.device = self.params[0].device
        if self.device.type != "cuda":
            raise lib.FnoError("FusedClipAdam runs on CUDA only")
        self.own_grads = own_grads
        self.hparams = torch.tensor([lr, eta_min, float(t_max), betas[0], betas[1], eps, weight_decay, clip_floor,
                                     clip_frac], dtype=torch.float32, device=self.device)
        self.state = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._built = False
        self.flat_grad: Optional[torch.Tensor] = None

    # -- torch.optim-like surface ----------------------------------------------------------------
    def zero_grad(self):
        if not self._built:
            for p in self.params:
                p.grad = None
        elif self.own_grads:
            self.flat_grad.zero_()
        else:
            for p in self.live:
                p.grad.zero_()

    def step(self):
        if not self._built:
            self._build()
        L = lib.load()
        lib._check(L.fno_clip_adam_step(self.chunks.data_ptr(), self.nchunks, self.partials.data_ptr(),
                                        self.state.data_ptr(), self.hparams.data_ptr(), lib._stream()),
                   "fno_clip_adam_step")

    @property
    def total_norm(self) -> torch.Tensor:
        return self.state[1]

    @property
    def lr(self) -> torch.Tensor:
        return self.state[3]

    # -- internals -------------------------------------------------------------------------------
    @staticmethod
    def _real(t: torch.Tensor) -> torch.Tensor:
        return torch.view_as_real(t) if t.is_complex() else t

    def _build(self):
        self.live = [p for p in self.params if p.grad is not None]
        if not self.live:
            raise lib.FnoError("FusedClipAdam.step() before any backward()")
        sizes = [self._real(p).numel() for p in self.live]
        padded = [(n + 3) & ~3 for n in sizes]          # 16-byte aligned slots (complex views need even offsets)
        total = sum(padded)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=self.device)
        if self.own_grads:
            self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        L = lib.load()
        chunk = L.fno_opt_chunk_floats()
        rec = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i4"), ("pad", "<i4")])
        assert rec.itemsize == L.fno_opt_chunk_bytes()
        rows = []
        off = 0
        for p, n, npad in zip(self.live, sizes, padded):
            if not p.is_contiguous():
                raise lib.FnoError("FusedClipAdam: parameters must be contiguous")
            if self.own_grads:
                view = self.flat_grad[off:off + n]
                gview = torch.view_as_complex(view.view(*p.shape, 2)) if p.is_complex() else view.view(p.shape)
                gview.copy_(p.grad)
                p.grad = gview
            gptr = self._real(p.grad).data_ptr()
            if not p.grad.is_contiguous():
                raise lib.FnoError("FusedClipAdam: gradients must be contiguous")
            pptr = p.data_ptr()
            for c0 in range(0, n, chunk):
                cn = min(chunk, n - c0)
                rows.append((pptr + 4 * c0, gptr + 4 * c0, self.exp_avg.data_ptr() + 4 * (off + c0),
                             self.exp_avg_sq.data_ptr() + 4 * (off + c0), cn, 0))
            off += npad
        table = np.array(rows, dtype=rec)
        self.nchunks = len(rows)
        self.chunks = torch.from_numpy(table.view(np.uint8).copy()).to(self.device)
        self.partials = torch.zeros(self.nchunks, dtype=torch.float32, device=self.device)
        self._built = True
