"""Drop-in spectral-convolution modules (reference: pdebench/models/fno/fno.py:35-92, :191-288).

Same constructor arguments, attributes, parameter names / shapes / dtypes (complex64
``weights1..2`` / ``weights1..4`` in ``[Ci, Co, m1, m2(, m3)]`` layout) and the same consumption
of the global torch RNG as the reference, so ``state_dict`` round-trips in both directions and a
model built under ``torch.manual_seed(s)`` starts from identical parameters.  ``forward`` runs the
hand-written sm_100a kernels (fno_b200.ops); there is no torch.fft / einsum path.
"""
from __future__ import annotations

import torch
from torch import nn

from . import lib, ops


class _SpectralConvBase(nn.Module):
    _ncorners = 2

    def __init__(self, in_channels: int, out_channels: int, *modes: int):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        for k, m in enumerate(modes, start=1):
            setattr(self, f"modes{k}", m)
        self.scale = 1 / (in_channels * out_channels)
        for k in range(1, self._ncorners + 1):
            # same draw as the reference: uniform [0, 1) real and imaginary parts, times scale
            init = self.scale * torch.rand(in_channels, out_channels, *modes, dtype=torch.cfloat)
            setattr(self, f"weights{k}", nn.Parameter(init))

    def _weights(self):
        return [getattr(self, f"weights{k}") for k in range(1, self._ncorners + 1)]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise lib.FnoError(
                f"{type(self).__name__} runs on CUDA sm_100a only (no CPU fallback); input is on {x.device}")
        return ops.spectral_conv(x, self._weights())


class SpectralConv1d(_SpectralConvBase):
    """1-D Fourier layer: pruned rfft -> per-mode channel mixing -> zero-padded irfft.  Named by the task's north_star;
    the reference tree has no 1-D layer, so constructor, parameter name / layout (`weights1` complex64 `[Ci, Co, modes1]`,
    `scale * rand`) and forward follow `SpectralConv2d_fast` (fno/fno.py:35-92) one dimension down."""
    _ncorners = 1

    def __init__(self, in_channels, out_channels, modes1):
        super().__init__(in_channels, out_channels, modes1)

    def compl_mul1d(self, input, weights):
        """(batch, in, x), (in, out, x) -> (batch, out, x)"""
        if not input.is_cuda:
            raise lib.FnoError("compl_mul1d runs on CUDA sm_100a only")
        return _Mix1dFn.apply(input, weights)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise lib.FnoError(f"SpectralConv1d runs on CUDA sm_100a only (no CPU fallback); input is on {x.device}")
        return ops.spectral_conv1d(x, self.weights1)


class _Mix1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, w):
        X, w = X.contiguous(), w.contiguous()
        ctx.save_for_backward(X, w)
        return lib.mix1d_fwd(X, w)

    @staticmethod
    def backward(ctx, g):
        X, w = ctx.saved_tensors
        return lib.mix1d_bwd(X, g.contiguous(), w, need_gx=ctx.needs_input_grad[0], need_gw=ctx.needs_input_grad[1])


class SpectralConv2d_fast(_SpectralConvBase):
    """2-D Fourier layer: pruned rfft2 -> per-mode channel mixing -> zero-padded irfft2."""
    _ncorners = 2

    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__(in_channels, out_channels, modes1, modes2)

    def compl_mul2d(self, input, weights):
        """(batch, in, x, y), (in, out, x, y) -> (batch, out, x, y); kept for API parity
        (fno.py:66-68).  Runs the K2 mixing kernel on one corner block."""
        return _corner_mix(input, weights)


class SpectralConv3d(_SpectralConvBase):
    """3-D Fourier layer; corners (x low|high, y low|high, z low) <-> weights1..4 (fno.py:274-285)."""
    _ncorners = 4

    def __init__(self, in_channels, out_channels, modes1, modes2, modes3):
        super().__init__(in_channels, out_channels, modes1, modes2, modes3)

    def compl_mul3d(self, input, weights):
        return _corner_mix(input, weights)


class _CornerMixFn(torch.autograd.Function):
    """einsum('bi...,io...->bo...') for ONE corner block through the K2 kernels (forward: fno_mix_fwd, backward:
    fno_mix_bwd), differentiable like the reference's einsum (fno.py:66-68, :255-257).  The block is presented as
    the low corner of a plan whose other corners are multiplied by zeros."""

    @staticmethod
    def _embed(inp, w):
        modes = tuple(w.shape[2:])
        nd = len(modes)
        spatial = tuple(2 * m for m in modes[:-1]) + (2 * (modes[-1] - 1) + 2,)
        plan = lib.get_plan(inp.device, spatial, modes)
        sl = (slice(None), slice(None)) + tuple(slice(0, m) for m in modes)
        zeros = torch.zeros_like(w)
        ws = [w.contiguous()] + [zeros] * ((2 if nd == 2 else 4) - 1)
        return plan, sl, ws

    @staticmethod
    def forward(ctx, inp, w):
        plan, sl, ws = _CornerMixFn._embed(inp, w)
        X = torch.zeros(tuple(inp.shape[:2]) + plan.spec_shape, dtype=torch.complex64, device=inp.device)
        X[sl] = inp
        Y = lib.mix_fwd(plan, X, ws)
        ctx.save_for_backward(X, w)
        return Y[sl].contiguous()

    @staticmethod
    def backward(ctx, g):
        X, w = ctx.saved_tensors
        plan, sl, ws = _CornerMixFn._embed(X, w)
        gY = torch.zeros((g.shape[0], w.shape[1]) + plan.spec_shape, dtype=torch.complex64, device=g.device)
        gY[sl] = g
        gX, gws = lib.mix_bwd(plan, X, gY, ws, need_gx=ctx.needs_input_grad[0], need_gw=ctx.needs_input_grad[1])
        return (gX[sl].contiguous() if gX is not None else None), (gws[0] if gws is not None else None)


def _corner_mix(inp: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    if not inp.is_cuda:
        raise lib.FnoError("compl_mul runs on CUDA sm_100a only")
    return _CornerMixFn.apply(inp, w)
