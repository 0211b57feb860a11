"""Drop-in FNO2d / FNO3d (reference: pdebench/models/fno/fno.py:95-188, :291-390).

Constructor kwargs, submodule names, parameter shapes/dtypes, RNG consumption order and
``forward(x, grid)`` semantics follow the reference exactly (state_dict keys: 22 for FNO2d, 50 for
FNO3d including the never-called ``bn0..3``).  The four Fourier layers -- spectral convolution,
1x1-conv bypass, add, exact GELU -- run as fused sm_100a kernel sequences (fno_b200.ops), and so
do the lift (statistics, normalisation, grid concat, fc0, permute, pad: one kernel writing the
trunk layout) and the projection head (unpad, fc1, GELU, fc2, de-normalisation: one kernel, the
128-wide hidden layer never reaches memory) -- SURVEY.md 8f rows f1/f2.
"""
from __future__ import annotations

import torch
from torch import nn

from . import lib, ops
from .spectral import SpectralConv2d_fast, SpectralConv3d


def _trunk(model, x: torch.Tensor) -> torch.Tensor:
    for layer in range(4):
        conv = getattr(model, f"conv{layer}")
        w = getattr(model, f"w{layer}")
        x = ops.fourier_layer(x, w.weight, w.bias, layer < 3, conv._weights())
    return x


def _lift(model, x, grid):
    """std_mean -> normalise -> [x_tv, grid] -> fc0 -> channel-first -> zero pad, one kernel writing
    the trunk layout (fno.py:140-159 / :343-360).  Returns (h, stats [B, 2, V], geometry)."""
    return ops.lift(x, grid, model.fc0.weight, model.fc0.bias, model.padding)


def _project(model, h, fc2, stats, geo):
    """unpad -> fc1 -> GELU -> fc2 -> de-normalise -> unsqueeze(-2) (fno.py:180-188 / :381-390)."""
    out = ops.head(h, model.fc1.weight, model.fc1.bias, fc2.weight, fc2.bias, stats, geo)
    return out.unsqueeze(-2)


def _check_cuda(model, x):
    if not x.is_cuda:
        raise lib.FnoError(
            f"{type(model).__name__} (fno_b200) runs on CUDA sm_100a only -- there is no CPU fallback; "
            f"input is on {x.device}")


class FNO2d(nn.Module):
    def __init__(self, num_channels, modes1=12, modes2=12, width=20, initial_step=10):
        super().__init__()
        self.modes1 = modes1
        self.modes2 = modes2
        self.width = width
        self.padding = 2
        self.fc0 = nn.Linear(initial_step * num_channels + 2, self.width)
        for layer in range(4):
            setattr(self, f"conv{layer}", SpectralConv2d_fast(self.width, self.width, modes1, modes2))
        for layer in range(4):
            setattr(self, f"w{layer}", nn.Conv2d(self.width, self.width, 1))
        self.fc1 = nn.Linear(self.width, 128)
        self._make_heads(num_channels)

    def _make_heads(self, num_channels):
        self.fc2 = nn.Linear(128, num_channels)

    def forward(self, x, grid):
        _check_cuda(self, x)
        h, stats, geo = _lift(self, x, grid)
        h = _trunk(self, h)
        return _project(self, h, self.fc2, stats, geo)


class FNO3d(nn.Module):
    def __init__(self, num_channels, modes1=8, modes2=8, modes3=8, width=20, initial_step=10):
        super().__init__()
        self.modes1 = modes1
        self.modes2 = modes2
        self.modes3 = modes3
        self.width = width
        self.padding = 6
        self.fc0 = nn.Linear(initial_step * num_channels + 3, self.width)
        for layer in range(4):
            setattr(self, f"conv{layer}", SpectralConv3d(self.width, self.width, modes1, modes2, modes3))
        for layer in range(4):
            setattr(self, f"w{layer}", nn.Conv3d(self.width, self.width, 1))
        for layer in range(4):
            # constructed but never called by the reference (fno.py:334-337): state_dict only
            setattr(self, f"bn{layer}", nn.BatchNorm3d(self.width))
        self.fc1 = nn.Linear(self.width, 128)
        self._make_heads(num_channels)

    def _make_heads(self, num_channels):
        self.fc2 = nn.Linear(128, num_channels)

    def forward(self, x, grid):
        _check_cuda(self, x)
        h, stats, geo = _lift(self, x, grid)
        h = _trunk(self, h)
        return _project(self, h, self.fc2, stats, geo)
