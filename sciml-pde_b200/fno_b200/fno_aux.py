"""Drop-in two-head FNO2d / FNO3d of the multiphysics joint-training path (reference:
pdebench/models/fno_aux/fno_aux.py:70-222, :325-475).

Differences from fno_b200.fno: ``fc2`` is replaced by ``fc2_primary`` / ``fc2_auxiliary``,
``shared_layers`` aliases the trunk modules (adding the duplicate ``shared_layers.N.*``
state_dict keys the reference has), and ``forward(x, grid, x_aux, grid_aux)`` returns
``(out_primary, out_auxiliary)``.  The reference runs the shared trunk twice, once per stream;
no op in the trunk couples samples, so here both streams go through ONE batched pass over
``cat([x, x_aux])`` -- half the kernel launches, identical results per sample.
"""
from __future__ import annotations

import torch
from torch import nn

from . import fno as _base


class FNO2d(_base.FNO2d):
    def __init__(self, num_channels, modes1=12, modes2=12, width=20, initial_step=10):
        super().__init__(num_channels, modes1, modes2, width, initial_step)
        self.shared_layers = nn.ModuleList([
            self.fc0, self.conv0, self.conv1, self.conv2, self.conv3,
            self.w0, self.w1, self.w2, self.w3, self.fc1,
        ])

    def _make_heads(self, num_channels):
        # same RNG position as the reference: right after fc1 (fno_aux.py:113-116)
        self.fc2_primary = nn.Linear(128, num_channels)
        self.fc2_auxiliary = nn.Linear(128, num_channels)

    def forward(self, x, grid, x_aux, grid_aux):
        return _two_head_forward(self, x, grid, x_aux, grid_aux)


class FNO3d(_base.FNO3d):
    def __init__(self, num_channels, modes1=8, modes2=8, modes3=8, width=20, initial_step=10):
        super().__init__(num_channels, modes1, modes2, modes3, width, initial_step)
        self.shared_layers = nn.ModuleList([
            self.fc0, self.conv0, self.conv1, self.conv2, self.conv3,
            self.w0, self.w1, self.w2, self.w3, self.fc1,
            self.bn0, self.bn1, self.bn2, self.bn3,
        ])

    def _make_heads(self, num_channels):
        self.fc2_primary = nn.Linear(128, num_channels)
        self.fc2_auxiliary = nn.Linear(128, num_channels)

    def forward(self, x, grid, x_aux, grid_aux):
        return _two_head_forward(self, x, grid, x_aux, grid_aux)


def _two_head_forward(model, x, grid, x_aux, grid_aux):
    _base._check_cuda(model, x)
    nb = x.shape[0]
    h, stats, geo = _base._lift(model, torch.cat((x, x_aux), dim=0), torch.cat((grid, grid_aux), dim=0))
    h = _base._trunk(model, h)
    out_p = _base._project(model, h[:nb], model.fc2_primary, stats[:nb].contiguous(), geo)
    out_a = _base._project(model, h[nb:], model.fc2_auxiliary, stats[nb:].contiguous(), geo)
    return out_p, out_a
