"""The reference's training step (fno/train.py:264-278, fno_aux/fno_train_aux.py:308-329) without
host synchronisation, plus its data-parallel form.

    im = model(xx, grid); loss = nrmse(im, yy).mean()
    zero_grad; backward; total_norm; clip = max(5, 0.1 * total_norm); clip_grad_norm_; Adam; sched

The reference evaluates ``max(5, 0.1 * total_norm)`` in Python (a device->host sync per step);
here it is a device-side clamp, numerically identical.  Under data parallelism the gradient
all-reduce (fno_b200.dp) completes before the norm is taken, so every rank clips and steps
identically.
"""
from __future__ import annotations

from typing import Optional

import torch

from .dp import BucketedGradAllReduce


def nrmse(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """fno/train.py:34-40: per-sample MSE over dims 1..3, normalised by the target's mean square."""
    dims = tuple(range(output.ndim))[1:4]
    num = (output - target).pow(2).mean(dims, keepdim=True)
    den = 1e-7 + target.pow(2).mean(dims, keepdim=True)
    return num / den


def global_grad_norm(params) -> torch.Tensor:
    norms = [torch.linalg.vector_norm(torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad)
             for p in params if p.grad is not None]
    return torch.linalg.vector_norm(torch.stack(norms))


class TrainStep:
    """One optimisation step of FNO2d/FNO3d (``aux=False``) or the two-head joint model."""

    def __init__(self, model, optimizer, scheduler=None, dp: Optional[BucketedGradAllReduce] = None,
                 auxiliary_weight: float = 0.0):
        self.model = model
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.dp = dp
        self.auxiliary_weight = auxiliary_weight
        self.params = [p for p in model.parameters()]

    def _backward_and_update(self, loss):
        if self.dp is not None:
            self.dp.zero_grad()
        else:
            self.optimizer.zero_grad()
        loss.backward()
        if self.dp is not None:
            self.dp.finish()
        total_norm = global_grad_norm(self.params)
        clip_value = torch.clamp(0.1 * total_norm, min=5.0)
        torch.nn.utils.clip_grad_norm_(self.params, clip_value)
        self.optimizer.step()
        if self.scheduler is not None:
            self.scheduler.step()
        return total_norm

    def __call__(self, xx, yy, grid):
        loss = nrmse(self.model(xx, grid), yy).mean()
        self._backward_and_update(loss)
        return loss.detach()

    def joint(self, xx, yy, grid, xx_aux, yy_aux, grid_aux):
        """fno_train_aux.py:308-329: loss = primary + auxiliary_weight * auxiliary."""
        out_p, out_a = self.model(xx, grid, xx_aux, grid_aux)
        lp = nrmse(out_p, yy).mean()
        la = nrmse(out_a, yy_aux).mean()
        loss = lp + self.auxiliary_weight * la
        self._backward_and_update(loss)
        return lp.detach(), la.detach()
