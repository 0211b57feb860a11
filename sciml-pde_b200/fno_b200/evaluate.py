"""Autoregressive rollout evaluation (reference: pdebench/models/metrics.py:337-344, :348-399) for
single-GPU and batch-sharded validation (BASELINE.json configs[4]).

    pred = model(xx, grid); xx = cat(xx[..., 1:, :], pred)        # repeated `rollout_test` times

The loop is host-driven as in the reference, but nothing leaves the device: predictions are fed back
through a device-side window shift, the RMSE / nRMSE accumulators stay on the device, and under
data parallelism each rank evaluates a disjoint shard of the validation set and the scalar
accumulators are summed with one all-reduce at the end.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


@torch.no_grad()
def rollout(model, xx: torch.Tensor, grid: torch.Tensor, steps: int) -> torch.Tensor:
    """Feeds the model its own predictions `steps` times; returns them stacked on the time axis
    ``[B, *spatial, steps, V]`` (metrics.py:341-344)."""
    preds = []
    for _ in range(steps):
        pred = model(xx, grid)
        preds.append(pred)
        xx = torch.cat((xx[..., 1:, :], pred), dim=-2)
    return torch.cat(preds, dim=-2)


def _metrics(pred: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-batch sums of RMSE and nRMSE over (sample, variable) -- metric_func's first two outputs
    (metrics.py:170-190): spatial mean of the squared error per (b, t, v), sqrt, normalised by the
    target's RMS, then averaged over t."""
    nd = pred.dim() - 3
    sp = tuple(range(1, 1 + nd))
    err = torch.sqrt(((pred - target) ** 2).mean(sp))            # [B, T, V]
    nrm = torch.sqrt((target ** 2).mean(sp))
    return err.mean(1).sum(), (err / nrm).mean(1).sum()


@torch.no_grad()
def evaluate_rollout(model, batches: Iterable[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], rollout_test: int,
                     group: Optional[dist.ProcessGroup] = None) -> dict:
    """`batches` yields (xx [B,*sp,T0,V], yy [B,*sp,>=rollout_test,V], grid) for THIS rank's shard.
    Returns global means of RMSE / nRMSE of the last rollout step and of the whole rollout."""
    dev = None
    acc = None
    for xx, yy, grid in batches:
        dev = xx.device
        if acc is None:
            acc = torch.zeros(5, dtype=torch.float64, device=dev)   # rmse_last, nrmse_last, rmse_all, nrmse_all, count
        preds = rollout(model, xx, grid, rollout_test)
        tgt = yy[..., :rollout_test, :]
        r_last, n_last = _metrics(preds[..., -1:, :], tgt[..., -1:, :])
        r_all, n_all = _metrics(preds, tgt)
        acc += torch.stack((r_last, n_last, r_all, n_all,
                            torch.tensor(float(xx.shape[0] * xx.shape[-1]), device=dev))).double()
    if acc is None:
        raise ValueError("evaluate_rollout: empty shard")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    n = acc[4].clamp_min(1.0)
    return {"rmse_last": float(acc[0] / n), "nrmse_last": float(acc[1] / n), "rmse_rollout": float(acc[2] / n),
            "nrmse_rollout": float(acc[3] / n), "samples_x_vars": int(acc[4])}
