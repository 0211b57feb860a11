"""On-device evaluation (reference: pdebench/models/metrics.py): `metric_func` (:164-306), the rollout loop of `metrics`
(:309-402, `val_type="rollout"`) and a batch-sharded variant for data-parallel validation (BASELINE.json configs[4]).

    pred = model(xx, grid); xx = cat(xx[..., 1:, :], pred)        # repeated `rollout_test` times

The loop is host-driven as in the reference, but every arithmetic op is a libfno_sm100 kernel and nothing leaves the
device until the end: the model is the drop-in FNO, the window shift is `fno_window_shift`, the six metrics (RMSE, nRMSE,
conserved variables, maximum, boundaries, Fourier bands) come from `fno_metric_func`, accumulators stay on the device, and
under data parallelism each rank evaluates a disjoint shard and the accumulators are summed with one all-reduce.
Plots (metrics.py:404-520) are out of scope.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist

from . import lib

METRIC_NAMES = ("RMSE", "nRMSE", "CSV", "Max", "BD")


@torch.no_grad()
def metric_func(pred: torch.Tensor, target: torch.Tensor, if_mean: bool = True, Lx: float = 1.0, Ly: float = 1.0,
                Lz: float = 1.0, iLow: int = 4, iHigh: int = 12, initial_step: int = 1):
    """Same signature and return order as the reference's `metric_func` (metrics.py:164-166) for 2-D / 3-D fields
    `[B, nx, ny(, nz), T, V]`: six 0-dim device tensors, the last one `[3]` (low / middle / high band)."""
    if not if_mean:
        raise lib.FnoError("metric_func: only if_mean=True runs on the device (the form the reference's loop uses)")
    out = lib.metric_func(pred.contiguous(), target.contiguous(), Lx, Ly, Lz, iLow, iHigh)
    return out[0], out[1], out[2], out[3], out[4], out[5:8]


@torch.no_grad()
def rollout(model, xx: torch.Tensor, grid: torch.Tensor, steps: int, keep: bool = True):
    """Feeds the model its own predictions `steps` times (metrics.py:341-344).  Returns the predictions stacked on the time
    axis ``[B, *spatial, steps, V]`` (``keep``) or only the last one."""
    preds = []
    xx = xx.contiguous()
    pred = None
    for _ in range(steps):
        pred = model(xx, grid)
        if keep:
            preds.append(pred)
        xx = lib.window_shift(xx, pred.contiguous())
    return torch.cat(preds, dim=-2) if keep else pred


@torch.no_grad()
def metrics(val_loader: Iterable[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], model, rollout_test: int, Lx: float = 1.0,
            Ly: float = 1.0, Lz: float = 1.0, initial_step: Optional[int] = None, val_type: str = "rollout",
            group: Optional[dist.ProcessGroup] = None) -> dict:
    """The FNO branch of the reference's `metrics(...)` (metrics.py:309-402) without its plots: per batch, roll the model
    out `rollout_test` times (``val_type="rollout"``) or predict once, score the LAST prediction against the last target
    frame with `metric_func`, and accumulate.  Returns the sums over batches (`sum`), their mean (`mean`) and the
    reference's own normalisation (`reference`: it divides by `itot`, the INDEX of the last batch -- metrics.py:348,
    :397-402 -- so a single-batch loader gives inf there).  One host synchronisation, at the end."""
    acc = None
    n = 0
    for xx, yy, grid in val_loader:
        if not xx.is_cuda:
            raise lib.FnoError("evaluate.metrics runs on CUDA tensors only")
        if val_type == "rollout":
            target = yy[..., -1:, :].contiguous()
            pred = rollout(model, xx, grid, rollout_test, keep=False)
        else:
            target = yy.contiguous()
            pred = model(xx, grid)
        out = lib.metric_func(pred.contiguous(), target, Lx, Ly, Lz).double()
        acc = out if acc is None else acc + out
        n += 1
    if acc is None:
        raise ValueError("evaluate.metrics: empty loader")
    cnt = torch.tensor([float(n)], dtype=torch.float64, device=acc.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = torch.cat((acc, cnt))
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        acc, cnt = packed[:-1], packed[-1:]
    vals = acc.cpu().tolist()
    n = int(cnt.item())

    def pack(div):
        d = {k: (vals[i] / div if div else float("inf")) for i, k in enumerate(METRIC_NAMES)}
        d["F"] = [(v / div if div else float("inf")) for v in vals[5:8]]
        return d

    return {"batches": n, "sum": pack(1.0), "mean": pack(float(n)), "reference": pack(float(n - 1)),
            "sum_l2_time": vals[8:], "val_l2_time": [v / (n - 1) if n > 1 else float("inf") for v in vals[8:]]}


@torch.no_grad()
def evaluate_rollout(model, batches: Iterable[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], rollout_test: int,
                     group: Optional[dist.ProcessGroup] = None) -> dict:
    """`batches` yields (xx [B,*sp,T0,V], yy [B,*sp,>=rollout_test,V], grid) for THIS rank's shard.  Returns the global,
    sample-weighted means of RMSE / nRMSE of the last rollout step and of the whole rollout (every frame scored, unlike the
    reference loop which scores the last one only)."""
    acc = None
    for xx, yy, grid in batches:
        preds = rollout(model, xx, grid, rollout_test)
        tgt = yy[..., :rollout_test, :].contiguous()
        last = lib.metric_func(preds[..., -1:, :].contiguous(), tgt[..., -1:, :].contiguous())
        full = lib.metric_func(preds, tgt)
        w = float(xx.shape[0] * xx.shape[-1])                       # metric_func means over (sample, variable, time)
        term = torch.stack((last[0], last[1], full[0], full[1])).double() * w
        term = torch.cat((term, torch.tensor([w], dtype=torch.float64, device=term.device)))
        acc = term if acc is None else acc + term
    if acc is None:
        raise ValueError("evaluate_rollout: empty shard")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    n = acc[4].clamp_min(1.0)
    return {"rmse_last": float(acc[0] / n), "nrmse_last": float(acc[1] / n), "rmse_rollout": float(acc[2] / n),
            "nrmse_rollout": float(acc[3] / n), "samples_x_vars": int(acc[4])}
