"""CPU restatement of the reference's evaluation metrics (pdebench/models/metrics.py:164-306 `metric_func`, 2-D and
3-D branches, and the rollout loop :337-399) in float64 torch.  TEST INFRASTRUCTURE ONLY: imported by tests/ to check
fno_b200.evaluate / fno_metric_func; pinned to the unmodified reference by tests/golden/metrics_small.npz
(oracle/make_golden_metrics.py).
"""
from __future__ import annotations

import math

import torch


def metric_func(pred, target, if_mean=True, Lx=1.0, Ly=1.0, Lz=1.0, iLow=4, iHigh=12):
    """pred, target [B, nx, ny(, nz), T, V] -> (RMSE, nRMSE, CSV, Max, BD, F[3]) (if_mean) or the per-(c, t) arrays."""
    pred, target = pred.double(), target.double()
    nd = pred.dim() - 3
    perm = (0, pred.dim() - 1) + tuple(range(1, pred.dim() - 1))          # metrics.py:183-188: channels first, time last
    pred, target = pred.permute(perm), target.permute(perm)
    nb, nc, nt = pred.shape[0], pred.shape[1], pred.shape[-1]
    sp = pred.shape[2:-1]
    npts = math.prod(sp)
    p, t = pred.reshape(nb, nc, -1, nt), target.reshape(nb, nc, -1, nt)
    err_mean = torch.sqrt(torch.mean((p - t) ** 2, dim=2))                 # :193-197
    err_RMSE = err_mean.mean(0)
    nrm = torch.sqrt(torch.mean(t ** 2, dim=2))
    err_nRMSE = (err_mean / nrm).mean(0)                                   # :199-200
    err_CSV = torch.sqrt(torch.mean((p.sum(2) - t.sum(2)) ** 2, dim=0)) / npts     # :202-220
    err_Max = (p - t).abs().amax(2).amax(0)                                # :222-228
    d2 = (pred - target) ** 2
    if nd == 2:                                                            # :234-243
        nx, ny = sp
        bx = d2[:, :, 0] + d2[:, :, -1]                                    # [nb, nc, ny, nt]
        by = d2[:, :, :, 0] + d2[:, :, :, -1]                              # [nb, nc, nx, nt]
        err_BD = torch.sqrt((bx.sum(-2) + by.sum(-2)) / (2 * nx + 2 * ny)).mean(0)
    else:                                                                  # :244-259 (sums over the channels, no batch mean)
        nx, ny, nz = sp
        bx = d2[:, :, 0] + d2[:, :, -1]
        by = d2[:, :, :, 0] + d2[:, :, :, -1]
        bz = d2[:, :, :, :, 0] + d2[:, :, :, :, -1]
        s = bx.reshape(nb, -1, nt).sum(-2) + by.reshape(nb, -1, nt).sum(-2) + bz.reshape(nb, -1, nt).sum(-2)
        err_BD = torch.sqrt(s / (2 * nx * ny + 2 * ny * nz + 2 * nz * nx))
    dims = list(range(2, 2 + nd))
    eF = torch.abs(torch.fft.fftn(pred - target, dim=dims)) ** 2           # fftn is linear: :268-269, :282-283
    half = [n // 2 for n in sp]
    nbins = min(half)
    err_F = torch.zeros(nb, nc, nbins, nt, dtype=torch.float64, device=pred.device)
    idx = torch.stack(torch.meshgrid(*[torch.arange(h, device=pred.device) for h in half], indexing="ij"), 0).double()
    it = torch.floor(torch.sqrt((idx ** 2).sum(0))).long()                 # :275, :290
    quad = eF[(slice(None), slice(None)) + tuple(slice(0, h) for h in half)]
    for b in range(nbins):
        m = it == b
        err_F[:, :, b] = quad[:, :, m].sum(2)
    L = Lx * Ly * (Lz if nd == 3 else 1.0)
    _err_F = torch.sqrt(err_F.mean(0)) / npts * L
    bands = torch.zeros(nc, 3, nt, dtype=torch.float64, device=pred.device)
    bands[:, 0] = _err_F[:, :iLow].mean(1)
    bands[:, 1] = _err_F[:, iLow:iHigh].mean(1)
    bands[:, 2] = _err_F[:, iHigh:].mean(1)
    if if_mean:
        return (err_RMSE.mean(), err_nRMSE.mean(), err_CSV.mean(), err_Max.mean(), err_BD.mean(), bands.mean(dim=(0, -1)))
    return err_RMSE, err_nRMSE, err_CSV, err_Max, err_BD, bands


def rollout_metrics(model, batches, rollout_test, initial_step, **kw):
    """metrics.py:337-399 (`val_type="rollout"`): feed the model its own predictions `rollout_test` times, score the LAST
    prediction against the last target frame, accumulate over batches.  Returns the six sums, the batch count and
    val_l2_time (the reference divides by `itot` = index of the last batch)."""
    acc = None
    n = 0
    l2t = None
    with torch.no_grad():
        for xx, yy, grid in batches:
            yy_last = yy[..., -1:, :]
            for _ in range(rollout_test):
                pred = model(xx, grid)
                xx = torch.cat((xx[..., 1:, :], pred), dim=-2)
            vals = metric_func(pred, yy_last, **kw)
            flat = torch.cat([v.reshape(-1) for v in vals])
            acc = flat if acc is None else acc + flat
            mean_dim = tuple(range(yy_last.dim() - 2)) + (-1,)
            term = torch.sqrt(torch.mean((pred.double() - yy_last.double()) ** 2, dim=mean_dim))
            l2t = term if l2t is None else l2t + term
            n += 1
    return acc, n, l2t
