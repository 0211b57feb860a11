"""Evaluation metrics: pdebench/models/metrics.py:164-306 `metric_func`.

* CPU (`-m "not gpu"`): oracle/metrics_port.py against outputs of the unmodified reference function
  (tests/golden/metrics_small.npz, oracle/make_golden_metrics.py) -- if_mean True and False, 2-D and 3-D.
* GPU: fno_metric_func / fno_window_shift through the C ABI against the port on the same fields and on seeded fields at
  the bench sizes; fno_b200.evaluate.metrics (rollout loop on the device) against the port's loop.
Tolerance: 1e-5 relative on the point-wise metrics; 2e-5 on the Fourier bands (fp32 DFT of the error field).
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import metrics_port as MP

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "metrics_small.npz")
CASES = ("2d", "2d_default", "3d")


def _kw(case):
    lx, ly, lz, lo, hi = GOLD[f"{case}_kw"]
    return dict(Lx=float(lx), Ly=float(ly), Lz=float(lz), iLow=int(lo), iHigh=int(hi))


def _close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    assert np.all(np.abs(a[m] - b[m]) <= tol * np.abs(b[m]) + 1e-9), (a, b)


@pytest.mark.parametrize("case", CASES)
def test_port_matches_reference_metric_func(case):
    pred, target = torch.from_numpy(GOLD[f"{case}_pred"]), torch.from_numpy(GOLD[f"{case}_target"])
    got = MP.metric_func(pred, target, if_mean=True, **_kw(case))
    # the reference computes in fp32 (its Fourier part as the difference of two fp32 FFTs): 2e-5
    _close(np.concatenate([np.atleast_1d(v.numpy()).reshape(-1) for v in got]), GOLD[f"{case}_mean"], 2e-5)
    arrs = MP.metric_func(pred, target, if_mean=False, **_kw(case))
    for name, v in zip(("rmse", "nrmse", "csv", "max", "bd", "f"), arrs):
        _close(v.numpy(), GOLD[f"{case}_{name}"], 5e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_metric_kernel_matches_golden_and_port(case):
    from fno_b200 import lib

    dev = torch.device("cuda", 0)
    pred, target = torch.from_numpy(GOLD[f"{case}_pred"]).to(dev), torch.from_numpy(GOLD[f"{case}_target"]).to(dev)
    out = lib.metric_func(pred, target, **_kw(case)).cpu().numpy()
    _close(out[:8], GOLD[f"{case}_mean"], 2e-5)
    port = MP.metric_func(pred.cpu(), target.cpu(), if_mean=True, **_kw(case))
    _close(out[:8], np.concatenate([np.atleast_1d(v.numpy()).reshape(-1) for v in port]), 1e-5)
    nd = pred.dim() - 3
    l2t = torch.sqrt(((pred - target).double() ** 2).mean(dim=tuple(range(1 + nd)) + (-1,))).cpu().numpy()
    _close(out[8:], l2t, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(8, 128, 128, 1, 2), (4, 130, 66, 2, 3), (2, 64, 64, 64, 1, 5), (2, 17, 9, 11, 3, 9)],
                         ids=["cfg1", "odd2d", "cfg4", "odd3d-tv27"])
def test_metric_kernel_full_size(shape):
    from fno_b200 import lib

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(7)
    target = torch.randn(shape, generator=g)
    pred = target + 0.05 * torch.randn(shape, generator=g) + 0.01
    out = lib.metric_func(pred.to(dev), target.to(dev)).cpu().numpy()
    port = MP.metric_func(pred, target, if_mean=True)
    _close(out[:8], np.concatenate([np.atleast_1d(v.numpy()).reshape(-1) for v in port]), 2e-5)


@pytest.mark.gpu
def test_window_shift():
    from fno_b200 import lib

    dev = torch.device("cuda", 0)
    xx = torch.randn(3, 9, 7, 10, 2, device=dev)
    pred = torch.randn(3, 9, 7, 1, 2, device=dev)
    assert torch.equal(lib.window_shift(xx, pred), torch.cat((xx[..., 1:, :], pred), dim=-2))
    with pytest.raises(lib.FnoError):
        lib.window_shift(xx, pred, out=xx)


@pytest.mark.gpu
def test_rollout_metrics_loop_matches_port():
    """fno_b200.evaluate.metrics (metrics.py:337-399 on the device) against the port's loop driving the same drop-in model."""
    from fno_b200 import evaluate
    from fno_b200.fno import FNO2d

    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    model = FNO2d(num_channels=2, modes1=6, modes2=6, width=12, initial_step=4).to(dev)
    g = torch.Generator().manual_seed(11)
    batches = []
    for _ in range(3):
        xx = torch.randn(2, 32, 32, 4, 2, generator=g).to(dev)
        yy = torch.randn(2, 32, 32, 7, 2, generator=g).to(dev)
        ax = torch.linspace(0, 1, 32)
        grid = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), -1).expand(2, 32, 32, 2).contiguous().to(dev)
        batches.append((xx, yy, grid))
    res = evaluate.metrics(batches, model, rollout_test=3, Lx=1.0, Ly=1.0, Lz=1.0, initial_step=4)
    acc, n, l2t = MP.rollout_metrics(model, batches, 3, 4)
    assert n == 3 and res["batches"] == 3
    got = np.array([res["sum"][k] for k in ("RMSE", "nRMSE", "CSV", "Max", "BD")] + list(res["sum"]["F"]))
    _close(got, acc.cpu().numpy(), 2e-5)
    _close(res["sum_l2_time"], l2t.cpu().numpy(), 2e-5)
    # the reference divides by itot = index of the last batch (metrics.py:348, :397-402)
    assert abs(res["reference"]["RMSE"] - res["sum"]["RMSE"] / 2) < 1e-12
