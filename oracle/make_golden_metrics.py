"""Calls the UNMODIFIED reference `metric_func` (/root/reference/pdebench/models/metrics.py:164-306) on small seeded 2-D and
3-D fields and stores inputs + outputs in tests/golden/metrics_small.npz.

Build container only:   python oracle/make_golden_metrics.py
TEST INFRASTRUCTURE ONLY.  matplotlib / mpl_toolkits are stubbed (the module imports them for its plots).
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/pdebench/models")


def fields(shape, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    target = torch.randn(shape, generator=g)
    # smooth-ish error with every wavenumber present, plus an offset so the conserved-variable metric is not ~0
    pred = target + 0.1 * torch.randn(shape, generator=g) + 0.02
    return pred, target


def main():
    if not REF.exists():
        raise SystemExit(f"{REF} not found: run in the build container")
    import torch

    for name in ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda *a, **k: None
    sys.path.insert(0, str(REF))
    import metrics as ref

    out = {}
    cases = {"2d": ((3, 16, 12, 3, 2), dict(Lx=1.0, Ly=2.0, iLow=2, iHigh=4)),
             "2d_default": ((2, 32, 32, 1, 3), dict()),
             "3d": ((2, 8, 6, 10, 2, 3), dict(Lx=1.0, Ly=0.5, Lz=2.0, iLow=1, iHigh=2))}
    for k, (shape, kw) in cases.items():
        pred, target = fields(shape, seed=len(shape) * 100 + shape[1])
        m = ref.metric_func(pred, target, if_mean=True, **kw)
        a = ref.metric_func(pred, target, if_mean=False, **kw)
        out[f"{k}_pred"] = pred.numpy()
        out[f"{k}_target"] = target.numpy()
        out[f"{k}_mean"] = np.concatenate([np.atleast_1d(v.cpu().numpy()).reshape(-1) for v in m]).astype(np.float64)
        for name, v in zip(("rmse", "nrmse", "csv", "max", "bd", "f"), a):
            out[f"{k}_{name}"] = v.cpu().numpy()
        out[f"{k}_kw"] = np.array([kw.get("Lx", 1.0), kw.get("Ly", 1.0), kw.get("Lz", 1.0), kw.get("iLow", 4), kw.get("iHigh", 12)])
    np.savez_compressed(ROOT / "tests" / "golden" / "metrics_small.npz", **out)
    print("wrote tests/golden/metrics_small.npz", {k: v.shape for k, v in out.items() if k.endswith("_mean")})


if __name__ == "__main__":
    main()
