#!/usr/bin/env python
"""lift_stats at the cfg-1 shape: parity vs torch.std_mean (fp64) and event-timed launches with L2 flushed."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "sciml-pde_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402

from fno_b200 import lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for V in (1, 2, 3, 4):
    x = torch.randn(128 if V <= 2 else 32, 128, 128, 10, V, device=dev) * 0.5 + 10.0
    st = lib.lift_stats(x).double()
    std, mean = torch.std_mean(x.double(), dim=(1, 2, 3))
    e1 = float((st[:, 0] - mean).abs().max() / mean.abs().max())
    e2 = float(((st[:, 1] - (std + 1e-7)) / std).abs().max())
    ts = []
    for _ in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lib.lift_stats(x); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"V={V} B={x.shape[0]}: mean err {e1:.1e} std err {e2:.1e}; {ts[len(ts)//2]:.1f} us, {x.numel()*4/ts[len(ts)//2]*1e-3:.0f} GB/s", flush=True)
