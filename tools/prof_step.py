#!/usr/bin/env python
"""One FNO2d forward + backward (no optimizer) of the bench workload: the ncu target for the
whole kernel set (lift, 4 Fourier layers, head).

    python tools/prof_step.py [--batch 128] [--iters 3]

Prints the libfno_sm100 launch count per iteration (for `ncu -s/-c`)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "sciml-pde_b200"):
    sys.path.insert(0, str(p))

import torch  # noqa: E402

from fno_b200 import data, lib  # noqa: E402
from fno_b200.fno import FNO2d  # noqa: E402
from fno_b200.train import nrmse  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--width", type=int, default=20)
ap.add_argument("--modes", type=int, default=12)
ap.add_argument("--res", type=int, default=128)
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(16)
model = FNO2d(num_channels=2, modes1=args.modes, modes2=args.modes, width=args.width, initial_step=10).to(dev)
xx, yy, grid = (t.to(dev) for t in data.synthetic_batch(args.batch, args.res, 10, 2, seed=0))
for it in range(args.iters):
    l0 = lib.launch_count()
    loss = nrmse(model(xx, grid), yy).mean()
    model.zero_grad(set_to_none=True)
    loss.backward()
    torch.cuda.synchronize()
    print(f"iter {it}: {lib.launch_count() - l0} libfno_sm100 launches, loss {loss.item():.6f}", flush=True)
