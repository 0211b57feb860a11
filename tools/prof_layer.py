#!/usr/bin/env python
"""One Fourier layer (forward + backward) of the bench workload through the C ABI: the ncu target.

    python tools/prof_layer.py [--batch 128] [--iters 3] [--width 20] [--res 128] [--modes 12]

Prints the number of libfno_sm100 launches per iteration so that `ncu -s/-c` can be set to skip
the warm-up iterations.  Used for profiles/*.csv (see profiles/README.md)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "sciml-pde_b200"):
    sys.path.insert(0, str(p))

import torch  # noqa: E402

from fno_b200 import lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--width", type=int, default=20)
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--modes", type=int, default=12)
ap.add_argument("--pad", type=int, default=2)
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
C, m, n = args.width, args.modes, args.res + args.pad
a = torch.randn(args.batch, C, n, n, device=dev, requires_grad=True)
wl = (torch.randn(C, C, 1, 1, device=dev) / C).requires_grad_()
bl = torch.randn(C, device=dev).requires_grad_()
ws = [(torch.rand(C, C, m, m, dtype=torch.cfloat, device=dev) / (C * C)).requires_grad_() for _ in range(2)]
g = torch.randn(args.batch, C, n, n, device=dev)
for it in range(args.iters):
    l0 = lib.launch_count()
    out = ops.fourier_layer(a, wl, bl, True, ws)
    out.backward(g)
    torch.cuda.synchronize()
    print(f"iter {it}: {lib.launch_count() - l0} libfno_sm100 launches", flush=True)
