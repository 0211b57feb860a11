#!/usr/bin/env python
"""Event-timed launches of the Fourier-layer entry points at the bench geometry (cfg 1 by default).

    python tools/time_layer.py [--batch 128] [--width 20] [--res 128] [--modes 12] [--iters 20]

Each entry is timed alone with CUDA events on the current stream after warm-up; between timed launches a
buffer larger than L2 is written so that no operand is L2-resident.  Prints microseconds and the achieved
algorithmic GB/s (SURVEY 8d byte counts)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "sciml-pde_b200"):
    sys.path.insert(0, str(p))

import torch  # noqa: E402

from fno_b200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--width", type=int, default=20)
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--modes", type=int, default=12)
ap.add_argument("--pad", type=int, default=2)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--tf32", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda", 0)
if args.tf32:
    lib.set_math_mode("tf32")
torch.manual_seed(0)
B, C, m, n = args.batch, args.width, args.modes, args.res + args.pad
plan = lib.get_plan(dev, (n, n), (m, m))
a = torch.randn(B, C, n, n, device=dev)
g = torch.randn(B, C, n, n, device=dev)
s = torch.randn(B, C, n, n, device=dev)
wl = torch.randn(C, C, 1, 1, device=dev) / C
bl = torch.randn(C, device=dev)
ws = [torch.rand(C, C, m, m, dtype=torch.cfloat, device=dev) / (C * C) for _ in range(2)]
X = lib.fwd_transform(plan, a)
Y = lib.mix_fwd(plan, X, ws)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
act = 4.0 * B * C * n * n
spec = 8.0 * B * C * 2 * m * m


def timeit(name, fn, nbytes):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:46s} {med:8.1f} us   {nbytes / med * 1e-3:8.1f} GB/s (algorithmic)   min {ts[0]:.1f}", flush=True)


s_out = torch.empty_like(a)
ds = torch.empty_like(a)
timeit("K1 fwd_transform", lambda: lib.fwd_transform(plan, a), act + spec)
timeit("K1 fwd_transform + gelu' (stores dS)", lambda: lib.fwd_transform(plan, g, preact=s, ds_out=ds, cmode=1, scale=1.0 / (n * n)),
       3 * act + spec)
timeit("K2 mix_fwd", lambda: lib.mix_fwd(plan, X, ws), 2 * spec + 8.0 * C * C * 2 * m * m)
timeit("K2 mix_bwd", lambda: lib.mix_bwd(plan, X, Y, ws), 4 * spec + 16.0 * C * C * 2 * m * m)
timeit("bypass pointwise_fwd", lambda: lib.pointwise_fwd(a, wl, bl), 2 * act)
lin = lib.pointwise_fwd(a, wl, bl)
timeit("K3 inv_transform + addend + gelu + preact (r1)",
       lambda: lib.inv_transform(plan, Y, addend=lin, s_out=s_out, out=lin, cmode=1, apply_gelu=True), spec + 3 * act)
timeit("K3 inv_transform + addend (r1 adjoint)", lambda: lib.inv_transform(plan, Y, addend=lin, out=lin, cmode=0, scale=1.0),
       spec + 2 * act)
timeit("bypass wgrad", lambda: lib.pointwise_wgrad(g, a, wl.shape), 2 * act)
timeit("bypass bwd (wgrad + dgrad, r1)", lambda: lib.pointwise_bwd(g, a, wl), 3 * act)
if lib.layer_fused_supported(plan, C):
    timeit("K3 fused tc: K3 + bypass + gelu + preact", lambda: lib.layer_inv_fused(plan, Y, a, wl, bl, s_out=s_out, apply_gelu=True),
           spec + 3 * act)
    timeit("K3 fused tc: K3 + bypass (no gelu, no preact)", lambda: lib.layer_inv_fused(plan, Y, a, wl, bl), spec + 2 * act)
    timeit("K3 fused tc adjoint: K3(gX) + Wl^T dS", lambda: lib.layer_inv_fused(plan, Y, g, wl, None, cmode=0, scale=1.0, transpose=True),
           spec + 2 * act)
