#!/usr/bin/env python
"""Debug driver: one fused-layer case per process (argv: B C H W m1 m2 [gelu] [adjoint])."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "sciml-pde_b200"):
    sys.path.insert(0, str(p))
import numpy as np
import torch
from fno_b200 import lib
from oracle import dft_oracle as O

B, C, H, W, m1, m2 = [int(v) for v in sys.argv[1:7]]
gelu = len(sys.argv) > 7 and sys.argv[7] == "1"
adj = len(sys.argv) > 8 and sys.argv[8] == "1"
rng = np.random.default_rng(1)
Y = (30 * (rng.standard_normal((B, C, 2 * m1, m2)) + 1j * rng.standard_normal((B, C, 2 * m1, m2)))).astype(np.complex64)
a = rng.standard_normal((B, C, H, W)).astype(np.float32)
wl = (rng.standard_normal((C, C, 1, 1)) / np.sqrt(C)).astype(np.float32)
bl = rng.standard_normal(C).astype(np.float32)
plan = lib.get_plan(torch.device("cuda", 0), (H, W), (m1, m2))
d = lambda x: torch.from_numpy(x).cuda()
if adj:
    out = lib.layer_inv_fused(plan, d(Y), d(a), d(wl), None, cmode=0, scale=1.0, transpose=True)
    torch.cuda.synchronize()
    ref = O.inv_transform(Y, (H, W), cmode=0, scale=1.0) + np.einsum("oi,bohw->bihw", wl[:, :, 0, 0].astype(np.float64), a)
else:
    s_out = torch.empty(B, C, H, W, device="cuda")
    out = lib.layer_inv_fused(plan, d(Y), d(a), d(wl), d(bl), s_out=s_out, cmode=1, apply_gelu=gelu)
    torch.cuda.synchronize()
    sref = O.inv_transform(Y, (H, W), cmode=1) + O.pointwise_conv(a, wl, bl)
    print("s err", O.rel_err(s_out.cpu().numpy(), sref))
    ref = O.gelu(sref) if gelu else sref
    spec = O.inv_transform(Y, (H, W), cmode=1)
    byp = O.pointwise_conv(a, wl, bl)
    got = s_out.cpu().numpy()
    print("  vs spectral only", O.rel_err(got, spec), " vs bypass only", O.rel_err(got, byp))
    e = np.abs(got - sref)
    idx = np.unravel_index(np.argmax(e), e.shape)
    print("  worst at", idx, got[idx], sref[idx], " per-w max err:", np.round(e.max(axis=(0, 1, 2))[:8], 4), "...", np.round(e.max(axis=(0, 1, 2))[-4:], 4))
print("out err", O.rel_err(out.cpu().numpy(), ref))
