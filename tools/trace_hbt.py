"""Debug: per-role clock64 stamps of head_bwd_tc_kernel (library built with -DFNO_TRACE)."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "sciml-pde_b200"):
    sys.path.insert(0, str(p))
import torch
from fno_b200 import lib
L = lib.load()
geo = lib.TrunkGeo((128, 128), 2)
B, Cw, V = 128, 20, 2
g = torch.Generator().manual_seed(1)
h = torch.randn((B, Cw) + geo.padded, generator=g).cuda()
W1 = (torch.randn(128, Cw, generator=g) / 4).cuda(); b1 = torch.randn(128, generator=g).cuda()
W2 = (torch.randn(V, 128, generator=g) / 11).cuda()
stats = torch.rand(B, 2, V, generator=g).cuda() + 0.5
dout = torch.randn(B, 128 * 128, V, generator=g).cuda()
for _ in range(3):
    lib.head_bwd(geo, h, dout, W1, b1, W2, stats)
torch.cuda.synchronize()
buf = (C.c_longlong * 512)()
L.fno_debug_trace.restype = C.c_int
L.fno_debug_trace.argtypes = [C.c_void_p]
assert L.fno_debug_trace(buf) == 0
t = [[buf[i * 32 + s] for s in range(32)] for i in range(16)]
t0 = min(v for row in t for v in row if v > 0)
names = {0: "mma:a_go", 1: "mma:bc_go", 2: "mma:bc_issued", 8: "epi:start", 9: "epi:computed", 10: "epi:bc_done(it-1)", 11: "epi:stored",
         16: "ld:step", 17: "ld:bh_free", 18: "ld:first_half", 19: "ld:bc_done(it-1)", 20: "ld:b3_staged", 21: "ld:step_end", 22: "dr:start", 23: "dr:waited", 24: "dr:ld_done", 25: "dr:stored", 26: "lr:start", 27: "lr:end"}
for i in range(16):
    ev = sorted((t[i][s] - t0, names[s]) for s in names if t[i][s] > 0)
    print(f"tile {i + 8}: " + "  ".join(f"{n}@{v}" for v, n in ev))
