"""Times K2 (mix_fwd / mix_bwd data / mix_bwd weight) alone at a given shape; run once with FNO_MIX_TC=0 (FP32 kernels)
and once without (tensor-core kernels at width 33..64).  A 256 MB buffer is rewritten between launches (L2 flush).
usage: python tools/time_mix.py [B C H W m1 m2]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "sciml-pde_b200")]
from fno_b200 import lib  # noqa: E402


def main():
    B, C, H, W, m1, m2 = (int(a) for a in sys.argv[1:7]) if len(sys.argv) >= 7 else (32, 64, 258, 258, 16, 16)
    dev = torch.device("cuda", 0)
    plan = lib.get_plan(dev, (H, W), (m1, m2))
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn((B, C) + plan.spec_shape, dtype=torch.complex64, device=dev, generator=g)
    gY = torch.randn((B, C) + plan.spec_shape, dtype=torch.complex64, device=dev, generator=g)
    ws = [torch.randn((C, C, m1, m2), dtype=torch.complex64, device=dev, generator=g) for _ in range(2)]
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    legs = {
        "mix_fwd": lambda: lib.mix_fwd(plan, X, ws),
        "mix_bwd_data": lambda: lib.mix_bwd(plan, None, gY, ws, need_gx=True, need_gw=False),
        "mix_bwd_weight": lambda: lib.mix_bwd(plan, X, gY, ws, need_gx=False, need_gw=True),
    }
    print(f"shape B={B} C={C} modes=({m1},{m2}) tensor cores: {lib.mix_tc_supported(plan, C, C)}")
    for name, fn in legs.items():
        for _ in range(3):
            fn()
        tot = 0.0
        n = 20
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        print(f"  {name:16s} {1e3 * tot / n:8.1f} us")


if __name__ == "__main__":
    main()
