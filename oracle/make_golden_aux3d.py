"""Golden vectors of the UNMODIFIED reference two-head FNO3d (pdebench/models/fno_aux/fno_aux.py:325-475): outputs
of both heads, selected gradients, every parameter, and the complete state_dict key / shape / dtype listing (100
keys: trunk, two heads, dead BatchNorm3d modules, aliased ``shared_layers.N.*``).

Build container only (the reference tree is absent on the GPU box):   python oracle/make_golden_aux3d.py
TEST INFRASTRUCTURE ONLY.  Writes tests/golden/aux3d_small.npz and tests/golden/aux3d_meta.json.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.make_golden import _import_reference, _model_case, fingerprint, seeded  # noqa: E402

OUT = ROOT / "tests" / "golden"
CTOR = dict(num_channels=3, modes1=3, modes2=3, modes3=3, width=6, initial_step=2)


def main():
    torch.set_num_threads(8)
    _, ref_aux = _import_reference()
    torch.manual_seed(16)
    m = ref_aux.FNO3d(**CTOR)
    meta = {"seed": 16, "ctor": CTOR, "rng_after_init": float(torch.rand(1)),
            "state_dict": {k: fingerprint(v) for k, v in m.state_dict().items()},
            "named_parameters": [k for k, _ in m.named_parameters()]}
    out = {}
    x = seeded((1, 8, 8, 8, 2, 3), 730)
    grid = torch.rand(1, 8, 8, 8, 3, generator=torch.Generator().manual_seed(731))
    xa = seeded((2, 8, 8, 8, 2, 3), 732)
    ga = torch.rand(2, 8, 8, 8, 3, generator=torch.Generator().manual_seed(733))
    out["aux3d_x"], out["aux3d_grid"], out["aux3d_xa"], out["aux3d_ga"] = x.numpy(), grid.numpy(), xa.numpy(), ga.numpy()
    _model_case(m, (x, grid, xa, ga), "aux3d", out,
                ["fc0.weight", "conv0.weights1", "conv1.weights4", "conv3.weights3", "w0.weight", "w3.bias",
                 "fc1.weight", "fc2_primary.weight", "fc2_auxiliary.bias"])
    meta["params_without_grad"] = [k for k, p in m.named_parameters() if p.grad is None]
    np.savez_compressed(OUT / "aux3d_small.npz", **out)
    (OUT / "aux3d_meta.json").write_text(json.dumps(meta, indent=1))
    print("aux3d_small.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw;", len(meta["state_dict"]), "state_dict keys")


if __name__ == "__main__":
    main()
