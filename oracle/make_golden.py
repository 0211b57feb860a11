"""Generates tests/golden/* from the UNMODIFIED reference modules (imported from /root/reference).

Run in the build container only (the reference tree does not exist on the GPU box):

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The reference ships no tests or golden vectors,
so these fixtures -- outputs of its own ``SpectralConv2d_fast`` / ``SpectralConv3d`` / ``FNO2d`` /
``FNO3d`` / ``fno_aux`` modules under the installed torch (2.11, CPU, fp32) -- are what pins the
oracles in ``oracle/`` and, through them, the CUDA path.  Inputs that are large are regenerated
from a seed with a CPU ``torch.Generator`` (bit-reproducible for a fixed torch build); only
sampled outputs are stored so the fixtures stay small.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/pdebench/models")
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _import_reference():
    if not REF.exists():
        raise SystemExit(f"{REF} not found: golden fixtures can only be generated in the build container")
    sys.path.insert(0, str(REF))
    import fno.fno as ref_fno  # noqa: WPS433
    import fno_aux.fno_aux as ref_aux  # noqa: WPS433
    return ref_fno, ref_aux


def seeded(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def sample_index(n, k, seed=1234):
    rng = np.random.default_rng(seed)
    return np.sort(rng.choice(n, size=min(k, n), replace=False))


SC2D_CASES = [
    # B, Ci, Co, H, W, m1, m2
    (2, 3, 4, 10, 9, 3, 4),      # odd W, Ci != Co
    (2, 4, 4, 12, 12, 6, 7),     # corners touch (2*m1 == H), Nyquist column kept (m2 == W/2+1)
    (1, 2, 3, 7, 16, 2, 5),      # odd H
    (2, 5, 5, 34, 34, 12, 12),   # reference modes on a small padded grid
]
SC3D_CASES = [
    # B, Ci, Co, D1, D2, D3, m1, m2, m3
    (2, 2, 3, 8, 6, 10, 3, 2, 4),
    (1, 3, 3, 8, 8, 8, 4, 4, 5),   # touching corners on both axes + Nyquist
    (1, 2, 2, 12, 10, 14, 4, 3, 4),
]


def gen_spectral(ref_fno):
    out = {}
    for ci, (B, Ci, Co, H, W, m1, m2) in enumerate(SC2D_CASES):
        torch.manual_seed(100 + ci)
        mod = ref_fno.SpectralConv2d_fast(Ci, Co, m1, m2)
        x = seeded((B, Ci, H, W), 200 + ci).requires_grad_(True)
        g = seeded((B, Co, H, W), 300 + ci)
        y = mod(x)
        y.backward(g)
        pre = f"sc2d_{ci}_"
        out[pre + "x"] = x.detach().numpy()
        out[pre + "g"] = g.numpy()
        out[pre + "y"] = y.detach().numpy()
        out[pre + "gx"] = x.grad.numpy()
        for k in (1, 2):
            w = getattr(mod, f"weights{k}")
            out[pre + f"w{k}"] = w.detach().numpy()
            out[pre + f"gw{k}"] = w.grad.numpy()
    for ci, (B, Ci, Co, D1, D2, D3, m1, m2, m3) in enumerate(SC3D_CASES):
        torch.manual_seed(400 + ci)
        mod = ref_fno.SpectralConv3d(Ci, Co, m1, m2, m3)
        x = seeded((B, Ci, D1, D2, D3), 500 + ci).requires_grad_(True)
        g = seeded((B, Co, D1, D2, D3), 600 + ci)
        y = mod(x)
        y.backward(g)
        pre = f"sc3d_{ci}_"
        out[pre + "x"] = x.detach().numpy()
        out[pre + "g"] = g.numpy()
        out[pre + "y"] = y.detach().numpy()
        out[pre + "gx"] = x.grad.numpy()
        for k in (1, 2, 3, 4):
            w = getattr(mod, f"weights{k}")
            out[pre + f"w{k}"] = w.detach().numpy()
            out[pre + f"gw{k}"] = w.grad.numpy()
    np.savez_compressed(OUT / "spectral_small.npz", **out)
    print("spectral_small.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


def _model_case(model, inputs, name, out, grad_keys):
    """Runs model(*inputs), backprops sum(out * g) and records outputs + selected gradients."""
    outs = model(*inputs)
    if not isinstance(outs, tuple):
        outs = (outs,)
    loss = 0.0
    for k, o in enumerate(outs):
        g = seeded(tuple(o.shape), 900 + k)
        out[f"{name}_out{k}"] = o.detach().numpy()
        out[f"{name}_g{k}"] = g.numpy()
        loss = loss + (o * g).sum()
    loss.backward()
    sd = dict(model.named_parameters())
    for key in grad_keys:
        gr = sd[key].grad
        out[f"{name}_grad_{key}"] = gr.detach().numpy()
    for k, v in model.state_dict().items():
        if k.startswith("shared_layers"):
            continue
        out[f"{name}_param_{k}"] = v.detach().numpy()


def gen_models(ref_fno, ref_aux):
    out = {}
    # small FNO2d
    torch.manual_seed(16)
    m = ref_fno.FNO2d(num_channels=2, modes1=4, modes2=4, width=8, initial_step=3)
    x = seeded((2, 14, 14, 3, 2), 700)
    grid = torch.rand(2, 14, 14, 2, generator=torch.Generator().manual_seed(701))
    out["fno2d_x"], out["fno2d_grid"] = x.numpy(), grid.numpy()
    _model_case(m, (x, grid), "fno2d", out,
                ["fc0.weight", "conv0.weights1", "conv3.weights2", "w1.weight", "w2.bias", "fc2.bias"])
    # small FNO3d
    torch.manual_seed(16)
    m = ref_fno.FNO3d(num_channels=3, modes1=3, modes2=3, modes3=3, width=6, initial_step=2)
    x = seeded((1, 8, 8, 8, 2, 3), 710)
    grid = torch.rand(1, 8, 8, 8, 3, generator=torch.Generator().manual_seed(711))
    out["fno3d_x"], out["fno3d_grid"] = x.numpy(), grid.numpy()
    _model_case(m, (x, grid), "fno3d", out,
                ["fc0.weight", "conv0.weights1", "conv1.weights4", "conv3.weights3", "w0.weight", "w3.bias", "fc1.weight"])
    # small aux FNO2d
    torch.manual_seed(16)
    m = ref_aux.FNO2d(num_channels=2, modes1=4, modes2=4, width=8, initial_step=3)
    x = seeded((2, 14, 14, 3, 2), 720)
    grid = torch.rand(2, 14, 14, 2, generator=torch.Generator().manual_seed(721))
    xa = seeded((6, 14, 14, 3, 2), 722)
    ga = torch.rand(6, 14, 14, 2, generator=torch.Generator().manual_seed(723))
    out["aux2d_x"], out["aux2d_grid"], out["aux2d_xa"], out["aux2d_ga"] = x.numpy(), grid.numpy(), xa.numpy(), ga.numpy()
    _model_case(m, (x, grid, xa, ga), "aux2d", out,
                ["fc0.weight", "conv0.weights1", "conv2.weights2", "w3.weight", "fc2_primary.weight", "fc2_auxiliary.bias"])
    np.savez_compressed(OUT / "models_small.npz", **out)
    print("models_small.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


def fingerprint(t: torch.Tensor):
    r = torch.view_as_real(t) if t.is_complex() else t
    r = r.double().flatten()
    return {
        "shape": list(t.shape), "dtype": str(t.dtype).replace("torch.", ""),
        "sum": float(r.sum()), "abs_sum": float(r.abs().sum()),
        "head": [float(v) for v in r[:4]],
    }


def gen_cfg1(ref_fno, ref_aux):
    """BASELINE.json configs[0]: FNO2d modes 12, width 20, initial_step 10, 2 channels, 128x128."""
    torch.manual_seed(16)
    m = ref_fno.FNO2d(num_channels=2, modes1=12, modes2=12, width=20, initial_step=10)
    meta = {"seed": 16, "ctor": dict(num_channels=2, modes1=12, modes2=12, width=20, initial_step=10),
            "params": {k: fingerprint(v) for k, v in m.state_dict().items()},
            "rng_after_init": float(torch.rand(1))}
    torch.manual_seed(16)
    ma = ref_aux.FNO2d(num_channels=2, modes1=12, modes2=12, width=20, initial_step=10)
    meta["aux_params"] = {k: fingerprint(v) for k, v in ma.state_dict().items()}
    torch.manual_seed(16)
    m3 = ref_fno.FNO3d(num_channels=5, modes1=12, modes2=12, modes3=12, width=20, initial_step=10)
    meta["fno3d_cfg4_params"] = {k: fingerprint(v) for k, v in m3.state_dict().items()
                                 if k.startswith(("fc0", "conv0", "w3", "bn2", "fc2"))}
    del m3

    # full-size forward/backward on B=2 seeded inputs; store sampled values only
    B = 2
    x = seeded((B, 128, 128, 10, 2), 800)
    lin = torch.linspace(-1 + 1 / 128, 1 - 1 / 128, 128)
    gx, gy = torch.meshgrid(lin, lin, indexing="ij")
    grid = torch.stack((gx, gy), dim=-1).unsqueeze(0).repeat(B, 1, 1, 1)
    yy = seeded((B, 128, 128, 1, 2), 801)
    out = m(x, grid)
    dims = (1, 2, 3)
    loss = (((out - yy) ** 2).mean(dims, keepdim=True) / (1e-7 + (yy ** 2).mean(dims, keepdim=True))).mean()
    loss.backward()
    idx = sample_index(out.numel(), 4096)
    arrays = {"out_idx": idx, "out_val": out.detach().flatten().numpy()[idx]}
    meta["cfg1_loss"] = float(loss.detach())
    meta["cfg1_out_sum"] = float(out.double().sum())
    meta["cfg1_grad_norms"] = {k: float(torch.norm(p.grad.detach(), 2)) for k, p in m.named_parameters()}
    # the same step through the fp64 port (oracle/fno_port.py): the fp32 reference's own gradient
    # norms deviate from these by up to ~4e-5 relative (its noise floor), so tight checks of the
    # CUDA path use the fp64 numbers and the fp32 ones are only a sanity band
    from oracle import fno_port as P
    p64 = P.as_leaves({k: v for k, v in m.state_dict().items()}, dtype=torch.float64)
    out64 = P.fno_forward(p64, x.double(), grid.double())
    loss64 = P.nrmse(out64, yy.double()).mean()
    loss64.backward()
    meta["cfg1_loss_fp64"] = float(loss64.detach())
    meta["cfg1_grad_norms_fp64"] = {k: float(torch.norm(p64[k].grad, 2)) for k, _ in m.named_parameters()}
    gidx = sample_index(m.conv1.weights1.numel(), 2048, seed=77)
    arrays["conv1_w1_grad_val_fp64"] = p64["conv1.weights1"].grad.flatten().numpy()[gidx]
    arrays["w2_weight_grad_fp64"] = p64["w2.weight"].grad.numpy()
    arrays["fc0_weight_grad_fp64"] = p64["fc0.weight"].grad.numpy()
    arrays["conv1_w1_grad_idx"] = gidx
    arrays["conv1_w1_grad_val"] = m.conv1.weights1.grad.flatten().numpy()[gidx]
    arrays["w2_weight_grad"] = m.w2.weight.grad.numpy()
    arrays["fc0_weight_grad"] = m.fc0.weight.grad.numpy()

    # cfg-1-size spectral operator alone (C = 20, 130x130 padded plane, modes 12)
    x = seeded((1, 20, 130, 130), 810).requires_grad_(True)
    g = seeded((1, 20, 130, 130), 811)
    y = m.conv0(x)
    m.conv0.weights1.grad = None
    m.conv0.weights2.grad = None
    y.backward(g)
    sidx = sample_index(y.numel(), 4096, seed=5)
    arrays["sc_idx"] = sidx
    arrays["sc_y"] = y.detach().flatten().numpy()[sidx]
    arrays["sc_gx"] = x.grad.flatten().numpy()[sidx]
    widx = sample_index(m.conv0.weights1.numel(), 2048, seed=6)
    arrays["sc_widx"] = widx
    arrays["sc_gw1"] = m.conv0.weights1.grad.flatten().numpy()[widx]
    arrays["sc_gw2"] = m.conv0.weights2.grad.flatten().numpy()[widx]
    np.savez_compressed(OUT / "cfg1_samples.npz", **arrays)
    (OUT / "cfg1_meta.json").write_text(json.dumps(meta, indent=1))
    print("cfg1_samples.npz / cfg1_meta.json written; loss", meta["cfg1_loss"])


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    ref_fno, ref_aux = _import_reference()
    gen_spectral(ref_fno)
    gen_models(ref_fno, ref_aux)
    gen_cfg1(ref_fno, ref_aux)


if __name__ == "__main__":
    main()
