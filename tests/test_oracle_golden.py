"""Pins both CPU oracles against the golden vectors produced by the unmodified reference modules
(oracle/make_golden.py).  Tolerance: the reference runs in fp32 (its own noise floor is ~3e-7
relative); the numpy oracle is fp64, so agreement must be at the 2e-6 level."""
import numpy as np
import pytest
import torch

from oracle import dft_oracle as O
from oracle import fno_port as P
from oracle.make_golden import SC2D_CASES, SC3D_CASES

TOL = 2e-6


@pytest.mark.parametrize("ci", range(len(SC2D_CASES)))
def test_dense_dft_oracle_matches_reference_2d(golden_spectral, ci):
    g = golden_spectral
    pre = f"sc2d_{ci}_"
    ws = [g[pre + "w1"], g[pre + "w2"]]
    y = O.spectral_conv_forward(g[pre + "x"], ws)
    gx, gws = O.spectral_conv_backward(g[pre + "x"], ws, g[pre + "g"])
    assert O.rel_err(y, g[pre + "y"]) < TOL
    assert O.rel_err(gx, g[pre + "gx"]) < TOL
    for k in range(2):
        assert O.rel_err(gws[k], g[pre + f"gw{k + 1}"]) < TOL


@pytest.mark.parametrize("ci", range(len(SC3D_CASES)))
def test_dense_dft_oracle_matches_reference_3d(golden_spectral, ci):
    g = golden_spectral
    pre = f"sc3d_{ci}_"
    ws = [g[pre + f"w{k}"] for k in range(1, 5)]
    y = O.spectral_conv_forward(g[pre + "x"], ws)
    gx, gws = O.spectral_conv_backward(g[pre + "x"], ws, g[pre + "g"])
    assert O.rel_err(y, g[pre + "y"]) < TOL
    assert O.rel_err(gx, g[pre + "gx"]) < TOL
    for k in range(4):
        assert O.rel_err(gws[k], g[pre + f"gw{k + 1}"]) < TOL


def test_transform_adjoint_identity():
    """<K1 x, Y> == <x, K1^H Y>: the identity that makes K3(c=1, scale=1) the backward of K1."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 11, 14))
    Y = rng.standard_normal((2, 3, 8, 5)) + 1j * rng.standard_normal((2, 3, 8, 5))
    lhs = np.sum((O.fwd_transform(x, (4, 5)) * Y.conj()).real)
    rhs = np.sum(x * O.inv_transform(Y, (11, 14), cmode=0, scale=1.0))
    assert abs(lhs - rhs) < 1e-9 * abs(lhs)


def test_fourier_layer_backward_matches_finite_differences():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((1, 2, 6, 8))
    ws = [0.3 * (rng.standard_normal((2, 2, 2, 3)) + 1j * rng.standard_normal((2, 2, 2, 3))) for _ in range(2)]
    wl = rng.standard_normal((2, 2, 1, 1))
    bl = rng.standard_normal(2)
    g = rng.standard_normal((1, 2, 6, 8))
    ga, gws, gwl, gbl = O.fourier_layer_backward(a, ws, wl, bl, True, g)

    def loss(a_, ws_, wl_, bl_):
        return float(np.sum(g * O.fourier_layer_forward(a_, ws_, wl_, bl_, True)[0]))

    eps = 1e-6
    for idx in [(0, 1, 2, 3), (0, 0, 5, 7)]:
        d = np.zeros_like(a); d[idx] = eps
        fd = (loss(a + d, ws, wl, bl) - loss(a - d, ws, wl, bl)) / (2 * eps)
        assert abs(fd - ga[idx]) < 1e-6 * max(1.0, abs(fd))
    idx = (1, 0, 1, 2)
    for part, pick in ((1.0, np.real), (1j, np.imag)):
        d = np.zeros_like(ws[1]); d[idx] = eps * part
        fd = (loss(a, [ws[0], ws[1] + d], wl, bl) - loss(a, [ws[0], ws[1] - d], wl, bl)) / (2 * eps)
        assert abs(fd - pick(gws[1][idx])) < 1e-6 * max(1.0, abs(fd))
    d = np.zeros_like(wl); d[1, 0, 0, 0] = eps
    fd = (loss(a, ws, wl + d, bl) - loss(a, ws, wl - d, bl)) / (2 * eps)
    assert abs(fd - gwl[1, 0, 0, 0]) < 1e-6 * max(1.0, abs(fd))
    d = np.zeros_like(bl); d[0] = eps
    fd = (loss(a, ws, wl, bl + d) - loss(a, ws, wl, bl - d)) / (2 * eps)
    assert abs(fd - gbl[0]) < 1e-6 * max(1.0, abs(fd))


def _params(golden, name):
    pre = f"{name}_param_"
    return {k[len(pre):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith(pre)}


@pytest.mark.parametrize("name,inputs,aux", [
    ("fno2d", ("fno2d_x", "fno2d_grid"), False),
    ("fno3d", ("fno3d_x", "fno3d_grid"), False),
    ("aux2d", ("aux2d_x", "aux2d_grid", "aux2d_xa", "aux2d_ga"), True),
])
def test_torch_port_matches_reference_models(golden_models, name, inputs, aux):
    g = golden_models
    # the port runs in fp64 here, so the residual is the reference's own fp32 rounding; bias/weight
    # gradients are long cancelling sums, hence the looser 1e-5 bound on them
    p = P.as_leaves(_params(g, name), dtype=torch.float64)
    args = [torch.from_numpy(g[k]).double() for k in inputs]
    outs = P.fno_aux_forward(p, *args) if aux else (P.fno_forward(p, *args),)
    loss = 0.0
    for k, o in enumerate(outs):
        assert O.rel_err(o.detach().numpy(), g[f"{name}_out{k}"]) < TOL
        loss = loss + (o * torch.from_numpy(g[f"{name}_g{k}"])).sum()
    loss.backward()
    pre = f"{name}_grad_"
    checked = 0
    for k in g.files:
        if k.startswith(pre):
            assert O.rel_err(p[k[len(pre):]].grad.numpy(), g[k]) < 1e-5, k
            checked += 1
    assert checked >= 6


def test_port_init_matches_reference_rng_order(golden_cfg1):
    meta, _ = golden_cfg1
    torch.manual_seed(meta["seed"])
    c = meta["ctor"]
    p = P.init_params(2, c["num_channels"], (c["modes1"], c["modes2"]), c["width"], c["initial_step"])
    assert list(p.keys()) == list(meta["params"].keys())
    for k, fp in meta["params"].items():
        t = p[k]
        r = (torch.view_as_real(t) if t.is_complex() else t).double().flatten()
        assert list(t.shape) == fp["shape"] and str(t.dtype).replace("torch.", "") == fp["dtype"]
        assert abs(float(r.sum()) - fp["sum"]) <= 1e-9 * max(1.0, abs(fp["sum"]))
        assert [float(v) for v in r[:4]] == fp["head"]
    assert float(torch.rand(1)) == meta["rng_after_init"]


def test_port_cfg1_forward_samples(golden_cfg1):
    """Full-size (128x128, width 20, modes 12) forward of the port vs sampled reference outputs."""
    from oracle.make_golden import seeded

    meta, arr = golden_cfg1
    torch.manual_seed(meta["seed"])
    c = meta["ctor"]
    p = P.init_params(2, c["num_channels"], (c["modes1"], c["modes2"]), c["width"], c["initial_step"])
    x = seeded((2, 128, 128, 10, 2), 800)
    lin = torch.linspace(-1 + 1 / 128, 1 - 1 / 128, 128)
    gx, gy = torch.meshgrid(lin, lin, indexing="ij")
    grid = torch.stack((gx, gy), dim=-1).unsqueeze(0).repeat(2, 1, 1, 1)
    with torch.no_grad():
        out = P.fno_forward(p, x, grid)
    assert O.rel_err(out.flatten().numpy()[arr["out_idx"]], arr["out_val"]) < TOL
