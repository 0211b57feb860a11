"""SpectralConv1d (named by north_star; absent from the reference tree, so PARITY IS UNPINNED against the reference --
oracle/dft_oracle.py::spectral_conv1d_* states the 2-D layer's algorithm, fno/fno.py:70-92, one dimension down).

* CPU: the fp64 dense-DFT oracle against torch.fft (the library the reference's 2-D layer calls) incl. autograd gradients.
* GPU: the module through the C ABI against the oracle: forward, input and weight gradients; odd N, Nyquist column
  (m = N/2 + 1), Ci != Co; constructor / parameter contract of the 2-D layer.  Tolerance 1e-5 (max-norm relative).
"""
import numpy as np
import pytest
import torch

from oracle import dft_oracle as O

CASES = [(3, 4, 5, 64, 12), (2, 20, 20, 256, 16), (2, 3, 2, 33, 17), (2, 2, 3, 16, 9), (1, 8, 8, 1024, 64), (4, 5, 5, 10, 1)]


def _data(B, Ci, Co, N, m, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Ci, N, generator=g)
    w = torch.rand(Ci, Co, m, dtype=torch.cfloat, generator=g) / (Ci * Co)
    gy = torch.randn(B, Co, N, generator=g)
    return x, w, gy


def _torch_ref(x, w, gy):
    x = x.double().requires_grad_()
    w = w.to(torch.complex128).requires_grad_()
    m, N = w.shape[-1], x.shape[-1]
    x_ft = torch.fft.rfft(x)
    out_ft = torch.zeros(x.shape[0], w.shape[1], N // 2 + 1, dtype=torch.complex128)
    out_ft[:, :, :m] = torch.einsum("bix,iox->box", x_ft[:, :, :m], w)
    y = torch.fft.irfft(out_ft, n=N)
    y.backward(gy.double())
    return y.detach().numpy(), x.grad.numpy(), w.grad.numpy()


@pytest.mark.parametrize("case", CASES)
def test_oracle_1d_matches_torch_fft(case):
    x, w, gy = _data(*case)
    y, gx, gw = _torch_ref(x, w, gy)
    assert O.rel_err(O.spectral_conv1d_forward(x.numpy(), w.numpy()), y) < 1e-12
    ogx, ogw = O.spectral_conv1d_backward(x.numpy(), w.numpy(), gy.numpy())
    assert O.rel_err(ogx, gx) < 1e-12
    assert O.rel_err(ogw, gw) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_spectral_conv1d_module(case):
    from fno_b200.spectral import SpectralConv1d

    B, Ci, Co, N, m = case
    x, w, gy = _data(*case)
    dev = torch.device("cuda", 0)
    layer = SpectralConv1d(Ci, Co, m).to(dev)
    assert layer.weights1.shape == (Ci, Co, m) and layer.weights1.dtype == torch.cfloat and layer.modes1 == m
    with torch.no_grad():
        layer.weights1.copy_(w)
    xd = x.to(dev).requires_grad_()
    y = layer(xd)
    y.backward(gy.to(dev))
    assert O.rel_err(y.detach().cpu().numpy(), O.spectral_conv1d_forward(x.numpy(), w.numpy())) < 1e-5
    ogx, ogw = O.spectral_conv1d_backward(x.numpy(), w.numpy(), gy.numpy())
    assert O.rel_err(xd.grad.cpu().numpy(), ogx) < 1e-5
    assert O.rel_err(layer.weights1.grad.cpu().numpy(), ogw) < 1e-5


@pytest.mark.gpu
def test_spectral_conv1d_contract():
    from fno_b200 import lib
    from fno_b200.spectral import SpectralConv1d, SpectralConv2d_fast

    torch.manual_seed(5)
    a = SpectralConv1d(4, 6, 8)
    torch.manual_seed(5)
    ref = (1 / (4 * 6)) * torch.rand(4, 6, 8, dtype=torch.cfloat)          # the 2-D layer's init rule (fno/fno.py:57-63)
    assert torch.equal(a.weights1.detach(), ref)
    assert list(a.state_dict()) == ["weights1"] and a.scale == SpectralConv2d_fast(4, 6, 2, 2).scale
    with pytest.raises(lib.FnoError):
        a(torch.randn(2, 4, 32))                                           # CPU input: no fallback
    dev = torch.device("cuda", 0)
    with pytest.raises(lib.FnoError):
        a.to(dev)(torch.randn(2, 4, 12, device=dev))                       # modes1 = 8 > 12 / 2 + 1
    X = torch.randn(2, 4, 8, dtype=torch.cfloat, device=dev, requires_grad=True)
    W = torch.randn(4, 6, 8, dtype=torch.cfloat, device=dev, requires_grad=True)
    out = a.compl_mul1d(X, W)
    ref_out = torch.einsum("bix,iox->box", X, W)
    assert torch.allclose(out, ref_out, atol=1e-5, rtol=1e-5)
    g = torch.randn_like(out)
    gX, gW = torch.autograd.grad(out, (X, W), g)
    rX, rW = torch.autograd.grad(ref_out, (X, W), g)
    assert torch.allclose(gX, rX, atol=1e-4, rtol=1e-5) and torch.allclose(gW, rW, atol=1e-4, rtol=1e-5)
