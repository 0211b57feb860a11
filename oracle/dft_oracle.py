"""CPU oracle #1 -- fp64 dense-DFT restatement of the FNO spectral-convolution path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker.

What it restates (reference = /root/reference/pdebench/models, read-only):

* ``fno/fno.py:66-68``   ``compl_mul2d``  -> :func:`mix_fwd`
* ``fno/fno.py:70-92``   ``SpectralConv2d_fast.forward`` -> :func:`spectral_conv2d_forward`
* ``fno/fno.py:255-257`` ``compl_mul3d``  -> :func:`mix_fwd` (same contraction, more mode axes)
* ``fno/fno.py:259-288`` ``SpectralConv3d.forward`` -> :func:`spectral_conv3d_forward`
* ``fno/fno.py:161-178`` one Fourier layer (spectral conv + 1x1 conv + exact GELU)
  -> :func:`fourier_layer_forward` / :func:`fourier_layer_backward`

The arithmetic of the reference lives in a third-party dependency that is not
vendored under /root/reference: PyTorch (pinned ``torch~=1.13.0`` in
``pyproject.toml:29``; 2.11.0 is what this image has).  This file therefore does
NOT call ``torch.fft``: every transform is written as an explicit dense DFT
matrix product in float64/complex128 (numpy), which is also exactly the
factorisation the CUDA kernels use:

    K1 fwd :  X  = F_H . x . F_W                      (rows K1 = low u high, cols q < m2)
    K2 fwd :  Y[b,o,k] = sum_i X[b,i,k] Wcat[i,o,k]
    K3 fwd :  y  = (1/HW) Re( F_H^H . Y . diag(c) . F_W^H )   c_0 = 1, c_q = 2 (1 at Nyquist)
    K3 bwd :  gY = (c/HW) * (F_H . g . F_W)
    K2 bwd :  gX = sum_o gY conj(W) ;  gW = sum_b conj(X) gY
    K1 bwd :  gx = Re( F_H^H . gX . F_W^H )

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so this oracle is pinned against outputs of the reference modules
themselves, generated in the build container by ``oracle/make_golden.py`` and
committed under ``tests/golden/`` (checked by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is in the image; fall back to math.erf if it ever is not
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


# ----------------------------------------------------------------------------
# DFT building blocks
# ----------------------------------------------------------------------------
def kept_rows(n: int, m: int) -> np.ndarray:
    """Signed-wrapped frequency indices kept along a full-complex axis of length ``n``:
    ``[0, m) u [n-m, n)`` in that storage order (fno.py:84-89 slices ``:m`` and ``-m:``)."""
    if 2 * m > n:
        raise ValueError(f"2*modes ({2 * m}) must be <= axis length ({n})")
    return np.concatenate([np.arange(m), np.arange(n - m, n)])


def dft_rows(n: int, m: int) -> np.ndarray:
    """F[k, h] = exp(-2 pi i k h / n) for the 2m kept rows -- shape [2m, n]."""
    k = kept_rows(n, m)[:, None].astype(np.float64)
    h = np.arange(n)[None, :].astype(np.float64)
    return np.exp(-2j * np.pi * k * h / n)


def dft_half_cols(n: int, m: int) -> np.ndarray:
    """F[w, q] = exp(-2 pi i q w / n) for q < m (the rfft axis) -- shape [n, m]."""
    if m > n // 2 + 1:
        raise ValueError(f"modes ({m}) must be <= n//2+1 ({n // 2 + 1})")
    w = np.arange(n)[:, None].astype(np.float64)
    q = np.arange(m)[None, :].astype(np.float64)
    return np.exp(-2j * np.pi * q * w / n)


def c2r_weights(n: int, m: int) -> np.ndarray:
    """Column weights of a complex-to-real inverse along an axis of length ``n`` when only
    the first ``m`` half-spectrum columns are non-zero: 1 for DC, 1 for Nyquist (n even),
    2 otherwise.  The imaginary part of the DC/Nyquist columns is dropped by taking Re."""
    c = np.full(m, 2.0)
    c[0] = 1.0
    if n % 2 == 0 and m - 1 == n // 2:
        c[-1] = 1.0
    return c


# ----------------------------------------------------------------------------
# K1 / K3: pruned transforms, any number of leading full axes + one half axis
# ----------------------------------------------------------------------------
def fwd_transform(x: np.ndarray, modes, cmode: int = 0, scale: float = 1.0) -> np.ndarray:
    """Pruned real-to-complex forward transform over the trailing ``len(modes)`` axes.

    x: [..., N1, ..., Nd] real.  Returns [..., 2*m1, ..., 2*m_{d-1}, m_d] complex128, i.e.
    only the retained modes, un-normalised (torch.fft ``norm="backward"``), optionally
    multiplied by ``scale`` and (``cmode=1``) by the C2R column weights of the last axis.
    """
    d = len(modes)
    x = np.asarray(x, dtype=np.float64)
    spatial = x.shape[-d:]
    out = x.astype(np.complex128)
    # last (half-spectrum) axis
    fw = dft_half_cols(spatial[-1], modes[-1])           # [W, m_d]
    out = out @ fw                                        # contracts the last axis
    # leading full axes
    for ax in range(d - 1):
        f = dft_rows(spatial[ax], modes[ax])              # [2m, N]
        axis = x.ndim - d + ax
        out = np.moveaxis(np.tensordot(f, out, axes=([1], [axis])), 0, axis)
    if cmode:
        out = out * c2r_weights(spatial[-1], modes[-1])
    return out * scale


def inv_transform(y: np.ndarray, spatial, cmode: int = 1, scale: float | None = None) -> np.ndarray:
    """Zero-padding inverse: retained modes [..., 2*m1, ..., m_d] -> real [..., N1, ..., Nd].

    ``cmode=1, scale=None`` reproduces ``irfftn(out_ft, s=spatial)`` on a spectrum that is
    zero outside the retained corners (fno.py:76-92).  ``cmode=0, scale=1`` is the adjoint
    used for the input gradient of the forward transform.
    """
    d = len(spatial)
    y = np.asarray(y, dtype=np.complex128)
    modes = [y.shape[y.ndim - d + ax] // 2 for ax in range(d - 1)] + [y.shape[-1]]
    if scale is None:
        scale = 1.0 / float(np.prod(spatial))
    out = y
    if cmode:
        out = out * c2r_weights(spatial[-1], modes[-1])
    for ax in range(d - 1):
        f = dft_rows(spatial[ax], modes[ax]).conj().T      # [N, 2m]
        axis = y.ndim - d + ax
        out = np.moveaxis(np.tensordot(f, out, axes=([1], [axis])), 0, axis)
    fw = dft_half_cols(spatial[-1], modes[-1]).conj().T    # [m_d, W]
    out = out @ fw
    return out.real * scale


# ----------------------------------------------------------------------------
# K2: per-mode channel mixing
# ----------------------------------------------------------------------------
def cat_corner_weights(weights) -> np.ndarray:
    """Assemble ``Wcat[i, o, 2*m1, (2*m2,) m_last]`` from the reference's corner parameters.

    2-D: (weights1, weights2) = (low rows, high rows)                     fno.py:84-89
    3-D: (w1, w2, w3, w4) = (x low y low, x high y low, x low y high, x high y high)
                                                                         fno.py:274-285
    """
    ws = [np.asarray(w, dtype=np.complex128) for w in weights]
    if len(ws) == 2:
        return np.concatenate(ws, axis=2)
    if len(ws) == 4:
        low_y = np.concatenate([ws[0], ws[1]], axis=2)
        high_y = np.concatenate([ws[2], ws[3]], axis=2)
        return np.concatenate([low_y, high_y], axis=3)
    raise ValueError("expected 2 (2-D) or 4 (3-D) corner weight tensors")


def split_corner_grads(gwcat: np.ndarray, n_corners: int):
    """Inverse of :func:`cat_corner_weights` for gradients."""
    if n_corners == 2:
        m1 = gwcat.shape[2] // 2
        return [gwcat[:, :, :m1], gwcat[:, :, m1:]]
    m1 = gwcat.shape[2] // 2
    m2 = gwcat.shape[3] // 2
    return [gwcat[:, :, :m1, :m2], gwcat[:, :, m1:, :m2], gwcat[:, :, :m1, m2:], gwcat[:, :, m1:, m2:]]


def mix_fwd(xs: np.ndarray, wcat: np.ndarray) -> np.ndarray:
    """out[b,o,...] = sum_i xs[b,i,...] * wcat[i,o,...]   (complex, no conjugate)."""
    return np.einsum("bi...,io...->bo...", xs, wcat)


def mix_bwd(xs: np.ndarray, gy: np.ndarray, wcat: np.ndarray):
    """(gX, gWcat) in torch's complex-gradient convention (grad = dL/dRe + i dL/dIm)."""
    gx = np.einsum("bo...,io...->bi...", gy, wcat.conj())
    gw = np.einsum("bi...,bo...->io...", xs.conj(), gy)
    return gx, gw


# ----------------------------------------------------------------------------
# SpectralConv forward/backward (2-D and 3-D share the code)
# ----------------------------------------------------------------------------
def _modes_of(weights):
    w0 = np.asarray(weights[0])
    return list(w0.shape[2:])


def spectral_conv_forward(x, weights, return_saved: bool = False):
    """y = irfftn(scatter(mix(rfftn(x)[corners])), s=spatial)  -- fno.py:70-92 / :259-288."""
    modes = _modes_of(weights)
    d = len(modes)
    spatial = x.shape[-d:]
    xs = fwd_transform(x, modes)
    wcat = cat_corner_weights(weights)
    ys = mix_fwd(xs, wcat)
    y = inv_transform(ys, spatial, cmode=1)
    if return_saved:
        return y, (xs, ys, wcat)
    return y


def spectral_conv_backward(x, weights, g):
    """Returns (gx, [gW per corner]) for L = sum(g * y)."""
    modes = _modes_of(weights)
    d = len(modes)
    spatial = x.shape[-d:]
    xs = fwd_transform(x, modes)
    wcat = cat_corner_weights(weights)
    gy = fwd_transform(g, modes, cmode=1, scale=1.0 / float(np.prod(spatial)))
    gxs, gw = mix_bwd(xs, gy, wcat)
    gx = inv_transform(gxs, spatial, cmode=0, scale=1.0)
    return gx, split_corner_grads(gw, len(weights))


def spectral_conv2d_forward(x, w1, w2):
    return spectral_conv_forward(x, [w1, w2])


def spectral_conv2d_backward(x, w1, w2, g):
    return spectral_conv_backward(x, [w1, w2], g)


def spectral_conv3d_forward(x, w1, w2, w3, w4):
    return spectral_conv_forward(x, [w1, w2, w3, w4])


def spectral_conv3d_backward(x, w1, w2, w3, w4, g):
    return spectral_conv_backward(x, [w1, w2, w3, w4], g)


# ----------------------------------------------------------------------------
# exact (erf) GELU, as F.gelu default -- fno.py:164,169,174
# ----------------------------------------------------------------------------
def gelu(s: np.ndarray) -> np.ndarray:
    s = np.asarray(s, dtype=np.float64)
    return 0.5 * s * (1.0 + _erf(s / math.sqrt(2.0)))


def gelu_grad(s: np.ndarray) -> np.ndarray:
    s = np.asarray(s, dtype=np.float64)
    return 0.5 * (1.0 + _erf(s / math.sqrt(2.0))) + s * np.exp(-0.5 * s * s) / math.sqrt(2.0 * math.pi)


# ----------------------------------------------------------------------------
# One Fourier layer: a' = act( SpectralConv(a) + Conv1x1(a) )     fno.py:161-178
# ----------------------------------------------------------------------------
def pointwise_conv(a: np.ndarray, wl: np.ndarray, bl: np.ndarray | None) -> np.ndarray:
    """nn.Conv{2,3}d(C, C, 1): out[b,o,...] = sum_i wl[o,i] a[b,i,...] + bl[o]."""
    w = np.asarray(wl, dtype=np.float64).reshape(wl.shape[0], wl.shape[1])
    out = np.einsum("oi,bi...->bo...", w, np.asarray(a, dtype=np.float64))
    if bl is not None:
        out = out + np.asarray(bl, dtype=np.float64).reshape((1, -1) + (1,) * (a.ndim - 2))
    return out


def fourier_layer_forward(a, weights, wl, bl, apply_gelu: bool):
    """Returns (a_next, s) with s the pre-activation."""
    s = spectral_conv_forward(a, weights) + pointwise_conv(a, wl, bl)
    return (gelu(s) if apply_gelu else s), s


def fourier_layer_backward(a, weights, wl, bl, apply_gelu: bool, g):
    """Gradients of L = sum(g * a_next) wrt (a, corner weights, wl, bl)."""
    _, s = fourier_layer_forward(a, weights, wl, bl, apply_gelu)
    ds = g * gelu_grad(s) if apply_gelu else np.asarray(g, dtype=np.float64)
    ga_spec, gws = spectral_conv_backward(a, weights, ds)
    w = np.asarray(wl, dtype=np.float64).reshape(wl.shape[0], wl.shape[1])
    ga = ga_spec + np.einsum("oi,bo...->bi...", w, ds)
    a64 = np.asarray(a, dtype=np.float64)
    ds2 = ds.reshape(ds.shape[0], ds.shape[1], -1)
    a2 = a64.reshape(a64.shape[0], a64.shape[1], -1)
    gwl = np.einsum("bop,bip->oi", ds2, a2).reshape(np.asarray(wl).shape)
    gbl = ds.sum(axis=tuple(i for i in range(ds.ndim) if i != 1))
    return ga, gws, gwl, gbl


def rel_err(got, ref) -> float:
    """max-abs-err / max-abs-ref -- the tolerance metric stated in SURVEY.md section 8c."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    denom = float(np.max(np.abs(ref)))
    if denom == 0.0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - ref)) / denom)


# ------------------------------------------------------------------------------------------------
# SpectralConv1d.  PARITY UNPINNED against the reference: /root/reference has no 1-D layer; north_star names
# it, so the oracle states the 2-D layer's algorithm (fno/fno.py:70-92) one dimension down,
#     y = irfft(pad(einsum("bix,iox->box", rfft(x)[..., :m], W)), n = N),
# as dense DFT products in float64, and tests/test_spectral1d_gpu.py additionally cross-checks it against torch.fft.
# ------------------------------------------------------------------------------------------------
def spectral_conv1d_forward(x: np.ndarray, w: np.ndarray, return_saved: bool = False):
    """x [B, Ci, N] real, w [Ci, Co, m] complex -> [B, Co, N]."""
    n, m = x.shape[-1], w.shape[-1]
    F = dft_half_cols(n, m)                                   # [N, m]: exp(-2 pi i q w / N)
    X = x.astype(np.float64) @ F                              # pruned rfft
    Y = np.einsum("bik,iok->bok", X, w.astype(np.complex128))
    c = c2r_weights(n, m)
    y = np.real((Y * c) @ np.conj(F).T) / n
    # torch.fft.irfft ignores Im of the DC (and Nyquist) column: Re(Y e^{i0}) does that by construction
    return (y, X) if return_saved else y


def spectral_conv1d_backward(x: np.ndarray, w: np.ndarray, g: np.ndarray):
    """Gradients (gx, gw) in PyTorch's complex convention (dL/dRe + i dL/dIm)."""
    n, m = x.shape[-1], w.shape[-1]
    F = dft_half_cols(n, m)
    _, X = spectral_conv1d_forward(x, w, return_saved=True)
    gY = (g.astype(np.float64) @ F) * (c2r_weights(n, m) / n)
    wc = w.astype(np.complex128)
    gX = np.einsum("bok,iok->bik", gY, np.conj(wc))
    gw = np.einsum("bik,bok->iok", np.conj(X), gY)
    gx = np.real(gX @ np.conj(F).T)
    return gx, gw
