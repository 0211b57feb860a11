"""fno_b200 -- B200 (sm_100a) implementation of the FNO spectral-convolution training path of
mehrdadmmz/SciML-PDE, behind the reference's own module API.

    from fno_b200.fno import FNO2d, FNO3d                 # drop-in for fno.fno
    from fno_b200.fno_aux import FNO2d, FNO3d             # drop-in for fno_aux.fno_aux
    from fno_b200.spectral import SpectralConv2d_fast, SpectralConv3d

The compute path is libfno_sm100.so (hand-written CUDA, C ABI in include/fno_sm100.h); it must
be built first (``__graft_entry__.build()``) and there is no CPU fallback.
"""
from . import lib  # noqa: F401
from .spectral import SpectralConv2d_fast, SpectralConv3d  # noqa: F401

__version__ = "0.1.0"
