"""C-ABI and host-side checks that need no GPU: libfno_sm100.so loads, exports every function
include/fno_sm100.h declares (and the ctypes binding binds exactly that set), refuses to compute
without a CUDA device instead of falling back to the CPU, and the host-side geometry / bucketing
helpers behave."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "fno_sm100.h"


def header_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(fno_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fno_b200 import lib

    handle = ctypes.CDLL(str(lib.lib_path()))
    names = header_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    assert sorted(lib.EXPORTED_SYMBOLS) == names      # the Python binding covers the whole ABI, nothing else
    L = lib.load()
    assert L.fno_version() >= 100 and L.fno_sm_arch() == 100
    assert L.fno_opt_chunk_bytes() == 40 and L.fno_opt_chunk_floats() > 0
    assert isinstance(L.fno_last_error(), bytes)


def test_no_cpu_fallback():
    """Product code must fail loudly without a CUDA device / on CPU tensors."""
    from fno_b200 import lib
    from fno_b200.fno import FNO2d
    from fno_b200.spectral import SpectralConv2d_fast

    with pytest.raises(lib.FnoError):
        SpectralConv2d_fast(2, 2, 2, 2)(torch.randn(1, 2, 8, 8))
    with pytest.raises(lib.FnoError):
        FNO2d(num_channels=1, modes1=2, modes2=2, width=4, initial_step=2)(torch.randn(1, 8, 8, 2, 1),
                                                                             torch.rand(1, 8, 8, 2))
    with pytest.raises(lib.FnoError):
        lib.get_plan(torch.device("cpu"), (8, 8), (2, 2))
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        rc = lib.load().fno_plan2d_create(0, 16, 16, 4, 4, ctypes.byref(h))
        assert rc < 0 and h.value is None           # FNO_E_CUDA: no device, no plan, no fallback
        assert b"no CPU fallback" in lib.load().fno_last_error() or rc == -2
    # argument validation happens before any device work
    assert lib.load().fno_plan2d_create(0, 8, 8, 5, 2, ctypes.byref(ctypes.c_void_p())) < 0   # 2*m1 > H


def test_trunk_geometry():
    from fno_b200 import lib

    g2 = lib.TrunkGeo((128, 128), 2)
    assert g2.ints == (128, 128, 130, 130) and g2.padded == (130, 130) and g2.npix == 128 * 128
    g3 = lib.TrunkGeo((64, 64, 64), 6)
    assert g3.ints == (64 * 64, 64, 64 * 64, 70) and g3.padded == (64, 64, 70)
    with pytest.raises(lib.FnoError):
        lib.TrunkGeo((8,), 2)


def test_state_dict_layout_matches_reference_contract():
    """Key sets / shapes / dtypes of SURVEY 8b (22 keys for FNO2d, 50 for FNO3d, 44 for the aux FNO2d)."""
    from fno_b200 import fno, fno_aux

    m = fno.FNO2d(num_channels=2, modes1=12, modes2=12, width=20, initial_step=10)
    sd = m.state_dict()
    assert len(sd) == 22
    assert sd["conv0.weights1"].shape == (20, 20, 12, 12) and sd["conv0.weights1"].dtype == torch.complex64
    assert sd["fc0.weight"].shape == (20, 22) and sd["w3.weight"].shape == (20, 20, 1, 1)
    assert sd["fc1.weight"].shape == (128, 20) and sd["fc2.weight"].shape == (2, 128)
    m3 = fno.FNO3d(num_channels=5, modes1=4, modes2=4, modes3=4, width=8, initial_step=2)
    assert len(m3.state_dict()) == 50 and m3.state_dict()["bn0.num_batches_tracked"].dtype == torch.int64
    ma = fno_aux.FNO2d(num_channels=2, modes1=4, modes2=4, width=8, initial_step=3)
    assert len(ma.state_dict()) == 44
    assert len(list(ma.parameters())) == 24          # shared_layers alias the trunk: de-duplicated
    ma.load_state_dict(ma.state_dict(), strict=True)


def test_dp_bucket_order():
    from fno_b200 import fno
    from fno_b200.dp import fno_bucket_names

    m = fno.FNO2d(num_channels=2, modes1=4, modes2=4, width=8, initial_step=3)
    groups = fno_bucket_names(m)
    assert [g[0].split(".")[0] for g in groups] == ["fc1", "conv3", "conv2", "conv1", "fc0"]
    flat = [n for g in groups for n in g]
    assert sorted(flat) == sorted(n for n, _ in m.named_parameters())


def test_epoch_indices_shard_every_global_batch_rank_strided():
    """Host logic of the device-resident dataset (fno_b200/data.py): N ranks at global batch G draw exactly the
    single-GPU batches, split rank-strided, ragged tail kept (drop_last=False)."""
    import torch

    from fno_b200 import data

    single = data.epoch_indices(23, 8, True, 16)
    assert [len(b) for b in single] == [8, 8, 7]
    assert sorted(torch.cat(single).tolist()) == list(range(23))
    for world in (2, 4):
        parts = [data.epoch_indices(23, 8, True, 16, 0, r, world) for r in range(world)]
        for step, whole in enumerate(single):
            merged = torch.stack([p[step] for p in parts if len(p[step]) == len(parts[0][step])], dim=1).flatten().tolist() \
                if all(len(p[step]) == len(parts[0][step]) for p in parts) else None
            union = sorted(torch.cat([p[step] for p in parts]).tolist())
            assert union == sorted(whole.tolist())
            if merged is not None:
                assert merged == whole.tolist()           # rank-strided: interleaving the ranks restores the order
    assert data.epoch_indices(23, 8, False, 0)[0].tolist() == list(range(8))
    assert data.epoch_indices(23, 8, True, 16, 1)[0].tolist() != single[0].tolist()   # next epoch reshuffles
