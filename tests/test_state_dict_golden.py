"""state_dict / RNG contract of the drop-in modules against fingerprints of the UNMODIFIED reference modules
(checkpoints must round-trip ``load_state_dict(strict=True)`` both ways).

Fixtures (generated in the build container from /root/reference, scripts committed under oracle/):
  tests/golden/cfg1_meta.json   ``params`` (fno.FNO2d cfg 1), ``aux_params`` (fno_aux.FNO2d cfg 1, 44 keys incl. the
                                aliased ``shared_layers.N.*``), ``fno3d_cfg4_params`` (fno.FNO3d cfg 4, sampled keys)
  tests/golden/aux3d_meta.json  the complete 100-key state_dict of a small fno_aux.FNO3d
Each fingerprint = shape, dtype, sum, |sum| and the first four scalars of the tensor drawn with seed 16: equal values
mean the constructor consumed the global RNG exactly like the reference constructor.  Constructing the modules
needs no GPU (the CUDA library is only touched by forward)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def _fingerprint(t):
    r = torch.view_as_real(t) if t.is_complex() else t
    r = r.double().flatten()
    return {"shape": list(t.shape), "dtype": str(t.dtype).replace("torch.", ""), "sum": float(r.sum()),
            "abs_sum": float(r.abs().sum()), "head": [float(v) for v in r[:4]]}


def _check(sd, golden, all_keys):
    if all_keys:
        assert list(sd.keys()) == list(golden.keys())                 # same keys in the same order
    for k, fp in golden.items():
        got = _fingerprint(sd[k])
        assert got["shape"] == fp["shape"] and got["dtype"] == fp["dtype"], k
        assert got["head"] == fp["head"], k                              # bit-identical draws
        assert abs(got["sum"] - fp["sum"]) <= 1e-9 * max(1.0, fp["abs_sum"]), k
        assert abs(got["abs_sum"] - fp["abs_sum"]) <= 1e-9 * max(1.0, fp["abs_sum"]), k


@pytest.fixture(scope="module")
def meta():
    return json.loads((GOLDEN / "cfg1_meta.json").read_text())


def test_fno2d_cfg1_state_dict_matches_reference(meta):
    from fno_b200.fno import FNO2d

    torch.manual_seed(meta["seed"])
    m = FNO2d(**meta["ctor"])
    _check(m.state_dict(), meta["params"], all_keys=True)
    assert float(torch.rand(1)) == meta["rng_after_init"]               # the constructor left the RNG where the reference does


def test_aux_fno2d_cfg1_state_dict_matches_reference(meta):
    from fno_b200.fno_aux import FNO2d

    torch.manual_seed(meta["seed"])
    m = FNO2d(**meta["ctor"])
    sd = m.state_dict()
    assert len(sd) == 44
    _check(sd, meta["aux_params"], all_keys=True)
    # the aliases are the trunk tensors themselves (fno_aux.py:118-121): optimizers see 24 unique parameters
    assert sd["shared_layers.1.weights1"].data_ptr() == sd["conv0.weights1"].data_ptr()
    assert len(list(m.parameters())) == 24


def test_fno3d_cfg4_state_dict_matches_reference(meta):
    from fno_b200.fno import FNO3d

    torch.manual_seed(meta["seed"])
    m = FNO3d(num_channels=5, modes1=12, modes2=12, modes3=12, width=20, initial_step=10)
    sd = m.state_dict()
    assert len(sd) == 50                                                 # incl. the dead bn0-3 (fno.py:334-337)
    assert sd["bn2.num_batches_tracked"].dtype == torch.int64
    _check(sd, meta["fno3d_cfg4_params"], all_keys=False)


def test_aux_fno3d_state_dict_matches_reference():
    from fno_b200.fno_aux import FNO3d

    am = json.loads((GOLDEN / "aux3d_meta.json").read_text())
    torch.manual_seed(am["seed"])
    m = FNO3d(**am["ctor"])
    sd = m.state_dict()
    assert len(sd) == 100
    _check(sd, am["state_dict"], all_keys=True)
    assert [k for k, _ in m.named_parameters()] == am["named_parameters"]
    assert float(torch.rand(1)) == am["rng_after_init"]
    # strict round trip through a reference-shaped checkpoint dict
    g = np.load(GOLDEN / "aux3d_small.npz")
    ck = {k[len("aux3d_param_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("aux3d_param_")}
    for k in sd:
        if k.startswith("shared_layers."):
            ck[k] = sd[k]        # the generator skipped the aliases; any tensor of the right shape loads
    m.load_state_dict(ck, strict=True)
