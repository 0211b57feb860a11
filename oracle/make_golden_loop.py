"""Runs the UNMODIFIED reference training loop (/root/reference/pdebench/models/fno/train.py,
`run_training`, training_type="single") on CPU with the reference FNO2d and a synthetic
PDEBench-shaped dataset, and stores the per-epoch numbers it prints in tests/golden/loop_cfg1.json.

Build container only (the reference tree is absent on the GPU box):   python oracle/make_golden_loop.py

TEST INFRASTRUCTURE ONLY.  Recipe of SURVEY.md appendix B: stub the modules the loop imports but never
uses in training (h5py, matplotlib), register the synthetic dataset under the module name the loop
imports its loader from (fno/train.py:9), disable wandb, call run_training(**kwargs) directly.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import re
import sys
import tempfile
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/pdebench/models")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sciml-pde_b200"))

CFG = dict(n=128, modes=12, width=20, initial_step=10, num_channels=2, batch_size=4, epochs=3,
           train_traj=3, val_traj=1, windows=8, train_seed=0, val_seed=1, learning_rate=1e-3)
# `--long`: 10 epochs x 20 iterations = 200 optimizer steps (the cosine LR decays to zero through the clip-active
# regime and the loss falls by more than 30 %): the fixture the fused / CUDA-graph step is held to
CFG_LONG = dict(CFG, epochs=10, train_traj=8, windows=10)
OUT_NAME = "loop_cfg1.json"
if "--long" in sys.argv:
    CFG = CFG_LONG
    OUT_NAME = "loop_cfg1_long.json"


def main():
    if not REF.exists():
        raise SystemExit(f"{REF} not found: run in the build container")
    os.environ["WANDB_MODE"] = "disabled"
    import torch

    import importlib.util
    spec = importlib.util.spec_from_file_location("fno_b200_data", ROOT / "sciml-pde_b200" / "fno_b200" / "data.py")
    data = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(data)          # the synthetic generator only: no CUDA library needed

    class FNODatasetMult(data.SyntheticWindows):
        def __init__(self, saved_folder=None, train_subsample=None, rollout_test=1, if_test=False):
            super().__init__(CFG["val_traj"] if if_test else CFG["train_traj"], CFG["n"], CFG["windows"],
                             CFG["initial_step"], CFG["num_channels"],
                             seed=CFG["val_seed"] if if_test else CFG["train_seed"], rollout=rollout_test)

    sys.path.insert(0, str(REF))
    for name in ["h5py", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda *a, **k: None
    import fno  # noqa: F401  (namespace package of the reference tree)
    stub = types.ModuleType("fno.utils_2d_ns_baseline_lie")
    stub.FNODatasetMult = FNODatasetMult
    sys.modules["fno.utils_2d_ns_baseline_lie"] = stub
    torch.set_num_threads(8)
    from fno.train import run_training

    buf = io.StringIO()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                      # the loop writes <name>_FNO.pt into cwd
        try:
            with contextlib.redirect_stdout(buf):
                run_training(if_training=True, continue_training=False, rollout_test=1, num_workers=0,
                             modes=CFG["modes"], width=CFG["width"], initial_step=CFG["initial_step"], t_train=101,
                             num_channels=CFG["num_channels"], batch_size=CFG["batch_size"], epochs=CFG["epochs"],
                             train_subsample=[8, 4, 12], learning_rate=CFG["learning_rate"], scheduler_step=100,
                             scheduler_gamma=0.5, model_update=1, FNO_model_flmn="golden", plot=False, channel_plot=0,
                             x_min=-1, x_max=1, y_min=-1, y_max=1, t_min=0, t_max=5, base_path="unused",
                             training_type="single", scheduler="cosine")
        finally:
            os.chdir(cwd)
    text = buf.getvalue()
    print(text)
    rows = re.findall(r"epoch: (\d+), loss: ([\d.eE+-]+),\s+trainL2: ([\d.eE+-]+), testL2: ([\d.eE+-]+)", text)
    if len(rows) != CFG["epochs"]:
        raise SystemExit("could not parse the reference loop's per-epoch lines")
    out = {"config": CFG, "seed": 16, "torch": torch.__version__,
           "source": "stdout of the unmodified /root/reference/pdebench/models/fno/train.py::run_training (CPU, fp32)",
           "printed_precision": "5 decimals (train.py:341-345)",
           "epochs": [{"epoch": int(e), "loss": float(l), "trainL2": float(t), "testL2": float(v)} for e, l, t, v in rows]}
    (ROOT / "tests" / "golden" / OUT_NAME).write_text(json.dumps(out, indent=1))
    print("wrote tests/golden/" + OUT_NAME)


if __name__ == "__main__":
    main()
