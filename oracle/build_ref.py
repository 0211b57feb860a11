"""Recipe for oracle/_ref/: the reference's own files for the training path, copied VERBATIM from /root/reference so that
the GPU box (where /root/reference does not exist) can run the UNMODIFIED loops -- `fno.train.run_training`
(fno/train.py:43-347) and `fno_aux.fno_train_aux.run_training` (fno_aux/fno_train_aux.py:43-430) -- against either the
reference models or the shadowed drop-in.

    python oracle/build_ref.py            # build container only; __graft_entry__.build() calls it when the tree is there

TEST INFRASTRUCTURE ONLY.  oracle/_ref/ is listed in .gitignore (no reference source enters the history) but not in
.gpurunignore (it travels to the GPU box like the built .so files).  Nothing in the product imports it.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/pdebench/models")
DST = ROOT / "oracle" / "_ref"
FILES = ["fno/fno.py", "fno/train.py", "fno_aux/fno_aux.py", "fno_aux/fno_train_aux.py", "metrics.py", "metrics_aux.py"]


def build(verbose: bool = True) -> bool:
    if not SRC.exists():
        if verbose:
            print(f"[oracle/_ref] {SRC} not present (GPU box): using the prebuilt copy", file=sys.stderr)
        return DST.exists()
    manifest = {}
    for rel in FILES:
        dst = DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(SRC / rel, dst)
        manifest[rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "sha256": manifest}, indent=1))
    if verbose:
        print(f"[oracle/_ref] copied {len(FILES)} reference files", file=sys.stderr)
    return True


if __name__ == "__main__":
    build()
