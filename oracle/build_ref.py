"""Recipe for oracle/_ref: the UNMODIFIED reference package, vendored outside of history.

    python oracle/build_ref.py          # needs /root/reference (the build container); no-op elsewhere

The reference is pure Python (there is nothing to compile): its hot path lives in models/{unet,base_flow,rectified_flow}.py.
This script copies those files byte for byte into oracle/_ref/models/ (git-ignored, NOT gpurun-ignored: the copy travels
to the GPU box like a built .so, the sources never enter this repository's history) and writes their sha256 next to them.
`bench.py --impl reference` and the `cpu_baseline` leg then time the reference's OWN code (`kind: "reference"`) on the host
cores instead of the functional port; tests/test_oracle_ref.py holds the port to it.

TEST / MEASUREMENT INFRASTRUCTURE: nothing in the product path imports oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["models/__init__.py", "models/unet.py", "models/base_flow.py", "models/rectified_flow.py"]


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "models")):
        if verbose:
            print(f"{SRC} not present: keeping whatever oracle/_ref already holds")
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "AlbertGoTri/rectified-flow-vision (checkout at /root/reference), unmodified", "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} files copied from {SRC}")
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json")) and all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def import_ref():
    """The vendored reference's `models` package (unmodified).  Raises if oracle/_ref was never built."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    import importlib.util
    spec = importlib.util.spec_from_file_location("rfv_reference_models", os.path.join(DST, "models", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(DST, "models")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rfv_reference_models"] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
