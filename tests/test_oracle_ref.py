"""The vendored, unmodified reference (oracle/_ref, built by oracle/build_ref.py in the build container) against the oracle
port that the parity tests use, and against the committed goldens it once generated.  Skipped where oracle/_ref is absent."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import build_ref
from tests import util

pytestmark = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not built (python oracle/build_ref.py)")


def test_ref_files_are_the_manifested_bytes():
    man = json.load(open(os.path.join(build_ref.DST, "MANIFEST.json")))["sha256"]
    assert sorted(man) == sorted(build_ref.FILES)
    for rel, sha in man.items():
        assert hashlib.sha256(open(os.path.join(build_ref.DST, rel), "rb").read()).hexdigest() == sha, rel


def test_ref_reproduces_golden_and_port_matches_it():
    from oracle import torch_port
    ref = build_ref.import_ref()
    c = util.manifest()["cases"]["small32"]
    torch.manual_seed(c["seed"])
    m = ref.BaseFlowModel(device="cpu", **c["kwargs"])
    m.eval()
    g = util.golden("small32")
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    with torch.no_grad():
        v = m(x, t)
        s2 = m.sample(noise=x, num_steps=2)
    assert np.abs(v.numpy() - g["v"]).max() <= 1e-5 and np.abs(s2.numpy() - g["sample_2"]).max() <= 1e-5
    P = {k: p.detach() for k, p in m.state_dict().items()}
    kw = c["kwargs"]
    arch = dict(model_channels=kw.get("model_channels", 64), channel_mult=tuple(kw.get("channel_mult", [1, 2, 4])),
                num_res_blocks=kw.get("num_res_blocks", 2))
    assert util.rel_l2(torch_port.unet_forward(P, x, t, **arch).numpy(), v.numpy()) <= 2e-5
    assert util.rel_l2(torch_port.euler_sample(P, x, 2, **arch).numpy(), s2.numpy()) <= 2e-5
