"""Pin the CPU oracles on NON-TRIVIAL weights: tests/golden/pert_*.npz are the unmodified reference's outputs after its
state_dict was rewritten by oracle/perturb.py (GroupNorm gamma ~ 1 + 0.5 N, beta ~ 0.5 N, biases + 0.2 N), at 32 / 64 / 128
pixels, including a 100-step integration and the reference's own loss.backward() gradients (oracle/make_golden_weights.py)."""
import numpy as np
import pytest
import torch

from oracle import torch_port as TP
from oracle import train_oracle as T
from oracle import unet_oracle as O
from tests import util


def _spec(kw):
    return O.UNetSpec(model_channels=kw.get("model_channels", 64), channel_mult=kw.get("channel_mult", [1, 2, 4]),
                      num_res_blocks=kw.get("num_res_blocks", 2))


@pytest.fixture(scope="module", params=["pert_small32", "pert_default64"])
def case(request):
    name = request.param
    m = util.perturbed_model(name)
    return name, m, util.golden(name), util.weights_manifest()["cases"][name]


def test_perturbation_is_not_trivial(case):
    name, m, g, info = case
    sd = m.state_dict()
    gam = sd["velocity_net.dec_blocks.0.norm1.weight"]
    bet = sd["velocity_net.dec_blocks.0.norm1.bias"]
    assert float((gam - 1).abs().mean()) > 0.2 and float(bet.abs().mean()) > 0.2
    assert float(sd["velocity_net.enc_blocks.0.conv1.bias"].abs().mean()) > 0.1


def test_numpy_oracle_velocity_and_layers(case):
    name, m, g, info = case
    taps = {}
    v = O.unet_forward(util.numpy_params(m), g["x"], g["t"], _spec(info["kwargs"]), taps=taps)
    assert util.rel_l2(v, g["v"]) < 2e-5 and util.max_rel(v, g["v"]) < 1e-4
    for lname, meta in info["taps"].items():
        a = taps[lname]
        assert list(a.shape) == meta["shape"], lname
        assert util.rel_l2(a.reshape(-1)[::997], g["tap_" + lname]) < 5e-5, lname


def test_numpy_oracle_euler_and_loss(case):
    name, m, g, info = case
    P, spec = util.numpy_params(m), _spec(info["kwargs"])
    for steps in ((1, 8, 100) if name == "pert_small32" else (1,)):
        assert util.rel_l2(O.euler_sample(P, g["x"], steps, spec), g[f"sample_{steps}"]) < 5e-5, steps
    assert abs(O.fm_loss(P, g["x"], g["x1"], g["t"], spec) - info["fm_loss"]) < 1e-4 * info["fm_loss"]
    if name == "pert_small32":
        assert abs(O.straightness(P, g["x"], g["x1"], 3, spec) - info["straightness_3"]) < 1e-4 * info["straightness_3"]


def test_torch_port_and_training_oracle(case):
    name, m, g, info = case
    P = {k: v.detach() for k, v in m.state_dict().items()}
    arch = util.arch_of(info["kwargs"])
    x, x1, t = (torch.from_numpy(g[k]) for k in ("x", "x1", "t"))
    with torch.no_grad():
        assert util.rel_l2(TP.unet_forward(P, x, t, **arch).numpy(), g["v"]) < 2e-5
        assert util.rel_l2(TP.euler_sample(P, x, 8, **arch).numpy(), g["sample_8"]) < 5e-5
    loss, grads = T.loss_and_grads(P, x, x1, t, **arch)
    assert abs(loss - info["fm_loss"]) < 1e-4 * info["fm_loss"]
    names = [str(n) for n in g["names"]]
    ref_norms = g["grad_norm_per_tensor"]
    for k, rn in zip(names, ref_norms):
        assert abs(float(grads[k].norm()) - rn) <= 1e-3 * max(rn, 1e-3 * ref_norms.max()), k
    n_full = 0
    for key in g.files:
        if key.startswith("grad_full/"):
            assert util.rel_l2(grads[key[10:]].numpy(), g[key]) < 1e-3, key
            n_full += 1
        elif key.startswith("grad_sampled/"):
            assert util.rel_l2(grads[key[13:]].numpy().reshape(-1)[::int(g["stride"])], g[key]) < 1e-3, key
    assert n_full >= 30   # every GroupNorm gradient is stored in full


def test_default64_hundred_steps_unperturbed():
    """100 Euler steps (the pair-generation step count, experiments/train_rectified.py:76-80) on the seeded default net."""
    m = util.seeded_model("default64")
    assert util.state_sha(m.state_dict()) == util.weights_manifest()["default64_100"]["state_sha256"]
    g = util.golden("default64")
    P = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        out = TP.euler_sample(P, torch.from_numpy(g["x"][:1]), 100).numpy()
    assert util.rel_l2(out, util.golden("default64_100")["sample_100"][:1]) < 1e-4


def test_128px_torch_port():
    name = "pert_default128"
    m = util.perturbed_model(name)
    g, info = util.golden(name), util.weights_manifest()["cases"][name]
    P = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        v = TP.unet_forward(P, torch.from_numpy(g["x"]), torch.from_numpy(g["t"])).numpy()
    assert util.rel_l2(v, g["v"]) < 2e-5
