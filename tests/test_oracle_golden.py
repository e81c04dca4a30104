"""Pin the CPU oracle (oracle/unet_oracle.py) against outputs of the reference itself (tests/golden/, generated
by oracle/make_golden.py from /root/reference).  fp32 numpy vs fp32 PyTorch: agreement to ~1e-5."""
import numpy as np
import pytest

from oracle import unet_oracle as O
from tests import util

CASES = ["default64", "small32"]


@pytest.fixture(scope="module", params=CASES)
def case(request):
    name = request.param
    m = util.seeded_model(name)
    return name, util.numpy_params(m), util.golden(name), util.spec_for(name), util.manifest()["cases"][name]


def test_seeded_init_matches_reference_sha(case):
    name, P, g, spec, info = case
    m = util.seeded_model(name)
    assert util.state_sha(m.state_dict()) == info["state_sha256"]
    assert sum(p.numel() for p in m.parameters()) == info["num_params"]


def test_velocity_and_layers_match_reference(case):
    name, P, g, spec, info = case
    taps = {}
    v = O.unet_forward(P, g["x"], g["t"], spec, taps=taps)
    assert util.rel_l2(v, g["v"]) < 2e-5
    assert util.max_rel(v, g["v"]) < 1e-4
    for lname, meta in info["taps"].items():
        a = taps[lname]
        assert list(a.shape) == meta["shape"], lname
        sample = a.reshape(-1)[::997]
        assert util.rel_l2(sample, g["tap_" + lname]) < 5e-5, lname
        assert abs(float(np.sqrt((a.astype(np.float64) ** 2).mean())) - meta["rms"]) < 1e-4 * max(meta["rms"], 1), lname


@pytest.mark.parametrize("steps", [1, 2, 4, 8])
def test_euler_sample_matches_reference(case, steps):
    name, P, g, spec, info = case
    if name == "default64" and steps == 8:
        pytest.skip("covered by small32 (keeps the CPU suite short)")
    x = O.euler_sample(P, g["x"], steps, spec)
    assert util.rel_l2(x, g[f"sample_{steps}"]) < 2e-5


def test_trajectory_matches_reference(case):
    name, P, g, spec, info = case
    traj = O.euler_sample(P, g["x"], 4, spec, return_trajectory=True, save_every=2)
    assert len(traj) == g["traj_4_2"].shape[0] == 3
    for a, b in zip(traj, g["traj_4_2"]):
        assert util.rel_l2(a, b) < 2e-5


def test_interpolation_loss_straightness(case):
    name, P, g, spec, info = case
    xt, target = O.get_interpolation(g["x"], g["x1"], g["t"])
    np.testing.assert_allclose(xt, g["xt"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(target, g["target"], rtol=0, atol=0)
    assert abs(O.fm_loss(P, g["x"], g["x1"], g["t"], spec) - info["fm_loss"]) < 1e-4 * info["fm_loss"]
    if name == "small32":
        s = O.straightness(P, g["x"], g["x1"], 3, spec)
        assert abs(s - info["straightness_3"]) < 1e-4 * info["straightness_3"]


def test_reference_interpolation_identities():
    """The only hot-path assertions the reference's own tests make (tests/test_utils.py:101-133)."""
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal((2, 3, 8, 8)).astype(np.float32)
    x1 = rng.standard_normal((2, 3, 8, 8)).astype(np.float32)
    for tv, want in ((0.0, x0), (1.0, x1), (0.5, (x0 + x1) / 2)):
        xt, tgt = O.get_interpolation(x0, x1, np.full(2, tv, np.float32))
        np.testing.assert_allclose(xt, want, atol=1e-6)
        assert tgt.shape == x0.shape


def test_flop_model_matches_survey():
    assert abs(O.unet_flops_per_image(size=64) / 1e9 - 12.7636) < 2e-3
    assert abs(O.unet_flops_per_image(size=128) / 1e9 - 51.8555) < 5e-3


def test_bf16_policy_is_close_to_fp32(case):
    """The bf16 storage policy the CUDA path uses stays within the stated end-to-end tolerance of fp32."""
    name, P, g, spec, info = case
    v = O.unet_forward(P, g["x"], g["t"], spec, policy=O.BF16_POLICY)
    assert util.rel_l2(v, g["v"]) < 3e-2
    assert util.max_rel(v, g["v"]) < 5e-2


def test_torch_port_matches_reference(case):
    """The functional-PyTorch port used as the CPU baseline is pinned against the same golden vectors."""
    import torch
    from oracle import torch_port
    name, P, g, spec, info = case
    kw = info["kwargs"]
    arch = dict(model_channels=kw.get("model_channels", 64), channel_mult=tuple(kw.get("channel_mult", [1, 2, 4])),
                num_res_blocks=kw.get("num_res_blocks", 2))
    Pt = {k: torch.from_numpy(v) for k, v in P.items()}
    v = torch_port.unet_forward(Pt, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), **arch).numpy()
    assert util.rel_l2(v, g["v"]) < 1e-5
    x = torch_port.euler_sample(Pt, torch.from_numpy(g["x"]), 2, **arch).numpy()
    assert util.rel_l2(x, g["sample_2"]) < 1e-5
