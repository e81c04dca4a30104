"""GPU parity on NON-TRIVIAL weights (north_star: "match the reference ... on the repo's committed checkpoints").

The reference's checkpoints are absent from the checkout (`.MISSING_LARGE_BLOBS`) and constructor-initialised weights leave
every GroupNorm at gamma = 1, beta = 0, so three further weight sets are pinned here, each through the C ABI:
  1. PERTURBED weights (oracle/perturb.py: GroupNorm gamma ~ 1 + 0.5 N, beta ~ 0.5 N, biases + 0.2 N) against goldens the
     UNMODIFIED reference produced after `load_state_dict` of the same tensors (tests/golden/pert_*.npz): velocity, per-layer
     taps, 1 / 8 / 100-step Euler, loss, straightness and the reference's own `loss.backward()` gradients at 32, 64 and 128 px.
  2. The unperturbed default net over 100 Euler steps (the pair-generation step count).
  3. NATIVELY TRAINED weights: train_rectified_flow on the GPU for a few dozen steps, then CUDA vs the CPU port of the
     reference (oracle/torch_port.py, itself pinned to the goldens) on whatever the training produced.
Tolerances are those of tests/test_gpu_parity.py / tests/test_gpu_train.py (bf16 storage, fp32 accumulate).
"""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

TOL_REF_L2, TOL_REF_MAX, TOL_POLICY_L2, MIN_PSNR = 3e-2, 5e-2, 2e-2, 45.0
TOL_LOSS, TOL_GNORM, TOL_GRAD_L2 = 5e-3, 1e-2, 5e-2
FLAG_KEEP_ACTS, FLAG_FUSE_GN, FLAG_NO_FUSE_GN, FLAG_NO_WA, FLAG_NO_UMMA = 4, 4096, 524288, 1048576, 1


@pytest.fixture(scope="module", params=["pert_small32", "pert_default64", "pert_default128"])
def case(request):
    return request.param


def _model(case, cls=None):
    import rectified_flow_vision_b200 as pkg
    return util.perturbed_model(case, device="cuda:0", cls=cls or pkg.RectifiedFlowModel)


def _engine(m, size, flags=0, micro_batch=4):
    from rectified_flow_vision_b200 import engine as E
    eng = E.Engine(m.velocity_net.arch(), size, torch.device("cuda:0"), micro_batch=micro_batch, flags=flags)
    eng.sync_weights(m.velocity_net)
    return eng


def test_velocity_on_perturbed_weights(case):
    m = _model(case)
    g = util.golden(case)
    m.eval()
    with torch.no_grad():
        v = m(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()).cpu().numpy()
    assert np.isfinite(v).all()
    l2, mx = util.rel_l2(v, g["v"]), util.max_rel(v, g["v"])
    print(f"{case}: velocity rel-L2 {l2:.3e}, max-rel {mx:.3e}")
    assert l2 <= TOL_REF_L2 and mx <= TOL_REF_MAX


@pytest.mark.parametrize("flags,name", [(0, "default plan"), (FLAG_FUSE_GN, "GroupNorm fused into every weights-as-A conv"),
                                        (FLAG_NO_FUSE_GN, "GroupNorm never fused"), (FLAG_NO_WA, "per-tap tcgen05 kernel"),
                                        (FLAG_NO_UMMA, "mma.sync kernels"), (8388608, "single-CTA 256-channel convs (no CTA pairs)")])
def test_every_kernel_plan_on_perturbed_weights(flags, name):
    """gamma / beta indexing of each GroupNorm code path (stand-alone apply, coefficient table of the fused conv, virtual
    concat) checked where it is visible: on weights with gamma != 1, beta != 0."""
    for case in ("pert_small32", "pert_default64"):
        m = _model(case)
        g = util.golden(case)
        eng = _engine(m, g["x"].shape[-1], flags=flags)
        v = eng.velocity(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()).cpu().numpy()
        assert util.rel_l2(v, g["v"]) <= TOL_REF_L2 and util.max_rel(v, g["v"]) <= TOL_REF_MAX, (name, case)


def test_layers_on_perturbed_weights():
    from oracle import unet_oracle as O
    for case in ("pert_small32", "pert_default64"):
        m = _model(case)
        g, info = util.golden(case), util.weights_manifest()["cases"][case]
        kw = info["kwargs"]
        spec = O.UNetSpec(model_channels=kw.get("model_channels", 64), channel_mult=kw.get("channel_mult", [1, 2, 4]),
                          num_res_blocks=kw.get("num_res_blocks", 2))
        taps = {}
        O.unet_forward(util.numpy_params(m), g["x"], g["t"], spec, policy=O.BF16_POLICY, taps=taps)
        eng = _engine(m, g["x"].shape[-1], flags=FLAG_KEEP_ACTS)
        eng.velocity(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda())
        for name, ref in taps.items():
            got = eng.debug_activation(name, ref.size).cpu().numpy().reshape(ref.shape)
            assert util.rel_l2(got, ref) <= TOL_POLICY_L2, (case, name, util.rel_l2(got, ref))
            # and against the reference's own fp32 activations (strided sample)
            assert util.rel_l2(got.reshape(-1)[::997], g["tap_" + name]) <= TOL_REF_L2, (case, name)


def test_euler_on_perturbed_weights(case):
    m = _model(case)
    g, info = util.golden(case), util.weights_manifest()["cases"][case]
    noise = torch.from_numpy(g["x"]).cuda()
    for steps in info["steps"]:
        out = m.sample(noise=noise, num_steps=steps).cpu().numpy()
        ps, l2 = util.psnr(out, g[f"sample_{steps}"]), util.rel_l2(out, g[f"sample_{steps}"])
        print(f"{case}: {steps}-step PSNR {ps:.1f} dB, rel-L2 {l2:.3e}")
        assert ps >= MIN_PSNR and l2 <= TOL_REF_L2, (steps, ps, l2)


def test_hundred_steps_on_seeded_weights():
    m = util.seeded_model("default64", device="cuda:0")
    g = util.golden("default64")
    out = m.sample(noise=torch.from_numpy(g["x"]).cuda(), num_steps=100).cpu().numpy()
    ref = util.golden("default64_100")["sample_100"]
    assert util.psnr(out, ref) >= MIN_PSNR and util.rel_l2(out, ref) <= TOL_REF_L2


def test_production_plan_with_golden_rows_inside_a_full_batch():
    """The plan the bench runs -- micro-batch 512, two lanes, ragged tail -- with the reference's golden rows embedded at
    positions that land in different micro-batches and lanes; plus the host-buffer path (generate_reflow_pairs)."""
    import rectified_flow_vision_b200 as pkg
    case = "pert_default64"
    m = _model(case)
    g = util.golden(case)
    gen = torch.Generator().manual_seed(11)
    noise = torch.randn(1030, 3, 64, 64, generator=gen)
    rows = (5, 1027)
    for r, src in zip(rows, g["x"]):
        noise[r] = torch.from_numpy(src)
    out = m.sample(noise=noise.cuda(), num_steps=8).cpu().numpy()
    for r, ref in zip(rows, g["sample_8"]):
        assert util.psnr(out[r], ref) >= MIN_PSNR and util.rel_l2(out[r], ref) <= TOL_REF_L2, r
    x0, x1 = pkg.generate_reflow_pairs(m, num_pairs=1030, num_steps=8, noise=noise.pin_memory())
    assert torch.equal(x0, noise)
    for r, ref in zip(rows, g["sample_8"]):
        assert util.psnr(x1[r].numpy(), ref) >= MIN_PSNR, r


def test_loss_and_straightness_on_perturbed_weights(case):
    m = _model(case)
    g, info = util.golden(case), util.weights_manifest()["cases"][case]
    x0, x1, t = (torch.from_numpy(g[k]).cuda() for k in ("x", "x1", "t"))
    loss = float(m._engine(x0.shape[-1]).fm_loss(x0, x1, t))
    assert abs(loss - info["fm_loss"]) <= 2e-2 * info["fm_loss"]
    s = m.compute_straightness(x0, x1, num_points=3)
    assert abs(s - info["straightness_3"]) <= 2e-2 * info["straightness_3"]


def test_gradients_on_perturbed_weights(case):
    """rfv_train_accumulate against the reference's own loss.backward() with gamma != 1, beta != 0: the GroupNorm backward's
    gamma factor, d(gamma) / d(beta) of all 30 sites (stored in full) and every other tensor's gradient norm."""
    m = _model(case)
    g, info = util.golden(case), util.weights_manifest()["cases"][case]
    x0, x1, t = (torch.from_numpy(g[k]).cuda() for k in ("x", "x1", "t"))
    eng = m.velocity_net.train_engine(x0.shape[-1], "cuda:0", micro_batch=2)
    eng.zero_grad()
    loss = float(eng.train_accumulate(x0, x1, t, dropout_p=0.0, seed=1).item())
    assert abs(loss - info["fm_loss"]) <= TOL_LOSS * info["fm_loss"], loss
    names = [str(n) for n in g["names"]]
    sd = dict(m.named_parameters())
    gn, worst, n_full = [], (0.0, ""), 0
    for k in names:
        gr = eng.get_grad(k, sd[k].numel()).cpu().numpy()
        assert np.isfinite(gr).all(), k
        gn.append(float(np.sqrt((gr.astype(np.float64) ** 2).sum())))
        for pre, sl in (("grad_full/", slice(None)), ("grad_sampled/", slice(None, None, int(g["stride"])))):
            if pre + k in g.files:
                err = util.rel_l2(gr.reshape(-1)[sl], g[pre + k].reshape(-1))
                worst = max(worst, (err, k))
                n_full += 1
                assert err <= TOL_GRAD_L2, (k, err)
    assert n_full >= 40
    gn, ref = np.array(gn), g["grad_norm_per_tensor"]
    total, total_ref = float(np.sqrt((gn ** 2).sum())), float(np.sqrt((ref ** 2).sum()))
    assert abs(total - total_ref) <= TOL_GNORM * total_ref, (total, total_ref)
    big = ref > 1e-3 * ref.max()
    rel = np.abs(gn - ref)[big] / ref[big]
    assert rel.max() <= TOL_GRAD_L2, (names[int(np.flatnonzero(big)[rel.argmax()])], float(rel.max()))
    print(f"{case}: loss {loss:.5f} (ref {info['fm_loss']:.5f}); |g| {total:.4f} (ref {total_ref:.4f}); worst tensor rel-L2 "
          f"{worst[0]:.3e} at {worst[1]}; worst norm rel {rel.max():.3e}")


@pytest.mark.parametrize("kw,size,steps_train", [
    (dict(image_size=32, model_channels=64, channel_mult=[1, 2], num_res_blocks=1), 32, 40),
    (dict(image_size=64), 64, 24)])
@util.retry_once
def test_natively_trained_weights(kw, size, steps_train):
    """Train on the GPU (train_rectified_flow: dropout 0.1, clip, AdamW), then hold the CUDA path to the CPU port of the
    reference on the TRAINED state_dict: velocity, 8-step Euler, loss and gradients."""
    import rectified_flow_vision_b200 as pkg
    from oracle import torch_port as TP
    from oracle import train_oracle as T
    torch.manual_seed(5)
    m = pkg.RectifiedFlowModel(device="cuda:0", **kw)
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(9)
    n = 16 * steps_train // 2
    x0 = torch.randn(n, 3, size, size, generator=gen)
    x1 = (0.5 * torch.randn(n, 3, size, size, generator=gen)).clamp(-1, 1)
    losses = pkg.train_rectified_flow(m, x0, x1, epochs=2, batch_size=16, lr=1e-3)
    assert len(losses) == 2 and np.isfinite(losses).all() and losses[1] < losses[0]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    moved = [float((sd[k] - before[k].cpu()).abs().max()) for k in sd]
    assert min(moved) > 0, "some tensor was never updated"
    gam = sd["velocity_net.dec_blocks.0.norm1.weight"]
    assert float((gam - 1).abs().max()) > 1e-3, "training left GroupNorm at its initialisation"
    arch = util.arch_of(kw)
    xt, x1t, t = torch.randn(2, 3, size, size, generator=gen), torch.randn(2, 3, size, size, generator=gen), torch.rand(2, generator=gen)
    m.eval()
    with torch.no_grad():
        v = m(xt.cuda(), t.cuda()).cpu().numpy()
        v_ref = TP.unet_forward(sd, xt, t, **arch).numpy()
        s8 = m.sample(noise=xt.cuda(), num_steps=8).cpu().numpy()
        s8_ref = TP.euler_sample(sd, xt, 8, **arch).numpy()
    assert util.rel_l2(v, v_ref) <= TOL_REF_L2 and util.max_rel(v, v_ref) <= TOL_REF_MAX, util.rel_l2(v, v_ref)
    assert util.psnr(s8, s8_ref) >= MIN_PSNR and util.rel_l2(s8, s8_ref) <= TOL_REF_L2
    loss_ref, grads = T.loss_and_grads(sd, xt, x1t, t, **arch)
    eng = m.velocity_net.train_engine(size, "cuda:0")
    eng.zero_grad()
    loss = float(eng.train_accumulate(xt.cuda(), x1t.cuda(), t.cuda(), dropout_p=0.0, seed=1).item())
    assert abs(loss - loss_ref) <= TOL_LOSS * loss_ref
    # Per-tensor bound: 5e-2 rel-L2 where the tensor carries a significant share of the gradient (>= 1e-2 of the largest
    # tensor norm); tensors below that are sums that cancel almost completely, so the bf16 rounding of the activation
    # gradients shows up relative to the un-cancelled magnitude -- they are held to the same ABSOLUTE error a significant
    # tensor would be allowed (5e-2 * 1e-2 * gmax).  The whole gradient's direction is checked on top.
    gmax = max(float(gr.norm()) for gr in grads.values())
    dot = nn_ = nr_ = 0.0
    for k, gr in grads.items():
        got = eng.get_grad(k, gr.numel()).cpu().numpy().reshape(gr.shape).astype(np.float64)
        ref = gr.numpy().astype(np.float64)
        err = float(np.sqrt(((got - ref) ** 2).sum()))
        bound = TOL_GRAD_L2 * max(float(gr.norm()), 1e-2 * gmax)
        assert err <= bound, (k, err / max(float(gr.norm()), 1e-30), float(gr.norm()) / gmax)
        dot += float((got * ref).sum()); nn_ += float((got ** 2).sum()); nr_ += float((ref ** 2).sum())
    cos = dot / np.sqrt(nn_ * nr_)
    assert cos >= 0.9995 and abs(np.sqrt(nn_ / nr_) - 1) <= TOL_GNORM, (cos, np.sqrt(nn_ / nr_))
