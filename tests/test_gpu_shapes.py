"""GPU parity on shapes the golden cases do not cover: images narrower than one output-conv / input-conv tile, grayscale and
4-channel images, two-level networks.  Every masked edge of the thin-conv tiles, the GroupNorm slice selection of the
single-pass backward (tiny HW) and the C_out = 4 fallback of the output conv runs here, compared with the CPU oracle
(oracle/unet_oracle.py, oracle/train_oracle.py) on seeded inputs AND with outputs of the reference itself on the same seeds
(tests/golden/shapes_*.npz from oracle/make_golden_shapes.py; tests/test_oracle_shapes_golden.py pins the oracles to them); plus
guard bands around caller buffers (compute-sanitizer is not available on the GPU pool, so out-of-bounds writes at the API
boundary are looked for with sentinels).

Tolerances as tests/test_gpu_parity.py / tests/test_gpu_train.py: velocity rel-L2 <= 3e-2 against the fp32 oracle, <= 2e-2
against the bf16-policy oracle; per-tensor gradient rel-L2 <= 5e-2."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

SHAPES = {
    # name: (image_size, in_channels, model_channels, channel_mult, num_res_blocks)   -- as oracle/make_golden_shapes.py
    "16px_two_levels": (16, 3, 64, [1, 2], 1),         # W = 16 < 32-pixel tile width; lowest level 8x8 = 64 tokens, head dim 32
    "gray_32px": (32, 1, 64, [1, 2, 4], 1),            # C_in = C_out = 1 (K = 9 -> 16 in the input conv, 9 Z columns)
    "four_channel_32px": (32, 4, 64, [1, 2, 4], 1),    # C_in = 4 (K = 36 -> 48), C_out = 4 -> tap-shifted output conv
}


def _build(name):
    import rectified_flow_vision_b200 as pkg
    from oracle.unet_oracle import UNetSpec
    S, cin, mc, mult, nres = SHAPES[name]
    torch.manual_seed(1234)                            # seed / input seed / batch of the golden files
    m = pkg.RectifiedFlowModel(image_size=S, in_channels=cin, model_channels=mc, channel_mult=mult, num_res_blocks=nres,
                               device="cpu")
    m.device = "cuda:0"
    m.to("cuda:0")
    spec = UNetSpec(in_channels=cin, model_channels=mc, out_channels=cin, channel_mult=mult, num_res_blocks=nres)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(5, cin, S, S, generator=g)
    x1 = torch.randn(5, cin, S, S, generator=g)
    t = torch.rand(5, generator=g)
    return m, spec, x, x1, t


@pytest.mark.parametrize("name", list(SHAPES))
def test_velocity_vs_oracle(name):
    from oracle import unet_oracle as O
    m, spec, x, _, t = _build(name)
    m.eval()
    with torch.no_grad():
        v = m(x.cuda(), t.cuda()).cpu().numpy()
    assert np.isfinite(v).all()
    P = util.numpy_params(m)
    ref = O.unet_forward(P, x.numpy(), t.numpy(), spec)
    pol = O.unet_forward(P, x.numpy(), t.numpy(), spec, policy=O.BF16_POLICY)
    assert util.rel_l2(v, ref) <= 3e-2, (name, util.rel_l2(v, ref))
    assert util.rel_l2(v, pol) <= 2e-2, (name, util.rel_l2(v, pol))
    g = np.load(f"{util.GOLD}/shapes_{name}.npz")      # the reference's own forward and 3-step Euler on the same seeds
    assert np.array_equal(g["x"], x.numpy()) and np.array_equal(g["t"], t.numpy())
    assert util.rel_l2(v, g["v"]) <= 3e-2, (name, util.rel_l2(v, g["v"]))
    with torch.no_grad():
        s3 = m.sample(x.cuda(), num_steps=3).cpu().numpy()
    assert util.rel_l2(s3, g["sample_3"]) <= 1e-2, (name, util.rel_l2(s3, g["sample_3"]))
    assert util.psnr(s3, g["sample_3"]) >= 45.0, name


@pytest.mark.parametrize("name", list(SHAPES))
def test_c_abi_euler_sample_leaves_guard_bands_alone(name):
    """rfv_euler_sample integrates x in place and writes trajectory snapshots: both live inside larger allocations here, and the
    bytes before and after them must not change; velocity output likewise.  Two runs agree (the only run-to-run freedom is
    the order of the fp32 statistic atomics), and the Python API (which clones the noise) gives the same samples."""
    import ctypes as C
    from rectified_flow_vision_b200 import engine as E
    m, _, x, _, t = _build(name)
    S = x.shape[-1]
    n = x.numel()
    G, FILL = 1024, 7.25
    eng = E.Engine(m.velocity_net.arch(), S, torch.device("cuda:0"), micro_batch=4)
    eng.sync_weights(m.velocity_net)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = []
    for _ in range(2):
        big = torch.full((n + 2 * G,), FILL, device="cuda:0")
        traj = torch.full((3 * n + 2 * G,), FILL, device="cuda:0")
        view = big[G:G + n].view(x.shape)
        view.copy_(x.cuda())
        rc = eng.lib.rfv_euler_sample(eng.h, C.c_void_p(view.data_ptr()), x.shape[0], 3, C.c_void_p(traj[G:].data_ptr()), 1, stream)
        assert rc == 0, eng.lib.rfv_last_error()
        torch.cuda.synchronize()
        for buf, used in ((big, n), (traj, 3 * n)):
            assert bool((buf[:G] == FILL).all()) and bool((buf[G + used:] == FILL).all()), name
        assert bool(torch.isfinite(view).all())
        assert torch.equal(traj[G + 2 * n:G + 3 * n].view(x.shape), view)      # last snapshot = final state
        outs.append(view.clone().cpu().numpy())
    assert util.rel_l2(outs[1], outs[0]) <= 5e-3, name
    vbig = torch.full((n + 2 * G,), FILL, device="cuda:0")
    xin, tin = x.cuda().contiguous(), t.cuda().contiguous()
    rc = eng.lib.rfv_velocity(eng.h, C.c_void_p(xin.data_ptr()), C.c_void_p(tin.data_ptr()), C.c_void_p(vbig[G:].data_ptr()),
                              x.shape[0], stream)
    assert rc == 0, eng.lib.rfv_last_error()
    torch.cuda.synchronize()
    assert bool((vbig[:G] == FILL).all()) and bool((vbig[G + n:] == FILL).all()), name
    assert torch.equal(xin.cpu(), x)
    m.eval()
    noise = x.cuda()
    keep = noise.clone()
    s = m.sample(noise, num_steps=3)
    assert torch.equal(noise, keep)
    assert util.rel_l2(s.cpu().numpy(), outs[0]) <= 5e-3, name


def test_unsupported_width_fails_loudly():
    """model_channels = 32 is outside what the kernels tile (64, 128, 256, ...): the engine says so instead of falling back."""
    import rectified_flow_vision_b200 as pkg
    from rectified_flow_vision_b200.engine import RfvError
    m = pkg.BaseFlowModel(image_size=32, model_channels=32, device="cpu")
    m.device = "cuda:0"
    m.to("cuda:0")
    m.eval()
    with pytest.raises(RfvError, match="model_channels"):
        m(torch.randn(2, 3, 32, 32, device="cuda:0"), torch.rand(2, device="cuda:0"))


@pytest.mark.parametrize("name", ["16px_two_levels", "gray_32px", "four_channel_32px"])
def test_training_gradients_vs_oracle(name):
    from oracle import train_oracle as T
    m, spec, x0, x1, t = _build(name)
    S, cin, mc, mult, nres = SHAPES[name]
    P = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    loss_ref, grads = T.loss_and_grads(P, x0, x1, t, model_channels=mc, channel_mult=tuple(mult), num_res_blocks=nres)
    eng = m.velocity_net.train_engine(S, "cuda:0", micro_batch=4)   # 5 rows: micro-batches of 4 + 1
    eng.zero_grad()
    loss = float(eng.train_accumulate(x0.cuda(), x1.cuda(), t.cuda(), dropout_p=0.0, seed=5).item())
    assert abs(loss - loss_ref) <= 5e-3 * loss_ref, (loss, loss_ref)
    gmax = max(float(g.norm()) for g in grads.values())
    for k, gr in grads.items():
        if float(gr.norm()) < 1e-3 * gmax:
            continue
        g = eng.get_grad(k, gr.numel()).cpu().numpy().reshape(gr.shape)
        assert np.isfinite(g).all(), k
        assert util.rel_l2(g, gr.numpy()) <= 5e-2, (name, k, util.rel_l2(g, gr.numpy()))
    # ... and against the reference's own loss.backward() on the same seeds
    gold = np.load(f"{util.GOLD}/shapes_{name}.npz")
    assert np.array_equal(gold["x"], x0.numpy()) and np.array_equal(gold["x1"], x1.numpy())
    assert abs(loss - float(gold["loss"])) <= 5e-3 * float(gold["loss"])
    ref_norm = gold["grad_norm_per_tensor"]
    for k, rn in zip([str(n) for n in gold["names"]], ref_norm):
        if rn < 1e-3 * ref_norm.max():
            continue
        full = eng.get_grad(k, grads[k].numel()).cpu().numpy().reshape(-1)
        assert abs(float(np.linalg.norm(full.astype(np.float64))) - rn) <= 5e-2 * rn, (name, k)
        if gold["grad_sampled/" + k].size >= 64:          # every 241st element: only meaningful on the large tensors
            assert util.rel_l2(full[::241], gold["grad_sampled/" + k]) <= 5e-2, (name, k, util.rel_l2(full[::241], gold["grad_sampled/" + k]))
