"""CPU: the training-step oracle (oracle/train_oracle.py) against golden vectors produced by the unmodified
reference's own step body (oracle/make_golden_train.py -> tests/golden/train_*.npz)."""
import numpy as np
import pytest
import torch

from tests import util


def _inputs(case):
    g = util.golden(case)
    return torch.from_numpy(g["x"]), torch.from_numpy(g["x1"]), torch.from_numpy(g["t"])


def _arch(case):
    kw = util.manifest()["cases"][case]["kwargs"]
    return dict(model_channels=kw.get("model_channels", 64), channel_mult=tuple(kw.get("channel_mult", [1, 2, 4])),
                num_res_blocks=kw.get("num_res_blocks", 2))


@pytest.mark.parametrize("case", ["small32", "default64"])
def test_train_oracle_matches_reference_step(case):
    from oracle import train_oracle as T
    tg = np.load(f"{util.GOLD}/train_{case}.npz")
    m = util.seeded_model(case)
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    names = [str(n) for n in tg["names"]]
    assert ["velocity_net." + k for k, _ in m.velocity_net.named_parameters()] == names
    x0, x1, t = _inputs(case)
    lr = float(tg["lr"])
    state, p0 = {}, {k: v.clone() for k, v in P.items()}
    for step in range(len(tg["losses"])):
        loss, grads = T.loss_and_grads(P, x0, x1, t, **_arch(case))
        assert abs(loss - tg["losses"][step]) <= 2e-4 * abs(tg["losses"][step]), (step, loss)
        if step == 0:
            norms = np.array([float(grads[k].norm()) for k in names])
            np.testing.assert_allclose(norms, tg["grad_norm_per_tensor"], rtol=2e-3, atol=1e-7)
            for key in tg.files:
                if key.startswith("grad_full/"):
                    assert util.rel_l2(grads[key[10:]].numpy(), tg[key]) <= 1e-3, key
                if key.startswith("grad_sampled/"):
                    got = grads[key[13:]].numpy().reshape(-1)[::int(tg["stride"])]
                    assert util.rel_l2(got, tg[key]) <= 1e-3, key
        total = T.adamw_step(P, grads, state, step + 1, lr=lr)
        assert abs(total - tg["grad_norms_total"][step]) <= 1e-3 * tg["grad_norms_total"][step]
    upd = np.array([float((P[k] - p0[k]).norm()) for k in names])
    np.testing.assert_allclose(upd, tg["update_norm"], rtol=2e-2, atol=1e-6)
    for key in tg.files:
        if key.startswith("update_full/"):
            assert util.rel_l2((P[key[12:]] - p0[key[12:]]).numpy(), tg[key]) <= 2e-2, key


def test_cosine_lr_matches_torch_scheduler():
    from rectified_flow_vision_b200.training import cosine_lr
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=3e-4)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, 7)
    for e in range(7):
        assert abs(opt.param_groups[0]["lr"] - cosine_lr(3e-4, e, 7)) < 1e-12
        opt.step()
        sch.step()
