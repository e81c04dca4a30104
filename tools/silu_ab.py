"""A/B of the two SiLU evaluations in gn_apply (RFV_FLAG_SILU_EXP): velocity error against the golden reference output."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tests import util
from rectified_flow_vision_b200 import engine as E

for case in ("small32", "default64"):
    m = util.seeded_model(case, device="cuda:0")
    g = util.golden(case)
    x, t = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()
    for fl in (0, 32768):
        eng = E.Engine(m.velocity_net.arch(), x.shape[-1], torch.device("cuda:0"), micro_batch=4, flags=fl)
        eng.sync_weights(m.velocity_net)
        v = eng.velocity(x, t).cpu().numpy()
        d = np.abs(v - g["v"])
        print(f"{case} flags={fl}: rel-L2 {util.rel_l2(v, g['v']):.3e}  max|err| {d.max():.3e}  max|v| {np.abs(g['v']).max():.3f}")
