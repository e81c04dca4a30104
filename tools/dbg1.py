import sys, os, torch
sys.path.insert(0, os.getcwd())
import rectified_flow_vision_b200 as pkg
from rectified_flow_vision_b200 import engine as E
flags = int(sys.argv[1]); size = int(sys.argv[2])
torch.manual_seed(0)
kw = dict(image_size=size) if size != 32 else dict(image_size=32, model_channels=64, channel_mult=[1, 2], num_res_blocks=1)
m = pkg.BaseFlowModel(device="cuda:0", **kw)
eng = E.Engine(m.velocity_net.arch(), size, torch.device("cuda:0"), micro_batch=4, flags=flags)
eng.sync_weights(m.velocity_net)
x = torch.randn(2, 3, size, size, device="cuda:0"); t = torch.rand(2, device="cuda:0")
eng.set_profiling(True)
try:
    v = eng.velocity(x, t)
    torch.cuda.synchronize()
    print("flags", flags, "size", size, "ok", float(v.abs().mean()))
except Exception as e:
    print("flags", flags, "size", size, "FAILED", str(e)[:200])
