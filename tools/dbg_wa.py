"""Debug helper: one Euler step at a given micro-batch / flags, prints elapsed time (run under `timeout`)."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rectified_flow_vision_b200 as pkg
from rectified_flow_vision_b200 import engine as E
mb, flags, size = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 64
torch.manual_seed(0)
m = pkg.BaseFlowModel(image_size=size, device="cuda:0")
eng = E.Engine(m.velocity_net.arch(), size, torch.device("cuda:0"), micro_batch=mb, flags=flags)
eng.sync_weights(m.velocity_net)
x = torch.randn(mb, 3, size, size, device="cuda:0")
torch.cuda.synchronize()
t0 = time.time()
eng.euler_sample(x, 1)
torch.cuda.synchronize()
print(f"mb={mb} flags={flags} size={size}: ok {time.time() - t0:.3f}s, |x|={x.float().norm().item():.4f}", flush=True)
