"""Run the denoiser forward alone (for ncu): python tools/run_unet.py [rows] [precision] [iters]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cld_b200 import default_algo_config
from cld_b200.dm_model import DmModel

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
dm = DmModel(default_algo_config(), {"image": (34, 224, 224)}, n_timesteps=100, precision=prec, max_rows=R).cuda()
eng = dm.engine(R)
x, cond = torch.randn(R, 52, 4).cuda(), torch.randn(R, 256).cuda()
t = torch.full((R,), 50).cuda()
for _ in range(2):
    eng.unet_forward(x, cond, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    eng.unet_forward(x, cond, t)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("rows %d %s: %.3f ms/forward, %.1f TFLOP/s" % (R, prec, ms, 119.23e6 * R / ms / 1e9))
