"""Time one PPO minibatch update on the GPU: cld_ppo_grad + cld_adam_step + weight re-pack, R rows.
usage: python tools/time_train.py [rows=128] [iters=20] [fp32|tf32] [graph]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cld_b200 import default_algo_config                     # noqa: E402
from cld_b200.dm_model import DmModel                        # noqa: E402
from cld_b200.trainer import FusedAdam                       # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = sys.argv[3] if len(sys.argv) > 3 else "fp32"
graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
algo = default_algo_config()
torch.manual_seed(0)
dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=16).cuda()
dm.train_precision = mode
for p in dm.model.parameters():
    p.requires_grad_(True)
opt = FusedAdam(dm, lr=1e-4, weight_decay=1e-5)
x1, x0, cond = torch.randn(R, 52, 4).cuda(), torch.randn(R, 52, 4).cuda(), torch.randn(R, 256).cuda()
t = torch.full((R,), 3, dtype=torch.long).cuda()
lp_old, reward = torch.randn(R).cuda(), torch.randn(R).cuda()


if graph:
    from cld_b200.trainer import GraphedPPOStep
    gs = GraphedPPOStep(dm, opt, R)
    for dst, src in zip(gs.buffers(), (x0, x1, lp_old, reward, cond)):
        dst.copy_(src)
    gs.t.copy_(t)


def step():
    if graph:
        gs(0.0)
        return
    dm.ppo_minibatch_grad(x1, x0, cond, t, lp_old, reward, 0.0, 0.2)
    opt.step()


for _ in range(3):
    step()
eng = dm.train_engine(R)
torch.cuda.synchronize()
n0 = eng.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter()
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
w1 = time.perf_counter()
ms = e0.elapsed_time(e1) / iters
flop = 3 * 119.23e6 * R     # forward + data gradient + weight gradient
print("%s%s rows %d: %.3f ms per PPO minibatch update (device), %.3f ms wall, %d launches per update, %.1f TFLOP/s fp32 (3 x forward FLOPs)"
      % (mode, " (CUDA graph)" if graph else "", R, ms, (w1 - w0) * 1e3 / iters, (eng.launch_count() - n0) // iters, flop / ms / 1e9))
