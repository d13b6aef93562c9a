"""Time the guidance-step kernels alone: python tools/time_guidance.py [scenes] [agents]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cld_b200 import default_algo_config, make_scenes
from cld_b200.dm_model import DmModel
from cld_b200.engine import default_guidance
from cld_b200.vae import VaeModel
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = int(sys.argv[2]) if len(sys.argv) > 2 else 16
R = S * A
algo = default_algo_config()
torch.manual_seed(0)
dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=100, precision="bf16", max_rows=R).cuda()
vae = VaeModel(algo).bind(dm)
aux, batch = make_scenes(S, A, seed=123, dense=True)
eng = dm.engine(R)
scene = eng.make_scene(batch, S, A, 1)
z = torch.randn(R, 52, 4).cuda()
cond, curr = aux["cond_feat"].cuda(), aux["curr_states"].cuda()
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("rows %d: decode+rollout %.3f ms | guidance step (decode+stash, losses, BPTT, update) %.3f ms" % (
    R, t(lambda: eng.decode_rollout(z, cond, curr)), t(lambda: eng.guidance_step(z, cond, curr, scene, default_guidance()))))
