"""Stress: tensor-core LSTM decode + guidance step for many row counts, repeated; checks finiteness and run-to-run determinism."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cld_b200 import default_algo_config, make_scenes
from cld_b200.dm_model import DmModel
from cld_b200.engine import default_guidance
from cld_b200.vae import VaeModel

algo = default_algo_config()
torch.manual_seed(0)
dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=100, precision="bf16", max_rows=8192).cuda()
VaeModel(algo).bind(dm)
eng = dm.engine(8192)
bad = 0
for (S, A) in [(1, 1), (1, 3), (1, 31), (1, 33), (3, 11), (7, 16), (64, 16), (255, 16), (128, 64)]:
    R = S * A
    aux, batch = make_scenes(S, A, seed=S * 100 + A, dense=True)
    scene = eng.make_scene(batch, S, A, 1)
    torch.manual_seed(R)
    z = torch.randn(R, 52, 4).cuda()
    cond, curr = aux["cond_feat"].cuda(), aux["curr_states"].cuda()
    ref = None
    for rep in range(6):
        act, traj = eng.decode_rollout(z, cond, curr)
        zo, g, l = eng.guidance_step(z, cond, curr, scene, default_guidance())
        torch.cuda.synchronize()
        cur = (act.clone(), traj.clone(), zo.clone(), g.clone())
        ok = all(torch.isfinite(t).all().item() for t in cur)
        if ref is None:
            ref = cur
        same = all(torch.equal(a, b) for a, b in zip(cur, ref))
        if not (ok and same):
            bad += 1
            print("R=%d rep %d: finite %s deterministic %s" % (R, rep, ok, same))
    print("R=%5d (S=%d, A=%d): ok" % (R, S, A))
print("stress done, %d problems" % bad)
