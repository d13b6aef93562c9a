"""cfg2 per-GPU shard: 512 scenes x 32 agents x 8 samples (131 072 rows), 50-step DDIM guided, bf16 mode: time one forward."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cld_b200 import default_algo_config, make_scenes
from cld_b200.dm_model import DmModel
from cld_b200.engine import default_guidance
from cld_b200.vae import VaeModel

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
A, N, T = 32, 8, 52
algo = default_algo_config(num_samp=N)
torch.manual_seed(0)
dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=100, precision="bf16", max_rows=65536).cuda()
dm.stride = 2
VaeModel(algo).bind(dm)
aux, batch = make_scenes(S, A, horizon=T, seed=123, dense=True)
aux = {k: v.cuda() for k, v in aux.items()}
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
R = S * A * N
x_init = torch.randn(R, T, 4, device="cuda")


def run():
    return dm(batch, aux, algo, x_init=x_init, noise=None, sampler="ddim", guidance=default_guidance(), want_indicators=True,
              agents_per_scene=A)


out = run()
torch.cuda.synchronize()
t0 = time.perf_counter()
out = run()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("cfg2 shard: %d scenes x %d agents x %d samples = %d rows: %.3f s per forward, %.1f scenarios/s, %.0f row-steps/s; finite %s; "
      "off-road rows %.3f, collision rows %.3f; peak memory %.1f GB" % (
          S, A, N, R, dt, S / dt, R * 50 / dt, bool(torch.isfinite(out["traj"]).all()), out["offroad"].any(dim=1).float().mean().item(),
          (out["coll"] > 0).float().mean().item(), torch.cuda.max_memory_allocated() / 1e9))
