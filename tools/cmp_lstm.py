"""Tensor-core LSTM decoder (bf16-precision mode) against the fp32 SIMT kernels on the same inputs:
python tools/cmp_lstm.py [scenes] [agents]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cld_b200 import default_algo_config, make_scenes
from cld_b200.dm_model import DmModel
from cld_b200.engine import default_guidance
from cld_b200.vae import VaeModel

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
A = int(sys.argv[2]) if len(sys.argv) > 2 else 16
R = S * A
algo = default_algo_config()


def build(simt):
    if simt:
        os.environ["CLD_LSTM_SIMT"] = "1"
    else:
        os.environ.pop("CLD_LSTM_SIMT", None)
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=100, precision="bf16", max_rows=R).cuda()
    vae = VaeModel(algo).bind(dm)
    eng = dm.engine(R)
    return dm, vae, eng


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


aux, batch = make_scenes(S, A, seed=123, dense=True)
torch.manual_seed(5)
z = torch.randn(R, 52, 4).cuda()
cond, curr = aux["cond_feat"].cuda(), aux["curr_states"].cuda()
dm_s, vae_s, eng_s = build(True)
dm_t, vae_t, eng_t = build(False)
out_s = eng_s.decode_rollout(z, cond, curr)
out_t = eng_t.decode_rollout(z, cond, curr)
torch.cuda.synchronize()
names = ["act", "traj"] if isinstance(out_s, (tuple, list)) else ["out"]
outs_s = out_s if isinstance(out_s, (tuple, list)) else [out_s]
outs_t = out_t if isinstance(out_t, (tuple, list)) else [out_t]
for n, a, b in zip(names, outs_t, outs_s):
    print("decode %-5s rel(tc vs simt) = %.3e  max abs %.3e  finite %s" % (n, rel(a, b), (a - b).abs().max().item(), bool(torch.isfinite(a).all())))
scene_s = eng_s.make_scene(batch, S, A, 1)
scene_t = eng_t.make_scene(batch, S, A, 1)
zs, gs, ls = eng_s.guidance_step(z, cond, curr, scene_s, default_guidance())
zt, gt, lt = eng_t.guidance_step(z, cond, curr, scene_t, default_guidance())
torch.cuda.synchronize()
nz = gs != 0
print("guidance: rel(loss) %.3e %.3e | rel(grad) %.3e | sign agreement %.6f | zero-set agreement %.6f | rel(z_out) %.3e" % (
    rel(lt[0], ls[0]), rel(lt[1], ls[1]), rel(gt, gs), (torch.sign(gt)[nz] == torch.sign(gs)[nz]).float().mean().item(),
    ((gt == 0) == (gs == 0)).float().mean().item(), rel(zt, zs)))


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for nm, eng, sc in (("simt", eng_s, scene_s), ("tc", eng_t, scene_t)):
    print("%-4s rows %d: decode+rollout %.3f ms | guidance step %.3f ms" % (
        nm, R, t(lambda: eng.decode_rollout(z, cond, curr)), t(lambda: eng.guidance_step(z, cond, curr, sc, default_guidance()))))
