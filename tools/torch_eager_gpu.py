"""The oracle's plain-PyTorch sampler (the reference's algorithm, ATen / cuDNN / cuBLAS kernels + autograd) run on the B200
itself: the "library kernels on the same box" comparison of SURVEY.md sec. 8(d).  python tools/torch_eager_gpu.py [scenes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import cld_oracle as O
from cld_b200 import default_algo_config, make_scenes
from cld_b200.dm_model import DmModel
from cld_b200.vae import VaeModel

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
A, N, T, NT, STRIDE = 16, 1, 52, 100, 2
dev = torch.device("cuda")
algo = default_algo_config()
torch.manual_seed(0)
dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=NT)
vae = VaeModel(algo)
unet_sd = {k: v.detach().to(dev) for k, v in dm.model.state_dict().items()}
dec_sd = {k: v.detach().to(dev) for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
sched = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in O.make_schedule(NT).items()} if isinstance(O.make_schedule(NT), dict) else O.make_schedule(NT)
aux, batch = make_scenes(S, A, horizon=T, seed=123, dense=True)
aux = {k: v.to(dev) for k, v in aux.items()}
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
torch.set_default_device(dev)      # the oracle creates its constants on the default device
torch.manual_seed(7)
x_init = torch.randn(S * A * N, T, 4, device=dev)
gd = dict(dec_sd=dec_sd, cond=aux["cond_feat"], curr=aux["curr_states"], batch=batch, A=A, N=N, cfg=O.DEFAULT_GUIDANCE)


def one():
    with torch.no_grad():
        out = O.sample(unet_sd, sched, aux["cond_feat"], x_init, None, NT, STRIDE, "ddim", guidance=gd)
        traj, _ = O.decode_rollout(dec_sd, out["pred_traj"], aux["cond_feat"], aux["curr_states"])
        O.indicators(traj[..., :2], batch)
    torch.cuda.synchronize()


one()
t0 = time.perf_counter()
one()
dt = time.perf_counter() - t0
print("PyTorch eager on the GPU (oracle port, fp32): %d scenes in %.2f s = %.2f scenarios/s" % (S, dt, S / dt))
