"""Experiment: two half-batches on two CUDA streams (two engines) against one engine on the full batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cld_b200 import default_algo_config, make_scenes
from cld_b200.dm_model import DmModel
from cld_b200.engine import default_guidance
from cld_b200.vae import VaeModel

S, A, T = 256, 16, 52
algo = default_algo_config(num_samp=1)


def build(S_):
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=100, precision="bf16", max_rows=S_ * A).cuda()
    dm.stride = 2
    VaeModel(algo).bind(dm)
    aux, batch = make_scenes(S_, A, horizon=T, seed=123, dense=True)
    aux = {k: v.cuda() for k, v in aux.items()}
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    x_init = torch.randn(S_ * A, T, 4, device="cuda")
    return dm, aux, batch, x_init


def run(dm, aux, batch, x_init):
    return dm(batch, aux, algo, x_init=x_init, noise=None, sampler="ddim", guidance=default_guidance(), want_indicators=True, agents_per_scene=A)


full = build(S)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
halves = [build(S // lanes) for _ in range(lanes)]
streams = [torch.cuda.Stream() for _ in range(lanes)]


def t(fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def lanes_fn():
    for st, h in zip(streams, halves):
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            run(*h)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)


a = t(lambda: run(*full))
b = t(lanes_fn)
print("one engine %d scenes: %.2f ms (%.0f scen/s) | %d lanes: %.2f ms (%.0f scen/s)" % (S, a, S / a * 1e3, lanes, b, S / b * 1e3))
