import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 device (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gold():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLD, name + ".npz")))
    return load


@pytest.fixture(scope="session")
def models_cpu():
    """Random-init parameter containers with the reference's seed-0 initialisation (CPU)."""
    import torch
    from cld_b200 import default_algo_config
    from cld_b200.dm_model import DmModel
    from cld_b200.vae import VaeModel

    def build(n_timesteps=10, **kw):
        algo = default_algo_config()
        torch.manual_seed(0)
        dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=n_timesteps, **kw)
        vae = VaeModel(algo)
        return dm, vae, algo
    return build


def guidance_ext_case(g):
    """Inputs of tests/golden/guidance_ext.npz (oracle/make_golden.py:guidance_ext_golden): scenes, targets, the oracle's
    guidance dict per term and for all six terms together."""
    import torch
    import cld_oracle as O
    from cld_b200.synthetic import make_scenes
    S, A, N = int(g["S"]), int(g["A"]), int(g["N"])
    aux, batch = make_scenes(S, A, horizon=52, seed=int(g["seed"]), dense=True)
    batch["target_pos"] = torch.tensor(g["target_pos"])
    batch["target_speed"] = torch.tensor(g["target_speed"])
    w = dict(zip(("target_pos", "target_speed", "acc_limit", "speed_limit"), [float(v) for v in g["weights"]]))
    base = dict(O.DEFAULT_GUIDANCE, agent_collision=0.0, map_collision=0.0, min_target_time=float(g["min_target_time"]),
                acc_limit_value=float(g["acc_limit_value"]), speed_limit_value=float(g["speed_limit_value"]))
    cfgs = {k: dict(base, **{k: w[k]}) for k in w}
    cfgs["all"] = dict(base, agent_collision=50.0, map_collision=1.0, **w)
    return S, A, N, aux, batch, torch.tensor(g["z"]), cfgs


def guidance_big_case(g):
    """Inputs of tests/golden/guidance_t104.npz: one scene of 64 agents x 8 samples, T = 104 (z regenerated from its seed)."""
    import torch
    from cld_b200.synthetic import make_scenes
    S, A, N, T = int(g["S"]), int(g["A"]), int(g["N"]), int(g["T"])
    aux, batch = make_scenes(S, A, horizon=T, seed=int(g["seed"]), dense=True)
    torch.manual_seed(int(g["z_seed"]))
    z = torch.randn(S * A * N, T, 4)
    assert abs(float(z.double().sum()) - float(g["z_sum"])) < 1e-6, "torch.randn stream changed: regenerate the golden"
    return S, A, N, T, aux, batch, z
