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
