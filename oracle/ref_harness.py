"""Import harness for the REAL reference (test infrastructure, container-only).

Only usable where /root/reference is mounted (the build container).  It is used by
``oracle/make_golden.py`` to (1) cross-check the in-repo restatement ``oracle/cld_oracle.py``
against the reference's own modules and (2) generate the committed golden vectors under
``tests/golden/``.  Nothing in the product path, in ``-m gpu`` tests, ``smoke()`` or ``bench.py``
imports this file (the reference does not exist on the GPU box).

Recipe follows SURVEY.md Appendix A: MagicMock meta-path finder for the absent third-party roots,
``Tensor.get_device`` shim for CPU, reference ``config.yaml`` -> ``ConfigBase``.
"""
import contextlib
import importlib.abc
import importlib.machinery
import io
import os
import sys
import types
from unittest import mock

REF = os.environ.get("CLD_REFERENCE", "/root/reference")
_STUB_ROOTS = {"trajdata", "pytorch_lightning", "matplotlib", "seaborn", "memory_profiler",
               "colorama", "h5py", "Pplan", "wandb", "pyemd", "shapely", "lightning_fabric"}


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


_installed = False


def available():
    return os.path.isdir(os.path.join(REF, "models")) and os.path.isdir(os.path.join(REF, "src", "tbsim"))


def install():
    """Make `models.*`, `tbsim.*` of the reference importable. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference not mounted at %s" % REF)
    import torch
    sys.meta_path.append(_StubFinder())          # appended: real modules win
    sys.path[:0] = [os.path.join(REF, "src"), REF]
    with contextlib.redirect_stdout(io.StringIO()):
        import this  # noqa: F401  (batch_utils does `from this import d`)
    _orig = torch.Tensor.get_device

    def _get_device(self):
        d = _orig(self)
        return "cpu" if d == -1 else d
    torch.Tensor.get_device = _get_device
    from tbsim.utils.batch_utils import set_global_batch_type
    set_global_batch_type("trajdata")
    _installed = True


def load_configs():
    install()
    import yaml
    from configs.custom_config import dict_to_config, ConfigBase
    with open(os.path.join(REF, "config.yaml")) as f:
        y = yaml.safe_load(f)
    algo = dict_to_config(ConfigBase, y["algo"])
    train = dict_to_config(ConfigBase, y["train"])
    return algo, train, y


def build_models(n_timesteps=100, seed=0):
    """Random-init reference DmModel + VaeModel (eval)."""
    install()
    import torch
    algo, train, _ = load_configs()
    from models.dm.dm_model import DmModel
    from models.vae.vae_model import VaeModel
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=n_timesteps).eval()
        vae = VaeModel(algo, train, {"image": (34, 224, 224)}).eval()
    return dm, vae, algo
