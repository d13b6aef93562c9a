"""Importable alias of the package directory `controllable-latent-diffusion-for-traffic-simulation_b200/`
(a hyphenated directory name cannot be imported directly)."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "controllable-latent-diffusion-for-traffic-simulation_b200")
__path__.insert(0, _PKG_DIR)
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
