"""TEST INFRASTRUCTURE ONLY -- shimmed import of the *unmodified* reference.

Imports ``/root/reference/drone_env.py`` (and, on request, ``SAC_agents`` /
``utils``) in THIS build container so that

  * ``oracle/make_golden.py`` can record golden input/output vectors of the
    reference's own ``drones.step()`` (committed under ``tests/golden/``), and
  * the restatements in ``oracle/np_oracle.py`` / ``oracle/drone_oracle.c`` can
    be validated against the live reference.

The reference files are never modified or copied.  The shims only make the
import succeed on this image:

  1. empty stub modules for ``matplotlib``/``IPython`` (plotting is off-path),
  2. ``np.infty = np.inf`` (used at reference ``drone_env.py:142``; removed in
     NumPy 2),
  3. stub ``turtle`` and ``autograd`` for ``utils.py:2,4-5`` (policy side).

``/root/reference`` does not exist on the GPU box.  What travels there is the git-ignored
staging copy ``oracle/_ref/`` written by ``oracle/make_ref.py`` (unmodified files); this module
falls back to it, so that ``bench.py``'s CPU-baseline legs and the one ``-m gpu`` test that drives
the reference's own agents can run the real reference beside the GPU.  The product package never
imports this module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("DRONESTEP_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "drone_env.py")) and \
        os.path.isfile(os.path.join(_STAGED, "drone_env.py")):
    REFERENCE_ROOT = _STAGED          # GPU box: the staged, unmodified copy (oracle/make_ref.py)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "drone_env.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def _install_shims() -> None:
    import numpy as np

    if not hasattr(np, "infty"):
        np.infty = np.inf  # reference drone_env.py:142
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.markers",
                 "matplotlib.animation", "IPython", "IPython.display"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "markers", "animation"):
        if not hasattr(mpl, sub):
            setattr(mpl, sub, sys.modules["matplotlib." + sub])
    ipy = sys.modules["IPython"]
    if not hasattr(ipy, "display"):
        ipy.display = sys.modules["IPython.display"]
    # policy side (utils.py:2 `from turtle import forward`, :4-5 autograd)
    try:
        importlib.import_module("turtle")
    except Exception:
        _stub("turtle", forward=None)
    try:
        importlib.import_module("autograd")
    except Exception:
        import numpy as _np
        ag = _stub("autograd", grad=lambda f: (lambda *a, **k: None))
        agnp = _stub("autograd.numpy")
        agnp.__dict__.update({k: getattr(_np, k) for k in dir(_np) if not k.startswith("__")})
        ag.numpy = agnp


def import_reference(module: str = "drone_env"):
    """Return the reference module ``module`` imported under a private name.

    The module is registered as ``_reference_<module>`` so it never shadows
    this repository's own ``drone_env`` drop-in.
    """
    if not reference_available():
        raise RuntimeError(
            f"reference checkout not found at {REFERENCE_ROOT}; golden vectors "
            "under tests/golden/ are the portable stand-in")
    _install_shims()
    private = "_reference_" + module
    if private in sys.modules:
        return sys.modules[private]
    path = os.path.join(REFERENCE_ROOT, module + ".py")
    spec = importlib.util.spec_from_file_location(private, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[private] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_policy_stack():
    """Import reference ``utils`` and ``SAC_agents`` under their own names.

    The pickled policies (``models/**/*.pth``) reference classes as
    ``utils.DiscreteSoftmaxNN`` etc., so these two must be importable by their
    real names while loading.  ``drone_env`` is NOT imported by either.
    """
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)  # append: never shadow repo modules
    utils = importlib.import_module("utils")
    sac = importlib.import_module("SAC_agents")
    return utils, sac
