"""B200-native drone_env.step() hot path.

Public surface:
  drone_env.drones        drop-in for the reference's class (E = 1, float64)
  BatchedDrones           E environments on one GPU, torch tensors in/out
  dist                    env sharding across ranks + all-reduce of episode aggregates
Everything computes through libdronestep.so (C ABI in include/dronestep.h).
"""
from . import _lib, formation  # noqa: F401
from ._lib import DroneStepError  # noqa: F401

__all__ = ["BatchedDrones", "DroneStepError", "drone_env", "formation", "dist"]


def __getattr__(name):
    # torch is imported lazily so that `import package` stays cheap for the ABI tests
    if name == "BatchedDrones":
        from .batched import BatchedDrones
        return BatchedDrones
    if name in ("drone_env", "dist", "batched"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
