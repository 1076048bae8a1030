"""Mirror of the reference's ``src`` package surface for the hot path (agents, losses, models,
core), backed by the CUDA library.  ``sess`` arguments are accepted and ignored."""
from . import core, utils, models, losses, agents  # noqa: F401

__all__ = ["core", "utils", "models", "losses", "agents"]
