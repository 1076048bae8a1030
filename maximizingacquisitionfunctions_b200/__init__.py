"""B200-native multi-start maximisation of Monte-Carlo batch acquisition functions.

Layout
  csrc/        CUDA kernels + the C ABI (include/maf.h) -> libmaf_b200.so
  _native.py   ctypes binding of the C ABI
  engine.py    NumPy-facing handle on one GPU context
  src/         mirror of the reference's ``src.models / src.losses / src.agents`` call surface
  dist.py      restart sharding across ranks + the single all-gather that picks the arg-min
"""
__version__ = "0.1.0"


def __getattr__(name):
    if name in ("Engine", "acq_params", "MafError"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
