"""Run the reference's scripts unmodified on the B200 path:

    python -m maximizingacquisitionfunctions_b200.compat.launcher /path/to/reference scripts/run_bayesopt.py \
        --task hartmann3 --loss negative_ei --subroutine sgd --parallelism 2 ...

`install(reference_root)` arranges ``sys.modules`` so that, when a reference script does
``import src, tasks`` / ``import tensorflow as tf`` / ``import matplotlib.pyplot``:
  * ``src`` (and ``src.agents / src.losses / src.models / src.core / src.utils``) is this package's mirror;
  * ``src.third_party.sobol_lib`` is the reference's own pure-NumPy Sobol generator, loaded from its file;
  * ``tensorflow`` is compat.tensorflow_shim unless a real TensorFlow is importable; the reference's
    ``tasks`` / ``helpers`` packages (NumPy objective functions, job manager, JSON storage) are used as they are;
  * ``matplotlib`` is stubbed if absent (plotting is out of scope);
  * ``yaml.load`` gets the Loader default that PyYAML >= 6 made mandatory (the script predates it).
"""
import importlib
import importlib.util
import os
import runpy
import sys
import types
from unittest.mock import MagicMock


def install(reference_root):
    reference_root = os.path.abspath(reference_root)
    from .. import src as mirror
    from . import tensorflow_shim

    try:
        import tensorflow  # noqa: F401  (a real TF, if any, is only used by the reference's helpers)
    except Exception:
        tf = types.ModuleType("tensorflow")
        tf.__dict__.update({k: v for k, v in vars(tensorflow_shim).items() if not k.startswith("__")})
        contrib, opt = types.ModuleType("tensorflow.contrib"), types.ModuleType("tensorflow.contrib.opt")
        opt.ScipyOptimizerInterface = None
        contrib.opt = opt
        tf.contrib = contrib
        sys.modules.update({"tensorflow": tf, "tensorflow.contrib": contrib, "tensorflow.contrib.opt": opt})
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors", "matplotlib.cm",
                     "matplotlib.patches", "matplotlib.lines"):
            sys.modules.setdefault(name, MagicMock())

    # src -> the mirror
    sys.modules["src"] = mirror
    for sub in ("agents", "losses", "models", "core", "utils"):
        sys.modules["src." + sub] = getattr(mirror, sub)
    spec = importlib.util.spec_from_file_location(
        "src.third_party.sobol_lib", os.path.join(reference_root, "src", "third_party", "sobol_lib.py"))
    sobol = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sobol)
    third = types.ModuleType("src.third_party")
    third.sobol_lib = sobol
    misc = types.ModuleType("src.misc")
    misc.random_feature_mapping = MagicMock()       # rfgp synthetic task only (out of scope)
    mirror.third_party, mirror.misc = third, misc
    sys.modules.update({"src.third_party": third, "src.third_party.sobol_lib": sobol, "src.misc": misc,
                        "src.misc.random_feature_mapping": misc.random_feature_mapping})

    import yaml
    if not getattr(yaml.load, "_maf_patched", False):
        _load = yaml.load

        def load(stream, Loader=None):
            return _load(stream, Loader=Loader or yaml.SafeLoader)
        load._maf_patched = True
        yaml.load = load

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)      # tasks/, helpers/
    return mirror


def load_script(reference_root, script="scripts/run_bayesopt.py", name="maf_reference_script"):
    """Import a reference script as a module (without running its __main__ block)."""
    install(reference_root)
    path = os.path.join(os.path.abspath(reference_root), script)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit(__doc__)
    root, script = argv[0], argv[1]
    install(root)
    sys.argv = [os.path.join(root, script)] + argv[2:]
    os.chdir(root)                               # the scripts read experiments/defaults/subroutines.yaml relatively
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
