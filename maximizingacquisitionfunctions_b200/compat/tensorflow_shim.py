"""The handful of TensorFlow 1.x symbols the reference's *scripts* touch directly
(scripts/run_bayesopt.py:18,417,419,473-474,566): a Session context manager whose ``run`` is a
no-op, ``set_random_seed``, ``get_collection`` / ``GraphKeys`` / ``variables_initializer`` (model
variables are plain NumPy state here) and ``logging.set_verbosity``.  Nothing numerical."""
import types


class Session(object):
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches=None, feed_dict=None, **kwargs):
        if hasattr(fetches, "run"):          # core.lazy_op (the eager stand-in for a graph node)
            return fetches.run(feed_dict)
        return None

    def close(self):
        pass


class GraphKeys(object):
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "trainable_variables"


def set_random_seed(seed):
    return None


def get_collection(key, scope=None):
    return []


def variables_initializer(var_list, name="init"):
    return None


def get_default_session():
    return None


logging = types.SimpleNamespace(set_verbosity=lambda level: None, ERROR="ERROR", INFO="INFO", WARN="WARN",
                                DEBUG="DEBUG", FATAL="FATAL")
float32, float64 = "float32", "float64"


def as_dtype(dtype):
    return types.SimpleNamespace(name=str(dtype), as_numpy_dtype=__import__("numpy").dtype(str(dtype)).type)
