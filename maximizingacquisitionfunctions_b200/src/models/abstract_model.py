from ..core import module


class abstract_model(module):
    """Interface of the surrogate models (reference: src/models/abstract_model.py)."""

    def predict(self, *args, **kwargs):
        raise NotImplementedError

    def fit(self, *args, **kwargs):
        raise NotImplementedError
