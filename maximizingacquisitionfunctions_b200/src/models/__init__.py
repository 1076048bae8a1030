from .abstract_model import abstract_model
from .gaussian_process import gaussian_process

__all__ = ["abstract_model", "gaussian_process"]
