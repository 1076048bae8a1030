from .module import module

__all__ = ["module"]
