from .module import lazy_op, lazy_ref, module

__all__ = ["module", "lazy_op", "lazy_ref"]
