"""Myopic acquisition losses of the mirror (same class names as the reference's src.losses; the nonmyopic losses --
entropy search, knowledge gradient -- are outside the accelerated path)."""
from .abstract_loss import abstract_loss
from .confidence_bound import confidence_bound
from .myopic_loss import myopic_loss
from .negative_ei import negative_ei
from .negative_pi import negative_pi
from .simple_regret import simple_regret

__all__ = [cls.__name__ for cls in (abstract_loss, myopic_loss, simple_regret, negative_ei, negative_pi, confidence_bound)]
