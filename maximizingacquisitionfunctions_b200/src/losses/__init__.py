__all__ = ["abstract_loss", "myopic_loss", "simple_regret", "negative_ei", "negative_pi", "confidence_bound"]
from .abstract_loss import abstract_loss
from .myopic_loss import myopic_loss
from .simple_regret import simple_regret
from .negative_ei import negative_ei
from .negative_pi import negative_pi
from .confidence_bound import confidence_bound
