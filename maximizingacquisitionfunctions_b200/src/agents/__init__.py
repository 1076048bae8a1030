__all__ = ["search_agent", "random_search", "bayesian_optimization"]
from .search_agent import search_agent
from .random_search import random_search
from .bayesian_optimization import bayesian_optimization
