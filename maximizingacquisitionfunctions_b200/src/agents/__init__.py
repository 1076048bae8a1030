"""Agents of the mirror (same class names as the reference's src.agents)."""
from .bayesian_optimization import bayesian_optimization
from .random_search import random_search
from .search_agent import search_agent

__all__ = [cls.__name__ for cls in (search_agent, random_search, bayesian_optimization)]
