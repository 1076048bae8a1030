"""Base class shared by models, losses and agents.

Mirrors the parts of the reference's ``src/core/module.py`` that define behaviour on the hot
path: the dtype, a per-object cache dictionary, nested-dict config merging (:86-98) and -- most
importantly -- the self-reseeding host RNG (:29-42,131-144) that defines "identical base
samples": every access to ``rng`` first advances the seed, ``seed <- max(1, (seed + 1) mod
(2**31 - 1))``, then reseeds one legacy MT19937 ``RandomState`` with it.  The TF graph-node
cache (``get_or_create_node/ref/var``) is a TensorFlow artefact with no counterpart here.
"""
from copy import deepcopy

import numpy as np


class module(object):
    def __init__(self, dtype="float64", cache=None, seed=None, name=None):
        self._name = self.__class__.__name__ if name is None else name
        self._dtype = dtype
        self._cache = dict() if cache is None else cache
        if seed is None:
            seed = np.random.choice(2 ** 31 - 1)
        self._seed = np.array([seed], dtype=np.int64)
        self._rng = np.random.RandomState(self._seed)

    def update_dict(self, src, updates=None, as_copy=False):
        """Recursive merge of `updates` into `src` (dict values merge, everything else overwrites)."""
        if as_copy:
            src = deepcopy(src)
        if updates is None:
            return src
        for key, new in updates.items():
            old = src.get(key, None)
            if isinstance(old, dict) and isinstance(new, dict):
                src[key] = self.update_dict(old, new, as_copy=False)
            else:
                src[key] = new
        return src

    @property
    def cache(self):
        return self._cache

    @property
    def dtype(self):
        return self._dtype

    @dtype.setter
    def dtype(self, new_dtype):
        assert isinstance(new_dtype, str)
        self._dtype = new_dtype

    @property
    def name(self):
        return self._name

    @property
    def seed(self):
        return self._seed[0]

    @seed.setter
    def seed(self, new_seed):
        self._seed[:] = new_seed

    @property
    def rng(self):
        self.seed = max(1, np.mod(self.seed + 1, 2 ** 31 - 1))
        self._rng.seed(self.seed)
        return self._rng
