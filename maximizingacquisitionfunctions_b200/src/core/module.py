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


class lazy_ref(object):
    """Placeholder token: the eager counterpart of the reference's cached tf.placeholder
    (module.get_or_create_ref, src/core/module.py:57-70); filled from the feed_dict by lazy_op.run."""

    def __init__(self, group, shape_or_rank=None, dtype=None):
        self.group, self.shape_or_rank, self.dtype = group, shape_or_rank, dtype

    def __repr__(self):
        return "lazy_ref({!r})".format(self.group)


class lazy_op(object):
    """Deferred call: the eager counterpart of a cached graph node (module.get_or_create_node, :44-54).
    ``run(feed_dict)`` substitutes lazy_ref arguments and calls the function -- what
    ``sess.run(op, feed_dict)`` does for scripts/run_bayesopt_incremental.py:150-163."""

    def __init__(self, fn, args=None, kwargs=None):
        self.fn, self.args, self.kwargs = fn, tuple(args or ()), dict(kwargs or {})

    def run(self, feed_dict=None):
        feed_dict = feed_dict or {}
        sub = lambda a: feed_dict[a] if isinstance(a, lazy_ref) else a
        kwargs = {k: sub(v) for k, v in self.kwargs.items() if k != "state_id"}
        return self.fn(*[sub(a) for a in self.args], **kwargs)

    __call__ = run


class module(object):
    def __init__(self, dtype="float64", cache=None, seed=None, name=None):
        self._name = self.__class__.__name__ if name is None else name
        self._dtype = dtype
        self._cache = dict() if cache is None else cache
        if seed is None:
            seed = np.random.choice(2 ** 31 - 1)
        self._seed = np.array([seed], dtype=np.int64)
        self._rng = np.random.RandomState(self._seed)

    def get_or_create_ref(self, group, shape_or_rank=None, dtype=None, cache=None, **kwargs):
        cache = self.cache if cache is None else cache
        key = ("ref", group, str(shape_or_rank))
        if key not in cache:
            cache[key] = lazy_ref(group, shape_or_rank, dtype or self.dtype)
        return cache[key]

    def get_or_create_node(self, group, fn, args=None, cache=None, kwargs=None, stateful=False, **unused):
        return lazy_op(fn, args, kwargs)

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
