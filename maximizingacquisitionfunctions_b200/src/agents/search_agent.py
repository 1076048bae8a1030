"""Base search agent: timer, input sampling, loss-proportional sub-sampling
(reference: src/agents/search_agent.py).  Host-side NumPy; not a GPU target."""
import itertools
import logging
from timeit import default_timer

import numpy as np
from scipy.special import comb as nCr

from ..core import module

logger = logging.getLogger(__name__)


class search_agent(module):
    def __init__(self, *args, timer=None, grids=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.timer = default_timer if timer is None else timer
        self.grids = dict() if grids is None else grids

    def suggest_inputs(self, num_suggest, inputs_old, outputs_old, *args, **kwargs):
        raise NotImplementedError("Search agents must implement suggest_inputs()")

    def suggest_answers(self, num_suggest, inputs_old, outputs_old, in_order=False):
        """Top-k best seen inputs (:49-58)."""
        idx = np.argpartition(np.ravel(outputs_old), num_suggest - 1)[:num_suggest]
        if in_order:
            idx = idx[np.argsort(np.ravel(outputs_old)[idx])]
        return inputs_old[idx], outputs_old[idx]

    def sample_inputs(self, shape, options=None, use_sobol=False, use_cache=False, dtype=None):
        if dtype is None:
            dtype = self.dtype
        if use_cache and options is None:
            key = (use_sobol, np.dtype(dtype).name, tuple(shape))
            if self.grids.get(key, None) is None:
                self.grids[key] = self._sample_inputs(shape, options, use_sobol, dtype)
            return self.grids[key]
        return self._sample_inputs(shape, options, use_sobol, dtype)

    def _sample_inputs(self, shape, options=None, use_sobol=False, dtype=None):
        """Input sets in [0,1]^d (uniform or Sobol) or distinct combinations of given options (:81-133).
        The reference's Sobol points come from its vendored Burkardt generator with a random skip;
        here scipy's unscrambled Sobol sequence with the same random skip is used."""
        if dtype is None:
            dtype = self.dtype
        shape = list(shape)
        if options is None:
            if use_sobol:
                num_sets = int(np.prod(shape[:-2])) if len(shape) > 2 else shape[0]
                set_size = int(np.prod(shape)) // num_sets
                if set_size > 40:
                    logger.warning("Sobol sequence not available for >40d spaces; using uniform random values instead")
                    inputs = self.rng.rand(*shape)
                else:
                    from scipy.stats import qmc
                    skip = int(self.rng.randint(2 ** 13 - 1))
                    eng = qmc.Sobol(d=set_size, scramble=False)
                    if skip:
                        eng.fast_forward(skip)
                    inputs = np.reshape(eng.random(num_sets), shape)
            else:
                inputs = self.rng.rand(*shape)
        else:
            assert options.shape[-1] == shape[-1], "Input dimensionality mismatch"
            num_options = len(options)
            num_sets, set_size = int(np.prod(shape[:-2])), shape[-2]
            num_combs = nCr(num_options, set_size, exact=True)
            if num_combs > num_sets:
                indices = np.empty([num_sets, set_size], dtype="int")
                counter = 0
                while counter < num_sets:
                    new = self.rng.choice(num_options, set_size, replace=False)
                    if not any(np.equal(indices[:counter], new).all(axis=1)):
                        indices[counter] = new
                        counter += 1
                inputs = options[indices]
            else:
                if num_combs != num_sets:
                    logger.warning("Requested more unique sets than options allow")
                inputs = np.array(tuple(itertools.combinations(options, set_size)))
        return np.asarray(inputs, dtype=dtype)

    def sample_propto_loss(self, options, losses, num_samples, top_k=1, eps=None):
        """Sub-sample options with probability proportional to (max loss - loss), the top-k always in
        (:136-183).  Returns (options[indices], indices)."""
        losses = np.asarray(losses)
        if eps is None:
            eps = np.finfo(losses.dtype).eps
        num_options = len(options)
        assert num_options >= len(losses), "Insufficient options provided"
        weights = np.max(losses) - losses
        mask = np.greater_equal(weights, eps)
        nnz = int(np.sum(mask))
        if top_k > 0 and nnz > 0:
            num_top = min(top_k, nnz)
            num_samples = num_samples - num_top
            top = np.argpartition(losses, num_top - 1)[:num_top]
            nnz -= num_top
            weights[top] = 0
            mask[top] = 0
        else:
            top = None
        if num_samples == 0:
            idx = np.empty([0], dtype="int")
        elif nnz == num_samples:
            idx = np.where(mask)[0]
        elif nnz < num_samples:
            nz = np.where(mask)[0]
            rest = self.rng.choice(np.where(np.logical_not(mask))[0], num_samples - nnz, replace=False)
            idx = np.hstack([nz, rest])
        else:
            total = np.sum(weights)
            if total > eps:
                weights /= np.sum(weights)
                weights[mask] = np.maximum(weights[mask], eps)
                weights /= np.sum(weights)      # numpy >= 1.x insists on an exactly normalised p
            else:
                weights = mask / np.sum(mask)
            idx = self.rng.choice(num_options, num_samples, p=weights, replace=False)
        if top is not None:
            idx = np.hstack([top, idx])
        return options[idx], idx
