"""Host-side base class of the agents: wall-clock timer, candidate sampling in the unit cube (uniform, Sobol, or
distinct subsets of a finite option list) and loss-weighted sub-sampling of options.  Interface and host-RNG
consumption follow src/agents/search_agent.py of the reference (one ``self.rng`` access per draw, same arguments),
so seeded runs pick the same candidates; nothing here runs on the GPU."""
import itertools
import logging
from timeit import default_timer

import numpy as np
from scipy.special import comb

from ..core import module

logger = logging.getLogger(__name__)

SOBOL_MAX_DIM = 40           # the reference's generator stops at 40 dimensions (:103-106)


def _split_shape(shape):
    """[..., set_size, input_dim] -> (number of sets, flattened size of one set) the way :96-101 counts them."""
    shape = list(shape)
    if len(shape) > 2:
        count = int(np.prod(shape[:-2]))
        return count, int(np.prod(shape)) // count
    return int(shape[0]), int(shape[1])


class search_agent(module):
    def __init__(self, *args, timer=None, grids=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.timer = timer or default_timer
        self.grids = {} if grids is None else grids           # cache of sampled grids, keyed by request

    # -- interface ---------------------------------------------------------------------------------------
    def suggest_inputs(self, num_suggest, inputs_old, outputs_old, *args, **kwargs):
        raise NotImplementedError("Search agents must implement suggest_inputs()")

    def suggest_answers(self, num_suggest, inputs_old, outputs_old, in_order=False):
        """The num_suggest best observations (:49-58)."""
        flat = np.ravel(outputs_old)
        best = np.argpartition(flat, num_suggest - 1)[:num_suggest]
        if in_order:
            best = best[np.argsort(flat[best])]
        return inputs_old[best], outputs_old[best]

    # -- candidate sampling ------------------------------------------------------------------------------
    def sample_inputs(self, shape, options=None, use_sobol=False, use_cache=False, dtype=None):
        dtype = self.dtype if dtype is None else dtype
        if not (use_cache and options is None):
            return self._sample_inputs(shape, options, use_sobol, dtype)
        key = (use_sobol, np.dtype(dtype).name, tuple(shape))
        cached = self.grids.get(key)
        if cached is None:
            cached = self.grids[key] = self._sample_inputs(shape, options, use_sobol, dtype)
        return cached

    def _sample_inputs(self, shape, options=None, use_sobol=False, dtype=None):
        dtype = self.dtype if dtype is None else dtype
        shape = list(shape)
        if options is not None:
            points = self._distinct_subsets(np.asarray(options), shape)
        elif use_sobol:
            points = self._sobol(shape)
        else:
            points = self.rng.rand(*shape)
        return np.asarray(points, dtype=dtype)

    def _sobol(self, shape):
        """Low-discrepancy sets with a random skip (:96-112).  The reference calls its vendored Burkardt generator
        ``i4_sobol_generate(width, count, skip).T``; when that module is loaded (compat.launcher registers the
        reference's own file as ``src.third_party.sobol_lib``) it is used, so seeded runs pick the reference's
        candidates.  Without a reference checkout SciPy's unscrambled Sobol sequence fast-forwarded by the same
        random skip stands in (different direction numbers: same discrepancy, different points)."""
        count, width = _split_shape(shape)
        if width > SOBOL_MAX_DIM:
            logger.warning("Sobol sequence not available for >40d spaces; using uniform random values instead")
            return self.rng.rand(*shape)
        skip = int(self.rng.randint(2 ** 13 - 1))
        import sys
        ref = sys.modules.get("src.third_party.sobol_lib")
        if ref is not None and hasattr(ref, "i4_sobol_generate"):
            return np.reshape(ref.i4_sobol_generate(width, count, skip).T, shape)
        import warnings
        from scipy.stats import qmc
        sequence = qmc.Sobol(d=width, scramble=False)
        if skip:
            sequence.fast_forward(skip)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)        # count is not a power of two in general
            points = sequence.random(count)
        return np.reshape(points, shape)

    def _distinct_subsets(self, options, shape):
        """`count` different size-`k` subsets of the option rows (:114-132): rejection sampling while there are more
        subsets than requested, otherwise all of them."""
        assert options.shape[-1] == shape[-1], "Input dimensionality mismatch"
        count, k = int(np.prod(shape[:-2])), shape[-2]
        available = comb(len(options), k, exact=True)
        if available <= count:
            if available != count:
                logger.warning("Requested more unique sets than options allow")
            return np.array(tuple(itertools.combinations(options, k)))
        picked = np.empty([count, k], dtype="int")
        filled = 0
        while filled < count:
            draw = self.rng.choice(len(options), k, replace=False)
            if not (picked[:filled] == draw).all(axis=1).any():
                picked[filled] = draw
                filled += 1
        return options[picked]

    # -- loss-weighted sub-sampling (:136-183) -------------------------------------------------------------
    def sample_propto_loss(self, options, losses, num_samples, top_k=1, eps=None):
        """num_samples options: the top_k lowest losses always, the rest drawn without replacement with probability
        proportional to (worst loss - loss); options no better than the worst only fill up when needed.
        Returns (options[chosen], chosen)."""
        losses = np.asarray(losses)
        eps = np.finfo(losses.dtype).eps if eps is None else eps
        assert len(options) >= len(losses), "Insufficient options provided"
        gap = losses.max() - losses                       # sampling weight before normalisation
        useful = gap >= eps
        remaining = int(useful.sum())

        leaders = None
        if top_k > 0 and remaining > 0:
            lead = min(top_k, remaining)
            leaders = np.argpartition(losses, lead - 1)[:lead]
            gap[leaders] = 0
            useful[leaders] = False
            remaining -= lead
            num_samples -= lead

        if num_samples == 0:
            drawn = np.empty([0], dtype="int")
        elif remaining == num_samples:
            drawn = np.flatnonzero(useful)
        elif remaining < num_samples:
            fillers = self.rng.choice(np.flatnonzero(~useful), num_samples - remaining, replace=False)
            drawn = np.concatenate([np.flatnonzero(useful), fillers])
        else:
            drawn = self.rng.choice(len(options), num_samples, p=self._probabilities(gap, useful, eps), replace=False)
        chosen = drawn if leaders is None else np.concatenate([leaders, drawn])
        return options[chosen], chosen

    @staticmethod
    def _probabilities(gap, useful, eps):
        total = gap.sum()
        if total <= eps:
            return useful / useful.sum()
        p = gap / total
        p[useful] = np.maximum(p[useful], eps)
        return p / p.sum()                                # current NumPy wants p normalised to the last bit
