"""Bayesian-optimisation agent: multi-start maximisation of the acquisition function
(reference: src/agents/bayesian_optimization.py).

Same method names, arguments, defaults and return shapes as the reference; evaluation is eager
(NumPy in / NumPy out through the CUDA library) so the TensorFlow plumbing -- placeholders,
variables, cached graph nodes, ``sess.run`` -- is replaced by a small ``pool_loss`` handle that owns
the trainable candidate sets.  ``sess`` arguments are accepted and ignored.

Supported inner optimisers: ``_optimize_inputs_sgd`` (TF-1 Adam + clip, on the device),
``_optimize_inputs_scipy`` (joint L-BFGS-B, SciPy on the host driving device value+gradient),
``_optimize_inputs_random``, ``_optimize_inputs_halving``.  The other derivative-free alternatives of the reference (CMA-ES,
NLopt) are outside the accelerated path.
"""
import logging

import numpy as np
from scipy.optimize import minimize as _scipy_minimize

from ... import dist as mdist
from ..utils import atleast_pools, join_pools
from .search_agent import search_agent

logger = logging.getLogger(__name__)


def _choice_without_replacement(rng, p, cdf0, size, raw0=None):
    """`rng.choice(len(p), size, p=p, replace=False)` of NumPy's legacy RandomState, draw for draw (numpy/random/mtrand.pyx:
    uniforms -> searchsorted on the cumulative weights, first occurrences kept, found indices zeroed and the rest redrawn),
    with the cumulative weights of the untouched `p` supplied by the caller (`cdf0` normalised, `raw0` the raw running sum).
    `rng` is ONE access of module.rng.

    Redraw rounds: NumPy re-sums all weights with the found indices zeroed (O(len(p)) per round: milliseconds at 2^20
    options, and with weights concentrated on few options almost every draw of several indices collides).  Here the
    zeroed running sum is raw0 minus a step function with one step per found index, so the position of a uniform is found
    with a few binary searches; if a uniform lands within 1e-9 of a cell boundary (where the two summation orders could
    disagree) the round is redone NumPy's way, so the outcome is always NumPy's."""
    found = np.zeros(size, dtype=np.int64)
    n_uniq, pp = 0, None
    while n_uniq < size:
        x = rng.rand(size - n_uniq)
        if n_uniq == 0:
            new = cdf0.searchsorted(x, side="right")
        else:
            new = _redraw_positions(x, raw0, p, found[:n_uniq]) if raw0 is not None else None
            if new is None:                                  # exact route
                if pp is None:
                    pp = p.copy()
                pp[found[:n_uniq]] = 0
                cdf = np.cumsum(pp)
                cdf /= cdf[-1]
                new = cdf.searchsorted(x, side="right")
        _, first = np.unique(new, return_index=True)
        first.sort()
        new = new.take(first)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found


def _redraw_positions(x, raw0, p, found):
    """searchsorted(cumsum(p with p[found] = 0) / total, x, 'right') from the running sum raw0 of the untouched p; None if any
    uniform is too close to a cell boundary to be sure of NumPy's answer."""
    fs = np.sort(found)
    cum = np.concatenate([[0.0], np.cumsum(p[fs])])         # weight removed before region r
    total = raw0[-1] - cum[-1]
    n = len(raw0)
    out = np.empty(len(x), dtype=np.int64)
    for k, xk in enumerate(x):
        target = xk * total
        ans = n
        for r in range(len(fs) + 1):
            hi = fs[r] if r < len(fs) else n
            i = int(raw0.searchsorted(target + cum[r], side="right"))
            if i < hi:
                ans = i
                break
        if ans >= n:
            return None
        adj = cum[int(np.searchsorted(fs, ans, side="right"))]
        above = raw0[ans] - adj - target
        below = target - ((raw0[ans - 1] - cum[int(np.searchsorted(fs, ans - 1, side="right"))]) if ans > 0 else 0.0)
        if min(above, below) < 1e-9 * total:
            return None
        out[k] = ans
    return out


class pool_loss(object):
    """Stand-in for the reference's (losses_op, tf.Variable inputs_new) pair: the trainable candidate
    sets ``value`` [S, q_new, d] plus everything needed to evaluate their losses (and gradient)."""

    def __init__(self, loss_fn, model, inputs_new, inputs_old, outputs_old, inputs_pend, loss_kwargs):
        self.loss_fn, self.model = loss_fn, model
        self.value = np.array(atleast_pools(inputs_new), dtype=np.float64)
        self.inputs_old, self.outputs_old, self.inputs_pend = inputs_old, outputs_old, inputs_pend
        self.loss_kwargs = {k: v for k, v in dict(loss_kwargs or {}).items() if k != "num_fantasies"}

    @property
    def shape(self):
        return self.value.shape

    def assign(self, new_value):
        self.value = np.array(atleast_pools(new_value), dtype=np.float64)

    def pools(self, inputs_new=None):
        return join_pools(self.value if inputs_new is None else inputs_new, self.inputs_pend)

    def evaluate(self, feed_dict=None, inputs_new=None, need_grad=False):
        """losses (reference shapes: [S] for MC, [S,1,1] closed form), extras and optionally the gradient."""
        pools, p = self.pools(inputs_new)
        rvs = None if feed_dict is None else feed_dict.get("fantasy_rvs", None)
        if pools.shape[-2] == 1:
            rvs = None                      # closed form (myopic_loss.evaluate :84-90)
        out = self.loss_fn.evaluate(pools, self.inputs_old, self.outputs_old, fantasy_rvs=rvs, model=self.model,
                                    num_pending=p, return_grad=need_grad, **self.loss_kwargs)
        return out


class bayesian_optimization(search_agent):
    def __init__(self, loss_fn, model, num_options=2 ** 16, risk=0, configs=None, **kwargs):
        super().__init__(**kwargs)
        self.loss_fn = loss_fn
        self.model = model
        self.num_options = num_options
        self.risk = risk
        self.modules = dict()
        self.configs = self.update_dict({
            "random": {"eval_limit": 2 ** 16},
            "sgd": {"method": "train.AdamOptimizer", "learning_rate": 25e-3, "use_averaging": False,
                    "step_limit": 1000, "batch_size": 128},
            "scipy": {"method": "L-BFGS-B", "maxiter": 1000, "ftol": 1e6 * np.finfo("float64").eps},
            "nlopt": {"method": "GN_DIRECT", "maxfun": 10000},
            "cmaes": {"method": "CMAEvolutionStrategy", "step_limit": 10000, "stop_early": False,
                      "sigma0": 0.5, "bounds": [0, 1]},
            "halving": {"eval_limit": 2 ** 20, "expon_step": 1, "min_samples": 128},
        }, configs)

    # ------------------------------------------------------------------------------------------
    # suggest_inputs (:105-168)
    # ------------------------------------------------------------------------------------------
    def suggest_inputs(self, num_suggest, inputs_old, outputs_old, sess=None, inputs_pend=None, num_options=None,
                       num_to_optimize=32, subroutine=None, options=None, use_sobol=False, allow_duplicates=True,
                       time_limit=np.inf, dtype=None, **kwargs):
        if num_options is None:
            num_options = self.num_options
        if dtype is None:
            dtype = self.dtype
        input_dim = inputs_old.shape[-1]
        start_time = self.timer()
        if subroutine not in [self._optimize_inputs_random, self._optimize_inputs_halving]:
            inputs_new = self.sample_starts(num_to_optimize, num_suggest, sess, inputs_old, outputs_old,
                                            inputs_pend, time_limit=time_limit / 10)
        else:
            inputs_new = self.sample_inputs([num_options, num_suggest, input_dim], options=options,
                                            use_sobol=use_sobol, dtype=dtype)

        # Restarts are independent (the loss is a per-restart vector, sharded_fn maps over them, src/misc/sharded_fn.py:
        # 106-132): under torchrun every rank draws the same starts from the same host-RNG stream, optimises and
        # re-evaluates its contiguous slice with the same base samples, and one all-gather of the per-rank
        # (loss, index, set) records picks the global arg-min with np.argmin's tie-break (:152).  The joint L-BFGS-B
        # subroutine couples the restarts of one rank only (its line search is shared), the others not at all.
        rank, world_size = mdist.active()
        lo, hi = mdist.shard_range(len(inputs_new), rank, world_size)
        if world_size > 1:
            inputs_new = inputs_new[lo:hi]

        if len(inputs_new):
            nodes, loss_seq = self.optimize_inputs(
                sess=sess, inputs_new=inputs_new, inputs_old=inputs_old, outputs_old=outputs_old,
                inputs_pend=inputs_pend, time_limit=time_limit - (self.timer() - start_time),
                subroutine=subroutine, **kwargs)
            inputs_new = np.array(nodes["inputs_new"].value)
        # re-evaluate the optimised sets with fresh base samples of size loss_fn.num_fantasies
        if world_size > 1:
            mdist.sync_seed(self.loss_fn)                   # same fresh base samples on every rank
            mdist.sync_seed(self)
        losses = np.empty([0])
        if len(inputs_new):
            losses_op, feed_dict, nodes = self.appraise_inputs(
                sess=sess, inputs_new=inputs_new, inputs_old=inputs_old, outputs_old=outputs_old,
                inputs_pend=inputs_pend, nodes=nodes, **{**kwargs, "inputs_as_var": False})[:3]
            losses = np.atleast_1d(np.squeeze(losses_op))

        if world_size > 1:
            eng = self.model.bind(inputs_old, outputs_old) if (np.ndim(inputs_old) == 2 and len(inputs_new)) else None
            num_pending = 0 if inputs_pend is None else len(inputs_pend)
            best_loss, self.last_argmin, suggestions = mdist.global_best(
                eng, losses, inputs_new, lo, pool_shape=(num_pending + num_suggest, input_dim))
            suggestions = np.array(suggestions)
            losses, argmin = np.array([best_loss]), 0
            mdist.sync_seed(self.loss_fn)                   # ranks with an empty slice drew nothing: realign the streams
            mdist.sync_seed(self)
        else:
            argmin = int(np.argmin(losses))
            self.last_argmin = argmin
            suggestions = inputs_new[argmin]
        if not allow_duplicates:
            exclude = np.vstack([a for a in (inputs_old, inputs_pend) if a is not None])
            replaced = 0
            for k, suggestion in enumerate(suggestions):
                was_dup = False
                while np.equal(exclude, suggestion).all(axis=1).any():
                    suggestion = self.sample_inputs([input_dim], options=options)
                    was_dup = True
                suggestions[k] = suggestion
                exclude = np.vstack([exclude, suggestion[None]])
                replaced += int(was_dup)
            if replaced:
                logger.info("Replaced {:d} duplicate input(s)".format(replaced))
        return suggestions, losses[argmin]

    # ------------------------------------------------------------------------------------------
    # appraise_inputs (:171-231)
    # ------------------------------------------------------------------------------------------
    def appraise_inputs(self, sess, inputs_new, inputs_old, outputs_old, inputs_pend=None, feed_dict=None,
                        model=None, loss_fn=None, nodes=None, loss_kwargs=None, predict_kwargs=None,
                        inputs_as_var=False, **kwargs):
        """Losses of candidate sets `inputs_new` [S, q_new, d] given observations and pending inputs.
        Returns (losses_op, feed_dict, nodes, extras): with ``inputs_as_var`` the first item is a
        ``pool_loss`` handle (the trainable variable + its loss), otherwise the evaluated losses."""
        if model is None:
            model = self.model
        if loss_fn is None:
            loss_fn = self.loss_fn
        feed_dict = dict() if feed_dict is None else dict(feed_dict)
        loss_kwargs = dict() if loss_kwargs is None else dict(loss_kwargs)
        if np.ndim(inputs_old) == 4 and inputs_pend is not None and len(inputs_pend):
            raise NotImplementedError("pending inputs together with fantasy-padded observations")
        nodes = self.prepare(sess, inputs_new, inputs_old, outputs_old, inputs_pend, feed_dict, nodes,
                             inputs_as_var)[0]
        # the training factor is cached on the device per (hyper-parameters, X0, Y0) (runtime.bind)
        model.bind(inputs_old, outputs_old)

        num_pending = 0 if inputs_pend is None else np.shape(inputs_pend)[-2]
        parallelism = num_pending + np.shape(inputs_new)[1]
        placeholders, prep_feed = loss_fn.prepare(sess, inputs_old, outputs_old, parallelism, self.model, **loss_kwargs)
        feed_dict = {**prep_feed, **feed_dict}       # user-specified terms take precedence

        handle = pool_loss(loss_fn, model, nodes["inputs_new_src"], inputs_old, outputs_old, inputs_pend, loss_kwargs)
        if inputs_as_var:
            nodes["inputs_new"] = handle
            return handle, feed_dict, nodes, None
        if handle.value.shape[0] == 0:
            nodes["inputs_new"] = handle
            return handle, feed_dict, nodes, None     # empty probe (sample_starts builds its loss this way)
        losses, extras = handle.evaluate(feed_dict)[:2]
        nodes["inputs_new"] = handle
        return losses, feed_dict, nodes, extras

    # ------------------------------------------------------------------------------------------
    # update (:234-274)
    # ------------------------------------------------------------------------------------------
    def update(self, sess, inputs_old, outputs_old, model=None, var_list=None, loss_fn=None, reset=True,
               fit_kwargs=None, loss_kwargs=None, verbose=2, **kwargs):
        if model is None:
            model = self.model
        if loss_fn is None:
            loss_fn = self.loss_fn
        tic = self.timer()
        if reset:
            model.reset_state()
        model.fit(sess, inputs_old, outputs_old, **(fit_kwargs or {}))
        var_values = [model.state[k] for k in ("mean", "log_lenscales", "log_amplitude", "log_noise")]
        if verbose > 0:
            logger.info("Fit model in {:.2e}s".format(self.timer() - tic))
        loss_fn.update(sess, inputs_old, outputs_old, **{"model": model, **(loss_kwargs or {})})
        return var_values

    # ------------------------------------------------------------------------------------------
    # optimize_inputs (:277-334)
    # ------------------------------------------------------------------------------------------
    def optimize_inputs(self, sess, inputs_new, inputs_old, outputs_old, inputs_pend=None, var_list=None,
                        loss_fn=None, batch_generator=None, config=None, subroutine=None, loss_kwargs=None,
                        time_limit=np.inf, **kwargs):
        start_time = self.timer()
        if loss_fn is None:
            loss_fn = self.loss_fn
        if loss_kwargs is None:
            loss_kwargs = dict()
        num_pending = 0 if inputs_pend is None else np.shape(inputs_pend)[-2]
        parallelism = num_pending + np.shape(inputs_new)[1]

        losses_op, feed_dict, nodes, extras = self.appraise_inputs(
            sess, inputs_new, inputs_old, outputs_old, inputs_pend=inputs_pend, inputs_as_var=True,
            loss_fn=loss_fn, loss_kwargs=loss_kwargs, **kwargs)

        if batch_generator is None:
            def batch_generator(batch_size, parallelism=parallelism):
                batch_kwargs = {**loss_kwargs, "num_fantasies": batch_size}
                return loss_fn.prepare(sess, inputs_old, outputs_old, parallelism, self.model, **batch_kwargs)[1]

        if subroutine is None:
            subroutine = self._optimize_inputs_scipy if parallelism == 1 else self._optimize_inputs_sgd

        loss_seq = subroutine(
            sess=sess, inputs_new=nodes["inputs_new"], losses_op=losses_op, feed_dict=feed_dict, var_list=var_list,
            batch_generator=batch_generator, time_limit=time_limit - (self.timer() - start_time),
            parallelism=parallelism, nodes=nodes, extras=extras, config=config, loss_fn=loss_fn)
        return nodes, loss_seq

    # ------------------------------------------------------------------------------------------
    # random search (:337-374)
    # ------------------------------------------------------------------------------------------
    def _optimize_inputs_random(self, sess, inputs_new, losses_op, feed_dict=None, config=None,
                                time_limit=np.inf, options=None, **kwargs):
        feed_dict = dict() if feed_dict is None else dict(feed_dict)
        config = self.update_dict(self.configs["random"], config, as_copy=True)
        eval_limit = config["eval_limit"]
        shard_shape = list(inputs_new.shape)
        shard_size = shard_shape[0]
        start_time = self.timer()
        inputs_seq, loss_seq, counter = [], [], 0
        while counter + shard_size <= eval_limit:
            if self.timer() - start_time > time_limit:
                break
            inputs_seq.append(self.rng.rand(*shard_shape))
            losses = losses_op.evaluate(feed_dict, inputs_new=inputs_seq[-1])[0]
            loss_seq.append(np.ravel(losses))
            counter += shard_size
        loss_seq = np.hstack(loss_seq)
        argmins = np.argpartition(loss_seq, shard_size - 1)[:shard_size]
        inputs_new.assign(np.vstack(inputs_seq)[argmins])
        logger.info("Random Search evaluated {:d} losses in {:.3e}s".format(len(loss_seq), self.timer() - start_time))
        return loss_seq

    # ------------------------------------------------------------------------------------------
    # stochastic gradient descent: Adam + clip on the device (:377-482)
    # ------------------------------------------------------------------------------------------
    def _optimize_inputs_sgd(self, sess, inputs_new, losses_op, optimizer=None, feed_dict=None, var_list=None,
                             batch_generator=None, config=None, time_limit=np.inf, global_step=None,
                             scope="sgd", reuse=None, block=None, **kwargs):
        feed_dict = dict() if feed_dict is None else dict(feed_dict)
        if not isinstance(inputs_new, pool_loss):
            raise TypeError("Candidates must be a trainable <pool_loss> handle")
        config = self.update_dict(self.configs["sgd"], config, as_copy=True)
        method = config.get("method", "train.AdamOptimizer")
        kinds = {"AdamOptimizer": "adam", "GradientDescentOptimizer": "gd", "MomentumOptimizer": "momentum"}
        if method.split(".")[-1] not in kinds:
            raise NotImplementedError("optimizers on the device: {} (got {})".format(sorted(kinds), method))
        kind = kinds[method.split(".")[-1]]
        # Adam: lr 25e-3; GradientDescent / Momentum(0.9): lr 1.0 decayed by (global_step + 1)^-0.7 (:405-415)
        lr = config.get("learning_rate", 25e-3 if kind == "adam" else 1.0)
        step_limit = config.get("step_limit", 1000)
        batch_size = config.get("batch_size", 128)
        use_averaging = config.get("use_averaging", False)

        handle = inputs_new
        pools, p = handle.pools()
        S, q, d = pools.shape
        if np.ndim(handle.inputs_old) == 4:
            return self._sgd_on_fantasies(handle, kind, lr, step_limit, time_limit)
        eng = handle.model.bind(handle.inputs_old, handle.outputs_old)
        prm = handle.loss_fn.acq_params(outputs_old=handle.outputs_old, **handle.loss_kwargs)
        eng.adam_begin(prm, pools, p_pending=p, lr=lr, method=kind, momentum=0.9)     # resets slots + step (:444-451)

        start_time = self.timer()
        # steps per library call (= between looks at the clock).  Without a time limit: 16.  With one: 4 to begin with,
        # then as many as fit a twentieth of the remaining time (at most 256), so that small problems -- a device step of
        # tens of microseconds -- are not bound by one host round trip every four steps.
        adaptive = block is None and np.isfinite(time_limit) and not use_averaging
        if block is None:
            block = 16 if np.isinf(time_limit) else 4
        if use_averaging:
            block = 1
        loss_seq, mean_x, done = [], None, 0
        world_size = mdist.active()[1]
        while done < step_limit:
            now = self.timer() - start_time
            timed_out = now >= time_limit
            if world_size > 1 and np.isfinite(time_limit):
                timed_out = mdist.any_rank(timed_out)        # ranks stop after the same step: same host-RNG state
            if timed_out:
                break
            if adaptive and done and world_size == 1:
                per_step = max(now / done, 1e-7)
                block = int(min(256, max(4, (time_limit - now) / 20 / per_step)))
            k = int(min(block, step_limit - done))
            if q == 1:                                      # deterministic closed-form loss, no base samples
                trace = eng.adam_steps(None, want_trace=True, steps=k)
            else:
                zs = []
                for _ in range(k):
                    if callable(batch_generator):
                        feed_dict.update(batch_generator(batch_size))
                    zs.append(np.asarray(feed_dict["fantasy_rvs"], dtype=np.float64))
                trace = eng.adam_steps(np.stack(zs), want_trace=True)
            loss_seq.append(trace)
            done += k
            if use_averaging:                               # Polyak mean of the iterates (:472-475)
                x = eng.adam_peek()
                mean_x = x if mean_x is None else mean_x + x
        x = eng.adam_end()
        if use_averaging and done:
            x = mean_x / done
        handle.assign(x[:, p:, :])
        loss_seq = np.concatenate([np.ravel(t) for t in loss_seq]) if loss_seq else np.empty([0])
        logger.info("SGD evaluated {:d} losses in {:.3e}s".format(len(loss_seq), self.timer() - start_time))
        return loss_seq

    def _sgd_on_fantasies(self, handle, kind, lr, step_limit, time_limit):
        """`--subroutine sgd` on fantasy-padded (rank-4) observations, which the reference accepts (its sgd route sees the
        mean over fantasies of the closed-form loss, src/models/gaussian_process.py:195-206): the same TF-1 Adam /
        GradientDescent / Momentum updates and clip as on the device, driven from the host on the device's
        fantasy-conditioned value + gradient (one call per step; the default subroutine for pools of one point, L-BFGS,
        runs entirely on the device)."""
        x = np.array(handle.value, dtype=np.float64)                    # [S, 1, d]
        eng = handle.model.bind(handle.inputs_old, handle.outputs_old)
        prm = handle.loss_fn.acq_params(outputs_old=None, **handle.loss_kwargs)
        m, v, acc = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        b1, b2, eps, b1p, b2p = 0.9, 0.999, 1e-8, 1.0, 1.0
        loss_seq, start_time = [], self.timer()
        for t in range(int(step_limit)):
            if self.timer() - start_time >= time_limit:
                break
            losses, g = eng.fantasies_value_and_grad(prm, x[:, 0, :])
            g = np.reshape(g, x.shape)
            loss_seq.append(np.ravel(losses))
            if kind == "adam":
                b1p, b2p = b1p * b1, b2p * b2
                m += (g - m) * (1 - b1)
                v += (g * g - v) * (1 - b2)
                x = x - lr * np.sqrt(1 - b2p) / (1 - b1p) * m / (np.sqrt(v) + eps)
            else:
                acc = (0.9 if kind == "momentum" else 0.0) * acc + g
                x = x - lr * float(t + 1) ** -0.7 * acc
            x = np.clip(x, 0.0, 1.0)
        handle.assign(x)
        loss_seq = np.concatenate(loss_seq) if loss_seq else np.empty([0])
        logger.info("SGD (fantasy-conditioned) evaluated {:d} losses in {:.3e}s".format(len(loss_seq), self.timer() - start_time))
        return loss_seq

    # ------------------------------------------------------------------------------------------
    # joint L-BFGS-B (:485-555)
    # ------------------------------------------------------------------------------------------
    def _optimize_inputs_scipy(self, sess, inputs_new, losses_op, feed_dict=None, var_list=None, config=None,
                               time_limit=np.inf, **kwargs):
        feed_dict = dict() if feed_dict is None else dict(feed_dict)
        if not isinstance(inputs_new, pool_loss):
            raise TypeError("Candidates must be a trainable <pool_loss> handle")
        config = self.update_dict(self.configs["scipy"], config, as_copy=True)
        method = config.pop("method", "L-BFGS-B")
        config.pop("var_to_bounds", None)
        on_device = config.pop("device", True)
        handle = inputs_new
        shape = handle.value.shape
        S = shape[0]
        inputs_seq, loss_seq, start_time = [], [], self.timer()

        # L-BFGS-B on the device: the objective mean_s loss_s is separable over restarts (they only share SciPy's line
        # search), so every restart runs its own box-constrained L-BFGS inside the library (maf_lbfgs_run), all restarts
        # in lock step and with no host round trip per evaluation.  config={"device": False} keeps the joint SciPy loop
        # below (host L-BFGS-B driving device value + gradient), as do other `method`s.
        eng = handle.model.bind(handle.inputs_old, handle.outputs_old)
        if on_device and method == "L-BFGS-B" and hasattr(eng, "lbfgs_run"):
            pools, p = handle.pools()
            fantasies = np.ndim(handle.inputs_old) == 4
            prm = handle.loss_fn.acq_params(outputs_old=None if fantasies else handle.outputs_old, **handle.loss_kwargs)
            rvs = None if pools.shape[-2] == 1 else feed_dict.get("fantasy_rvs", None)
            if pools.shape[-2] > 1 and rvs is None:
                raise ValueError("base samples (feed_dict['fantasy_rvs']) are required for pools of more than one point")
            X, losses, trace, info = eng.lbfgs_run(
                prm, pools, None if rvs is None else np.asarray(rvs, dtype=np.float64), p_pending=p, fantasies=fantasies,
                maxiter=int(config.get("maxiter", 1000)), ftol=float(config.get("ftol", 1e6 * np.finfo("float64").eps)),
                pgtol=float(config.get("gtol", 1e-5)), time_limit=time_limit)
            handle.assign(X[:, p:, :])
            loss_seq = np.ravel(trace) if trace is not None else np.ravel(losses)
            logger.info("L-BFGS (device) evaluated {:d} losses in {:.3e}s".format(len(loss_seq), self.timer() - start_time))
            return loss_seq

        class _Timeout(Exception):
            pass

        def fun(x):
            xs = x.reshape(shape)
            losses, _, grad = handle.evaluate(feed_dict, inputs_new=xs, need_grad=True)
            losses = np.ravel(losses)
            loss_seq.append(losses)                       # loss_callback of the reference (:505-509)
            inputs_seq.append(xs.copy())
            if self.timer() - start_time >= time_limit:
                raise _Timeout("Runtime limit exhausted.")
            return float(np.mean(losses)), np.ravel(grad) / S     # loss_op = reduce_mean(losses_op) (:512)

        try:
            _scipy_minimize(fun, np.ravel(handle.value), jac=True, method=method,
                            bounds=[(0.0, 1.0)] * handle.value.size, options=config)
        except _Timeout:
            pass

        # best S (iterate, restart) pairs over the whole trajectory (:541-550)
        loss_seq = np.ravel(loss_seq)
        inputs_seq = np.reshape(inputs_seq, (-1,) + tuple(shape[1:]))
        argmins = np.argpartition(loss_seq, S - 1)[:S]
        handle.assign(inputs_seq[argmins])
        logger.info("ScipyOptimize evaluated {:d} losses in {:.3e}s".format(len(loss_seq), self.timer() - start_time))
        return loss_seq

    # ------------------------------------------------------------------------------------------
    # successive halving (:697-935), non-random-feature path
    # ------------------------------------------------------------------------------------------
    def _optimize_inputs_halving(self, sess, inputs_new, losses_op, feed_dict=None, config=None, time_limit=np.inf,
                                 options=None, batch_generator=None, parallelism=None, nodes=None, extras=None,
                                 loss_fn=None, cost_history=None, **kwargs):
        """Random options are scored with a growing number of base samples (128, 256, ... num_fantasies); after every
        round the better half survives.  Round 0 runs the full path (posterior + Monte-Carlo) on the device and
        keeps the posterior moments of every option; later rounds re-use them (``maf_mc_from_moments``: Cholesky
        + sampling only, the reference's ``loss_kwargs={'posteriors': ...}`` graph, :835-848) with the NEW base
        samples only and fold the estimates together weighted by their sample counts (:872-883)."""
        if not isinstance(inputs_new, pool_loss):
            raise TypeError("Candidates must be a trainable <pool_loss> handle")
        if loss_fn is None:
            loss_fn = self.loss_fn
        if getattr(loss_fn, "use_rand_features", False):
            raise NotImplementedError("random-feature sample paths are outside the accelerated path (SURVEY 2)")
        if cost_history is None:
            if not hasattr(self, "halving_costs"):
                self.halving_costs = dict()
            cost_history = self.halving_costs
        feed_dict = dict() if feed_dict is None else dict(feed_dict)
        config = self.update_dict(self.configs["halving"], config, as_copy=True)
        handle = inputs_new
        shard_shape = list(handle.shape)
        shard_size = shard_shape[0]
        num_pending = 0 if handle.inputs_pend is None else len(handle.inputs_pend)
        num_suggest = shard_shape[-2]
        if parallelism is None:
            parallelism = num_pending + num_suggest

        max_samples = config.get("max_samples", self.loss_fn.num_fantasies)
        expon_step = config.get("expon_step", 1)
        if parallelism > 1:
            lower, upper = int(np.log2(config.get("min_samples", 128))), int(np.log2(max_samples))
            if expon_step < 0:
                expon_step = upper - lower
            log_sizes = list(range(lower, upper + 1, max(int(expon_step), 1)))
            log_sizes[-1] = upper
            batch_sizes = [int(2 ** ls) for ls in log_sizes]
        else:
            batch_sizes = [max_samples]                     # closed form: a single round

        try:                                               # allocate_time (:764-783)
            allocs = [cost_history[(num_pending, num_suggest, b)] / b for b in batch_sizes]
        except KeyError:
            allocs = np.ones(len(batch_sizes))
        time_allocs = np.multiply(time_limit / np.sum(allocs), allocs)

        eng = handle.model.bind(handle.inputs_old, handle.outputs_old)
        prm = handle.loss_fn.acq_params(outputs_old=handle.outputs_old, **handle.loss_kwargs)

        def run_random_search(time_limit, eval_limit, options, posteriors, rvs):
            inputs_seq, loss_seq, kept, costs = [], [], [], []
            counter, run_start = 0, self.timer()
            while counter + shard_size <= eval_limit:
                tic = self.timer()
                if self.timer() - run_start > time_limit:
                    break
                if options is None:
                    inputs_seq.append(self.rng.rand(*shard_shape))
                    out = handle.evaluate({"fantasy_rvs": rvs}, inputs_new=inputs_seq[-1])
                    losses_k, ex = out[0], out[1]
                    kept.append(ex)
                else:
                    sl = slice(counter, counter + shard_size)
                    inputs_seq.append(options[sl])
                    mu, cov = posteriors[0][sl], posteriors[1][sl]
                    losses_k, _, _, chol = eng.mc_from_moments(prm, mu, cov, np.asarray(rvs, dtype=np.float64))
                    kept.append((mu, cov, chol))
                loss_seq.append(np.ravel(losses_k))
                counter += shard_size
                costs.append(self.timer() - tic)
            post = None
            if kept and kept[0] is not None and kept[0][0] is not None:
                post = [np.vstack([np.asarray(k[i]) for k in kept]) if kept[0][i] is not None else None for i in range(3)]
            return np.vstack(inputs_seq), np.hstack(loss_seq), costs, post

        losses, posteriors, counters, bsize_prev = None, None, [], 0
        start_time = self.timer()
        for k, (batch_size, time_alloc) in enumerate(zip(batch_sizes, time_allocs)):
            feed_dict.update(batch_generator(batch_size - bsize_prev))
            rvs = feed_dict.get("fantasy_rvs", None)
            eval_limit = config["eval_limit"] if options is None else len(options)
            options, losses_new, costs, posteriors = run_random_search(time_alloc, eval_limit, options, posteriors, rvs)
            cost_history[(num_pending, num_suggest, batch_size)] = float(np.mean(costs)) if costs else 0.0
            if losses is None:
                losses = losses_new
            else:
                num = len(losses_new)
                w_old, w_new = bsize_prev, batch_size - bsize_prev
                losses = (w_old * losses[:num] + w_new * losses_new) / (w_old + w_new)
            bsize_prev = batch_size
            num_options = len(options)
            counters.append(num_options)
            if k + 1 < len(batch_sizes):
                if time_limit == np.inf and num_options > shard_size:
                    top_k = max(shard_size, num_options // (2 ** int(expon_step)))
                    leaders = np.argpartition(losses, top_k - 1)[:top_k]
                    indices = leaders[np.argsort(losses[leaders])]
                else:
                    indices = np.argsort(losses)
                options, losses = options[indices], losses[indices]
                posteriors = [None if t is None else t[indices] for t in posteriors]

        argmins = np.argpartition(losses, shard_size - 1)[:shard_size]
        handle.assign(options[argmins])
        logger.info("Successive Halving evaluated {:d} options in {:.3e}s".format(counters[0], self.timer() - start_time))
        return losses

    def _optimize_inputs_cmaes(self, *args, **kwargs):
        raise NotImplementedError("CMA-ES is outside the accelerated path (SURVEY 2, row 1)")

    def _optimize_inputs_nlopt(self, *args, **kwargs):
        raise NotImplementedError("NLopt is outside the accelerated path (SURVEY 2, row 1)")

    # ------------------------------------------------------------------------------------------
    # suggest_answers (:938-1048)
    # ------------------------------------------------------------------------------------------
    def suggest_answers(self, num_suggest, inputs_old, outputs_old, sess=None, risk=None, num_starts=2 ** 10,
                        num_options=2 ** 16, config=None, time_limit=np.inf, in_order=False, model=None):
        """Top-k minimisers of the (risk-penalised) posterior mean  mu - risk * sqrt(var)."""
        from ..losses import confidence_bound, simple_regret
        if risk is None:
            risk = self.risk
        if model is None:
            model = self.model
        start_time = self.timer()
        config = self.update_dict(self.configs["scipy"], config, as_copy=True)
        input_dim = inputs_old.shape[-1]
        if risk != 0:
            answer_loss = confidence_bound(beta=float(risk) ** 2, lower=risk > 0, model=model, seed=int(self.seed))
        else:
            answer_loss = simple_regret(model=model, seed=int(self.seed))

        num_old = len(inputs_old)
        src = np.vstack([inputs_old, self.sample_inputs([max(num_options - num_old, 0), input_dim], use_sobol=True)])
        losses = np.ravel(answer_loss.evaluate(src[:, None, :], inputs_old, outputs_old, model=model)[0])
        starts = self.sample_propto_loss(src, losses, min(num_starts, len(src)), top_k=min(32, num_options, num_starts))[0]

        handle = pool_loss(answer_loss, model, starts[:, None, :], inputs_old, outputs_old, None, None)
        self._optimize_inputs_scipy(sess, handle, handle, dict(), config=config,
                                    time_limit=time_limit - (self.timer() - start_time))
        answers = handle.value[:, 0, :]
        means, sigma2s = model.predict(answers, inputs_old, outputs_old)
        means = np.squeeze(means, axis=-1)
        indices = np.argpartition(means, num_suggest - 1)[:num_suggest]
        if in_order:
            indices = indices[np.argsort(means[indices])]
        return answers[indices], (means[indices], sigma2s[indices])

    # ------------------------------------------------------------------------------------------
    # prepare (:1051-1125)
    # ------------------------------------------------------------------------------------------
    def prepare(self, sess, inputs_new, inputs_old, outputs_old, inputs_pend=None, feed_dict=None, nodes=None,
                inputs_as_var=False):
        nodes = dict() if nodes is None else nodes
        feed_dict = dict() if feed_dict is None else feed_dict
        assert np.ndim(inputs_new) > 1, "Candidates must be at least 2-dimensional"
        assert np.ndim(inputs_old) > 1, "Observations must be at least 2-dimensional"
        assert np.ndim(inputs_old) == np.ndim(outputs_old), "Observations have mismatched rank"
        assert np.shape(outputs_old)[-1] == 1, "Multiple objectives not yet supported"
        nodes["inputs_old"], nodes["outputs_old"] = inputs_old, outputs_old
        if inputs_pend is None:
            inputs_pend = np.empty((0, np.shape(inputs_new)[-1]))
        nodes["inputs_pend"] = inputs_pend
        nodes["inputs_new_src"] = np.asarray(inputs_new, dtype=np.float64)
        return nodes, feed_dict

    # ------------------------------------------------------------------------------------------
    # sample_starts (:1128-1204)
    # ------------------------------------------------------------------------------------------
    def sample_starts(self, num_starts, num_suggest, sess, inputs_old, outputs_old, inputs_pend=None, feed_dict=None,
                      losses_op=None, eval_limit=2 ** 20, time_limit=np.inf, shard_size=256, shards_per_call=64,
                      **kwargs):
        """Starting sets drawn with probability proportional to marginal utility: closed-form (q = 1)
        losses on uniform candidates, `shard_size` per host-RNG access exactly like the reference's loop
        (:1167-1172); `shards_per_call` shards are sent to the device in one call."""
        if inputs_pend is not None and len(inputs_pend):
            means_pend = self.model.predict(inputs_pend, inputs_old, outputs_old)[0]      # fantasise at the means
            inputs_old = np.vstack([inputs_old, inputs_pend])
            outputs_old = np.vstack([outputs_old, means_pend])
        input_dim = inputs_old.shape[-1]
        if losses_op is None:
            losses_op = self.appraise_inputs(sess=sess, inputs_new=np.empty([0, 1, input_dim]), inputs_old=inputs_old,
                                             outputs_old=outputs_old, feed_dict=feed_dict, **kwargs)[0]
        losses = []
        total_shards, k = eval_limit // shard_size, 0
        # every shard is one access to the reseeding host RNG, exactly like the reference's loop; the uniforms land in one
        # preallocated block (no per-call stacking) and the number of shards per device sweep doubles while the clock allows
        options = np.empty((total_shards * shard_size, input_dim))
        rank, world_size = mdist.active()
        g_call = max(1, int(shards_per_call))
        start_time = self.timer()
        while k < total_shards:
            timed_out = k > 0 and self.timer() - start_time > time_limit      # at least one sweep: the draws need weights
            if world_size > 1 and np.isfinite(time_limit):
                timed_out = mdist.any_rank(timed_out)        # every rank leaves the sweep after the same shard
            if timed_out:
                break
            g = int(min(g_call, total_shards - k))
            g = 1 << (g.bit_length() - 1)                    # a power of two: sweeps stay exact multiples of the loss's shard size
            block = options[k * shard_size:(k + g) * shard_size]
            for i in range(g):
                block[i * shard_size:(i + 1) * shard_size] = self.rng.rand(shard_size, input_dim)
            cand = block[:, None, :]
            if world_size > 1:                               # independent candidates: each rank scores its slice
                lo, hi = mdist.shard_range(len(cand), rank, world_size)
                l = losses_op.evaluate(None, inputs_new=cand[lo:hi])[0] if hi > lo else np.empty([0])
                l = mdist.all_gather_concat(np.ravel(l), len(cand))
            else:
                l = losses_op.evaluate(None, inputs_new=cand)[0]
            losses.append(np.atleast_1d(np.squeeze(l)))
            k += g
            t_call = (self.timer() - start_time) / max(k, 1) * g          # seconds per sweep at this size
            if world_size == 1 and g_call < 1024 and 4 * t_call < time_limit - (self.timer() - start_time):
                g_call *= 2                                               # (ranks keep the fixed size: same shards everywhere)
        losses = np.hstack(losses) if losses else np.empty([0])
        options = options[:k * shard_size]
        num_options = len(options)
        logger.info("Evaluated {:d} marginal losses in {:.3e}s".format(num_options, self.timer() - start_time))

        weights = np.clip(np.max(losses) - losses, np.finfo(losses.dtype).eps, np.inf)
        weights /= np.sum(weights)
        # `num_starts` draws of `num_suggest` distinct indices with probability proportional to the weights, from the same
        # host-RNG stream and with the same outcome as the reference's loop of `self.rng.choice(num_options, num_suggest,
        # p=weights, replace=False)` (:1186-1199) -- but the cumulative weights are summed once instead of once per draw
        # (2^20 options x thousands of starts), and the duplicate test is a set lookup instead of a scan of all earlier
        # draws: at 8192 starts the reference's loop costs ~100 s of host time per suggest_inputs call, this one ~0.1 s.
        raw0 = np.cumsum(weights)
        cdf0 = raw0 / raw0[-1]
        indices, fillers, seen = [], [], set()
        while len(indices) + len(fillers) < num_starts:
            new = _choice_without_replacement(self.rng, weights, cdf0, num_suggest, raw0)
            key = new.tobytes()
            if key in seen:
                fillers.append(self.rng.rand(num_suggest, input_dim))
            else:
                seen.add(key)
                indices.append(new)
        starts = options[np.stack(indices)]
        if len(fillers):
            starts = np.vstack([starts, np.stack(fillers)])
        return starts
