#!/usr/bin/env python
"""Golden vectors for acquisition values, posterior moments and gradients from the REFERENCE's own loss / model code.

Runs only in the build container (needs /root/reference).  TensorFlow 1.x cannot be installed here, so the
reference modules

    src/models/gaussian_process.py, src/models/kernels.py, src/losses/{myopic_loss,negative_ei,negative_pi,
    simple_regret,confidence_bound}.py, src/utils/{tensor,math,linalg,stats,cache}_ops.py, src/misc/sharded_fn.py,
    src/core/{module,stateful_module}.py

are imported UNMODIFIED with `tests/golden/tf_eager_shim/tensorflow` standing in for `tensorflow` (every tf.* op runs
eagerly on torch CPU float64; see that file).  What is exercised is the reference's own call chain

    loss.monte_carlo / loss.closed_form -> sharded_subroutine -> sharded_fn (tf.map_fn over shards)
      -> model.predict -> _predict_nd (covar_fn, kernels.*, pdist2, compute_cholesky, utils.cholesky_solve,
         broadcast_matmul, mean_fn, noise_fn) -> utils.jitter_cholesky -> _monte_carlo (draw_samples, level_fn,
         reduce_samples, soft_less, transform_residuals) / _closed_form (normal_pdf, normal_cdf, safe_div, tile_as)

and the gradient of sum(losses) with respect to the pools (what optimizer.minimize differentiates,
src/agents/bayesian_optimization.py:420-426) comes from torch autograd through those same calls.

Output: tests/golden/ref_acq.npz (a few hundred KB).  tests/test_oracle.py checks the oracle against it on the CPU,
tests/test_gpu_parity.py checks the CUDA path against it on the GPU.
"""
import os
import sys

import numpy as np

REF = os.environ.get("MAF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, acquisition class, constructor kwargs, kernel, (n, d, q, S, M), shard_limit, extras
    ("ei_m52", "negative_ei", {}, "matern52", (24, 3, 3, 6, 16), 3, {"level": "median"}),
    ("ei_se", "negative_ei", {}, "squared_exponential", (40, 2, 4, 6, 32), 6, {}),
    ("ei_m32_winc", "negative_ei", {}, "matern32", (20, 4, 2, 4, 24), 2, {"weights": True, "incumbents": True}),
    ("pi_soft_m52", "negative_pi", {"temperature": 0.01}, "matern52", (24, 3, 3, 6, 16), 6, {}),
    ("pi_hard_se", "negative_pi", {}, "squared_exponential", (24, 3, 3, 6, 16), 2, {"level": "median"}),
    ("sr_m52", "simple_regret", {}, "matern52", (30, 6, 5, 4, 20), 4, {}),
    ("sr_m32_inc", "simple_regret", {}, "matern32", (24, 3, 3, 6, 16), 3, {"incumbents": True}),
    ("lcb_m52", "confidence_bound", {"beta": 2.0, "lower": True}, "matern52", (24, 3, 3, 6, 16), 3, {}),
    ("ucb_se", "confidence_bound", {"beta": 3.0, "lower": False}, "squared_exponential", (28, 2, 4, 4, 16), 4, {}),
    # q = 1: closed forms (src/losses/myopic_loss.py:84-90)
    ("cf_ei_m52", "negative_ei", {}, "matern52", (24, 3, 1, 8, 0), 4, {"level": "median"}),
    ("cf_pi_se", "negative_pi", {}, "squared_exponential", (24, 3, 1, 8, 0), 8, {}),
    ("cf_sr_m32", "simple_regret", {}, "matern32", (24, 3, 1, 8, 0), 8, {}),
    ("cf_lcb_m52", "confidence_bound", {"beta": 2.0, "lower": True}, "matern52", (24, 3, 1, 8, 0), 8, {}),
]


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "tf_eager_shim"))
    sys.path.insert(0, REF)
    import warnings
    warnings.simplefilter("ignore")
    import tensorflow as tf                                   # the shim
    assert tf.__version__ == "1.x-eager-shim"
    from src import losses, models                            # the reference's packages, unmodified
    assert os.path.abspath(losses.__file__).startswith(os.path.abspath(REF))
    return tf, losses, models


def main():
    import torch
    tf, losses, models = _import_reference()
    out = {"cases": np.array([c[0] for c in CASES])}
    for k, (name, cls, ctor, kernel_id, (n, d, q, S, M), shard_limit, extras) in enumerate(CASES):
        rs = np.random.RandomState(1000 + k)
        X0 = rs.rand(n, d)
        Y0 = rs.randn(n, 1)
        pools = rs.rand(S, q, d)
        pools[0, 0] = X0[3] + 1e-3 * rs.randn(d)              # one candidate next to a training point
        ls = 0.3 + 0.5 * rs.rand(d)
        amp, noise, mean = 0.8 + 0.4 * rs.rand(), 1e-3, 0.1 * rs.randn()
        model = models.gaussian_process(kernel_id=kernel_id, dtype="float64", seed=1, name="gp_%s" % name,
                                        log_lenscales=np.log(ls), log_amplitude=np.log(amp),
                                        log_noise=np.log(noise), mean=mean)
        loss_fn = getattr(losses, cls)(model=model, seed=2, name="loss_%s" % name, **ctor)
        tX0, tY0 = tf.Tensor(torch.as_tensor(X0)), tf.Tensor(torch.as_tensor(Y0))
        tp = torch.as_tensor(pools).clone().requires_grad_(True)
        kwargs = {"sharding_config": {"shard_limit": shard_limit}}
        rec = dict(X0=X0, Y0=Y0, pools=pools, lenscales=ls, amplitude=amp, noise=noise, mean=mean, jitter=model.jitter)
        if extras.get("level") == "median":                   # explicit threshold (levels= of _monte_carlo / _closed_form)
            rec["level"] = float(np.median(Y0))
            kwargs["levels"] = tf.Tensor(torch.full((1, 1), rec["level"], dtype=torch.float64))
        if q > 1:
            z = rs.randn(M, q)
            rec["z"] = z
            if extras.get("weights"):
                w = rs.rand(M)
                rec["weights"] = w / w.sum()
                kwargs["weights"] = tf.Tensor(torch.as_tensor(rec["weights"]))
            if extras.get("incumbents"):
                rec["incumbents"] = Y0.min() + 0.5 * rs.rand(M)
                kwargs["incumbents"] = tf.Tensor(torch.as_tensor(rec["incumbents"]))
            res = loss_fn.monte_carlo(tf.Tensor(tp), tX0, tY0, fantasy_rvs=tf.Tensor(torch.as_tensor(z)),
                                      parallelism=q, **kwargs)
        else:
            res = loss_fn.closed_form(tf.Tensor(tp), tX0, tY0, **kwargs)
        vals, (means, cov, chol) = res                        # myopic_loss.py:179
        if vals.t.requires_grad:
            vals.t.sum().backward()                           # gradient of a vector loss = gradient of its sum
            grad = tp.grad.numpy()
        else:                                                 # hard PI: tf.less has no gradient
            grad = np.zeros_like(pools)
        rec.update(losses=vals.numpy(), means=means.numpy(), cov=cov.numpy(), chol=chol.numpy(), grad=grad)
        if q > 1:
            assert rec["losses"].shape == (S,) and rec["cov"].shape == (S, q, q)
        else:
            assert rec["losses"].shape == (S, 1, 1) and rec["cov"].shape == (S, 1, 1)      # negative_ei.py:44-49
        for key, v in rec.items():
            out["%s/%s" % (name, key)] = np.asarray(v)
        out["%s/meta" % name] = np.array([cls, kernel_id, repr(sorted(ctor.items()))])
        print("%-12s %-18s losses %s  |grad| max %.3e" % (name, cls, np.array2string(rec["losses"].ravel()[:3], precision=6),
                                                            np.abs(rec["grad"]).max()))
    # ---- negative log posterior of gaussian_process.fit (log_likelihood :78-106 + default priors :55-62) and its
    # gradient with respect to every state variable (what ScipyOptimizerInterface differentiates, :283-325)
    nll_names = []
    for k, kernel_id in enumerate(("matern52", "squared_exponential", "matern32")):
        rs = np.random.RandomState(2000 + k)
        n, d = 36, 3
        X0, Y0 = rs.rand(n, d), rs.randn(n, 1)
        hyp = dict(log_lenscales=np.log(0.3 + 0.5 * rs.rand(d)), log_amplitude=0.2 * rs.randn(), log_noise=-4.0 - rs.rand(),
                   mean=0.1 * rs.randn())
        model = models.gaussian_process(kernel_id=kernel_id, dtype="float64", seed=1, name="gpnll_%s" % kernel_id, **hyp)
        for v in model.state.values():
            v.t.requires_grad_(True)
        nll = model.log_likelihood(tf.Tensor(torch.as_tensor(X0)), tf.Tensor(torch.as_tensor(Y0)), as_negative=True)
        nll.t.backward()
        name = "nll_%s" % kernel_id
        nll_names.append(name)
        out["%s/X0" % name], out["%s/Y0" % name] = X0, Y0
        out["%s/kernel_id" % name] = np.array(kernel_id)
        out["%s/jitter" % name] = np.asarray(model.jitter)
        out["%s/value" % name] = nll.numpy()
        for key, v in hyp.items():
            out["%s/%s" % (name, key)] = np.asarray(v)
            out["%s/grad_%s" % (name, key)] = model.state[key].t.grad.numpy()
        print("%-24s nll %.12f  d/dlog_noise %.6e" % (name, float(nll.numpy()), float(model.state["log_noise"].t.grad)))
    out["nll_cases"] = np.array(nll_names)

    # ---- fantasy-padded observations of the incremental greedy script (scripts/run_bayesopt_incremental.py:150-179):
    # inputs_old [F,1,n,d], outputs_old [F,1,n,1], q = 1 closed forms through the rank-4 branch of _predict_nd
    # (gaussian_process.py:195-206) with the precomputed rank-4 Cholesky factor (appraise_inputs :188-198), then the
    # mean over the fantasy axis (appraise_inputs :227-229)
    fan_names = []
    for k, (cls, ctor) in enumerate((("negative_ei", {}), ("negative_pi", {}), ("simple_regret", {}),
                                     ("confidence_bound", {"beta": 2.0, "lower": True}))):
        rs = np.random.RandomState(3000 + k)
        n, d, S, F = 18, 2, 7, 5
        X0 = rs.rand(n, d)
        Y = rs.randn(1, n) + 0.3 * rs.randn(F, n)              # F fantasised output vectors over the same inputs
        pools = rs.rand(S, 1, d)
        pools[:3, 0] = np.clip(X0[np.argmin(Y.mean(0))] + 0.03 * rs.randn(3, d), 0, 1)   # near the incumbent: EI / PI of order 1
        ls = 0.3 + 0.5 * rs.rand(d)
        amp, noise, mean = 0.8 + 0.4 * rs.rand(), 0.05, 0.1 * rs.randn()
        name = "fan_%s" % cls
        model = models.gaussian_process(kernel_id="matern52", dtype="float64", seed=1, name="gp_%s" % name,
                                        log_lenscales=np.log(ls), log_amplitude=np.log(amp), log_noise=np.log(noise), mean=mean)
        loss_fn = getattr(losses, cls)(model=model, seed=2, name="loss_%s" % name, **ctor)
        X0_4 = np.tile(X0[None, None], [F, 1, 1, 1])
        Y0_4 = Y[:, None, :, None]
        tX0, tY0 = tf.Tensor(torch.as_tensor(X0_4)), tf.Tensor(torch.as_tensor(Y0_4))
        chol4 = model.compute_cholesky(tX0, noisy=True)
        tp = torch.as_tensor(pools).clone().requires_grad_(True)
        vals, _extras = loss_fn.closed_form(tf.Tensor(tp), tX0, tY0, predict_kwargs={"chol": chol4})
        assert vals.shape.ndims == 4 and tuple(vals.shape) == (F, S, 1, 1)
        vals = tf.reduce_mean(vals, 0)
        vals.t.sum().backward()
        fan_names.append(name)
        for key, v in dict(X0=X0, Y=Y, pools=pools, lenscales=ls, amplitude=amp, noise=noise, mean=mean, jitter=model.jitter,
                           losses=vals.numpy(), grad=tp.grad.numpy()).items():
            out["%s/%s" % (name, key)] = np.asarray(v)
        out["%s/meta" % name] = np.array([cls, "matern52", repr(sorted(ctor.items()))])
        print("%-24s losses %s |grad| max %.3e" % (name, np.array2string(vals.numpy().ravel()[:3], precision=6), np.abs(tp.grad.numpy()).max()))
    out["fantasy_cases"] = np.array(fan_names)

    # ---- pending points: pools = concat(tile(expand_to(inputs_pend, 3), [S,1,1]), inputs_new, axis=-2) as
    # bayesian_optimization.prepare builds them (:1106-1123; same tf / utils calls), gradient w.r.t. inputs_new only
    from src import utils as ref_utils
    rs = np.random.RandomState(5000)
    n, d, S, M, p_pend, q_new = 26, 3, 5, 24, 2, 3
    X0, Y0 = rs.rand(n, d), rs.randn(n, 1)
    ls = 0.3 + 0.5 * rs.rand(d)
    amp, noise, mean = 0.9, 1e-3, -0.05
    Xpend, Xnew, z = rs.rand(p_pend, d), rs.rand(S, q_new, d), rs.randn(M, p_pend + q_new)
    Xpend = np.clip(X0[np.argsort(Y0[:, 0])[-p_pend:]] + 0.05 * rs.randn(p_pend, d), 0, 1)   # pending jobs at poor locations
    model = models.gaussian_process(kernel_id="matern52", dtype="float64", seed=1, name="gp_pend", log_lenscales=np.log(ls),
                                    log_amplitude=np.log(amp), log_noise=np.log(noise), mean=mean)
    pend_names = []
    for cls, ctor in (("negative_ei", {}), ("confidence_bound", {"beta": 2.0, "lower": True})):
        loss_fn = getattr(losses, cls)(model=model, seed=2, name="loss_pend_%s" % cls, **ctor)
        tn = torch.as_tensor(Xnew).clone().requires_grad_(True)
        inputs_new = tf.Tensor(tn)
        shape_n = tf.shape(inputs_new)
        tiled = tf.tile(ref_utils.expand_to(tf.Tensor(torch.as_tensor(Xpend)), 3), (shape_n.__getitem__(0), 1, 1))
        pools = tf.concat((tiled, inputs_new), axis=-2)
        kw = {"levels": tf.Tensor(torch.full((1, 1), float(np.median(Y0)), dtype=torch.float64))} if cls == "negative_ei" else {}
        vals, _extras = loss_fn.monte_carlo(pools, tf.Tensor(torch.as_tensor(X0)), tf.Tensor(torch.as_tensor(Y0)),
                                            fantasy_rvs=tf.Tensor(torch.as_tensor(z)), parallelism=p_pend + q_new, **kw)
        vals.t.sum().backward()
        name = "pend_%s" % cls
        pend_names.append(name)
        for key, v in dict(X0=X0, Y0=Y0, lenscales=ls, amplitude=amp, noise=noise, mean=mean, jitter=model.jitter, Xpend=Xpend,
                           Xnew=Xnew, z=z, level=float(np.median(Y0)), losses=vals.numpy(), grad=tn.grad.numpy()).items():
            out["%s/%s" % (name, key)] = np.asarray(v)
        out["%s/meta" % name] = np.array([cls, "matern52", repr(sorted(ctor.items()))])
        print("%-24s losses %s |grad| max %.3e" % (name, np.array2string(vals.numpy()[:3], precision=6), np.abs(tn.grad.numpy()).max()))
    out["pending_cases"] = np.array(pend_names)

    # ---- gaussian_process.draw_samples (:328-349) with explicit base samples: at a point set (rank 2, first greedy
    # step of the incremental script) and under fantasy-padded observations (rank 4, later steps)
    rs = np.random.RandomState(4000)
    n, d, m, ns, F = 20, 3, 4, 6, 5
    X0, Y0 = rs.rand(n, d), rs.randn(n, 1)
    ls = 0.3 + 0.5 * rs.rand(d)
    amp, noise, mean = 1.1, 1e-3, 0.05
    model = models.gaussian_process(kernel_id="matern52", dtype="float64", seed=1, name="gp_draw", log_lenscales=np.log(ls),
                                    log_amplitude=np.log(amp), log_noise=np.log(noise), mean=mean)
    X1, eps = rs.rand(m, d), rs.randn(ns, m)
    smp = model.draw_samples(ns, tf.Tensor(torch.as_tensor(X1)), tf.Tensor(torch.as_tensor(X0)), tf.Tensor(torch.as_tensor(Y0)),
                             base_rvs=tf.Tensor(torch.as_tensor(eps)))
    assert tuple(smp.shape) == (ns, m)
    Yf = Y0.reshape(1, n) + 0.3 * rs.randn(F, n)
    x1, eps1 = rs.rand(1, d), rs.randn(1, 1)
    smp4 = model.draw_samples(1, tf.Tensor(torch.as_tensor(x1)), tf.Tensor(torch.as_tensor(np.tile(X0[None, None], [F, 1, 1, 1]))),
                              tf.Tensor(torch.as_tensor(Yf[:, None, :, None])), base_rvs=tf.Tensor(torch.as_tensor(eps1)))
    assert tuple(smp4.shape) == (F, 1, 1, 1)
    for key, v in dict(X0=X0, Y0=Y0, lenscales=ls, amplitude=amp, noise=noise, mean=mean, X1=X1, eps=eps, samples=smp.numpy(),
                       Yf=Yf, x1=x1, eps1=eps1, samples4=smp4.numpy()).items():
        out["draw/%s" % key] = np.asarray(v)
    print("draw_samples rank 2 %s rank 4 %s" % (smp.numpy()[0, :2], smp4.numpy().ravel()[:2]))
    path = os.path.join(HERE, "ref_acq.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
