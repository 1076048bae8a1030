"""Host-side helpers.  The reference's ``src/utils`` is a library of TensorFlow ops; on this path
they are fused inside the CUDA kernels (pdist2 / kernels -> cross_fwd_kernel, cholesky_solve /
broadcast_matmul -> gemm_nt_f64_kernel + posterior_kernel, jitter_cholesky / normal_pdf / normal_cdf /
soft_less / safe_div -> mc_acq_kernel / closed_form_kernel).  Only shape utilities that the host
code needs are provided here."""
import numpy as np


def atleast_pools(inputs_new):
    """[S, q, d] view of candidate sets given as [q, d] or [S, q, d]."""
    a = np.asarray(inputs_new, dtype=np.float64)
    if a.ndim == 2:
        a = a[None]
    assert a.ndim == 3, "candidates must be [S, q, d]"
    return a


def join_pools(inputs_new, inputs_pend):
    """pools = concat(tile(pend), new) along the set axis (bayesian_optimization.prepare :1106-1123)."""
    new = atleast_pools(inputs_new)
    if inputs_pend is None or len(inputs_pend) == 0:
        return np.ascontiguousarray(new), 0
    pend = np.asarray(inputs_pend, dtype=np.float64).reshape(-1, new.shape[-1])
    tiled = np.broadcast_to(pend[None], (new.shape[0],) + pend.shape)
    return np.ascontiguousarray(np.concatenate([tiled, new], axis=-2)), pend.shape[0]
