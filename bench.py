#!/usr/bin/env python
"""Benchmark of the hot path: qEI value + gradient evaluations per second.

    python bench.py --gpus N --steps K --warmup W              # the CUDA path
    python bench.py --impl reference --steps K --warmup W      # the reference op sequence on host cores

Metric (BASELINE.json): (restart x q x MC-sample) acquisition value+gradient evaluations per
second for qEI at n_train=4096, d=6, q=8, 128 MC samples.  One "step" = one value+gradient
pass over all of this rank's restarts (S per GPU, weak scaling: restarts are independent
units) followed by the rank-local arg-min and, for N > 1, the single all-gather that picks
the global best.  Prints ONE JSON line (rank 0).

  value : inputs (X, z) resident in HBM, timed with CUDA events on the library's stream.
  e2e   : the same through the host-pointer C-ABI call (maf_acq_value_grad) with pinned host
          buffers: H2D of X and z, D2H of losses and gradient inside the timed region.
  roofline : the dominant kernel, the contraction V^T = K10 L0^-T / Kbar = Vbar^T L0^-1.  Default engine: exact INT8
          digit slicing on tcgen05 (ozaki_i8_gemm_v4_kernel, 28 int8 digit-plane products per FP64 product) against
          2 x the measured cuBLAS bf16 figure of MEASURED_PEAKS.json; --gemm f64: the FP64 DMMA kernel against the
          repo's own cuBLAS DGEMM measurement.  Algorithmic flops 2 n^2 q per restart per value+grad,
          durations from CUDA events around every launch of that kernel on the library's stream.
  adam_loop : the same metric through maf_adam_steps (value+gradient, Adam update and clip on the device).
  cpu_baseline : the torch-CPU float64 restatement of the reference op sequence (oracle/), timed
          on this box's host cores on a bounded sample (rank 0, N=1 only).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# level: the improvement threshold of qEI.  The reference takes min(y0) (negative_ei.py:31-34); on this synthetic training
# set (y0 ~ N(0,1) noise, n=4096) no posterior sample of a random candidate reaches it, so qEI and its gradient are
# identically 0, the reverse pass multiplies all-zero digit planes and the power-capped part runs it at higher clocks.
# The bench therefore places the threshold at the median of y0 (about half of the samples improve, every restart has a
# non-zero gradient): same kernels, same launches, realistic operand activity.  --level min restores the reference's choice.
WORKLOAD = dict(n=4096, d=6, q=8, M=128, S_per_gpu=65536, kernel="matern52", acq="ei", seed=0, level="median")
METRIC = "qEI value+grad evals/sec (restarts x q x MC) @ n=4096,d=6,q=8"
UNIT = "evals/s"


def load_fp64_peak():
    """FP64 roofline denominator.  MEASURED_PEAKS.json (driver-written) has only HBM and bf16
    entries, so the FP64 figure is this repo's own cuBLAS DGEMM measurement on the same pool
    (scripts/microbench_fp64.cu -> profiles/fp64_peaks.json); else nominal."""
    p = os.path.join(ROOT, "profiles", "fp64_peaks.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["cublas_dgemm_tflops"]), "own cuBLAS DGEMM measurement (profiles/fp64_peaks.json); MEASURED_PEAKS.json has no fp64 entry"
        except Exception:
            pass
    return 37.0, "nominal B200 fp64 (no measurement found)"


OZ_PAIRS = 28          # digit-plane pair products of the FP64-grade configuration (ns=7, lmax=8)
DEFAULT_GEMM = "i8"


def load_i8_peak():
    """INT8 tensor-pipe denominator: 2 x the driver-measured cuBLAS bf16 figure of MEASURED_PEAKS.json (kind::i8
    runs at twice the bf16 rate on tcgen05; sustained figure, the kernel is timed inside a long step), plus this
    repo's own raw tcgen05 kind::i8 issue measurement (scripts/microbench_tcgen05.cu) for context."""
    issue = None
    try:
        issue = float(json.load(open(os.path.join(ROOT, "profiles", "i8_peaks.json")))["tcgen05_i8_issue_tops"])
    except Exception:
        pass
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return 2.0 * float(d["bf16_tflops_sustained"]), "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (of measured)", issue
    except Exception:
        return 2.0 * 1400.0, "2 x 1.4 PFLOP/s sustained bf16 (of fallback)", issue


def load_traffic(kind):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kind]
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (nvidia-smi, 200 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload_level(w, y0):
    """None = the reference's min(y0); 'median' = median of y0; a number = itself."""
    lv = w.get("level", "min")
    if lv in (None, "min"):
        return None
    if lv == "median":
        return float(np.median(y0))
    return float(lv)


def make_inputs(S, w=WORKLOAD):
    from maximizingacquisitionfunctions_b200.synthetic import synthetic_inputs
    return synthetic_inputs(w["n"], w["d"], w["q"], S, w["M"], seed=w["seed"])


def oracle_state(w, X0, y0):
    """CPU-baseline legs only: the oracle's view of the same model (the one place bench.py touches oracle/)."""
    from oracle import maf_oracle as orc
    gp = orc.GP(log_lenscales=math.log(math.sqrt(w["d"]) / 4), log_amplitude=0.0, log_noise=math.log(1e-3),
                mean=0.0, kernel_id=w["kernel"], jitter=1e-6)
    return orc, orc.prepare_train(gp, X0, y0), orc.Acq(w["acq"], level=workload_level(w, y0))


def cpu_reference_rate(w, threads, target_s=12.0, s_first=256):
    """evals/s of the torch-CPU restatement of the reference op sequence (value + autograd
    gradient, 4096-restart shards like sharded_fn) on a bounded sample of the same workload."""
    import torch
    torch.set_num_threads(threads)
    X0, y0, X, z = make_inputs(max(s_first, 4096), w)
    orc, ts, acq = oracle_state(w, X0, y0)
    t0 = time.perf_counter()
    orc.value_and_grad(ts, acq, X[:s_first], z)
    t1 = time.perf_counter() - t0
    S = int(min(4096, max(s_first, s_first * target_s / max(t1, 1e-3))))
    S = max(s_first, S // s_first * s_first)
    t0 = time.perf_counter()
    orc.value_and_grad(ts, acq, X[:S], z)
    dt = time.perf_counter() - t0
    return S * w["q"] * w["M"] / dt, S, dt


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: TensorFlow 1.x cannot be installed)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    w = dict(WORKLOAD)
    w["level"] = args.level
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    S = 512
    X0, y0, X, z = make_inputs(S, w)
    orc, ts, acq = oracle_state(w, X0, y0)
    for _ in range(args.warmup):
        orc.value_and_grad(ts, acq, X[:64], z)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.value_and_grad(ts, acq, X, z)
    dt = time.perf_counter() - t0
    val = args.steps * S * w["q"] * w["M"] / dt
    sample = "S=%d restarts per step (of %d), n=%d,d=%d,q=%d,M=%d, value+autograd gradient" % (
        S, w["S_per_gpu"], w["n"], w["d"], w["q"], w["M"])
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "qEI n=4096 d=6 q=8 M=128 matern52, level=%s of y0 (torch-CPU restatement of the reference TF op sequence; TF 1.x not installable)" % w["level"],
                   "S_per_step": S},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


AGENT_CONFIGS = [
    # BASELINE configs 1-3 through the product API (agent.suggest_inputs: starts by the closed-form sweep, Adam on the
    # device, re-evaluation with 1024 base samples, arg-min); a fixed optimizer budget instead of the scripts' time limit
    ("C1 Hartmann-3 qEI q=2 n=32 S=64 M=128", dict(task="hartmann3", d=3, n=32, q=2, S=64, M=128, loss="negative_ei", kw={}, steps=100)),
    ("C2 Branin qPI(tau=0.01) q=4 n=256 S=1024 M=256", dict(task="braninhoo", d=2, n=256, q=4, S=1024, M=256, loss="negative_pi", kw={"temperature": 0.01}, steps=50)),
    ("C3 Hartmann-6 qEI q=8 n=1024 S=4096 M=128", dict(task="hartmann6", d=6, n=1024, q=8, S=4096, M=128, loss="negative_ei", kw={}, steps=25)),
]


def agent_suggest_inputs_times(device):
    """Wall time of agent.suggest_inputs (second call; the first one warms buffers and graphs up) for C1-C3."""
    from maximizingacquisitionfunctions_b200 import runtime, tasks
    from maximizingacquisitionfunctions_b200.src import agents, losses, models
    out = []
    for name, c in AGENT_CONFIGS:
        rng = np.random.RandomState(0)
        X0 = rng.rand(c["n"], c["d"])
        Y0 = getattr(tasks, c["task"])(X0) + math.sqrt(1e-3) * rng.randn(c["n"], 1)
        model = models.gaussian_process(kernel_id="matern52", seed=1, log_lenscales=np.log(math.sqrt(c["d"]) / 4) * np.ones(c["d"]),
                                        log_amplitude=0.0, log_noise=np.log(1e-3), mean=0.0, device=device)
        loss_fn = getattr(losses, c["loss"])(seed=2, num_fantasies=1024, **c["kw"])
        agent = agents.bayesian_optimization(loss_fn=loss_fn, model=model, seed=3, timer=time.perf_counter)
        cfg = {"step_limit": c["steps"], "batch_size": c["M"]}
        wall = []
        for _ in range(2):
            t0 = time.perf_counter()
            agent.suggest_inputs(c["q"], X0, Y0, None, num_to_optimize=c["S"], subroutine=agent._optimize_inputs_sgd, config=cfg)
            wall.append(time.perf_counter() - t0)
        out.append({"config": name, "adam_steps": c["steps"], "wall_s": wall[1], "first_call_wall_s": wall[0],
                    "evals_per_s": c["steps"] * c["S"] * c["q"] * c["M"] / wall[1]})
        runtime.release_engines()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--restarts", type=int, default=WORKLOAD["S_per_gpu"], help="restarts per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-agent", action="store_true", help="skip the agent-level suggest_inputs timings (C1-C3)")
    ap.add_argument("--level", type=str, default=WORKLOAD["level"],
                    help="qEI improvement threshold: 'median' (of y0; default, non-degenerate), 'min' (reference default) or a number")
    ap.add_argument("--gemm", type=str, default=os.environ.get("MAF_BENCH_GEMM", DEFAULT_GEMM), choices=["f64", "i8"],
                    help="contraction engine: f64 = DMMA, i8 = exact INT8 digit slicing on tcgen05 (FP64-grade)")
    ap.add_argument("--dtype", type=str, default="float64", choices=["float64", "float32"],
                    help="float32 = the reference's dtype='float32' mode (jitter 1e-4, 1e-4 bar): 15 digit-plane products")
    ap.add_argument("--scaling", type=str, default="weak", choices=["weak", "strong"],
                    help="weak: --restarts per GPU (default, the driver's scaling run); strong: BASELINE config 5, --restarts "
                         "in TOTAL (65536) sharded over the ranks, 1024 base samples")
    ap.add_argument("--mc-samples", type=int, default=None, help="base samples M (default 128; 1024 with --scaling strong)")
    ap.add_argument("--digits", type=str, default=None,
                    help="F,B: digit planes of the forward / reverse INT8 contraction (default 7,7; float32 mode 5,5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from maximizingacquisitionfunctions_b200 import dist as mdist
    from maximizingacquisitionfunctions_b200.engine import Engine, acq_params

    rank, ws, local = mdist.world()
    if ws > 1:
        mdist.init_process_group("nccl")
    import torch
    torch.cuda.set_device(local)

    w = dict(WORKLOAD)
    w["level"] = args.level
    strong = args.scaling == "strong"
    w["M"] = int(args.mc_samples) if args.mc_samples else (1024 if strong else w["M"])
    n, d, q, M = w["n"], w["d"], w["q"], w["M"]
    if strong:
        # BASELINE config 5: a fixed set of restarts partitioned contiguously over the ranks (dist.shard_range)
        S_total = int(args.restarts)
        lo, hi = mdist.shard_range(S_total, rank, ws)
        S, offset = hi - lo, lo
    else:
        # every rank owns S restarts (weak scaling); its shard of the global restart index space
        S = int(args.restarts)
        S_total, offset = ws * S, rank * S
    X0, y0, _, z = make_inputs(1, w)
    X = np.random.RandomState(w["seed"] + 2 + 1000 * rank).rand(S, q, d)
    eng = Engine(local, args.dtype)
    eng.set_gemm_mode(1 if args.gemm == "i8" else 0)
    f32 = args.dtype == "float32"
    ns_f, ns_b = (5, 5) if f32 else (7, 7)
    if args.digits:
        ns_f, ns_b = (int(v) for v in args.digits.split(","))
        eng.set_digits(ns_f, ns_b)
    pairs_f, pairs_b = ns_f * (ns_f + 1) // 2, ns_b * (ns_b + 1) // 2     # products with level i + i' <= ns + 1
    pairs = 0.5 * (pairs_f + pairs_b)
    eng.set_model(w["kernel"], np.full(d, math.sqrt(d) / 4), 1.0, 1e-3, 0.0, 1e-4 if f32 else 1e-6)
    t0 = time.perf_counter()
    eng.set_train(X0, y0)
    setup_s = time.perf_counter() - t0
    prm = acq_params(w["acq"], level=workload_level(w, y0))

    # ---- value: inputs resident in HBM -------------------------------------------------------
    dX = eng.alloc(X.nbytes).upload(X)
    dz = eng.alloc(z.nbytes).upload(z)
    dloss = eng.alloc(S * 8)
    dgrad = eng.alloc(X.nbytes)

    Xshape = np.broadcast_to(0.0, (S, q, d))

    def step_dev():
        eng.value_and_grad_dev(prm, S, q, 0, dX, M, dz, dloss, dgrad)
        if ws > 1:
            # selection record assembled on the device, ONE NCCL all-gather, one read-back of the W x (2 + q d) table
            val, idx, _ = mdist.global_best(eng, None, Xshape, offset)
        else:
            idx, val = eng.argmin_last()          # device kernel + 16-byte D2H
        return idx, val

    for _ in range(args.warmup):
        step_dev()
    eng.level_switch_stats(reset=True)
    eng.profile_enable(True)
    eng.profile_reset()
    mdist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count
    eng.timer_start()
    for _ in range(args.steps):
        step_dev()
    ms = eng.timer_stop()
    torch.cuda.synchronize()
    mdist.barrier()
    launches = eng.launch_count - launches0
    lsw = eng.level_switch_stats()
    if lsw["restarts"]:
        # level switch on: restarts tested run one plane less first (ns - 1), the recomputed ones again on all planes;
        # a direction the switch has parked (more than 40 % recomputed) runs on all planes directly
        N_eval = float(S * args.steps)
        cheap_f, cheap_b = (ns_f - 1) * ns_f // 2, (ns_b - 1) * ns_b // 2
        pairs_f = (lsw["tested_forward"] * cheap_f + (lsw["recomputed_forward"] + N_eval - lsw["tested_forward"]) * pairs_f) / N_eval
        pairs_b = (lsw["tested_reverse"] * cheap_b + (lsw["recomputed_reverse"] + N_eval - lsw["tested_reverse"]) * pairs_b) / N_eval
        pairs = 0.5 * (pairs_f + pairs_b)
    clocks = sampler.stop()
    prof = eng.profile_read()
    eng.profile_enable(False)
    ms = mdist.max_over_ranks(ms)
    value = S_total * q * M * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (FP64 DMMA contraction) --------------------------------
    gemm_ms = prof["gemm_fwd"][0] + prof["gemm_bwd"][0]
    gemm_launches = prof["gemm_fwd"][1] + prof["gemm_bwd"][1]
    npad = (n + 127) // 128 * 128
    flops_per_step = 2.0 * npad * npad * q * S            # n^2 q per triangular GEMM, two GEMMs (this rank's restarts)
    peak64, peak64_src = load_fp64_peak()
    fp64_equiv = flops_per_step * args.steps / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
    common = {"kernel_ms_per_launch": gemm_ms / max(gemm_launches, 1), "kernel_launches_per_step": gemm_launches / args.steps,
              "kernel_ms_per_step": gemm_ms / args.steps, "kernel_share_of_step": gemm_ms / ms,
              "per_kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()}}
    if args.gemm == "i8":
        # every algorithmic FP64 flop is OZ_PAIRS exact int8 digit-plane products on the tcgen05 i8 pipe
        peak8, peak8_src, peak8_issue = load_i8_peak()
        achieved = fp64_equiv * pairs
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak8, "unit": "TOP/s", "frac": achieved / peak8,
                    "traffic": load_traffic("i8"), "kernel": ("ozaki_i8_gemm_v4_kernel<%d,%d,1> (first pass of the level switch, every restart) and <%d,%d,1> (second pass, the recomputed restarts on compact rows)" % (ns_f - 1, ns_f, ns_f, ns_f + 1) if lsw["restarts"] else
                               "ozaki_i8_gemm_v4_kernel<%d,%d,1> forward / <%d,%d,1> reverse" % (ns_f, ns_f + 1, ns_b, ns_b + 1)) +
                              " (tcgen05.mma.cta_group::2 kind::i8, TMA, TMEM; persistent CTA pairs, register-stash epilogue)",
                    "peak_source": peak8_src,
                    "peak_note": "the driver's sustained cuBLAS bf16 figure is itself power-capped (1335 MHz median), so "
                                 "frac can exceed 1 when this kernel holds higher clocks under the same 1 kW cap; "
                                 "frac_of_own_tcgen05_i8_issue_peak is against the raw kind::i8 issue rate measured "
                                 "with an MMA-only loop (profiles/i8_peaks.json)",
                    "frac_of_own_tcgen05_i8_issue_peak": achieved / peak8_issue if peak8_issue else None,
                    "algorithmic": "n^2 q flops per restart per launch x %.2f (forward) / %.2f (reverse) exact int8 digit-plane products per FP64 product (ns=%d/%d planes; level switch: %s)" % (pairs_f, pairs_b, ns_f, ns_b, "on" if lsw["restarts"] else "off"),
                    "products": {"forward": pairs_f, "reverse": pairs_b},
                    # what the int8 pipe allows with every other kernel free: issue rate / (products per FP64 flop)
                    "pipe_ceiling": None if not peak8_issue else {
                        "fp64_equivalent_tflops": peak8_issue / pairs,
                        "evals_per_s_per_gpu": M * peak8_issue * 1e12 / (npad * npad * (pairs_f + pairs_b)),
                        "note": "raw tcgen05 kind::i8 issue rate (profiles/i8_peaks.json) / products; target of BASELINE.json is 1e9"},
                    "fp64_equivalent_tflops": fp64_equiv, "fp64_dgemm_peak_tflops": peak64,
                    "fp64_equivalent_over_dgemm_peak": fp64_equiv / peak64}
    else:
        roofline = {"bound": "tensor", "achieved": fp64_equiv, "peak": peak64, "unit": "TFLOP/s", "frac": fp64_equiv / peak64,
                    "traffic": load_traffic("f64"), "kernel": "gemm_nt_f64_kernel (DMMA.8x8x4)", "peak_source": peak64_src}
    roofline.update(common)

    # ---- e2e: host buffers through the C ABI -----------------------------------------------------
    hX = eng.pinned_empty((S, q, d))
    hz = eng.pinned_empty((M, q))
    hloss = eng.pinned_empty((S,))
    hgrad = eng.pinned_empty((S, q, d))
    hX[...] = X
    hz[...] = z
    e2e_steps = max(2, args.steps)
    eng.value_and_grad(prm, hX, hz, out_loss=hloss, out_grad=hgrad)
    mdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.value_and_grad(prm, hX, hz, out_loss=hloss, out_grad=hgrad)
        i = int(np.argmin(hloss))
        if ws > 1:
            mdist.select_global_best(float(hloss[i]), offset + i, hX[i])
    torch.cuda.synchronize()
    e2e_s = mdist.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": S_total * q * M * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(hX.nbytes + hz.nbytes),
           "d2h_bytes_per_step": int(hloss.nbytes + hgrad.nbytes), "steps": e2e_steps}

    # ---- Adam loop on the device (steps x S x q x M / time), fresh z per step, same S --------------
    adam_steps = 4                                          # >= 4: one captured step replayed as a CUDA graph
    zs = np.random.RandomState(w["seed"] + 7).randn(adam_steps, M, q)
    eng.adam_begin(prm, X)
    eng.adam_steps(zs)                                      # warm-up call (buffers, capture)
    mdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.adam_steps(zs)
    torch.cuda.synchronize()
    adam_s = mdist.max_over_ranks(time.perf_counter() - t0)
    eng.adam_end()
    adam_loop = {"value": S_total * q * M * adam_steps / adam_s, "unit": UNIT, "steps": adam_steps,
                 "note": "maf_adam_steps: value+gradient, Adam update and clip per step on the device (one captured step replayed as a CUDA graph); z uploaded once per call"}

    # ---- CPU baseline (rank 0, N=1) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, S_cpu, dt = cpu_reference_rate(w, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "S=%d restarts (of %d), one value+autograd-gradient pass in %.1f s, same n/d/q/M/z" % (S_cpu, S, dt)}

    agent_times = None
    if rank == 0 and ws == 1 and not args.no_agent:
        try:
            agent_times = agent_suggest_inputs_times(local)
        except Exception as exc:                            # reported, never fatal for the headline
            agent_times = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": ("f32-grade (contractions: 5 int8 digit planes per operand = 39-bit fixed point; everything else f64)" if f32 else
                      "f64" if args.gemm == "f64" else "f64 (contractions: exact int8 digit planes, int32 accumulate, FP64 recombination)"),
            "data": "synthetic",
            "config": {"workload": ("qEI value+grad, n=%d d=%d q=%d M=%d matern52, level=%s of y0, " % (n, d, q, M, w["level"])) +
                                   ("S=%d restarts in total sharded over the ranks (BASELINE config 5)" % S_total if strong else "S=%d restarts per GPU" % S),
                       "contraction": args.gemm,
                       "parallelism": "restarts sharded x%d, replicated training factor, one device-side all-gather of the selection records per step" % ws,
                       "l2_policy": "working set per chunk (K10 + V^T = %.1f GB) >> 126 MB L2; no flush needed" % (2 * min(S * q, 65536) * npad * 8 / 1e9),
                       "setup_s_not_timed": setup_s},
            "level_switch": None if not lsw["restarts"] else {
                "restarts": int(S * args.steps), "tested_forward": lsw["tested_forward"], "recomputed_forward": lsw["recomputed_forward"],
                "tested_reverse": lsw["tested_reverse"], "recomputed_reverse": lsw["recomputed_reverse"],
                "tol": lsw["tol"], "calibrated_err_cov_abs": lsw["err_cov_abs"], "calibrated_err_grad_per_scale": lsw["err_grad_per_scale"],
                "products_forward_avg": pairs_f, "products_reverse_avg": pairs_b,
                "note": "both contractions first keep %d of %d digit planes (%d exact products); restarts whose own result does not tolerate the calibrated error of that level within tol are recomputed on all planes (%d) inside the same call" % (ns_f - 1, ns_f, (ns_f - 1) * ns_f // 2, ns_f * (ns_f + 1) // 2)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "adam_loop": adam_loop, "agent_suggest_inputs": agent_times, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if ws > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
