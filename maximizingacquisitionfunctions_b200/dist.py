"""Multi-GPU plumbing: restarts are independent units (src/misc/sharded_fn.py:106-132 maps
over them; the loss is a per-restart vector), so they are partitioned contiguously across
ranks with a replicated training set and NO collective inside the optimisation loop.  After
the final re-evaluation one all-gather of (best loss, global restart index, X*[q,d]) per rank
selects the global arg-min with np.argmin's tie-break (lowest index;
src/agents/bayesian_optimization.py:152).

torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process => (0,1,0))."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def init_process_group(backend: Optional[str] = None):
    """Initialise torch.distributed from the torchrun env if WORLD_SIZE > 1 (idempotent)."""
    rank, ws, local = world()
    if ws == 1:
        return None
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    return dist


def shard_range(S: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous partition of S restarts: the first (S mod W) ranks get one extra."""
    base, rem = divmod(S, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_best(losses: np.ndarray, offset: int) -> Tuple[float, int]:
    """First minimum of this rank's losses, NaNs never winning; returns (value, GLOBAL index).
    An empty shard or an all-NaN shard reports (+inf, -1)."""
    losses = np.asarray(losses, dtype=np.float64).reshape(-1)
    if losses.size == 0 or np.all(np.isnan(losses)):
        return float("inf"), -1
    i = int(np.nanargmin(losses))
    return float(losses[i]), offset + i


def select_global_best(best_loss: float, best_index: int, best_X: np.ndarray, device=None):
    """One all-gather of the per-rank (loss, global index, X*[q,d]) records; every rank returns the
    same (loss, index, X*).  Ties go to the lowest global index."""
    rank, ws, _ = world()
    best_X = np.asarray(best_X, dtype=np.float64)
    if ws == 1:
        return float(best_loss), int(best_index), best_X
    import torch
    import torch.distributed as dist
    rec = np.concatenate([[best_loss, float(best_index)], best_X.reshape(-1)])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(rec).to(device)
    out = torch.empty(ws * rec.size, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, t)
    out = out.cpu().numpy().reshape(ws, rec.size)
    return pick_record(out, best_X.shape)


def active() -> Tuple[int, int]:
    """(rank, world_size) of the restart sharding the agent layer applies: the torchrun environment when it names more
    than one rank (the process group is initialised on first use), else (0, 1).  MAF_SHARD_RESTARTS=0 switches it off."""
    rank, ws, _ = world()
    if ws == 1 or os.environ.get("MAF_SHARD_RESTARTS", "1") == "0":
        return 0, 1
    init_process_group()
    return rank, ws


def _device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def any_rank(flag: bool) -> bool:
    """True on every rank if `flag` is true on any (keeps data-dependent loop exits -- time limits -- in lock step, so
    that the host-RNG streams, and with them the base samples, stay identical across ranks)."""
    rank, ws, _ = world()
    if ws == 1:
        return bool(flag)
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))


def all_gather_concat(local: np.ndarray, total: int) -> np.ndarray:
    """Concatenation over ranks of the contiguous shards `shard_range(total, rank, W)` of a float64 vector."""
    rank, ws, _ = world()
    local = np.ascontiguousarray(local, dtype=np.float64).reshape(-1)
    if ws == 1:
        return local
    import torch
    import torch.distributed as dist
    width = -(-total // ws)
    buf = np.zeros(width)
    buf[:local.size] = local
    t = torch.from_numpy(buf).to(_device())
    out = torch.empty(ws * width, dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    out = out.cpu().numpy().reshape(ws, width)
    parts = []
    for r in range(ws):
        lo, hi = shard_range(total, r, ws)
        parts.append(out[r, :hi - lo])
    return np.concatenate(parts)


def sync_seed(module) -> None:
    """Give every rank the furthest-advanced host-RNG seed of `module` (src/core/module.py:140-144 advances the seed on
    every access), so the next draw -- the base samples of the final re-evaluation -- is the same everywhere."""
    rank, ws, _ = world()
    if ws == 1:
        return
    module.seed = int(max_over_ranks(float(module.seed)))


def global_best(engine, losses_local: np.ndarray, X_local: np.ndarray, offset: int, pool_shape=None):
    """Global first minimum over the ranks' shards: (loss, global index, X*[q_new, d]) on every rank, by ONE all-gather
    of the per-rank records [loss, global index, pool[q][d]].  With NCCL the record of a CUDA engine is assembled on the
    device from the losses and candidate sets of its last evaluation (maf_best_record_dev) and gathered without leaving
    the device; the W x (2 + q d) table is read back once.  CPU engines (tests, gloo) build the record on the host."""
    rank, ws, _ = world()
    X_local = np.asarray(X_local, dtype=np.float64)
    q_new, d = X_local.shape[1:] if X_local.ndim == 3 else pool_shape
    q = pool_shape[0] if pool_shape is not None else q_new
    if ws == 1:
        val, gidx = local_best(losses_local, offset)
        return val, gidx, X_local[gidx - offset]
    import torch
    import torch.distributed as dist
    on_device = (dist.get_backend() == "nccl" and engine is not None and hasattr(engine, "best_record_dev")
                 and getattr(engine, "last_eval_shape", None) == (len(X_local), q, d))
    if on_device:
        rec = torch.empty(2 + q * d, dtype=torch.float64, device=_device())
        engine.best_record_dev(offset, rec.data_ptr())
        engine.sync()                                       # the library's stream -> visible to torch's stream
    else:
        val, gidx = local_best(losses_local, offset)
        pool = np.zeros((q, d))
        if gidx >= 0:
            pool[q - q_new:] = X_local[gidx - offset]
        rec = torch.from_numpy(np.concatenate([[val, float(gidx)], pool.reshape(-1)])).to(_device())
    out = torch.empty(ws * rec.numel(), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(out, rec)
    table = out.cpu().numpy().reshape(ws, rec.numel())
    loss, idx, pool = pick_record(table, (q, d))
    return loss, idx, pool[q - q_new:]


def pick_record(records: np.ndarray, x_shape) -> Tuple[float, int, np.ndarray]:
    """arg-min over gathered records [W, 2 + q*d] with lowest-index tie-break; index -1 = empty."""
    best = None
    for r in records:
        loss, idx = float(r[0]), int(r[1])
        if idx < 0 or np.isnan(loss):
            continue
        if best is None or loss < best[0] or (loss == best[0] and idx < best[1]):
            best = (loss, idx, r[2:].reshape(x_shape).copy())
    if best is None:
        raise ValueError("no rank reported a finite loss")
    return best


def max_over_ranks(value: float) -> float:
    rank, ws, _ = world()
    if ws == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    rank, ws, _ = world()
    if ws > 1:
        import torch.distributed as dist
        dist.barrier()
