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
