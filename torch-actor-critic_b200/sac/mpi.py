"""Data-parallel helpers with the names of the reference's sac/mpi.py, on torch.distributed instead of mpi4py:
one process per GPU, NCCL over NVLink for device tensors (gloo for CPU tensors in tests).

Inside a training burst the gradient averaging does not go through this module at all: by default the engine sums the
ranks' flat gradient arenas INSIDE the persistent update kernel (CUDA-IPC peer loads over NVLink in fixed rank order,
`Engine.peer_init_from_torch`); `SACB_COMM=nccl` selects ncclAllReduce between the backward and Adam launches instead
(`Engine.comm_init_from_torch`).  These functions serve the host-side callers (parameter broadcast at start, scalar
statistics) and keep the reference API.
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist


def _ensure_group():
    """Join the default process group when launched under torchrun (RANK/WORLD_SIZE in the environment)."""
    if dist.is_available() and not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return dist.is_available() and dist.is_initialized()


def mpi_fork(n, bind_to_core=False):
    """Re-launch the script as n workers (reference sac/mpi.py:10-34 uses `mpirun -np n`; here torchrun, one
    process per GPU).  The parent waits for the workers and exits."""
    n = int(n)
    if n <= 1:
        return
    if os.getenv("IN_MPI") is None:
        env = os.environ.copy()
        env.update(MKL_NUM_THREADS="1", OMP_NUM_THREADS="1", IN_MPI="1")
        args = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511")] + sys.argv
        p = subprocess.Popen(args, env=env)
        try:
            p.wait()
        except KeyboardInterrupt:
            p.kill()
        sys.exit()


def proc_id():
    return dist.get_rank() if _ensure_group() else 0


def num_procs():
    return dist.get_world_size() if _ensure_group() else 1


def _comm_device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce(x: torch.Tensor, op=None):
    """In-place all-reduce of a tensor (sum by default)."""
    if num_procs() > 1:
        dist.all_reduce(x, op=op or dist.ReduceOp.SUM)
    return x


def mpi_op(x, op):
    x, scalar = ([x], True) if np.isscalar(x) else (x, False)
    arr = np.asarray(x, dtype=np.float32)
    if num_procs() > 1:
        t = torch.from_numpy(arr.copy()).to(_comm_device())
        dist.all_reduce(t, op=op)
        arr = t.cpu().numpy()
    return arr[0] if scalar else arr


def mpi_sum(x):
    return mpi_op(x, dist.ReduceOp.SUM)


def mpi_avg(x):
    """Average a scalar or vector over the workers."""
    return mpi_sum(x) / num_procs()


def setup_pytorch_for_mpi():
    """Avoid slowdown from oversubscribed CPU threads (reference sac/mpi.py:67-74)."""
    if torch.get_num_threads() == 1:
        return
    torch.set_num_threads(max(int(torch.get_num_threads() / num_procs()), 1))


def mpi_avg_grads(module):
    """Average `.grad` of every parameter across workers, as ONE flat all-reduce (the reference issues one
    Allreduce per parameter, sac/mpi.py:77-85)."""
    if num_procs() == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= num_procs()
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def broadcast(x: torch.Tensor, root=0):
    if num_procs() > 1:
        dist.broadcast(x, src=root)


def sync_params(module):
    """Broadcast rank 0's parameters (reference sac/mpi.py:89-98).  Arena-backed modules broadcast their flat arena
    in one message."""
    if num_procs() == 1:
        return
    arena = getattr(module, "arena", None)
    if isinstance(arena, torch.Tensor):
        broadcast(arena)
        return
    for p in module.parameters():
        broadcast(p.data)


def mpi_statistics_scalar(x, with_min_and_max=False):
    """Mean / std (and optional min / max) of a scalar sample across workers (reference sac/mpi.py:101-115)."""
    x = np.array(x, dtype=np.float32)
    global_sum, global_n = mpi_sum([np.sum(x), len(x)])
    mean = global_sum / global_n
    global_sum_sq = mpi_sum(np.sum((x - mean) ** 2))
    std = np.sqrt(global_sum_sq / global_n)
    if with_min_and_max:
        global_min = mpi_op(np.min(x) if len(x) > 0 else np.inf, op=dist.ReduceOp.MIN)
        global_max = mpi_op(np.max(x) if len(x) > 0 else -np.inf, op=dist.ReduceOp.MAX)
        return mean, std, global_min, global_max
    return mean, std
