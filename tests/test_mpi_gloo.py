"""world_size-2 gloo tests (CPU) of the host-side data-parallel helpers that replace sac/mpi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch_actor_critic_b200.sac import mpi
    try:
        assert mpi.proc_id() == rank and mpi.num_procs() == world
        # sync_params: everyone ends with rank 0's parameters (sac/mpi.py:89-98)
        torch.manual_seed(100 + rank)
        net = nn.Sequential(nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 3))
        ref0 = None
        if rank == 0:
            ref0 = [p.detach().clone() for p in net.parameters()]
        mpi.sync_params(net)
        gathered = [None] * world
        dist.all_gather_object(gathered, [p.detach().clone() for p in net.parameters()])
        for a, b in zip(gathered[0], gathered[1]):
            assert torch.equal(a, b)
        if rank == 0:
            for a, b in zip(ref0, gathered[0]):
                assert torch.equal(a, b)
        # mpi_avg_grads: per-parameter mean over ranks (sac/mpi.py:77-85)
        x = torch.full((4, 5), float(rank + 1))
        net(x).sum().backward()
        mine = [p.grad.clone() for p in net.parameters()]
        allg = [None] * world
        dist.all_gather_object(allg, mine)
        mpi.mpi_avg_grads(net)
        for i, p in enumerate(net.parameters()):
            want = (allg[0][i] + allg[1][i]) / 2
            assert torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7)
        # scalar helpers
        assert mpi.mpi_sum(float(rank + 1)) == 3.0
        assert mpi.mpi_avg(float(rank + 1)) == 1.5
        v = mpi.mpi_avg(np.array([rank, 2 * rank], np.float32))
        assert np.allclose(v, [0.5, 1.0])
        mean, std, lo, hi = mpi.mpi_statistics_scalar([1.0 + rank, 3.0 + rank], with_min_and_max=True)
        allv = np.array([1, 3, 2, 4], np.float32)
        assert np.isclose(mean, allv.mean()) and np.isclose(std, allv.std()) and lo == 1 and hi == 4
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_mpi_helpers_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_single_process_defaults():
    from torch_actor_critic_b200.sac import mpi
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        pytest.skip("under torchrun")
    assert mpi.proc_id() == 0 and mpi.num_procs() == 1
    net = nn.Linear(2, 2)
    mpi.sync_params(net)          # early return, sac/mpi.py:94-95
    mpi.mpi_avg_grads(net)        # early return, sac/mpi.py:79-80
    assert mpi.mpi_avg(3.0) == 3.0
    mpi.mpi_fork(1)               # n <= 1: no-op, sac/mpi.py:14-15
