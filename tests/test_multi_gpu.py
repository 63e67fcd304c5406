"""Multi-GPU data-parallel parity at world = 2, 4, 8 (run with -m gpu on a box with enough GPUs; skipped otherwise).

Each rank owns its own replay shard and batch (one reference "MPI rank" per GPU, sac/algorithm.py:203, sac/mpi.py);
the engine all-reduces the flat critic / actor gradient arenas with NCCL between backward and Adam.  Checked against
the float64 oracle run with rank-averaged gradients, and replicas must stay bit-identical.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from golden_helpers import assert_close

pytestmark = pytest.mark.gpu

SHAPES = {
    "toy": dict(O=11, A=3, H=(64, 48), B=32, N=512, K=3),
    "c1": dict(O=17, A=6, H=(256, 256), B=256, N=4096, K=3),        # BASELINE configs[1]/[3] shapes (HalfCheetah, batch 256 per GPU)
}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_data(rank, cfg):
    O, A, B, N, K = cfg["O"], cfg["A"], cfg["B"], cfg["N"], cfg["K"]
    rng = np.random.default_rng(100 + rank)
    arrays = dict(state=rng.standard_normal((N, O)).astype(np.float32), actions=rng.uniform(-1, 1, (N, A)).astype(np.float32),
                  rewards=rng.standard_normal(N).astype(np.float32), next_state=rng.standard_normal((N, O)).astype(np.float32),
                  done=rng.random(N) < 0.1)
    idx = np.stack([rng.choice(N, B, replace=False) for _ in range(K)])
    eps = rng.standard_normal((K, 2, B, A)).astype(np.float32)
    return arrays, idx, eps


def _worker(rank, world, port, flags, comm, q, cfg, dead_rank=-1):
    try:
        O, A, H, B, N, K = cfg["O"], cfg["A"], cfg["H"], cfg["B"], cfg["N"], cfg["K"]
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from oracle import sac_numpy as onp
        from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
        actor, critic = onp.init_params(O, A, H, np.random.default_rng(7 + rank))     # ranks start DIFFERENT ...
        eng = Engine(O, A, list(H), batch_size=B, max_burst=K, device=rank, act_limit=1.0, rank=rank, world=world, flags=flags)
        pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
        for k, v in actor_views(eng.layout, pa, O, A, list(H)).items():
            v.copy_(torch.from_numpy(actor[k]))
        for k, v in critic_views(eng.layout, pc, O, A, list(H)).items():
            v.copy_(torch.from_numpy(critic[k]))
        pt = pc.clone()
        eng.bind_params(pa, pc, pt)
        eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
        eng.comm_init_from_torch()
        eng.sync_params()                                                             # ... sync_params makes them rank 0's
        if comm == "peer":
            eng.peer_init_from_torch()                                                # fused in-kernel NVLink exchange
        arrays, idx, eps = _rank_data(rank, cfg)
        dev = eng.tdev()
        t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in arrays.items()}
        eng.bind_buffer(t["state"], t["actions"], t["rewards"], t["next_state"], t["done"])
        eng.set_buffer_state(0, N)
        if dead_rank >= 0:
            # a rank that never joins the exchange: its peers must time out with SACB_ERR_PEER instead of spinning forever
            from torch_actor_critic_b200 import _lib
            if rank == dead_rank:
                q.put((rank, "ok", "dead"))
            else:
                eng.update(1, idx=idx[:1], eps=eps[:1])
                try:
                    eng.peer_status()
                    q.put((rank, "ok", "no error raised"))
                except _lib.SacbError as ex:
                    q.put((rank, "ok", ex.code, str(ex)))
            dist.barrier()
            dist.destroy_process_group()
            return
        lq, lp = eng.update(K, idx=idx, eps=eps)
        torch.cuda.synchronize()
        q.put((rank, "ok", lq.cpu().numpy(), lp.cpu().numpy(), pa.cpu().numpy(), pc.cpu().numpy(), pt.cpu().numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


def _oracle(world, sync_actor, cfg):
    O, A, H, K = cfg["O"], cfg["A"], cfg["H"], cfg["K"]
    from oracle import sac_numpy as onp
    hp = onp.HyperParams(act_limit=1.0)
    actor, critic = onp.init_params(O, A, H, np.random.default_rng(7))               # rank 0's parameters
    actor, critic = onp.cast_params(actor, np.float64), onp.cast_params(critic, np.float64)
    # replicas: identical critic/target; actors diverge only when the actor gradient is NOT synced (reference quirk 3)
    actors = [onp.cast_params(actor, np.float64) for _ in range(world)]
    target = onp.cast_params(critic, np.float64)
    opt_q = onp.AdamState.zeros_like(critic)
    opt_pi = [onp.AdamState.zeros_like(actor) for _ in range(world)]
    data = [_rank_data(r, cfg) for r in range(world)]
    losses_q = [[] for _ in range(world)]
    losses_pi = [[] for _ in range(world)]
    for k in range(K):
        batches = []
        for r in range(world):
            arrays, idx, eps = data[r]
            ii = idx[k]
            batches.append(dict(states=arrays["state"][ii].astype(np.float64), actions=arrays["actions"][ii].astype(np.float64),
                                rewards=arrays["rewards"][ii].astype(np.float64), next_states=arrays["next_state"][ii].astype(np.float64),
                                done=arrays["done"][ii].astype(np.float64)))
        gq = []
        for r in range(world):
            l, g, _ = onp.q_loss_and_grads(actors[r], critic, target, batches[r], data[r][2][k, 0].astype(np.float64), hp)
            losses_q[r].append(l); gq.append(g)
        avg = {n: sum(g[n] for g in gq) / world for n in gq[0]}                      # mpi_avg_grads(critic), sac/algorithm.py:138
        onp.adam_step(critic, avg, opt_q, hp)
        gp = []
        for r in range(world):
            l, g, _ = onp.pi_loss_and_grads(actors[r], critic, batches[r], data[r][2][k, 1].astype(np.float64), hp)
            losses_pi[r].append(l); gp.append(g)
        if sync_actor:
            avgp = {n: sum(g[n] for g in gp) / world for n in gp[0]}
            gp = [avgp] * world
        for r in range(world):
            onp.adam_step(actors[r], gp[r], opt_pi[r], hp)
        onp.polyak_update(critic, target, hp.polyak)
    return actors, critic, target, losses_q, losses_pi


def _flatten(eng_layout_views, params):
    return np.concatenate([np.asarray(params[k], np.float64).reshape(-1) for k in eng_layout_views])


def _spawn(world, flags, comm, cfg, dead_rank=-1, env=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        procs = [ctx.Process(target=_worker, args=(r, world, port, flags, comm, q, cfg, dead_rank)) for r in range(world)]
        for p in procs:
            p.start()
        res = sorted([q.get(timeout=600) for _ in procs], key=lambda x: x[0])
        for p in procs:
            p.join(timeout=120)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for r in res:
        assert r[1] == "ok", f"rank {r[0]}: {r[1]}"
    return res


CASES = [("toy", 2, True, "nccl"), ("toy", 2, False, "nccl"), ("toy", 2, True, "peer"),
         ("c1", 2, True, "peer"), ("c1", 4, True, "peer"), ("c1", 8, True, "peer"), ("c1", 4, True, "nccl"), ("toy", 8, True, "peer")]


@pytest.mark.parametrize("shape,world,sync_actor,comm", CASES)
def test_data_parallel_update_matches_oracle(shape, world, sync_actor, comm):
    """comm = "nccl": ncclAllReduce between the backward and Adam launches; comm = "peer": the exchange fused into the
    persistent kernel (system-scope flags + NVLink loads of the peers' gradient arenas, one launch per burst).  Checked
    against the float64 oracle with rank-averaged gradients (mpi_avg_grads, sac/mpi.py:77-85) at the BASELINE shapes, and
    every replica must hold bit-identical parameters afterwards."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from torch_actor_critic_b200 import _lib
    from torch_actor_critic_b200.engine import actor_views, critic_views, layout_for
    from oracle.sac_numpy import ACTOR_KEYS, CRITIC_KEYS
    cfg = SHAPES[shape]
    O, A, H, K = cfg["O"], cfg["A"], cfg["H"], cfg["K"]
    flags = 0 if sync_actor else _lib.FLAG_NO_ACTOR_GRAD_SYNC
    res = _spawn(world, flags, comm, cfg)
    actors, critic, target, lq, lp = _oracle(world, sync_actor, cfg)
    lay = layout_for(O, A, list(H))
    for r in range(world):
        _, _, glq, glp, pa, pc, pt = res[r]
        assert_close(glq, lq[r], 1e-3, 1e-4, f"loss_q rank {r}")
        assert_close(glp, lp[r], 1e-3, 1e-4, f"loss_pi rank {r}")
        av = actor_views(lay, torch.from_numpy(pa), O, A, list(H))
        cv = critic_views(lay, torch.from_numpy(pc), O, A, list(H))
        tv = critic_views(lay, torch.from_numpy(pt), O, A, list(H))
        for k in ACTOR_KEYS:
            assert_close(av[k].numpy(), actors[r][k], 1e-5, 3e-4 * K, f"actor {k} rank {r}")
        for k in CRITIC_KEYS:
            assert_close(cv[k].numpy(), critic[k], 1e-5, 3e-4 * K, f"critic {k} rank {r}")
            assert_close(tv[k].numpy(), target[k], 1e-5, 1e-5, f"target {k} rank {r}")
    # replicas never drift: every rank sums the same gradients in the same (rank) order
    for r in range(1, world):
        assert np.array_equal(res[0][5], res[r][5]) and np.array_equal(res[0][6], res[r][6]), f"critic/target replica {r} drifted"
        if sync_actor:
            assert np.array_equal(res[0][4], res[r][4]), f"actor replica {r} drifted"
        else:
            assert not np.array_equal(res[0][4], res[r][4])       # the reference's behaviour: actors diverge


def test_peer_exchange_times_out_on_a_dead_rank():
    """Rank 1 never joins the burst: rank 0 must come back with SACB_ERR_PEER (naming rank 1) instead of hanging."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from torch_actor_critic_b200 import _lib
    res = _spawn(2, 0, "peer", SHAPES["toy"], dead_rank=1, env={"SACB_PEER_TIMEOUT_MS": "400"})
    assert res[1][2] == "dead"
    assert res[0][2] == _lib.SACB_ERR_PEER, res[0]
    assert "rank 1" in res[0][3]
