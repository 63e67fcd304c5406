"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on the CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

`gym`, `mlflow`, `mpi4py` are not installed, so stub modules are placed in sys.modules (rank 0 / size 1:
`mpi_avg_grads` and `sync_params` early-return, sac/mpi.py:79-80,94-95) and `sac.algorithm`, `networks.linear`,
`buffer.replay_buffer` are then imported unmodified.  Determinism recipe: parameters come from a
numpy RNG and are loaded with load_state_dict; batch indices are injected by building `Batch` from explicit
indices; the rsample noise is reproduced by `torch.manual_seed(s)` immediately before each update call
(`Normal.rsample` draws `torch.empty(shape).normal_()`), and recorded.
"""
import os
import random
import sys
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import numpy as np
import torch

REF = "/root/reference"


from oracle.ref_stubs import install_stubs as _install_stubs


def install_stubs():
    _install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)


def import_reference():
    install_stubs()
    import sac.algorithm as alg
    from buffer.replay_buffer import Batch, ReplayBuffer
    from networks.linear import Actor, Critic, DoubleCritic
    return alg, Batch, ReplayBuffer, Actor, Critic, DoubleCritic


def make_buffer_arrays(rng, n, obs_dim, act_dim, act_limit, p_done=0.05):
    """Synthetic transitions (BASELINE.md section 3 recipe)."""
    return dict(
        state=rng.standard_normal((n, obs_dim)).astype(np.float32),
        actions=rng.uniform(-act_limit, act_limit, (n, act_dim)).astype(np.float32),
        rewards=rng.standard_normal(n).astype(np.float32),
        next_state=rng.standard_normal((n, obs_dim)).astype(np.float32),
        done=(rng.random(n) < p_done),
    )


def sd_to_np(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def run_case(name, obs_dim, act_dim, hidden, batch, n_buf, n_updates, act_limit, seed, full):
    from oracle.sac_numpy import init_params
    alg, Batch, ReplayBuffer, Actor, Critic, DoubleCritic = import_reference()
    from copy import deepcopy
    rng = np.random.default_rng(seed)
    actor_np, critic_np = init_params(obs_dim, act_dim, hidden, rng)
    arrays = make_buffer_arrays(rng, n_buf, obs_dim, act_dim, act_limit)
    idx = np.stack([rng.choice(n_buf, size=batch, replace=False) for _ in range(n_updates)]).astype(np.int64)

    torch.set_num_threads(1)
    actor = Actor(obs_dim, act_dim, list(hidden), act_limit=act_limit)
    critic = DoubleCritic(obs_dim, act_dim, list(hidden))
    actor.load_state_dict({k: torch.tensor(v) for k, v in actor_np.items()})
    critic.load_state_dict({k: torch.tensor(v) for k, v in critic_np.items()})
    target = deepcopy(critic)                                   # sac/algorithm.py:194-196
    for p in target.parameters():
        p.requires_grad = False
    pi_opt = torch.optim.Adam(actor.parameters(), lr=3e-4)      # main.py:93-95
    q_opt = torch.optim.Adam(critic.parameters(), lr=3e-4)
    sac = alg.SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=batch,
                  start_steps=0, steps_per_epoch=1, max_ep_len=1, update_after=0, update_every=1, save_every=1)

    buf = ReplayBuffer(n_buf, obs_dim, act_dim)
    buf.state[:], buf.actions[:], buf.rewards[:] = arrays["state"], arrays["actions"], arrays["rewards"]
    buf.next_state[:], buf.done[:] = arrays["next_state"], arrays["done"]
    buf.size = n_buf

    out = dict(obs_dim=obs_dim, act_dim=act_dim, hidden=np.array(hidden), batch=batch, n_buf=n_buf,
               n_updates=n_updates, act_limit=act_limit, seed=seed, idx=idx)
    eps_all = np.zeros((n_updates, 2, batch, act_dim), np.float32)
    captured = {}

    def hook_actor(_m, _inp, outp):
        if outp[1] is not None:
            captured.setdefault("actor", []).append((outp[0].detach().numpy().copy(), outp[1].detach().numpy().copy()))

    def hook_critic(tag):
        def fn(_m, _inp, outp):
            captured.setdefault(tag, []).append((outp[0].detach().numpy().copy(), outp[1].detach().numpy().copy()))
            if outp[0].requires_grad:
                for j, o in enumerate(outp):
                    o.register_hook(lambda g, j=j: captured.setdefault(tag + "_dq", []).append((j, g.numpy().copy())))
        return fn

    actor.register_forward_hook(hook_actor)
    critic.register_forward_hook(hook_critic("critic"))
    target.register_forward_hook(hook_critic("target"))

    per_step = []
    for k in range(n_updates):
        ii = idx[k]
        samples = Batch(torch.FloatTensor(buf.state[ii]), torch.FloatTensor(buf.actions[ii]),
                        torch.FloatTensor(buf.rewards[ii]), torch.FloatTensor(buf.next_state[ii]),
                        torch.FloatTensor(buf.done[ii]))               # == ReplayBuffer.sample with injected idx
        captured.clear()
        sq, sp = 1000 + 2 * k, 1001 + 2 * k
        torch.manual_seed(sq)
        eps_all[k, 0] = torch.empty(batch, act_dim).normal_().numpy()
        torch.manual_seed(sp)
        eps_all[k, 1] = torch.empty(batch, act_dim).normal_().numpy()

        torch.manual_seed(sq)
        loss_q = sac.update_critic(q_opt, actor, critic, target, samples)
        critic_grads = {n: p.grad.numpy().copy() for n, p in critic.named_parameters()}
        a2, logp_next = captured["actor"][0]
        q1t, q2t = captured["target"][0]
        q1, q2 = captured["critic"][0]
        dq = dict(captured["critic_dq"])
        torch.manual_seed(sp)
        loss_pi = sac.update_policy(pi_opt, actor, critic, samples)
        actor_grads = {n: p.grad.numpy().copy() for n, p in actor.named_parameters()}
        pi, logp_pi = captured["actor"][1]
        q1pi, q2pi = captured["critic"][1]
        alg.update_targets(critic, target, sac.polyak)
        per_step.append(dict(loss_q=float(loss_q), loss_pi=float(loss_pi), a2=a2, logp_next=logp_next, q1_target=q1t,
                             q2_target=q2t, q1=q1, q2=q2, dq1=dq[0], dq2=dq[1], pi=pi, logp_pi=logp_pi,
                             q1_pi=q1pi, q2_pi=q2pi, critic_grads=critic_grads, actor_grads=actor_grads,
                             actor=sd_to_np(actor), critic=sd_to_np(critic), target=sd_to_np(target)))

    out["eps"] = eps_all
    out["loss_q"] = np.array([s["loss_q"] for s in per_step], np.float32)
    out["loss_pi"] = np.array([s["loss_pi"] for s in per_step], np.float32)
    for key in ("a2", "logp_next", "q1_target", "q2_target", "q1", "q2", "dq1", "dq2", "pi", "logp_pi", "q1_pi", "q2_pi"):
        out[f"step0.{key}"] = per_step[0][key]
    # backup recovered from the reference's own dL/dq: dq = 2 (q - backup)/B  (sac/algorithm.py:70)
    out["step0.backup"] = (per_step[0]["q1"].astype(np.float64)
                           - batch * per_step[0]["dq1"].astype(np.float64) / 2).astype(np.float32)

    def put(prefix, d, full_):
        for n, v in d.items():
            if full_:
                out[f"{prefix}.{n}"] = v
            else:   # summaries: sum, abs-sum, strided subsample
                flat = v.reshape(-1).astype(np.float64)
                out[f"{prefix}.{n}.sum"] = flat.sum()
                out[f"{prefix}.{n}.abssum"] = np.abs(flat).sum()
                out[f"{prefix}.{n}.sub"] = v.reshape(-1)[:: max(1, v.size // 512)][:512].copy()

    if full:
        for n, v in actor_np.items():
            out[f"init.actor.{n}"] = v
        for n, v in critic_np.items():
            out[f"init.critic.{n}"] = v
        for n, v in arrays.items():
            out[f"buffer.{n}"] = v
    put("step0.critic_grads", per_step[0]["critic_grads"], full)
    put("step0.actor_grads", per_step[0]["actor_grads"], full)
    for tag in ("actor", "critic", "target"):
        put(f"step0.{tag}", per_step[0][tag], full)
        put(f"final.{tag}", per_step[-1][tag], full)
    if full:
        for nm, opt, mod in (("pi_opt", pi_opt, actor), ("q_opt", q_opt, critic)):
            for (n, p) in mod.named_parameters():
                st = opt.state[p]
                out[f"final.{nm}.exp_avg.{n}"] = st["exp_avg"].numpy().copy()
                out[f"final.{nm}.exp_avg_sq.{n}"] = st["exp_avg_sq"].numpy().copy()
                out[f"final.{nm}.step.{n}"] = float(st["step"])
        # extra forward-only contracts: deterministic / no-logprob / unbatched (tests/test_linear.py:7-16)
        x = torch.tensor(arrays["state"][:4])
        with torch.no_grad():
            pd, lpd = actor(x, deterministic=True)
            pn, none = actor(x, deterministic=True, with_logprob=False)
            assert none is None
            pu, lpu = actor(x[0], deterministic=True)
            qa, qb = critic(x, torch.tensor(arrays["actions"][:4]))
        out["fwd.x"], out["fwd.act"] = x.numpy(), arrays["actions"][:4]
        out["fwd.pi_det"], out["fwd.logp_det"] = pd.numpy(), lpd.numpy()
        out["fwd.pi_unbatched"], out["fwd.logp_unbatched"] = pu.numpy(), np.asarray(lpu.numpy())
        out["fwd.q1"], out["fwd.q2"] = qa.numpy(), qb.numpy()
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.1f} KiB, loss_q={out['loss_q']}, loss_pi={out['loss_pi']}")


def run_buffer_case():
    """ReplayBuffer.store/sample (buffer/replay_buffer.py:29-54): ring wrap, size saturation, python-RNG draw."""
    _, Batch, ReplayBuffer, *_ = import_reference()
    rng = np.random.default_rng(7)
    n, O, A, n_store, B = 10, 3, 2, 25, 6
    arrays = make_buffer_arrays(rng, n_store, O, A, 1.0, p_done=0.3)
    buf = ReplayBuffer(n, O, A)
    sizes, ptrs = [], []
    for i in range(n_store):
        buf.store(arrays["state"][i], arrays["actions"][i], arrays["rewards"][i], arrays["next_state"][i],
                  arrays["done"][i])
        sizes.append(buf.size)
        ptrs.append(buf.ptr)
    random.seed(123)
    idx = random.sample(range(buf.size), B)
    random.seed(123)
    b = buf.sample(B)
    out = dict(n=n, obs_dim=O, act_dim=A, n_store=n_store, batch=B, sizes=np.array(sizes), ptrs=np.array(ptrs),
               idx=np.array(idx), **{f"in.{k}": v for k, v in arrays.items()},
               **{f"final.{k}": getattr(buf, k).copy() for k in ("state", "actions", "rewards", "next_state", "done")},
               **{f"batch.{k}": getattr(b, k).numpy() for k in ("states", "actions", "rewards", "next_states", "done")})
    try:
        ReplayBuffer(4, O, A).sample(1)
        out["empty_sample_error"] = "none"
    except ValueError as e:
        out["empty_sample_error"] = "ValueError"
    path = os.path.join(ROOT, "tests", "golden", "replay_buffer.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.1f} KiB")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    run_case("small", obs_dim=5, act_dim=3, hidden=(32, 32), batch=16, n_buf=64, n_updates=4, act_limit=2.0,
             seed=11, full=True)
    run_case("c1", obs_dim=17, act_dim=6, hidden=(256, 256), batch=256, n_buf=4096, n_updates=3, act_limit=1.0,
             seed=0, full=False)
    run_case("ragged", obs_dim=11, act_dim=3, hidden=(48, 24), batch=37, n_buf=200, n_updates=2, act_limit=0.4,
             seed=5, full=True)
    run_buffer_case()
