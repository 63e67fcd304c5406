"""GPU parity tests (run with -m gpu on the B200 box): the CUDA engine, called through the C ABI, against
  (1) the golden vectors produced by the unmodified reference (tests/golden/*.npz) and
  (2) the float64 numpy oracle on the same seeded inputs.

Stated tolerance (north_star: "within a stated fp32 tolerance"): the engine computes in fp32 with FFMA and a
different summation order than torch's CPU GEMMs, so single-step quantities are compared at rtol 2e-4 / atol 2e-5
(losses, Q values, log-probs), gradients at rtol 1e-3 / atol 5e-6, one-Adam-step parameters at atol 2e-6
(|dp| = lr = 3e-4), and short trajectories at a few lr (Adam's g/sqrt(v) amplifies sign flips of ~0 gradients).
Integer / byte work (ring pointer arithmetic, gather, index sets) is bit-exact.
"""
import numpy as np
import pytest
import torch

from golden_helpers import GOLDEN, assert_close, batch_from, load_case, summary
from oracle import sac_numpy as onp
from oracle.sac_numpy import ACTOR_KEYS, CRITIC_KEYS

pytestmark = pytest.mark.gpu


def _engine_for(g, actor, critic, arrays, hp, K=None, flags=0):
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    O, A, B = int(g["obs_dim"]), int(g["act_dim"]), int(g["batch"])
    H = [int(h) for h in g["hidden"]]
    K = K or int(g["n_updates"])
    eng = Engine(O, A, H, batch_size=B, max_burst=K, device=0, act_limit=hp.act_limit, flags=flags)
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    av, cv = actor_views(eng.layout, pa, O, A, H), critic_views(eng.layout, pc, O, A, H)
    for k in ACTOR_KEYS:
        av[k].copy_(torch.from_numpy(actor[k]))
    for k in CRITIC_KEYS:
        cv[k].copy_(torch.from_numpy(critic[k]))
    pt = pc.clone()
    tv = critic_views(eng.layout, pt, O, A, H)
    eng.bind_params(pa, pc, pt)
    adam = [eng.new_actor_arena(), eng.new_actor_arena(), eng.new_critic_arena(), eng.new_critic_arena()]
    eng.bind_adam(adam[0], adam[1], 0, adam[2], adam[3], 0)
    dev = eng.tdev()
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in arrays.items()}
    eng.bind_buffer(t["state"], t["actions"], t["rewards"], t["next_state"], t["done"])
    eng.set_buffer_state(0, len(arrays["rewards"]))
    return eng, av, cv, tv, adam


def _check(g, prefix, name, value, rtol, atol):
    value = np.asarray(value)
    if f"{prefix}.{name}" in g:
        assert_close(value, g[f"{prefix}.{name}"], rtol, atol, f"{prefix}.{name}")
    else:
        s, a, sub = summary(value)
        assert_close(sub, g[f"{prefix}.{name}.sub"], rtol, atol, f"{prefix}.{name}.sub")
        assert abs(s - float(g[f"{prefix}.{name}.sum"])) <= atol * value.size + rtol * float(g[f"{prefix}.{name}.abssum"])


@pytest.mark.parametrize("name", ["small", "ragged", "c1"])
@pytest.mark.parametrize("mode,tf32x3,fold", [(0, 1, 1), (1, 1, 1), (0, 0, 1), (0, 2, 1), (1, 2, 1), (0, 1, 0), (1, 1, 0), (0, 0, 0)])
def test_update_matches_reference_golden(name, mode, tf32x3, fold, monkeypatch):
    """mode 0 = persistent kernel, 1 = one launch per phase; tf32x3: 0 = plain fp32 FFMA, 1 = mma.sync tensor-core
    GEMMs with 3xTF32 error compensation, 2 = tcgen05 (TMEM accumulators) with 3xTF32; fold 1 = the folded 8-phase program
    (row jobs and input layers generated inside the consumer GEMM tiles; applies to tf32x3 0/1), 0 = the first 16-phase
    program.  Same tolerances for all."""
    from torch_actor_critic_b200.engine import actor_views, critic_views
    monkeypatch.setenv("SACB_FOLD", str(fold))
    g, actor, critic, arrays, hp = load_case(name)
    eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp)
    eng.set_mode(mode)
    eng.set_precision(tf32x3)
    O, A = int(g["obs_dim"]), int(g["act_dim"])
    H = [int(h) for h in g["hidden"]]
    K = int(g["n_updates"])
    # ---- first update: every intermediate of sac/algorithm.py:61-72,37-41 ----
    lq, lp = eng.update(1, idx=g["idx"][:1], eps=g["eps"][:1])
    assert_close(lq.cpu().numpy(), g["loss_q"][:1], 2e-4, 2e-5, "loss_q[0]")
    assert_close(lp.cpu().numpy(), g["loss_pi"][:1], 2e-4, 2e-5, "loss_pi[0]")
    for key in ("a2", "logp_next", "q1_target", "q2_target", "backup", "q1", "q2", "pi", "logp_pi", "q1_pi", "q2_pi"):
        assert_close(eng.debug_read(key), g[f"step0.{key}"], 2e-4, 5e-5, key)
    ga = actor_views(eng.layout, torch.from_numpy(eng.debug_read("grad_actor")), O, A, H)
    gc = critic_views(eng.layout, torch.from_numpy(eng.debug_read("grad_critic")), O, A, H)
    for k in ACTOR_KEYS:
        _check(g, "step0.actor_grads", k, ga[k].numpy(), 1e-3, 5e-6)
    for k in CRITIC_KEYS:
        _check(g, "step0.critic_grads", k, gc[k].numpy(), 1e-3, 5e-6)
    for k in ACTOR_KEYS:
        _check(g, "step0.actor", k, av[k].cpu().numpy(), 1e-6, 2e-6)
    for k in CRITIC_KEYS:
        _check(g, "step0.critic", k, cv[k].cpu().numpy(), 1e-6, 2e-6)
        _check(g, "step0.target", k, tv[k].cpu().numpy(), 1e-6, 1e-7)
    # ---- remaining updates in one burst ----
    if K > 1:
        lq, lp = eng.update(K - 1, idx=g["idx"][1:], eps=g["eps"][1:])
        assert_close(lq.cpu().numpy(), g["loss_q"][1:], 1e-3, 1e-4, "loss_q[1:]")
        assert_close(lp.cpu().numpy(), g["loss_pi"][1:], 1e-3, 1e-4, "loss_pi[1:]")
    for k in ACTOR_KEYS:
        _check(g, "final.actor", k, av[k].cpu().numpy(), 1e-5, 3e-4 * K)
    for k in CRITIC_KEYS:
        _check(g, "final.critic", k, cv[k].cpu().numpy(), 1e-5, 3e-4 * K)
        _check(g, "final.target", k, tv[k].cpu().numpy(), 1e-5, 3e-4 * K * 0.005 * K + 1e-6)
    assert eng.adam_steps() == (K, K)
    assert eng.launch_count() > 0


def test_persistent_and_phased_modes_bit_identical():
    g, actor, critic, arrays, hp = load_case("ragged")
    outs = []
    for mode in (0, 1):
        eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp)
        eng.set_mode(mode)
        lq, lp = eng.update(int(g["n_updates"]), idx=g["idx"], eps=g["eps"])
        torch.cuda.synchronize()
        outs.append((lq.cpu(), lp.cpu(), eng._keep["actor"].cpu(), eng._keep["critic"].cpu(), eng._keep["target"].cpu()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_adam_moments_match_reference():
    from torch_actor_critic_b200.engine import actor_views, critic_views
    g, actor, critic, arrays, hp = load_case("small")
    eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp)
    O, A, H = int(g["obs_dim"]), int(g["act_dim"]), [int(h) for h in g["hidden"]]
    eng.update(int(g["n_updates"]), idx=g["idx"], eps=g["eps"])
    ma, va = actor_views(eng.layout, adam[0], O, A, H), actor_views(eng.layout, adam[1], O, A, H)
    mc = critic_views(eng.layout, adam[2], O, A, H)
    for k in ACTOR_KEYS:
        assert_close(ma[k].cpu().numpy(), g[f"final.pi_opt.exp_avg.{k}"], 2e-3, 2e-6, k)
        assert_close(va[k].cpu().numpy(), g[f"final.pi_opt.exp_avg_sq.{k}"], 4e-3, 1e-9, k)
    for k in CRITIC_KEYS:
        assert_close(mc[k].cpu().numpy(), g[f"final.q_opt.exp_avg.{k}"], 2e-3, 2e-6, k)


def test_canonical_logprob_flag_matches_oracle():
    """SACB_FLAG_CANONICAL_LOGPROB: the parenthesised correction (not what the reference computes)."""
    from torch_actor_critic_b200 import _lib
    g, actor, critic, arrays, hp = load_case("small")
    hp.logprob_reference_form = False
    eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp, flags=_lib.FLAG_CANONICAL_LOGPROB)
    lq, lp = eng.update(1, idx=g["idx"][:1], eps=g["eps"][:1])
    a64, c64, t64 = (onp.cast_params(p, np.float64) for p in (actor, critic, critic))
    o = onp.sac_update(a64, c64, t64, onp.AdamState.zeros_like(a64), onp.AdamState.zeros_like(c64),
                       batch_from(arrays, g["idx"][0], np.float64), g["eps"][0, 0].astype(np.float64),
                       g["eps"][0, 1].astype(np.float64), hp)
    assert_close(lq.cpu().numpy(), [o["loss_q"]], 2e-4, 2e-5)
    assert_close(lp.cpu().numpy(), [o["loss_pi"]], 2e-4, 2e-5)
    assert abs(float(lp) - float(g["loss_pi"][0])) > 1e-2          # and it does differ from the reference form


def test_piecewise_api_equals_burst():
    """SAC.update_critic / update_policy / update_targets on an explicit Batch == the fused burst (same arithmetic)."""
    import torch.optim as optim
    from torch_actor_critic_b200.buffer.replay_buffer import Batch
    from torch_actor_critic_b200.networks.linear import Actor, DoubleCritic
    from torch_actor_critic_b200.sac.algorithm import SAC, update_targets
    from copy import deepcopy
    g, actor_np, critic_np, arrays, hp = load_case("small")
    O, A, B = int(g["obs_dim"]), int(g["act_dim"]), int(g["batch"])
    H = [int(h) for h in g["hidden"]]
    actor = Actor(O, A, H, act_limit=hp.act_limit)
    critic = DoubleCritic(O, A, H)
    actor.load_state_dict({k: torch.tensor(v) for k, v in actor_np.items()})
    critic.load_state_dict({k: torch.tensor(v) for k, v in critic_np.items()})
    target = deepcopy(critic)
    assert isinstance(target, DoubleCritic) and target.arena.data_ptr() != critic.arena.data_ptr()
    pi_opt = optim.Adam(actor.parameters(), lr=3e-4)
    q_opt = optim.Adam(critic.parameters(), lr=3e-4)
    sac = SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=B, start_steps=0,
              steps_per_epoch=1, max_ep_len=1, update_after=0, update_every=4, save_every=1)
    for k in range(int(g["n_updates"])):
        b = batch_from(arrays, g["idx"][k])
        samples = Batch(*(torch.tensor(b[n]).cuda() for n in ("states", "actions", "rewards", "next_states", "done")))
        lq = sac.update_critic(q_opt, actor, critic, target, samples, eps=g["eps"][k, 0])
        lp = sac.update_policy(pi_opt, actor, critic, samples, eps=g["eps"][k, 1])
        update_targets(critic, target, sac.polyak)
        assert lq.dim() == 0 and lp.dim() == 0
        assert_close(lq.cpu().numpy(), g["loss_q"][k], 1e-3, 1e-4, "loss_q")
        assert_close(lp.cpu().numpy(), g["loss_pi"][k], 1e-3, 1e-4, "loss_pi")
    sd_a, sd_c, sd_t = actor.state_dict(), critic.state_dict(), target.state_dict()
    K = int(g["n_updates"])
    for k in ACTOR_KEYS:
        assert_close(sd_a[k].cpu().numpy(), g[f"final.actor.{k}"], 1e-5, 3e-4 * K, k)
    for k in CRITIC_KEYS:
        assert_close(sd_c[k].cpu().numpy(), g[f"final.critic.{k}"], 1e-5, 3e-4 * K, k)
        assert_close(sd_t[k].cpu().numpy(), g[f"final.target.{k}"], 1e-5, 1e-5, k)
    # optimizer state is visible to torch (checkpoint path, sac/algorithm.py:176-178)
    st = q_opt.state_dict()["state"]
    assert len(st) == 12 and float(st[0]["step"]) == K
    assert_close(st[0]["exp_avg"].cpu().numpy(), g["final.q_opt.exp_avg.q1.layers.0.weight"], 2e-3, 2e-6)
    assert float(pi_opt.state_dict()["state"][0]["step"]) == K


def test_forward_modules_match_reference():
    """Actor / DoubleCritic forward contracts incl. the reference's own shape tests (tests/test_linear.py:7-41)."""
    from torch_actor_critic_b200.networks.linear import Actor, Critic, DoubleCritic
    g, _, _, _, hp = load_case("small")
    O, A = int(g["obs_dim"]), int(g["act_dim"])
    H = [int(h) for h in g["hidden"]]
    actor = Actor(O, A, H, act_limit=hp.act_limit)
    critic = DoubleCritic(O, A, H)
    actor.load_state_dict({k: torch.tensor(g[f"final.actor.{k}"]) for k in ACTOR_KEYS})
    critic.load_state_dict({k: torch.tensor(g[f"final.critic.{k}"]) for k in CRITIC_KEYS})
    x = torch.tensor(g["fwd.x"])
    pi, logp = actor(x, deterministic=True)
    assert_close(pi.cpu().numpy(), g["fwd.pi_det"], 1e-4, 1e-5)
    assert_close(logp.cpu().numpy(), g["fwd.logp_det"], 1e-4, 5e-5)
    pn, none = actor(x, deterministic=True, with_logprob=False)
    assert none is None and torch.equal(pn, pi)
    pu, lu = actor(x[0], deterministic=True)
    assert list(pu.size()) == [A] and list(lu.size()) == []
    assert_close(pu.cpu().numpy(), g["fwd.pi_unbatched"], 1e-4, 1e-5)
    q1, q2 = critic(x, torch.tensor(g["fwd.act"]))
    assert_close(q1.cpu().numpy(), g["fwd.q1"], 1e-4, 1e-5)
    assert_close(q2.cpu().numpy(), g["fwd.q2"], 1e-4, 1e-5)
    # stochastic path: injected eps reproduces the oracle; device-drawn eps is ~N(0,1)
    eps = torch.randn(4, A)
    p2, l2 = actor(x, eps=eps)
    ao = {k: g[f"final.actor.{k}"] for k in ACTOR_KEYS}
    po, lo, _ = onp.actor_forward(ao, g["fwd.x"], eps.numpy(), hp)
    assert_close(p2.cpu().numpy(), po, 1e-4, 1e-5)
    assert_close(l2.cpu().numpy(), lo, 1e-4, 5e-5)
    # the reference's shape tests, verbatim sizes
    a = Actor(10, 20, [256, 256])
    p, lp = a(torch.ones((10,)))
    assert list(p.size()) == [20] and list(lp.size()) == []
    c = Critic(5, 10, [256, 256])
    assert list(c(torch.ones((2, 5)), torch.ones((2, 10))).size()) == [2]
    dc = DoubleCritic(5, 10, [256, 256])
    v1, v2 = dc(torch.ones((3, 5)), torch.ones((3, 10)))
    assert list(v1.size()) == [3] and list(v2.size()) == [3]
    # more rows than one kernel pass (chunking) and a 1000-row batch
    big = torch.randn(1000, 10)
    pb, lb = a(big, deterministic=True)
    ps, ls = a(big[300:301], deterministic=True)
    # (the single row runs through the fp32 cluster acting kernel, the batch through the 3xTF32 GEMM tiles)
    assert pb.shape == (1000, 20) and torch.allclose(pb[300], ps[0], rtol=1e-4, atol=1e-5) and torch.allclose(lb[300], ls[0], rtol=1e-4, atol=1e-4)


def test_replay_buffer_matches_reference_bit_exact():
    """store ring semantics + gather vs the reference's ReplayBuffer (golden replay_buffer.npz)."""
    import os
    from torch_actor_critic_b200.buffer.replay_buffer import Batch, ReplayBuffer
    g = dict(np.load(os.path.join(GOLDEN, "replay_buffer.npz")))
    buf = ReplayBuffer(int(g["n"]), int(g["obs_dim"]), int(g["act_dim"]))
    for i in range(int(g["n_store"])):
        buf.store(g["in.state"][i], g["in.actions"][i], g["in.rewards"][i], g["in.next_state"][i], g["in.done"][i])
        assert buf.size == g["sizes"][i] and buf.ptr == g["ptrs"][i]
    for k in ("state", "actions", "rewards", "next_state", "done"):
        assert np.array_equal(getattr(buf, k).cpu().numpy(), g[f"final.{k}"]), k
    b = buf.sample(int(g["batch"]), idx=g["idx"])
    assert isinstance(b, Batch)
    for k in ("states", "actions", "rewards", "next_states", "done"):
        v = getattr(b, k)
        assert v.dtype == torch.float32 and np.array_equal(v.cpu().numpy(), g[f"batch.{k}"]), k
    with pytest.raises(ValueError):
        ReplayBuffer(4, 3, 2).sample(1)                          # empty buffer, as random.sample raises
    with pytest.raises(ValueError):
        buf.sample(int(g["n"]) + 1)
    # vectorised store == repeated store (wrap included)
    buf2 = ReplayBuffer(int(g["n"]), int(g["obs_dim"]), int(g["act_dim"]))
    buf2.store_many(g["in.state"], g["in.actions"], g["in.rewards"], g["in.next_state"], g["in.done"])
    for k in ("state", "actions", "rewards", "next_state", "done"):
        assert np.array_equal(getattr(buf2, k).cpu().numpy(), g[f"final.{k}"]), k
    assert (buf2.ptr, buf2.size) == (buf.ptr, buf.size)


def test_device_sampler_without_replacement_and_uniform():
    from torch_actor_critic_b200.buffer.replay_buffer import ReplayBuffer
    n, B = 300, 256
    buf = ReplayBuffer(n, 2, 1, seed=3)
    obs = np.arange(n, dtype=np.float32)[:, None].repeat(2, 1)
    buf.store_many(obs, np.zeros((n, 1), np.float32), np.arange(n, dtype=np.float32), obs, np.zeros(n, bool))
    counts = np.zeros(n)
    for _ in range(200):
        b = buf.sample(B)
        r = b.rewards.cpu().numpy().astype(int)
        assert len(np.unique(r)) == B                           # duplicate-free, like random.sample
        assert np.array_equal(b.states.cpu().numpy()[:, 0].astype(int), r)      # rows stay together
        counts[r] += 1
    expect = 200 * B / n
    assert np.all(np.abs(counts - expect) < 6 * np.sqrt(expect * (1 - B / n)) + 1)   # uniform inclusion probability
    full = buf.sample(n)                                          # B == size: a permutation (dense regime, Fisher-Yates)
    assert np.array_equal(np.sort(full.rewards.cpu().numpy()), np.arange(n))
    # sparse regime (size >= 4B): rejection sampler
    n2 = 5000
    buf2 = ReplayBuffer(n2, 2, 1, seed=4)
    obs2 = np.arange(n2, dtype=np.float32)[:, None].repeat(2, 1)
    buf2.store_many(obs2, np.zeros((n2, 1), np.float32), np.arange(n2, dtype=np.float32), obs2, np.zeros(n2, bool))
    counts2 = np.zeros(n2)
    for _ in range(400):
        r = buf2.sample(B).rewards.cpu().numpy().astype(int)
        assert len(np.unique(r)) == B
        counts2[r] += 1
    e2 = 400 * B / n2
    assert np.all(np.abs(counts2 - e2) < 6 * np.sqrt(e2) + 1) and counts2.min() > 0


def _c1_setup(N=1_000_000, seed=0):
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    O, A, H, B = 17, 6, [256, 256], 256
    rng = np.random.default_rng(seed)
    actor, critic = onp.init_params(O, A, H, rng)
    eng = Engine(O, A, H, batch_size=B, max_burst=50, device=0, act_limit=1.0, seed=seed)
    dev = eng.tdev()
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    for k, v in actor_views(eng.layout, pa, O, A, H).items():
        v.copy_(torch.from_numpy(actor[k]))
    for k, v in critic_views(eng.layout, pc, O, A, H).items():
        v.copy_(torch.from_numpy(critic[k]))
    pt = pc.clone()
    eng.bind_params(pa, pc, pt)
    eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
    gen = torch.Generator(device=dev).manual_seed(seed)
    s = torch.randn(N, O, device=dev, generator=gen); s2 = torch.randn(N, O, device=dev, generator=gen)
    a = torch.rand(N, A, device=dev, generator=gen) * 2 - 1; r = torch.randn(N, device=dev, generator=gen)
    d = torch.rand(N, device=dev, generator=gen) < 1e-3
    eng.bind_buffer(s, a, r, s2, d)
    eng.set_buffer_state(0, N)
    return eng, pa, pc, pt


def test_full_size_c1_properties():
    """BASELINE config 1 (1e6-transition HBM buffer, B=256, MLP 256x256): size-independent properties."""
    eng, pa, pc, pt = _c1_setup()
    pc0, pt0, pa0 = pc.clone(), pt.clone(), pa.clone()
    lq, lp = eng.update(1)
    torch.cuda.synchronize()
    assert torch.isfinite(lq).all() and torch.isfinite(lp).all()
    # Polyak identity (sac/algorithm.py:80-81) holds exactly in fp32: targ' = fl(fl(0.995*targ) + fl(0.005*critic'))
    pk, pcf = torch.tensor(0.995, dtype=torch.float32).cuda(), torch.tensor(1.0 - 0.995, dtype=torch.float32).cuda()
    assert torch.equal(pt, pt0 * pk + pcf * pc)
    # first Adam step moves every parameter with a non-zero gradient by lr (bias-corrected m/sqrt(v) = sign(g))
    gc = torch.from_numpy(eng.debug_read("grad_critic")).cuda()
    nz = gc.abs() > 1e-4          # |g| >> adam eps (1e-8): the step is lr * g/(|g| + eps*sqrt(bc2)) ~= lr
    assert nz.float().mean() > 0.05
    assert torch.allclose((pc - pc0)[nz].abs(), torch.full_like(pc[nz], 3e-4), rtol=2e-3, atol=2e-8)
    assert torch.equal((pc - pc0)[nz].sign(), -gc[nz].sign())
    ga = torch.from_numpy(eng.debug_read("grad_actor")).cuda()
    nza = ga.abs() > 1e-4
    assert torch.allclose((pa - pa0)[nza].abs(), torch.full_like(pa[nza], 3e-4), rtol=2e-3, atol=2e-8)
    # a 50-update burst with device-drawn indices and noise: finite, and deterministic for a fixed seed
    lq1, lp1 = eng.update(50)
    eng2, pa2, pc2, pt2 = _c1_setup()
    eng2.update(1)
    lq2, lp2 = eng2.update(50)
    torch.cuda.synchronize()
    assert torch.isfinite(lq1).all() and torch.isfinite(lp1).all()
    assert torch.equal(lq1, lq2) and torch.equal(lp1, lp2) and torch.equal(pc, pc2) and torch.equal(pa, pa2)


def test_bursts_of_odd_lengths_persistent_equals_phased():
    """The next batch is gathered while the last phase of an update still reads this one (parity double buffer of the s'
    staging tile, fused-Adam program).  Bursts of odd and even lengths, back to back, must leave exactly the same
    parameters whether an update runs inside the persistent kernel or as one launch per phase."""
    states = []
    for mode in (0, 1):
        eng, pa, pc, pt = _c1_setup(N=50_000, seed=3)
        eng.set_mode(mode)
        losses = []
        for n in (7, 50, 1, 13, 2):
            lq, lp = eng.update(n)
            losses.append(torch.cat([lq, lp]).clone())
        torch.cuda.synchronize()
        states.append((pa.clone(), pc.clone(), pt.clone(), torch.cat(losses)))
    for x, y in zip(*states):
        assert torch.isfinite(x).all() and torch.equal(x, y)


def test_soak_10k_updates_stay_finite_and_target_tracks_critic():
    """200 bursts of 50 updates with device-drawn batches and noise on random-reward data: nothing blows up, the losses
    stay in the range the data allows (rewards ~ N(0,1)), and the Polyak target stays within a few percent of the critic."""
    eng, pa, pc, pt = _c1_setup(N=200_000, seed=5)
    last = None
    for _ in range(200):
        last = eng.update(50)
    torch.cuda.synchronize()
    lq, lp = last
    assert torch.isfinite(lq).all() and torch.isfinite(lp).all()
    assert torch.isfinite(pa).all() and torch.isfinite(pc).all() and torch.isfinite(pt).all()
    assert 0.1 < float(lq.mean()) < 50.0
    assert float((pt - pc).norm() / pc.norm()) < 0.2
    assert tuple(int(x) for x in eng.adam_steps()) == (10_000, 10_000)


def test_full_size_c1_first_update_matches_oracle():
    """One update at BASELINE config-1 sizes with injected indices/noise vs the float64 oracle."""
    eng, pa, pc, pt = _c1_setup(N=200_000, seed=1)
    from torch_actor_critic_b200.engine import actor_views, critic_views
    O, A, H, B = 17, 6, [256, 256], 256
    rng = np.random.default_rng(5)
    idx = rng.choice(200_000, B, replace=False)[None]
    eps = rng.standard_normal((1, 2, B, A)).astype(np.float32)
    s, a, r, s2, d = (t.cpu().numpy() for t in eng._keep["buf"])
    actor = {k: v.cpu().numpy().astype(np.float64) for k, v in actor_views(eng.layout, pa, O, A, H).items()}
    critic = {k: v.cpu().numpy().astype(np.float64) for k, v in critic_views(eng.layout, pc, O, A, H).items()}
    target = onp.cast_params(critic, np.float64)
    lq, lp = eng.update(1, idx=idx, eps=eps)
    batch = dict(states=s[idx[0]].astype(np.float64), actions=a[idx[0]].astype(np.float64), rewards=r[idx[0]].astype(np.float64),
                 next_states=s2[idx[0]].astype(np.float64), done=d[idx[0]].astype(np.float64))
    hp = onp.HyperParams(act_limit=1.0)
    o = onp.sac_update(actor, critic, target, onp.AdamState.zeros_like(actor), onp.AdamState.zeros_like(critic), batch,
                       eps[0, 0].astype(np.float64), eps[0, 1].astype(np.float64), hp)
    assert_close(lq.cpu().numpy(), [o["loss_q"]], 1e-4, 1e-5, "loss_q")
    assert_close(lp.cpu().numpy(), [o["loss_pi"]], 1e-4, 1e-5, "loss_pi")
    for key in ("backup", "q1", "q2", "logp_next", "logp_pi", "q1_pi", "q2_pi"):
        assert_close(eng.debug_read(key), o[key], 2e-4, 5e-5, key)
    gc = critic_views(eng.layout, torch.from_numpy(eng.debug_read("grad_critic")), O, A, H)
    for k in CRITIC_KEYS:
        assert_close(gc[k].numpy(), o["critic_grads"][k], 1e-3, 5e-6, k)
    ga = actor_views(eng.layout, torch.from_numpy(eng.debug_read("grad_actor")), O, A, H)
    for k in ACTOR_KEYS:
        assert_close(ga[k].numpy(), o["actor_grads"][k], 1e-3, 5e-6, k)


def test_errors_are_loud():
    from torch_actor_critic_b200 import _lib
    from torch_actor_critic_b200.engine import Engine
    eng = Engine(3, 2, [8, 8], batch_size=4, max_burst=2, device=0)
    with pytest.raises(_lib.SacbError) as ei:
        eng.update(1)
    assert ei.value.code == _lib.SACB_ERR_STATE
    with pytest.raises(NotImplementedError):
        Engine(3, 2, [8, 8, 8])
    with pytest.raises(_lib.SacbError) as ei:
        Engine(3, 2, [8, 8], batch_size=20000)                       # beyond the on-device sampler's 16384 rows
    assert ei.value.code == _lib.SACB_ERR_INVALID


@pytest.mark.parametrize("tf32x3", [1, 0, 2])
def test_humanoid_shaped_update_matches_oracle(tf32x3):
    """BASELINE config 2 shapes (obs=376, act=17, MLP 512x512, batch=1024) in fp32 storage: one update vs the float64
    oracle.  Exercises multi-tile heads (2A = 34 > 32), 16 partial-Q tiles and K = 393 / 512 / 1024 reductions."""
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    O, A, H, B, N = 376, 17, [512, 512], 1024, 4096
    rng = np.random.default_rng(3)
    actor, critic = onp.init_params(O, A, H, rng)
    arrays = dict(state=rng.standard_normal((N, O)).astype(np.float32), actions=rng.uniform(-0.4, 0.4, (N, A)).astype(np.float32),
                  rewards=rng.standard_normal(N).astype(np.float32), next_state=rng.standard_normal((N, O)).astype(np.float32),
                  done=rng.random(N) < 0.05)
    idx = rng.choice(N, B, replace=False)[None]
    eps = rng.standard_normal((1, 2, B, A)).astype(np.float32)
    eng = Engine(O, A, H, batch_size=B, max_burst=2, device=0, act_limit=0.4)
    eng.set_precision(tf32x3)
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    for k, v in actor_views(eng.layout, pa, O, A, H).items():
        v.copy_(torch.from_numpy(actor[k]))
    for k, v in critic_views(eng.layout, pc, O, A, H).items():
        v.copy_(torch.from_numpy(critic[k]))
    pt = pc.clone()
    eng.bind_params(pa, pc, pt)
    eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}
    eng.bind_buffer(t["state"], t["actions"], t["rewards"], t["next_state"], t["done"])
    eng.set_buffer_state(0, N)
    lq, lp = eng.update(1, idx=idx, eps=eps)
    hp = onp.HyperParams(act_limit=0.4)
    a64, c64, t64 = (onp.cast_params(p, np.float64) for p in (actor, critic, critic))
    o = onp.sac_update(a64, c64, t64, onp.AdamState.zeros_like(a64), onp.AdamState.zeros_like(c64),
                       batch_from(arrays, idx[0], np.float64), eps[0, 0].astype(np.float64), eps[0, 1].astype(np.float64), hp)
    assert_close(lq.cpu().numpy(), [o["loss_q"]], 2e-4, 2e-5, "loss_q")
    assert_close(lp.cpu().numpy(), [o["loss_pi"]], 2e-4, 2e-5, "loss_pi")
    for key in ("backup", "q1", "q2", "logp_next", "logp_pi", "a2", "pi"):
        assert_close(eng.debug_read(key), o[key], 2e-4, 1e-4, key)
    for key in ("q1_pi", "q2_pi"):      # evaluated with the critic AFTER its Adam step (|dp| = lr: near-zero gradients can flip sign)
        assert_close(eng.debug_read(key), o[key], 2e-3, 1e-3, key)
    # Gradients: rtol 2e-3 / atol 5e-6, except that a pre-activation which is ~0 (|x| < 1e-6) may land on either side of
    # the ReLU in fp32, which moves ONE column of dh by one sample's contribution (<= ~1e-3 here).  At these sizes (1.4M
    # parameters, 1024 samples) allow that for at most 0.2% of a tensor's entries.
    def grads_close(got, want, name):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        err = np.abs(got - want)
        bad = err > 5e-6 + 2e-3 * np.abs(want)
        assert bad.mean() <= 2e-3 and err.max() <= 2e-3, f"{name}: {bad.sum()} of {bad.size} entries off, max err {err.max():.3e}"
    gc = critic_views(eng.layout, torch.from_numpy(eng.debug_read("grad_critic")), O, A, H)
    for k in CRITIC_KEYS:
        grads_close(gc[k].numpy(), o["critic_grads"][k], k)
    ga = actor_views(eng.layout, torch.from_numpy(eng.debug_read("grad_actor")), O, A, H)
    for k in ACTOR_KEYS:
        grads_close(ga[k].numpy(), o["actor_grads"][k], k)


@pytest.mark.parametrize("shape", ["humanoid", "ragged", "ragged8"])
@pytest.mark.parametrize("mode", [0, 1])
def test_bf16_tensor_core_path_matches_bf16_oracle(shape, mode):
    """BASELINE config 3 ("Humanoid-shaped ..., bf16 tensor-core MLP path"): precision 3 = tcgen05 GEMMs with bf16
    operands (rounded to nearest-even on the way into shared memory), fp32 accumulation in TMEM, fp32 master weights,
    activations and Adam.  Checked against the float64 oracle with the SAME operand rounding (oracle.operand_rounding),
    so the comparison only has to absorb accumulation order plus the rare element whose bf16 rounding flips because its
    fp32 value differs in the last bits (relative 2^-9 on that element).  Stated tolerance: losses rtol 1e-3; per-sample
    quantities rtol 2e-3 / atol 2e-3.  Gradients: every layer re-rounds its input to bf16, so a last-bit difference
    upstream flips roundings downstream and two correct implementations agree only down to the bf16 noise floor of a
    K-term sum (~2^-9 / sqrt(K) of the tensor's scale): |err| <= 1e-2 |g| + 2e-2 rms(g) for at least 99% of a tensor's
    entries and max |err| <= 0.1 max |g|.  "ragged": tile tails in every dimension (M < 128, N < 64, K not a multiple of 16, unaligned
    rows -> element-wise staging)."""
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    # "humanoid" and "ragged8" (obs_dim a multiple of 8: every operand has a 16-byte-aligned bf16 shadow) run on the TMA-fed
    # tiles (csrc/sacb_tma.cuh), "ragged8" with tile tails in every dimension; "ragged" (obs_dim 11) cannot be mapped and
    # takes the staged cp.async tiles
    if shape == "humanoid":
        O, A, H, B, N, lim = 376, 17, [512, 512], 1024, 4096, 0.4
    elif shape == "ragged8":
        O, A, H, B, N, lim = 24, 5, [72, 40], 96, 700, 1.0
    else:
        O, A, H, B, N, lim = 11, 3, [72, 40], 96, 700, 1.0
    rng = np.random.default_rng(5)
    actor, critic = onp.init_params(O, A, H, rng)
    arrays = dict(state=rng.standard_normal((N, O)).astype(np.float32), actions=rng.uniform(-lim, lim, (N, A)).astype(np.float32),
                  rewards=rng.standard_normal(N).astype(np.float32), next_state=rng.standard_normal((N, O)).astype(np.float32),
                  done=rng.random(N) < 0.05)
    idx = rng.choice(N, B, replace=False)[None]
    eps = rng.standard_normal((1, 2, B, A)).astype(np.float32)
    eng = Engine(O, A, H, batch_size=B, max_burst=2, device=0, act_limit=lim)
    eng.set_precision(3)
    eng.set_mode(mode)
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    for k, v in actor_views(eng.layout, pa, O, A, H).items():
        v.copy_(torch.from_numpy(actor[k]))
    for k, v in critic_views(eng.layout, pc, O, A, H).items():
        v.copy_(torch.from_numpy(critic[k]))
    pt = pc.clone()
    eng.bind_params(pa, pc, pt)
    eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}
    eng.bind_buffer(t["state"], t["actions"], t["rewards"], t["next_state"], t["done"])
    eng.set_buffer_state(0, N)
    assert bool(eng.debug_read("operand_path")[0]) == (shape != "ragged")
    lq, lp = eng.update(1, idx=idx, eps=eps)
    hp = onp.HyperParams(act_limit=lim)
    a64, c64, t64 = (onp.cast_params(p, np.float64) for p in (actor, critic, critic))
    with onp.operand_rounding("bf16"):
        o = onp.sac_update(a64, c64, t64, onp.AdamState.zeros_like(a64), onp.AdamState.zeros_like(c64),
                           batch_from(arrays, idx[0], np.float64), eps[0, 0].astype(np.float64), eps[0, 1].astype(np.float64), hp)
        # the same update WITHOUT operand rounding must be measurably different, or this test proves nothing
    o32 = onp.sac_update(*(onp.cast_params(p, np.float64) for p in (actor, critic, critic)), onp.AdamState.zeros_like(a64),
                         onp.AdamState.zeros_like(c64), batch_from(arrays, idx[0], np.float64), eps[0, 0].astype(np.float64),
                         eps[0, 1].astype(np.float64), hp)
    assert np.abs(o32["q1"] - o["q1"]).max() > 1e-4
    assert_close(lq.cpu().numpy(), [o["loss_q"]], 1e-3, 1e-4, "loss_q")
    assert_close(lp.cpu().numpy(), [o["loss_pi"]], 1e-3, 1e-4, "loss_pi")
    assert abs(float(lq) - o["loss_q"]) < 0.5 * abs(o32["loss_q"] - o["loss_q"]) + 1e-5      # closer to the bf16 oracle than to fp32
    for key in ("backup", "q1", "q2", "logp_next", "logp_pi", "a2", "pi", "q1_pi", "q2_pi"):
        assert_close(eng.debug_read(key), o[key], 2e-3, 2e-3, key)

    def grads_close(got, want, name):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        err = np.abs(got - want)
        rms = float(np.sqrt(np.mean(want * want)))
        bad = err > 2e-2 * rms + 1e-2 * np.abs(want)
        assert bad.mean() <= 1e-2 and err.max() <= 0.1 * np.abs(want).max() + 1e-7, \
            f"{name}: {bad.sum()} of {bad.size} entries off, max err {err.max():.3e}, rms {rms:.3e}"
    gc = critic_views(eng.layout, torch.from_numpy(eng.debug_read("grad_critic")), O, A, H)
    for k in CRITIC_KEYS:
        grads_close(gc[k].numpy(), o["critic_grads"][k], k)
    ga = actor_views(eng.layout, torch.from_numpy(eng.debug_read("grad_actor")), O, A, H)
    for k in ACTOR_KEYS:
        grads_close(ga[k].numpy(), o["actor_grads"][k], k)


@pytest.mark.parametrize("shape", ["humanoid", "ragged8"])
def test_bf16_tma_operands_match_staged_operands(shape, monkeypatch):
    """Precision 3 twice on the same seeds: GEMM operands fetched by the TMA from the bf16 shadows (csrc/sacb_tma.cuh, the
    default) and staged through cp.async + conversion (SACB_TMA=0).  Both round the same fp32 values to bf16 and accumulate
    in fp32; they differ in the order of the K sum (the TMA tiles keep four accumulators), i.e. by fp32 rounding, which
    the next layer's bf16 rounding can amplify to one bf16 ulp (2^-8 relative) on isolated activations.  Stated tolerance
    over a 3-update burst: losses rtol 1e-3, per-sample quantities 5e-3; updated parameters within 2e-5 for >= 97% of the
    entries and within 6 lr for all (see below)."""
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    O, A, H, B, N = (376, 17, [512, 512], 1024, 4096) if shape == "humanoid" else (24, 5, [72, 40], 96, 700)
    rng = np.random.default_rng(11)
    actor, critic = onp.init_params(O, A, H, rng)
    arrays = dict(state=rng.standard_normal((N, O)).astype(np.float32), actions=rng.uniform(-1, 1, (N, A)).astype(np.float32),
                  rewards=rng.standard_normal(N).astype(np.float32), next_state=rng.standard_normal((N, O)).astype(np.float32),
                  done=rng.random(N) < 0.05)
    idx = np.stack([rng.choice(N, B, replace=False) for _ in range(3)])
    eps = rng.standard_normal((3, 2, B, A)).astype(np.float32)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}
    out = {}
    for tma in (1, 0):
        monkeypatch.setenv("SACB_TMA", str(tma))
        eng = Engine(O, A, H, batch_size=B, max_burst=4, device=0, act_limit=1.0)
        eng.set_precision(3)
        pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
        for k, v in actor_views(eng.layout, pa, O, A, H).items():
            v.copy_(torch.from_numpy(actor[k]))
        for k, v in critic_views(eng.layout, pc, O, A, H).items():
            v.copy_(torch.from_numpy(critic[k]))
        pt = pc.clone()
        eng.bind_params(pa, pc, pt)
        eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
        eng.bind_buffer(t["state"], t["actions"], t["rewards"], t["next_state"], t["done"])
        eng.set_buffer_state(0, N)
        assert bool(eng.debug_read("operand_path")[0]) == bool(tma)
        lq, lp = eng.update(3, idx=idx, eps=eps)
        out[tma] = dict(lq=lq.cpu().numpy(), lp=lp.cpu().numpy(), pa=pa.cpu().numpy(), pc=pc.cpu().numpy(), pt=pt.cpu().numpy(),
                        q1=eng.debug_read("q1"), backup=eng.debug_read("backup"), pi=eng.debug_read("pi"))
    a, b = out[1], out[0]
    assert_close(a["lq"], b["lq"], 1e-3, 1e-5, "loss_q")
    assert_close(a["lp"], b["lp"], 1e-3, 1e-5, "loss_pi")
    for key in ("q1", "backup", "pi"):
        assert_close(a[key], b[key], 5e-3, 5e-3, key)
    # parameters: Adam's first steps are +-lr whatever the gradient's size, so an entry whose (tiny) gradient changes sign
    # between the two summation orders moves by up to 2 lr per step; everything else agrees to fp32 rounding
    for key in ("pa", "pc", "pt"):
        d = np.abs(a[key] - b[key])
        assert (d > 2e-5).mean() <= 3e-2 and d.max() <= 3 * 2 * 3e-4 * 1.001, f"{key}: {(d > 2e-5).mean():.3%} off, max {d.max():.2e}"


@pytest.mark.parametrize("precision", [1, 3])
def test_packed_ring_gather_matches_struct_of_arrays(precision):
    """The product's replay ring is the packed `[N, pitch]` record array (ReplayBuffer); most parity tests bind the five
    separate arrays.  Same transitions, same indices and noise through both layouts must give bit-identical bursts -- in the
    default kernel (records land through cp.async.bulk, element scatter) and in the tcgen05 / TMA-operand kernel (records
    land in the idle operand ring, 128-bit scatter + bf16 shadows), Humanoid-sized records (771 floats)."""
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    O, A, H, B, N = 376, 17, [512, 512], 1024, 4096
    rng = np.random.default_rng(21)
    actor, critic = onp.init_params(O, A, H, rng)
    arrays = dict(state=rng.standard_normal((N, O)).astype(np.float32), actions=rng.uniform(-1, 1, (N, A)).astype(np.float32),
                  rewards=rng.standard_normal(N).astype(np.float32), next_state=rng.standard_normal((N, O)).astype(np.float32),
                  done=rng.random(N) < 0.05)
    idx = np.stack([rng.choice(N, B, replace=False) for _ in range(3)])
    eps = rng.standard_normal((3, 2, B, A)).astype(np.float32)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}
    pitch = (2 * O + A + 2 + 15) // 16 * 16
    rec = torch.zeros(N, pitch, device="cuda")
    rec[:, :O] = t["state"]; rec[:, O:O + A] = t["actions"]; rec[:, O + A] = t["rewards"]
    rec[:, O + A + 1:2 * O + A + 1] = t["next_state"]; rec[:, 2 * O + A + 1] = t["done"].float()
    out = {}
    for packed in (True, False):
        eng = Engine(O, A, H, batch_size=B, max_burst=4, device=0, act_limit=1.0)
        eng.set_precision(precision)
        pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
        for k, v in actor_views(eng.layout, pa, O, A, H).items():
            v.copy_(torch.from_numpy(actor[k]))
        for k, v in critic_views(eng.layout, pc, O, A, H).items():
            v.copy_(torch.from_numpy(critic[k]))
        pt = pc.clone()
        eng.bind_params(pa, pc, pt)
        eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
        if packed:
            eng.bind_buffer_packed(rec)
        else:
            eng.bind_buffer(t["state"], t["actions"], t["rewards"], t["next_state"], t["done"])
        eng.set_buffer_state(0, N)
        lq, lp = eng.update(3, idx=idx, eps=eps)
        out[packed] = (lq.cpu().numpy(), lp.cpu().numpy(), pa.cpu().numpy(), pc.cpu().numpy(), pt.cpu().numpy())
    for a, b in zip(out[True], out[False]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [1, 3, 4])
def test_acting_path_cluster_kernel_matches_oracle(n):
    """`action, _ = actor(state)` per env step (sac/algorithm.py:233): n <= 4 rows run as ONE launch of an 8-CTA
    thread-block cluster (activations exchanged through distributed shared memory).  Checked against the float64 oracle
    and against the batched phase-program path (n > 4)."""
    from torch_actor_critic_b200.networks.linear import Actor
    O, A, H = 17, 6, [256, 256]
    rng = np.random.default_rng(n)
    actor_np, _ = onp.init_params(O, A, H, rng)
    actor = Actor(O, A, H, act_limit=1.0)
    actor.load_state_dict({k: torch.tensor(v) for k, v in actor_np.items()})
    x = rng.standard_normal((n, O)).astype(np.float32)
    eps = rng.standard_normal((n, A)).astype(np.float32)
    hp = onp.HyperParams(act_limit=1.0)
    a64 = onp.cast_params(actor_np, np.float64)
    for det in (True, False):
        pi, logp = actor(torch.tensor(x), deterministic=det, eps=None if det else torch.tensor(eps))
        po, lo, _ = onp.actor_forward(a64, x.astype(np.float64), eps.astype(np.float64), hp, deterministic=det)
        assert_close(pi.cpu().numpy(), po, 1e-5, 1e-6, "pi")
        assert_close(logp.cpu().numpy(), lo, 1e-5, 2e-5, "logp")
    # same rows through the batched path (40 rows -> phase programs), first n rows must agree
    xb = np.concatenate([x, rng.standard_normal((40 - n, O)).astype(np.float32)]) if n < 40 else x
    pb, lb = actor(torch.tensor(xb), deterministic=True)
    ps, ls = actor(torch.tensor(x), deterministic=True)
    assert torch.allclose(pb[:n], ps, rtol=1e-5, atol=1e-6) and torch.allclose(lb[:n], ls, rtol=1e-5, atol=2e-5)
    # device-drawn noise: different draws on consecutive calls, actions inside the limit
    p1, _ = actor(torch.tensor(x)); p2, _ = actor(torch.tensor(x))
    assert not torch.equal(p1, p2) and float(p1.abs().max()) <= 1.0
    pn, none = actor(torch.tensor(x), deterministic=True, with_logprob=False)
    assert none is None and torch.equal(pn, ps)


@pytest.mark.parametrize("name,mode", [("small", 0), ("ragged", 0), ("small", 1)])
def test_pi_from_states_flag_matches_oracle(name, mode):
    """SACB_FLAG_PI_FROM_STATES: the canonical actor loss pi = actor(states) (the reference evaluates the policy at
    next_states, sac/algorithm.py:37).  A K-step trajectory against the float64 oracle with pi_from_next_states=False,
    and the piecewise update_policy program."""
    from torch_actor_critic_b200 import _lib
    g, actor, critic, arrays, hp = load_case(name)
    hp.pi_from_next_states = False
    K = int(g["n_updates"])
    eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp, flags=_lib.FLAG_PI_FROM_STATES)
    eng.set_mode(mode)
    lq, lp = eng.update(K, idx=g["idx"], eps=g["eps"])
    a64, c64, t64 = (onp.cast_params(p, np.float64) for p in (actor, critic, critic))
    oa, oq = onp.AdamState.zeros_like(a64), onp.AdamState.zeros_like(c64)
    rq, rp = [], []
    for k in range(K):
        o = onp.sac_update(a64, c64, t64, oa, oq, batch_from(arrays, g["idx"][k], np.float64), g["eps"][k, 0].astype(np.float64),
                           g["eps"][k, 1].astype(np.float64), hp)
        rq.append(o["loss_q"]); rp.append(o["loss_pi"])
    assert_close(lq.cpu().numpy(), rq, 1e-3, 1e-4, "loss_q")
    assert_close(lp.cpu().numpy(), rp, 1e-3, 1e-4, "loss_pi")
    for k in ACTOR_KEYS:
        assert_close(av[k].cpu().numpy(), a64[k], 1e-5, 3e-4 * K, f"actor {k}")
    for k in CRITIC_KEYS:
        assert_close(cv[k].cpu().numpy(), c64[k], 1e-5, 3e-4 * K, f"critic {k}")
    assert abs(float(lp[0]) - float(g["loss_pi"][0])) > 1e-4        # it does differ from the reference's pi(next_states)


def test_burst_longer_than_one_launch_is_split():
    """update_every > 64 (e.g. 100, accepted by the reference): Engine.update splits the burst into launches of <= 64
    updates; the Adam step counters carry across, so the result equals the same updates issued one by one."""
    g, actor, critic, arrays, hp = load_case("small")
    B, A = int(g["batch"]), int(g["act_dim"])
    n, K = len(arrays["rewards"]), 70
    rng = np.random.default_rng(3)
    idx = np.stack([rng.choice(n, B, replace=False) for _ in range(K)])
    eps = rng.standard_normal((K, 2, B, A)).astype(np.float32)
    outs = []
    for chunk in (K, 1):
        eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp, K=100)     # max_burst is clamped to 64
        assert eng.max_burst == 64
        if chunk == K:
            lq, lp = eng.update(K, idx=idx, eps=eps)
        else:
            parts = [eng.update(1, idx=idx[k:k + 1], eps=eps[k:k + 1]) for k in range(K)]
            lq, lp = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
        torch.cuda.synchronize()
        assert eng.adam_steps() == (K, K)
        outs.append((lq.cpu(), lp.cpu(), eng._keep["actor"].cpu(), eng._keep["critic"].cpu(), eng._keep["target"].cpu()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_optimizer_state_loaded_after_first_update_is_adopted():
    """opt.load_state_dict() AFTER the engine bound the optimizer replaces opt.state's tensors: the binding must notice and
    adopt the loaded moments / step count instead of silently stepping its old arenas."""
    import torch.optim as optim
    from copy import deepcopy
    from torch_actor_critic_b200.buffer.replay_buffer import ReplayBuffer
    from torch_actor_critic_b200.networks.linear import Actor, DoubleCritic
    from torch_actor_critic_b200.sac.algorithm import SAC
    g, actor_np, critic_np, arrays, hp = load_case("small")
    O, A, B = int(g["obs_dim"]), int(g["act_dim"]), int(g["batch"])
    H = [int(h) for h in g["hidden"]]

    def fresh():
        actor, critic = Actor(O, A, H, act_limit=hp.act_limit), DoubleCritic(O, A, H)
        actor.load_state_dict({k: torch.tensor(v) for k, v in actor_np.items()})
        critic.load_state_dict({k: torch.tensor(v) for k, v in critic_np.items()})
        buf = ReplayBuffer(len(arrays["rewards"]), O, A)
        buf.store_many(arrays["state"], arrays["actions"], arrays["rewards"], arrays["next_state"], arrays["done"])
        sac = SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=B, start_steps=0, steps_per_epoch=1,
                  max_ep_len=1, update_after=0, update_every=4, save_every=1)
        return sac, actor, critic, deepcopy(critic), buf, optim.Adam(actor.parameters(), lr=3e-4), optim.Adam(critic.parameters(), lr=3e-4)

    idx, eps = g["idx"], g["eps"]
    sac, actor, critic, target, buf, pi_opt, q_opt = fresh()
    sac.update_burst(2, actor, critic, target, buf, pi_opt, q_opt, idx=idx[:2], eps=eps[:2])
    snap = dict(a=deepcopy(actor.state_dict()), c=deepcopy(critic.state_dict()), t=deepcopy(target.state_dict()),
                pi=deepcopy(pi_opt.state_dict()), q=deepcopy(q_opt.state_dict()))
    sac.update_burst(1, actor, critic, target, buf, pi_opt, q_opt, idx=idx[2:3], eps=eps[2:3])
    want = (actor.arena.clone(), critic.arena.clone())
    # roll everything back to the snapshot IN PLACE on the already-bound objects, then repeat the third update
    actor.load_state_dict(snap["a"]); critic.load_state_dict(snap["c"]); target.load_state_dict(snap["t"])
    pi_opt.load_state_dict(snap["pi"]); q_opt.load_state_dict(snap["q"])
    sac.update_burst(1, actor, critic, target, buf, pi_opt, q_opt, idx=idx[2:3], eps=eps[2:3])
    torch.cuda.synchronize()
    assert torch.equal(actor.arena, want[0]) and torch.equal(critic.arena, want[1])
    assert float(q_opt.state_dict()["state"][0]["step"]) == 3 == float(pi_opt.state_dict()["state"][0]["step"])
