"""GPU parity of the VISUAL SAC engine (run with -m gpu on the B200 box), through the C ABI (`sacv_*`), against
  (1) golden vectors produced by the unmodified reference's visual path (tests/golden/visual_*.npz: VisualReplayBuffer,
      VisualActor / VisualDoubleCritic, SAC.update_critic / update_policy / update_targets) and
  (2) the float64 torch oracle (oracle/visual_torch.py) on the same seeded inputs.

Stated tolerance: the convolutions / dense layers run as tensor-core GEMMs with 3xTF32 error compensation and fp32
accumulation (fp32-level accuracy, different summation order than torch's CPU kernels): losses rtol 5e-4, per-sample
Q / log-prob values rtol 5e-4 atol 1e-4, gradients rtol 2e-3 (a small fraction of near-zero entries behind ReLU boundaries
excepted), parameters after one Adam step within 2e-5 (a few noise-decided entries up to 2 lr), short trajectories K*lr.
Integer / byte work (ring arithmetic, uint8 frame gather) is bit-exact.
"""
import numpy as np
import pytest
import torch

from golden_helpers import assert_close
from oracle import visual_torch as vt
from visual_helpers import check_summary, close_frac, load_visual_case

pytestmark = pytest.mark.gpu


def _engine_for(g, actor, critic, arrays, hp, precision=1):
    from torch_actor_critic_b200.visual_engine import VisualEngine, vactor_views, vcritic_views
    O, A, B = int(g["obs_dim"]), int(g["act_dim"]), int(g["batch"])
    vis = tuple(int(v) for v in g["vis_dim"]); H = [int(h) for h in g["hidden"]]
    eng = VisualEngine(O, A, vis, H, batch_size=B, device=0, act_limit=hp.act_limit, precision=precision)
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    av, cv = vactor_views(eng.layout, pa, O, A, vis[0], H), vcritic_views(eng.layout, pc, O, A, vis[0], H)
    for k in vt.actor_keys():
        av[k].copy_(torch.from_numpy(actor[k]))
    for k in vt.critic_keys():
        cv[k].copy_(torch.from_numpy(critic[k]))
    pt = pc.clone()
    tv = vcritic_views(eng.layout, pt, O, A, vis[0], H)
    eng.bind_params(pa, pc, pt)
    adam = [eng.new_actor_arena(), eng.new_actor_arena(), eng.new_critic_arena(), eng.new_critic_arena()]
    eng.bind_adam(adam[0], adam[1], 0, adam[2], adam[3], 0)
    dev, n = eng.tdev(), len(arrays["rewards"])
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in arrays.items()}
    eng.bind_buffer(t["features"], t["frames"], t["actions"], t["rewards"], t["next_features"], t["next_frames"], t["done"])
    eng.set_buffer_state(0, n)
    eng._keep["ring"] = t
    return eng, av, cv, tv


@pytest.mark.parametrize("name", ["visual_small", "visual_raw", "visual_c5"])
def test_visual_update_matches_reference_golden(name):
    from torch_actor_critic_b200.visual_engine import vactor_views, vcritic_views
    g, actor, critic, arrays, hp = load_visual_case(name)
    eng, av, cv, tv = _engine_for(g, actor, critic, arrays, hp)
    O, A = int(g["obs_dim"]), int(g["act_dim"])
    vis = tuple(int(v) for v in g["vis_dim"]); H = [int(h) for h in g["hidden"]]
    K = int(g["n_updates"])
    lq, lp = eng.update(1, idx=g["idx"][:1], eps=g["eps"][:1])
    assert_close(lq.cpu().numpy(), g["loss_q"][:1], 5e-4, 2e-5, "loss_q[0]")
    assert_close(lp.cpu().numpy(), g["loss_pi"][:1], 5e-4, 2e-5, "loss_pi[0]")
    for key in ("a2", "logp_next", "q1_target", "q2_target", "q1", "q2", "pi", "logp_pi", "q1_pi", "q2_pi"):
        assert_close(eng.debug_read(key), g[f"step0.{key}"], 5e-4, 1e-4, key)
    ga = vactor_views(eng.layout, torch.from_numpy(eng.debug_read("grad_actor")), O, A, vis[0], H)
    gc = vcritic_views(eng.layout, torch.from_numpy(eng.debug_read("grad_critic")), O, A, vis[0], H)
    for k in vt.critic_keys():
        scale = float(np.abs(g[f"step0.critic_grads.{k}"]).max()) if f"step0.critic_grads.{k}" in g else float(g[f"step0.critic_grads.{k}.abssum"]) / gc[k].numel()
        check_summary(g, "step0.critic_grads", k, gc[k].numpy(), 2e-3, 1e-6 + 2e-3 * scale, assert_close, frac=0.01, atol_out=10 * scale + 1e-5)
    for k in vt.actor_keys():
        scale = float(np.abs(g[f"step0.actor_grads.{k}"]).max()) if f"step0.actor_grads.{k}" in g else float(g[f"step0.actor_grads.{k}.abssum"]) / ga[k].numel()
        check_summary(g, "step0.actor_grads", k, ga[k].numpy(), 2e-3, 1e-6 + 2e-3 * scale, assert_close, frac=0.01, atol_out=10 * scale + 1e-5)
    for k in vt.critic_keys():
        check_summary(g, "step0.critic", k, cv[k].cpu().numpy(), 1e-6, 2e-5, assert_close, frac=0.03, atol_out=6.1e-4)
        check_summary(g, "step0.target", k, tv[k].cpu().numpy(), 1e-6, 1e-6, assert_close, frac=0.03, atol_out=4e-6)
    for k in vt.actor_keys():
        check_summary(g, "step0.actor", k, av[k].cpu().numpy(), 1e-6, 2e-5, assert_close, frac=0.03, atol_out=6.1e-4)
    if K > 1:
        lq, lp = eng.update(K - 1, idx=g["idx"][1:], eps=g["eps"][1:])
        assert_close(lq.cpu().numpy(), g["loss_q"][1:], 2e-3, 2e-4, "loss_q[1:]")
        assert_close(lp.cpu().numpy(), g["loss_pi"][1:], 2e-3, 2e-4, "loss_pi[1:]")
        for k in vt.critic_keys():
            check_summary(g, "final.critic", k, cv[k].cpu().numpy(), 1e-5, 3e-4 * K, assert_close)
        for k in vt.actor_keys():
            check_summary(g, "final.actor", k, av[k].cpu().numpy(), 1e-5, 3e-4 * K, assert_close)
    assert eng.adam_steps() == (K, K) and eng.launch_count() > 0


def test_visual_update_matches_float64_oracle_and_tf32_is_coarser():
    """One update against the float64 oracle; the plain-TF32 arithmetic (precision 0) lands further away than 3xTF32."""
    g, actor, critic, arrays, hp = load_visual_case("visual_small")
    a64, c64 = vt.to_torch(actor), vt.to_torch(critic)
    t64 = {k: v.clone() for k, v in c64.items()}
    o = vt.visual_sac_update(a64, c64, t64, vt.AdamState(), vt.AdamState(), vt.torch_batch(arrays, g["idx"][0]),
                             torch.tensor(g["eps"][0, 0], dtype=torch.float64), torch.tensor(g["eps"][0, 1], dtype=torch.float64), hp)
    errs = {}
    for precision in (1, 0):
        eng, av, cv, tv = _engine_for(g, actor, critic, arrays, hp, precision=precision)
        lq, lp = eng.update(1, idx=g["idx"][:1], eps=g["eps"][:1])
        q1 = eng.debug_read("q1")
        errs[precision] = (abs(float(lq) - o["loss_q"]), np.abs(q1 - o["q1"].numpy()).max())
        if precision == 1:
            assert_close(lq.cpu().numpy(), [o["loss_q"]], 3e-4, 1e-5, "loss_q vs f64")
            assert_close(lp.cpu().numpy(), [o["loss_pi"]], 3e-4, 1e-5, "loss_pi vs f64")
            assert_close(q1, o["q1"].numpy(), 3e-4, 5e-5, "q1 vs f64")
    assert errs[0][1] > errs[1][1]                              # TF32's 10-bit mantissa shows; 3xTF32 does not
    assert errs[0][0] < 2e-2 * abs(o["loss_q"])                 # ... but it is still the same computation


def test_visual_replay_buffer_matches_reference_semantics():
    """VisualReplayBuffer.store / sample with MultiObservation objects (visual_replay_buffer.py:33-66): ring wrap, size
    saturation, uint8 frame ring bit-exact against the stored frames, ValueError when size < batch."""
    from torch_actor_critic_b200.buffer.visual_replay_buffer import VisualBatch, VisualReplayBuffer
    from torch_actor_critic_b200.environments import MultiObservation
    rng = np.random.default_rng(3)
    N, O, A, V = 10, 4, 2, (3, 12, 9)
    arr = vt.make_visual_arrays(rng, 25, O, A, V, 1.0, p_done=0.3)
    buf = VisualReplayBuffer(N, A)
    with pytest.raises(ValueError):
        buf.sample(1)
    for i in range(25):
        obs = MultiObservation(torch.tensor(arr["features"][i]), torch.tensor(arr["frames"][i].astype(np.float32)))
        nobs = MultiObservation(torch.tensor(arr["next_features"][i]), torch.tensor(arr["next_frames"][i].astype(np.float32)))
        buf.store(obs, arr["actions"][i], arr["rewards"][i], nobs, arr["done"][i])
        assert buf.size == min(i + 1, N) and buf.ptr == (i + 1) % N
    idx = np.array([0, 9, 4, 5, 2, 7])
    b = buf.sample(6, idx=idx)
    assert isinstance(b, VisualBatch) and isinstance(b.states, MultiObservation)
    src = np.array([20, 19, 24, 15, 22, 17])                   # slot s holds the last transition i with i % 10 == s
    assert np.array_equal(b.states.features.cpu().numpy(), arr["features"][src])
    assert np.array_equal(b.states.frame.cpu().numpy(), arr["frames"][src].astype(np.float32))
    assert np.array_equal(b.next_states.frame.cpu().numpy(), arr["next_frames"][src].astype(np.float32))
    assert np.array_equal(b.actions.cpu().numpy(), arr["actions"][src]) and np.array_equal(b.rewards.cpu().numpy(), arr["rewards"][src])
    assert np.array_equal(b.done.cpu().numpy(), arr["done"][src].astype(np.float32))
    assert b.states.frame.shape == (6, *V) and b.states.frame.dtype == torch.float32
    d = buf.sample(10)                                          # device draw: a permutation of the ring
    got = sorted(map(tuple, d.states.features.cpu().numpy().round(6)))
    assert got == sorted(map(tuple, arr["features"][15:25].round(6)))
    with pytest.raises(ValueError):
        buf.sample(11)
    with pytest.raises(ValueError):
        buf.store(MultiObservation(torch.zeros(O), torch.full(V, 0.5)), arr["actions"][0], 0.0, MultiObservation(torch.zeros(O), torch.zeros(V)), False)


def test_visual_modules_forward_and_piecewise_api():
    """VisualActor / VisualDoubleCritic forward contracts (tests/test_convolutional.py of the reference: shapes, a single
    unbatched MultiObservation) and SAC.update_critic / update_policy / update_targets on an explicit VisualBatch == the
    fused burst on the same samples."""
    import torch.optim as optim
    from copy import deepcopy
    from torch_actor_critic_b200.buffer.visual_replay_buffer import VisualBatch, VisualReplayBuffer
    from torch_actor_critic_b200.environments import MultiObservation
    from torch_actor_critic_b200.networks.convolutional import VisualActor, VisualDoubleCritic, calculate_size
    from torch_actor_critic_b200.sac.algorithm import SAC, update_targets
    g, actor_np, critic_np, arrays, hp = load_visual_case("visual_small")
    O, A, B = int(g["obs_dim"]), int(g["act_dim"]), int(g["batch"])
    vis = tuple(int(v) for v in g["vis_dim"]); H = [int(h) for h in g["hidden"]]
    assert calculate_size(vis, [32, 64, 64], [8, 4, 3], [4, 2, 1]) == int(g["calculate_size"])
    assert calculate_size((3, 84, 84), [32, 64, 64], [8, 4, 3], [4, 2, 1]) == 3136

    def fresh():
        actor = VisualActor(O, A, vis, H, hp.act_limit)
        critic = VisualDoubleCritic(O, A, vis, H)
        assert [n for n, _ in actor.named_parameters()] == vt.actor_keys() and [n for n, _ in critic.named_parameters()] == vt.critic_keys()
        actor.load_state_dict({k: torch.tensor(v) for k, v in actor_np.items()})
        critic.load_state_dict({k: torch.tensor(v) for k, v in critic_np.items()})
        return actor, critic, deepcopy(critic), optim.Adam(actor.parameters(), lr=3e-4), optim.Adam(critic.parameters(), lr=3e-4)

    actor, critic, target, pi_opt, q_opt = fresh()
    one = MultiObservation(torch.tensor(arrays["features"][0]), torch.tensor(arrays["frames"][0].astype(np.float32)))
    p1, l1 = actor(one, deterministic=True)
    assert list(p1.shape) == [A] and list(l1.shape) == []
    assert_close(p1.cpu().numpy(), g["fwd.pi_unbatched"], 5e-4, 5e-5, "pi unbatched")
    assert_close(l1.cpu().numpy(), g["fwd.logp_unbatched"], 5e-4, 1e-4, "logp unbatched")
    qa, qb = critic(one, torch.tensor(arrays["actions"][0]))
    assert_close(qa.cpu().numpy().reshape(-1), np.asarray(g["fwd.q1_unbatched"]).reshape(-1), 5e-4, 5e-5, "q1 unbatched")
    assert_close(qb.cpu().numpy().reshape(-1), np.asarray(g["fwd.q2_unbatched"]).reshape(-1), 5e-4, 5e-5, "q2 unbatched")
    sac = SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=B, start_steps=0, steps_per_epoch=1,
              max_ep_len=1, update_after=0, update_every=2, save_every=1)
    K = int(g["n_updates"])
    for k in range(K):
        ii = g["idx"][k]
        t = lambda a: torch.tensor(np.asarray(a[ii], np.float32)).cuda()      # noqa: E731
        samples = VisualBatch(MultiObservation(t(arrays["features"]), t(arrays["frames"])), t(arrays["actions"]), t(arrays["rewards"]),
                              MultiObservation(t(arrays["next_features"]), t(arrays["next_frames"])), t(arrays["done"]))
        lq = sac.update_critic(q_opt, actor, critic, target, samples, eps=g["eps"][k, 0])
        lp = sac.update_policy(pi_opt, actor, critic, samples, eps=g["eps"][k, 1])
        update_targets(critic, target, sac.polyak)
        assert_close(lq.cpu().numpy(), g["loss_q"][k], 2e-3, 2e-4, "loss_q")
        assert_close(lp.cpu().numpy(), g["loss_pi"][k], 2e-3, 2e-4, "loss_pi")
    assert float(q_opt.state_dict()["state"][0]["step"]) == K == float(pi_opt.state_dict()["state"][0]["step"])
    # the same samples through the HBM VisualReplayBuffer + fused burst
    actor2, critic2, target2, pi2, q2 = fresh()
    buf = VisualReplayBuffer(len(arrays["rewards"]), A)
    buf.store_many(arrays["features"], arrays["frames"], arrays["actions"], arrays["rewards"], arrays["next_features"], arrays["next_frames"], arrays["done"])
    sac2 = SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=B, start_steps=0, steps_per_epoch=1,
               max_ep_len=1, update_after=0, update_every=K, save_every=1)
    sac2.update_burst(K, actor2, critic2, target2, buf, pi2, q2, idx=g["idx"], eps=g["eps"])
    torch.cuda.synchronize()
    assert torch.equal(actor.arena, actor2.arena) and torch.equal(critic.arena, critic2.arena) and torch.equal(target.arena, target2.arena)
    for k in vt.actor_keys():
        check_summary(g, "final.actor", k, actor.state_dict()[k].cpu().numpy(), 1e-5, 3e-4 * K, assert_close)
