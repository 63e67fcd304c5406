"""Drop-in surface end to end (GPU): the reference's `main.py` flow -- ReplayBuffer, Actor / DoubleCritic, two Adam
optimizers, SAC(**params), sac.train(...) -- against a gym-style fake environment (gym / MuJoCo are not installable
here), plus the checkpoint round trip of sac/algorithm.py:164-180 / main.py:28-51 through state_dicts."""
import numpy as np
import pytest
import torch
import torch.optim as optim

pytestmark = pytest.mark.gpu


class _Box:
    def __init__(self, dim, high):
        self.shape = (dim,)
        self.high = np.full(dim, high, np.float32)
        self._rng = np.random.default_rng(0)

    def sample(self):
        return self._rng.uniform(-self.high, self.high).astype(np.float32)


class FakeEnv:
    """Old-gym API as used by the reference (reset() -> obs; step(a) -> (obs, reward, done, info))."""

    def __init__(self, obs_dim=7, act_dim=3, horizon=40):
        self.observation_space = _Box(obs_dim, np.inf)
        self.action_space = _Box(act_dim, 1.0)
        self.horizon, self.t = horizon, 0
        self.rng = np.random.default_rng(1)
        self.state = None

    def reset(self):
        self.t = 0
        self.state = self.rng.standard_normal(self.observation_space.shape[0]).astype(np.float32)
        return self.state

    def step(self, action):
        action = np.asarray(action, np.float32)
        assert action.shape == self.action_space.shape and np.all(np.abs(action) <= 1.0 + 1e-5)
        self.t += 1
        self.state = (0.9 * self.state + 0.1 * self.rng.standard_normal(self.state.shape)).astype(np.float32)
        reward = float(-np.square(action).sum() + self.state[0])
        return self.state, reward, self.t >= self.horizon, {}

    def render(self):
        pass


def _build(env, batch):
    from torch_actor_critic_b200.buffer.replay_buffer import ReplayBuffer
    from torch_actor_critic_b200.networks.linear import Actor, DoubleCritic
    from torch_actor_critic_b200.sac.algorithm import SAC
    obs_dim, act_dim = env.observation_space.shape[0], env.action_space.shape[0]
    act_limit = env.action_space.high[0]                                   # main.py:59
    actor = Actor(obs_dim, act_dim, [64, 64], act_limit=act_limit)
    critic = DoubleCritic(obs_dim, act_dim, [64, 64])
    pi_opt = optim.Adam(actor.parameters(), lr=3e-4)                       # main.py:93-95
    q_opt = optim.Adam(critic.parameters(), lr=3e-4)
    buffer = ReplayBuffer(5000, obs_dim, act_dim)
    sac = SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=2, batch_size=batch, steps_per_epoch=400,
              start_steps=100, update_after=150, update_every=50, max_ep_len=5000, save_every=10)     # main.py:147-160 (scaled down)
    return sac, actor, critic, buffer, pi_opt, q_opt


def test_sac_train_runs_the_reference_flow(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    env = FakeEnv()
    sac, actor, critic, buffer, pi_opt, q_opt = _build(env, batch=32)
    a0 = {k: v.clone() for k, v in actor.state_dict().items()}
    c0 = {k: v.clone() for k, v in critic.state_dict().items()}
    metrics = sac.train(start_epoch=0, env=env, actor=actor, critic=critic, buffer=buffer, pi_opt=pi_opt, q_opt=q_opt,
                        render=False, logging=True)
    assert buffer.size == 800 and buffer.ptr == 800                       # every env step was stored
    assert np.isfinite(metrics["loss_q"]) and np.isfinite(metrics["loss_pi"]) and metrics["step"] == 800
    assert any(not torch.equal(a0[k], v) for k, v in actor.state_dict().items())
    assert any(not torch.equal(c0[k], v) for k, v in critic.state_dict().items())
    # 12 bursts of 50 updates happened (steps 200, 250, ..., 750): Adam step counters visible through torch
    assert float(q_opt.state_dict()["state"][0]["step"]) == 600 == float(pi_opt.state_dict()["state"][0]["step"])
    # target network moved towards the critic but is not equal to it
    tgt = sac.target_critic.state_dict()
    assert all(torch.isfinite(v).all() for v in tgt.values())
    assert any(not torch.equal(tgt[k], v) for k, v in critic.state_dict().items())
    # checkpoint artifacts with the reference's names (sac/algorithm.py:172-180): whole modules + auxiliaries, loadable
    ck = tmp_path / "sacb200_checkpoints" / "epoch_0"
    actor_l = torch.load(ck / "actor.pt", weights_only=False); critic_l = torch.load(ck / "critic.pt", weights_only=False)
    aux = torch.load(ck / "auxiliaries.pt", weights_only=False)
    assert type(actor_l).__name__ == "Actor" and type(critic_l).__name__ == "DoubleCritic" and aux["epoch"] == 0
    assert set(aux) == {"pi_opt", "q_opt", "epoch"} and actor_l.arena.is_cuda
    assert set(actor_l.state_dict()) == set(actor.state_dict())
    s0 = torch.as_tensor(env.reset())
    assert torch.equal(actor_l(s0, deterministic=True)[0], actor_l(s0, deterministic=True)[0])
    # acting path: deterministic action for evaluation (run_agent.py:37) stays inside the action limits
    act, _ = actor(env.reset(), deterministic=True)
    assert act.shape == (3,) and float(act.abs().max()) <= 1.0


def test_checkpoint_round_trip_resumes_bit_exact(tmp_path):
    """state_dict()s of the modules and optimizers (the content of the reference's checkpoint) restore a session that
    continues bit-identically: same burst, same injected indices/noise -> same parameters."""
    from torch_actor_critic_b200.buffer.replay_buffer import ReplayBuffer
    from copy import deepcopy
    env = FakeEnv()
    rng = np.random.default_rng(0)
    n, B, K = 600, 32, 5
    obs = rng.standard_normal((n, 7)).astype(np.float32); nobs = rng.standard_normal((n, 7)).astype(np.float32)
    act = rng.uniform(-1, 1, (n, 3)).astype(np.float32); rew = rng.standard_normal(n).astype(np.float32)
    done = rng.random(n) < 0.05
    idx = np.stack([rng.choice(n, B, replace=False) for _ in range(2 * K)])
    eps = rng.standard_normal((2 * K, 2, B, 3)).astype(np.float32)

    def fresh():
        sac, actor, critic, buffer, pi_opt, q_opt = _build(env, batch=B)
        buffer.store_many(obs, act, rew, nobs, done)
        return sac, actor, critic, buffer, pi_opt, q_opt

    torch.manual_seed(0)
    sac, actor, critic, buffer, pi_opt, q_opt = fresh()
    target = deepcopy(critic)
    sac.update_burst(K, actor, critic, target, buffer, pi_opt, q_opt, idx=idx[:K], eps=eps[:K])
    ckpt = {"actor": deepcopy(actor.state_dict()), "critic": deepcopy(critic.state_dict()), "target": deepcopy(target.state_dict()),
            "pi_opt": deepcopy(pi_opt.state_dict()), "q_opt": deepcopy(q_opt.state_dict())}
    torch.save(ckpt, tmp_path / "ckpt.pt")
    sac.update_burst(K, actor, critic, target, buffer, pi_opt, q_opt, idx=idx[K:], eps=eps[K:])

    ckpt = torch.load(tmp_path / "ckpt.pt")
    sac2, actor2, critic2, buffer2, pi_opt2, q_opt2 = fresh()
    actor2.load_state_dict(ckpt["actor"]); critic2.load_state_dict(ckpt["critic"])
    target2 = deepcopy(critic2); target2.load_state_dict(ckpt["target"])
    pi_opt2.load_state_dict(ckpt["pi_opt"]); q_opt2.load_state_dict(ckpt["q_opt"])      # main.py:38-42
    sac2.update_burst(K, actor2, critic2, target2, buffer2, pi_opt2, q_opt2, idx=idx[K:], eps=eps[K:])
    torch.cuda.synchronize()
    assert torch.equal(actor.arena, actor2.arena) and torch.equal(critic.arena, critic2.arena) and torch.equal(target.arena, target2.arena)
    assert float(q_opt2.state_dict()["state"][0]["step"]) == 2 * K
