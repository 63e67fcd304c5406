"""Time the UNMODIFIED reference hot loop (oracle/_ref, see oracle/build_ref.py) on the host cores.

TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's `--impl reference` arm and `cpu_baseline` leg.  The loop is
sac/algorithm.py:273-281 verbatim in structure: per update `buffer.sample(batch_size)` (python RNG index draw + numpy
fancy indexing + FloatTensor construction, buffer/replay_buffer.py:45-54), `SAC.update_critic`, `SAC.update_policy`,
`update_targets`, and the two per-update loss readbacks (`.cpu().detach().numpy()`, :280-281).
"""
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


class ReferenceLoop:
    def __init__(self, arrays, obs_dim, act_dim, hidden, batch, act_limit, threads):
        import torch
        from copy import deepcopy
        from .ref_stubs import import_reference
        alg, Batch, ReplayBuffer, Actor, Critic, DoubleCritic = import_reference(REF_ROOT)
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        self.alg, self.batch = alg, batch
        self.actor = Actor(obs_dim, act_dim, list(hidden), act_limit=act_limit)
        self.critic = DoubleCritic(obs_dim, act_dim, list(hidden))
        self.target = deepcopy(self.critic)                     # sac/algorithm.py:194-196
        for p in self.target.parameters():
            p.requires_grad = False
        self.pi_opt = torch.optim.Adam(self.actor.parameters(), lr=3e-4)       # main.py:93-95
        self.q_opt = torch.optim.Adam(self.critic.parameters(), lr=3e-4)
        self.sac = alg.SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=batch, start_steps=0,
                           steps_per_epoch=1, max_ep_len=1000, update_after=0, update_every=50, save_every=1)
        torch.set_num_threads(threads)                          # SAC.__init__ -> setup_pytorch_for_mpi may touch it
        n = len(arrays["rewards"])
        buf = ReplayBuffer.__new__(ReplayBuffer)                # adopt the synthetic arrays without a 2nd 165 MB copy
        buf.state, buf.actions, buf.rewards = arrays["state"], arrays["actions"], arrays["rewards"]
        buf.next_state, buf.done = arrays["next_state"], arrays["done"]
        buf.ptr, buf.size, buf.max_size = 0, n, n
        self.buf = buf

    def one_update(self):
        alg, sac = self.alg, self.sac
        batch = self.buf.sample(self.batch)                                     # sac/algorithm.py:275
        loss_q = sac.update_critic(self.q_opt, self.actor, self.critic, self.target, batch)   # :276
        loss_pi = sac.update_policy(self.pi_opt, self.actor, self.critic, batch)                # :277
        alg.update_targets(self.critic, self.target, sac.polyak)                                # :278
        loss_pi.cpu().detach().numpy(); loss_q.cpu().detach().numpy()                           # :280-281
        return loss_q, loss_pi

    def time_updates(self, n_updates, warmup, max_seconds):
        """(seconds, updates_done).  At least `n_updates` unless `max_seconds` of timed work run out first."""
        for _ in range(warmup):
            self.one_update()
        t0 = time.perf_counter()
        done = 0
        while done < n_updates:
            self.one_update()
            done += 1
            if done % 10 == 0 and time.perf_counter() - t0 > max_seconds:
                break
        return time.perf_counter() - t0, done


def reference_rate(arrays, obs_dim, act_dim, hidden, batch, act_limit, threads, n_updates, warmup, max_seconds):
    loop = ReferenceLoop(arrays, obs_dim, act_dim, hidden, batch, act_limit, threads)
    dt, done = loop.time_updates(n_updates, warmup, max_seconds)
    return done / dt, done


def reference_visual_rate(arrays, obs_dim, act_dim, vis_dim, hidden, batch, act_limit, threads, n_updates, warmup, max_seconds):
    """updates/s of the UNMODIFIED reference's visual hot loop (VisualReplayBuffer.sample -> SAC.update_critic ->
    SAC.update_policy -> update_targets with VisualActor / VisualDoubleCritic, main.py:63-87) on the host."""
    import torch
    from copy import deepcopy
    from .ref_stubs import import_reference_visual
    (alg, VisualBatch, VisualReplayBuffer, MultiObservation, VisualActor, VisualCritic, VisualDoubleCritic, _cs, _cnn) = import_reference_visual(REF_ROOT)
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    actor = VisualActor(obs_dim, act_dim, tuple(vis_dim), list(hidden), act_limit)
    critic = VisualDoubleCritic(obs_dim, act_dim, tuple(vis_dim), list(hidden))
    target = deepcopy(critic)
    for p in target.parameters():
        p.requires_grad = False
    pi_opt = torch.optim.Adam(actor.parameters(), lr=3e-4)
    q_opt = torch.optim.Adam(critic.parameters(), lr=3e-4)
    sac = alg.SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=batch, start_steps=0, steps_per_epoch=1,
                  max_ep_len=1000, update_after=0, update_every=50, save_every=1)
    torch.set_num_threads(threads)
    n = len(arrays["rewards"])
    buf = VisualReplayBuffer(n, act_dim)
    for i in range(n):
        obs = MultiObservation(torch.FloatTensor(arrays["features"][i]), torch.FloatTensor(arrays["frames"][i].astype(np.float32)))
        nobs = MultiObservation(torch.FloatTensor(arrays["next_features"][i]), torch.FloatTensor(arrays["next_frames"][i].astype(np.float32)))
        buf.store(obs, arrays["actions"][i], arrays["rewards"][i], nobs, arrays["done"][i])

    def one():
        b = buf.sample(batch)
        lq = sac.update_critic(q_opt, actor, critic, target, b)
        lp = sac.update_policy(pi_opt, actor, critic, b)
        alg.update_targets(critic, target, sac.polyak)
        lp.cpu().detach().numpy(); lq.cpu().detach().numpy()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done = 0
    while done < n_updates:
        one()
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    return done / (time.perf_counter() - t0), done
