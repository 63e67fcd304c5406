"""torch-CPU oracle / CPU baseline of the SAC update -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference delegates all arithmetic to torch on the CPU: nn.Linear, autograd,
torch.optim.Adam, Normal.rsample.  This file restates the hot loop `sac/algorithm.py:274-281` with the
same torch CPU operators (functional form, own structure) so that
  * it can be timed on the GPU box's host cores as the CPU baseline (`kind: "port"`), since
    /root/reference itself does not exist there, and
  * it cross-checks `oracle.sac_numpy`'s hand-derived gradients against autograd.
It is pinned against golden vectors produced by the unmodified reference (tests/golden/).
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch
import torch.nn.functional as F

from .sac_numpy import ACTOR_KEYS, CRITIC_KEYS, HyperParams


def to_torch(params: dict, requires_grad=True) -> dict:
    return {k: torch.tensor(np.asarray(v), dtype=torch.float32).requires_grad_(requires_grad)
            for k, v in params.items()}


def actor_forward(p, x, hp: HyperParams, eps=None, deterministic=False, with_logprob=True):
    """networks/linear.py:32-53.  eps=None draws it exactly as Normal.rsample does
    (torch.distributions.utils._standard_normal -> torch.empty(shape).normal_())."""
    h = F.relu(F.linear(x, p["layers.0.weight"], p["layers.0.bias"]))
    h = F.relu(F.linear(h, p["layers.1.weight"], p["layers.1.bias"]))
    mu = F.linear(h, p["mu_layer.weight"], p["mu_layer.bias"])
    log_std = torch.clip(F.linear(h, p["log_std_layer.weight"], p["log_std_layer.bias"]),
                         hp.log_min_std, hp.log_max_std)
    std = torch.exp(log_std)
    u = mu
    if not deterministic:
        if eps is None:
            eps = torch.empty(mu.shape, dtype=mu.dtype).normal_()
        u = mu + eps * std
    pi = torch.tanh(u) * hp.act_limit
    if not with_logprob:
        return pi, None
    var = std ** 2
    logp = (-((u - mu) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)
    if hp.logprob_reference_form:
        logp = logp - (2 * torch.as_tensor(np.log(2)) - u - F.softplus(-2 * u)).sum(dim=-1)
    else:
        logp = logp - (2 * (math.log(2) - u - F.softplus(-2 * u))).sum(dim=-1)
    return pi, logp


def critic_forward(p, q, s, a):
    """networks/linear.py:61-69."""
    x = torch.cat([s, a], dim=-1)
    x = F.relu(F.linear(x, p[f"{q}.layers.0.weight"], p[f"{q}.layers.0.bias"]))
    x = F.relu(F.linear(x, p[f"{q}.layers.1.weight"], p[f"{q}.layers.1.bias"]))
    return F.linear(x, p[f"{q}.layers.2.weight"], p[f"{q}.layers.2.bias"]).squeeze(-1)


def q_loss(actor, critic, target, batch, hp, eps=None):
    """sac/algorithm.py:46-74."""
    with torch.no_grad():
        a2, logp = actor_forward(actor, batch["next_states"], hp, eps)
        qt = torch.min(critic_forward(target, "q1", batch["next_states"], a2),
                       critic_forward(target, "q2", batch["next_states"], a2))
        backup = hp.reward_scale * batch["rewards"] + hp.gamma * (1 - batch["done"]) * (qt - hp.alpha * logp)
    q1 = critic_forward(critic, "q1", batch["states"], batch["actions"])
    q2 = critic_forward(critic, "q2", batch["states"], batch["actions"])
    return ((q1 - backup) ** 2).mean() + ((q2 - backup) ** 2).mean(), dict(backup=backup, q1=q1, q2=q2)


def pi_loss(actor, critic, batch, hp, eps=None):
    """sac/algorithm.py:30-43 (policy at next_states, Q at (states, pi): quirk 2)."""
    xin = batch["next_states"] if hp.pi_from_next_states else batch["states"]
    pi, logp = actor_forward(actor, xin, hp, eps)
    q = torch.min(critic_forward(critic, "q1", batch["states"], pi),
                  critic_forward(critic, "q2", batch["states"], pi))
    return (hp.alpha * logp - q).mean(), dict(pi=pi, logp_pi=logp)


class TorchCpuSAC:
    """The reference's hot loop on torch CPU: sample -> update_critic -> update_policy -> update_targets."""

    def __init__(self, actor_np, critic_np, hp: HyperParams, target_np=None):
        self.hp = hp
        self.actor = to_torch(actor_np)
        self.critic = to_torch(critic_np)
        src = critic_np if target_np is None else target_np
        self.target = to_torch(src, requires_grad=False)              # deepcopy + freeze, sac/algorithm.py:194-196
        self.pi_opt = torch.optim.Adam([self.actor[k] for k in ACTOR_KEYS], lr=hp.lr,
                                       betas=(hp.beta1, hp.beta2), eps=hp.adam_eps)   # main.py:93-95
        self.q_opt = torch.optim.Adam([self.critic[k] for k in CRITIC_KEYS], lr=hp.lr,
                                      betas=(hp.beta1, hp.beta2), eps=hp.adam_eps)

    def update(self, batch, eps_q=None, eps_pi=None):
        hp = self.hp
        # update_critic, sac/algorithm.py:115-141
        self.q_opt.zero_grad()
        loss_q, aux_q = q_loss(self.actor, self.critic, self.target, batch, hp, eps_q)
        loss_q.backward()
        critic_grads = {k: self.critic[k].grad.detach().clone() for k in CRITIC_KEYS}
        self.q_opt.step()
        # update_policy, :143-162
        for k in CRITIC_KEYS:
            self.critic[k].requires_grad_(False)
        self.pi_opt.zero_grad()
        loss_pi, aux_pi = pi_loss(self.actor, self.critic, batch, hp, eps_pi)
        loss_pi.backward()
        actor_grads = {k: self.actor[k].grad.detach().clone() for k in ACTOR_KEYS}
        self.pi_opt.step()
        for k in CRITIC_KEYS:
            self.critic[k].requires_grad_(True)
        # update_targets, :77-81
        with torch.no_grad():
            for k in CRITIC_KEYS:
                self.target[k].mul_(hp.polyak)
                self.target[k].add_((1 - hp.polyak) * self.critic[k])
        out = dict(loss_q=loss_q.detach(), loss_pi=loss_pi.detach(), critic_grads=critic_grads,
                   actor_grads=actor_grads)
        out.update({k: v.detach() for k, v in aux_q.items()})
        out.update({k: v.detach() for k, v in aux_pi.items()})
        return out

    def update_fast(self, batch):
        """Same arithmetic without the bookkeeping copies: what the timed CPU baseline runs."""
        hp = self.hp
        self.q_opt.zero_grad()
        loss_q, _ = q_loss(self.actor, self.critic, self.target, batch, hp)
        loss_q.backward()
        self.q_opt.step()
        for k in CRITIC_KEYS:
            self.critic[k].requires_grad_(False)
        self.pi_opt.zero_grad()
        loss_pi, _ = pi_loss(self.actor, self.critic, batch, hp)
        loss_pi.backward()
        self.pi_opt.step()
        for k in CRITIC_KEYS:
            self.critic[k].requires_grad_(True)
        with torch.no_grad():
            for k in CRITIC_KEYS:
                self.target[k].mul_(hp.polyak)
                self.target[k].add_((1 - hp.polyak) * self.critic[k])
        return loss_q, loss_pi


class HostReplay:
    """buffer/replay_buffer.py:17-54 on host numpy arrays, including the python-RNG index draw (:46)."""

    def __init__(self, state, actions, rewards, next_state, done, size=None):
        self.state, self.actions, self.rewards, self.next_state, self.done = state, actions, rewards, next_state, done
        self.size = len(rewards) if size is None else size

    def sample(self, batch_size):
        idx = random.sample(range(self.size), batch_size)
        return dict(states=torch.FloatTensor(self.state[idx]), actions=torch.FloatTensor(self.actions[idx]),
                    rewards=torch.FloatTensor(self.rewards[idx]), next_states=torch.FloatTensor(self.next_state[idx]),
                    done=torch.FloatTensor(self.done[idx]))


def time_cpu_updates(replay: HostReplay, sac: TorchCpuSAC, batch_size, n_updates, warmup, update_every=50,
                     max_seconds=None):
    """Time the burst loop sac/algorithm.py:273-281 (incl. the two per-update loss readbacks, :280-281).
    Returns (seconds, updates_done); stops early once `max_seconds` of timed work have elapsed."""
    import time
    t_w = time.perf_counter()
    for _ in range(warmup):
        lq, lp = sac.update_fast(replay.sample(batch_size))
        if max_seconds is not None and time.perf_counter() - t_w > max_seconds:
            break
    t0 = time.perf_counter()
    done = 0
    while done < n_updates:
        for _ in range(min(update_every, n_updates - done)):
            lq, lp = sac.update_fast(replay.sample(batch_size))
            lp.cpu().detach().numpy(); lq.cpu().detach().numpy()
            done += 1
            if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
                return time.perf_counter() - t0, done
    return time.perf_counter() - t0, done
