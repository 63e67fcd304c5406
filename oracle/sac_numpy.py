"""numpy oracle of one SAC gradient update -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it restates.  Gradients are derived by
hand (no autograd) so that this file is an independent check of the backward
formulas the CUDA kernels implement.  `dtype` selects float32 (same precision
as the reference's torch-CPU path) or float64 (ground truth for tolerances).

Parameter dictionaries use the reference's `state_dict` names:
  actor : layers.0.weight [H1,O]  layers.0.bias [H1]  layers.1.weight [H2,H1]  layers.1.bias [H2]
          mu_layer.weight [A,H2]  mu_layer.bias [A]   log_std_layer.weight [A,H2]  log_std_layer.bias [A]
  critic: q{1,2}.layers.0.weight [H1,O+A] .bias [H1]; q{1,2}.layers.1.weight [H2,H1] .bias [H2];
          q{1,2}.layers.2.weight [1,H2] .bias [1]
(networks/linear.py:23-27,59,75-76; networks/core.py:6-10).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

ACTOR_KEYS = (
    "layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias",
    "mu_layer.weight", "mu_layer.bias", "log_std_layer.weight", "log_std_layer.bias",
)
CRITIC_KEYS = tuple(
    f"{q}.layers.{i}.{wb}" for q in ("q1", "q2") for i in (0, 1, 2) for wb in ("weight", "bias")
)

LN2 = math.log(2.0)
HALF_LN_2PI = math.log(math.sqrt(2.0 * math.pi))


@dataclass
class HyperParams:
    """Scalars of the update (main.py:147-160, main.py:93-95, networks/linear.py:19-21)."""
    alpha: float = 0.2
    gamma: float = 0.99
    polyak: float = 0.995
    reward_scale: float = 1.0
    lr: float = 3e-4
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    log_min_std: float = -20.0
    log_max_std: float = 2.0
    act_limit: float = 10.0
    # reference quirks; True reproduces the reference
    logprob_reference_form: bool = True     # networks/linear.py:51 (un-parenthesised 2*ln2 - u - softplus(-2u))
    pi_from_next_states: bool = True        # sac/algorithm.py:37,148-153


def init_params(obs_dim: int, act_dim: int, hidden, rng: np.random.Generator, dtype=np.float32):
    """nn.Linear default init U(+-1/sqrt(fan_in)) for W and b (networks/core.py:8), drawn from a numpy RNG
    so fixtures can be regenerated without torch."""
    h1, h2 = hidden

    def lin(fan_out, fan_in):
        k = 1.0 / math.sqrt(fan_in)
        return (rng.uniform(-k, k, size=(fan_out, fan_in)).astype(dtype),
                rng.uniform(-k, k, size=(fan_out,)).astype(dtype))

    actor = {}
    actor["layers.0.weight"], actor["layers.0.bias"] = lin(h1, obs_dim)
    actor["layers.1.weight"], actor["layers.1.bias"] = lin(h2, h1)
    actor["mu_layer.weight"], actor["mu_layer.bias"] = lin(act_dim, h2)
    actor["log_std_layer.weight"], actor["log_std_layer.bias"] = lin(act_dim, h2)
    critic = {}
    for q in ("q1", "q2"):
        critic[f"{q}.layers.0.weight"], critic[f"{q}.layers.0.bias"] = lin(h1, obs_dim + act_dim)
        critic[f"{q}.layers.1.weight"], critic[f"{q}.layers.1.bias"] = lin(h2, h1)
        critic[f"{q}.layers.2.weight"], critic[f"{q}.layers.2.bias"] = lin(1, h2)
    return actor, critic


# --------------------------------------------------------------------------------------------
# GEMM operand rounding (the "bf16 tensor-core MLP path" of BASELINE.json configs[2])
# --------------------------------------------------------------------------------------------
_OPERAND_ROUNDING = None          # None: products in the array dtype; "bf16": both operands rounded to bfloat16 first


class operand_rounding:
    """Context manager: inside it every dense-layer product (forward, dX, dW) of the hidden / head layers rounds BOTH
    operands to bfloat16 (round-to-nearest-even) and accumulates in the array dtype -- the arithmetic of the CUDA
    engine's precision-3 path.  The critics' scalar output layer ([1,H2]: a dot product, not a tensor-core GEMM) and
    everything element-wise stay in the array dtype, as in the engine."""

    def __init__(self, mode):
        assert mode in (None, "bf16")
        self.mode = mode

    def __enter__(self):
        global _OPERAND_ROUNDING
        self.prev, _OPERAND_ROUNDING = _OPERAND_ROUNDING, self.mode

    def __exit__(self, *exc):
        global _OPERAND_ROUNDING
        _OPERAND_ROUNDING = self.prev


def round_bf16(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32.  NaN/Inf pass through."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    r = ((u.astype(np.uint64) + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    r = np.where(np.isfinite(x), r, u & np.uint32(0xFFFF0000) | np.where(np.isnan(x), np.uint32(0x00400000), np.uint32(0)))
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def _mm(a, b):
    if _OPERAND_ROUNDING == "bf16":
        # products of two bfloat16 values are exact in float32; the accumulation runs in the array dtype
        return round_bf16(a).astype(a.dtype) @ round_bf16(b).astype(a.dtype)
    return a @ b


def _softplus(x):
    """torch F.softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x))."""
    with np.errstate(over="ignore"):
        return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0)))).astype(x.dtype)


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)


# --------------------------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------------------------
def actor_forward(p, x, eps, hp: HyperParams, deterministic=False):
    """Actor.forward, networks/linear.py:32-53.  `eps` is the N(0,1) draw of Normal.rsample (:45).

    Returns (pi_action [B,A], logp [B], cache)."""
    dt = x.dtype
    h1 = np.maximum(_mm(x, p["layers.0.weight"].T) + p["layers.0.bias"], 0)        # :33-35
    h2 = np.maximum(_mm(h1, p["layers.1.weight"].T) + p["layers.1.bias"], 0)
    mu = _mm(h2, p["mu_layer.weight"].T) + p["mu_layer.bias"]                      # :37
    ls_raw = _mm(h2, p["log_std_layer.weight"].T) + p["log_std_layer.bias"]        # :38
    ls = np.clip(ls_raw, dt.type(hp.log_min_std), dt.type(hp.log_max_std))         # :39
    std = np.exp(ls)                                                               # :40
    u = mu if deterministic else (mu + eps * std).astype(dt)                       # :43-45
    th = np.tanh(u)
    pi = (th * dt.type(hp.act_limit)).astype(dt)                                   # :46
    # Normal.log_prob(u): -((u-mu)^2)/(2 var) - log(std) - log(sqrt(2 pi))            :50
    z = u - mu
    var = std * std
    logp_gauss = (-(z * z) / (2 * var) - np.log(std) - dt.type(HALF_LN_2PI)).sum(-1)
    sp = _softplus(-2 * u)
    if hp.logprob_reference_form:
        corr = (dt.type(2 * LN2) - u - sp).sum(-1)                                 # :51 as written
    else:
        corr = (2 * (dt.type(LN2) - u - sp)).sum(-1)                               # canonical
    logp = (logp_gauss - corr).astype(dt)
    cache = dict(x=x, h1=h1, h2=h2, mu=mu, ls_raw=ls_raw, ls=ls, std=std, u=u, th=th, eps=eps)
    return pi, logp, cache


def critic_forward(p, q, s, a):
    """Critic.forward, networks/linear.py:61-69, for sub-critic `q` in {"q1","q2"}."""
    x = np.concatenate([s, a], axis=-1)                                             # :62
    h1 = np.maximum(_mm(x, p[f"{q}.layers.0.weight"].T) + p[f"{q}.layers.0.bias"], 0)   # :63-65
    h2 = np.maximum(_mm(h1, p[f"{q}.layers.1.weight"].T) + p[f"{q}.layers.1.bias"], 0)
    out = (h2 @ p[f"{q}.layers.2.weight"].T + p[f"{q}.layers.2.bias"])[..., 0]      # :67-68
    return out, dict(x=x, h1=h1, h2=h2)


def double_critic_forward(p, s, a):
    """DoubleCritic.forward, networks/linear.py:78-79."""
    q1, c1 = critic_forward(p, "q1", s, a)
    q2, c2 = critic_forward(p, "q2", s, a)
    return q1, q2, c1, c2


# --------------------------------------------------------------------------------------------
# backward helpers (hand-derived)
# --------------------------------------------------------------------------------------------
def _critic_backward(p, q, cache, dq, need_param_grads=True):
    """Back-propagate dL/dq [B] through one Critic.  Returns (param grads dict, dL/dx [B,O+A])."""
    W1 = p[f"{q}.layers.1.weight"]
    W0 = p[f"{q}.layers.0.weight"]
    W2 = p[f"{q}.layers.2.weight"]
    dh2 = (dq[:, None] * W2) * (cache["h2"] > 0)              # relu'(pre) = (out > 0)
    dh1 = _mm(dh2, W1) * (cache["h1"] > 0)
    dx = _mm(dh1, W0)
    g = {}
    if need_param_grads:
        g[f"{q}.layers.2.weight"] = (dq[None, :] @ cache["h2"])
        g[f"{q}.layers.2.bias"] = dq.sum(keepdims=True)
        g[f"{q}.layers.1.weight"] = _mm(dh2.T, cache["h1"])
        g[f"{q}.layers.1.bias"] = dh2.sum(0)
        g[f"{q}.layers.0.weight"] = _mm(dh1.T, cache["x"])
        g[f"{q}.layers.0.bias"] = dh1.sum(0)
    return g, dx


# --------------------------------------------------------------------------------------------
# losses + gradients
# --------------------------------------------------------------------------------------------
def q_loss_and_grads(actor, critic, target, batch, eps, hp: HyperParams):
    """eval_q_loss (sac/algorithm.py:46-74) + loss_q.backward() (:137), w.r.t. critic params.

    batch = dict(states, actions, rewards, next_states, done) float arrays (buffer/replay_buffer.py:8-14)."""
    s, a, r, s2, d = (batch[k] for k in ("states", "actions", "rewards", "next_states", "done"))
    dt = s.dtype
    a2, logp, _ = actor_forward(actor, s2, eps, hp)                                 # :61
    q1t, q2t, _, _ = double_critic_forward(target, s2, a2)                          # :62
    qt = np.minimum(q1t, q2t)                                                       # :63
    backup = (dt.type(hp.reward_scale) * r
              + (dt.type(hp.gamma) * (1 - d)) * (qt - dt.type(hp.alpha) * logp)).astype(dt)   # :65-67
    q1, q2, c1, c2 = double_critic_forward(critic, s, a)                            # :69
    B = s.shape[0]
    loss = (((q1 - backup) ** 2).mean() + ((q2 - backup) ** 2).mean()).astype(dt)   # :70-72
    dq1 = (2 * (q1 - backup) / B).astype(dt)
    dq2 = (2 * (q2 - backup) / B).astype(dt)
    g1, _ = _critic_backward(critic, "q1", c1, dq1)
    g2, _ = _critic_backward(critic, "q2", c2, dq2)
    grads = {**g1, **g2}
    grads = {k: grads[k].reshape(critic[k].shape).astype(dt) for k in CRITIC_KEYS}
    aux = dict(a2=a2, logp_next=logp, q1_target=q1t, q2_target=q2t, backup=backup, q1=q1, q2=q2)
    return loss, grads, aux


def pi_loss_and_grads(actor, critic, batch, eps, hp: HyperParams):
    """eval_pi_loss (sac/algorithm.py:30-43) + loss_pi.backward() (:156), w.r.t. actor params.

    Quirk 2: the policy is evaluated at next_states but Q at (states, pi) (:37-38,148-153)."""
    s, s2 = batch["states"], batch["next_states"]
    dt = s.dtype
    xin = s2 if hp.pi_from_next_states else s
    pi, logp, c = actor_forward(actor, xin, eps, hp)                                # :37
    q1, q2, c1, c2 = double_critic_forward(critic, s, pi)                           # :38
    qmin = np.minimum(q1, q2)                                                       # :40
    B = s.shape[0]
    loss = (dt.type(hp.alpha) * logp - qmin).mean().astype(dt)                      # :41
    # d loss / d q_i : -1/B to the smaller one, split on ties (torch minimum backward)
    gq = dt.type(-1.0 / B)
    dq1 = np.where(q1 < q2, gq, np.where(q1 == q2, gq / 2, 0)).astype(dt)
    dq2 = np.where(q2 < q1, gq, np.where(q1 == q2, gq / 2, 0)).astype(dt)
    _, dx1 = _critic_backward(critic, "q1", c1, dq1, need_param_grads=False)
    _, dx2 = _critic_backward(critic, "q2", c2, dq2, need_param_grads=False)
    O = s.shape[1]
    dpi = (dx1 + dx2)[:, O:]                                                        # action columns of cat([s,a])
    glogp = dt.type(hp.alpha / B)
    th, u, std, eps_ = c["th"], c["u"], c["std"], c["eps"]
    # pi = act_limit*tanh(u);  d(-corr)/du = tanh(u) (reference form) or 2*tanh(u) (canonical)
    dcorr = th if hp.logprob_reference_form else 2 * th
    g_u = dpi * dt.type(hp.act_limit) * (1 - th * th) + glogp * dcorr
    # gaussian term: total derivative w.r.t. mu is 0, w.r.t. log_std is -1 (eps held fixed)
    g_mu = g_u
    inside = (c["ls_raw"] >= hp.log_min_std) & (c["ls_raw"] <= hp.log_max_std)      # clip backward, inclusive
    g_ls = (g_u * eps_ * std - glogp) * inside
    h2, h1, x = c["h2"], c["h1"], c["x"]
    g = {}
    g["mu_layer.weight"] = _mm(g_mu.T, h2)
    g["mu_layer.bias"] = g_mu.sum(0)
    g["log_std_layer.weight"] = _mm(g_ls.T, h2)
    g["log_std_layer.bias"] = g_ls.sum(0)
    dh2 = (_mm(g_mu, actor["mu_layer.weight"]) + _mm(g_ls, actor["log_std_layer.weight"])) * (h2 > 0)
    g["layers.1.weight"] = _mm(dh2.T, h1)
    g["layers.1.bias"] = dh2.sum(0)
    dh1 = _mm(dh2, actor["layers.1.weight"]) * (h1 > 0)
    g["layers.0.weight"] = _mm(dh1.T, x)
    g["layers.0.bias"] = dh1.sum(0)
    grads = {k: g[k].astype(dt) for k in ACTOR_KEYS}
    aux = dict(pi=pi, logp_pi=logp, q1_pi=q1, q2_pi=q2)
    return loss, grads, aux


# --------------------------------------------------------------------------------------------
# optimizer + target update
# --------------------------------------------------------------------------------------------
@dataclass
class AdamState:
    """torch.optim.Adam state (exp_avg, exp_avg_sq, step) for one parameter dict."""
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)
    step: int = 0

    @classmethod
    def zeros_like(cls, params):
        return cls({k: np.zeros_like(v) for k, v in params.items()},
                   {k: np.zeros_like(v) for k, v in params.items()}, 0)


def adam_step(params, grads, st: AdamState, hp: HyperParams):
    """torch.optim.adam._single_tensor_adam (non-capturable branch; torch 2.11), called at
    sac/algorithm.py:139,157 with lr=3e-4, betas=(0.9,0.999), eps=1e-8 (main.py:93-95).  In place."""
    st.step += 1
    t = st.step
    bc1 = 1.0 - hp.beta1 ** t          # python doubles, as in torch
    bc2 = 1.0 - hp.beta2 ** t
    step_size = hp.lr / bc1
    bc2_sqrt = bc2 ** 0.5
    for k, p in params.items():
        dt = p.dtype
        g = grads[k]
        m, v = st.m[k], st.v[k]
        m += dt.type(1.0 - hp.beta1) * (g - m)                      # exp_avg.lerp_(grad, 1-beta1)
        v *= dt.type(hp.beta2)                                      # exp_avg_sq.mul_(beta2)
        v += dt.type(1.0 - hp.beta2) * g * g                        #   .addcmul_(grad, grad, value=1-beta2)
        denom = np.sqrt(v) / dt.type(bc2_sqrt) + dt.type(hp.adam_eps)
        p += dt.type(-step_size) * m / denom                        # param.addcdiv_(exp_avg, denom, value=-step_size)


def polyak_update(source, target, polyak: float):
    """update_targets, sac/algorithm.py:77-81: targ.mul_(polyak); targ.add_((1-polyak)*src).  In place."""
    for k, src in source.items():
        dt = src.dtype
        t = target[k]
        t *= dt.type(polyak)
        t += dt.type(1.0 - polyak) * src


# --------------------------------------------------------------------------------------------
# one full gradient update = a3+a9+a11+a12 of SURVEY.md section 8
# --------------------------------------------------------------------------------------------
def sac_update(actor, critic, target, opt_pi: AdamState, opt_q: AdamState, batch, eps_q, eps_pi,
               hp: HyperParams):
    """One iteration of the hot loop sac/algorithm.py:275-278 on an already-gathered batch:
    update_critic (:115-141) -> update_policy (:143-162, sees the updated critic) -> update_targets (:77-81).
    Mutates params/opt state in place; returns every intermediate the parity tests compare."""
    loss_q, gq, aux_q = q_loss_and_grads(actor, critic, target, batch, eps_q, hp)
    adam_step(critic, gq, opt_q, hp)
    loss_pi, gpi, aux_pi = pi_loss_and_grads(actor, critic, batch, eps_pi, hp)
    adam_step(actor, gpi, opt_pi, hp)
    polyak_update(critic, target, hp.polyak)
    out = dict(loss_q=loss_q, loss_pi=loss_pi, critic_grads=gq, actor_grads=gpi)
    out.update(aux_q)
    out.update(aux_pi)
    return out


# --------------------------------------------------------------------------------------------
# replay buffer (buffer/replay_buffer.py:17-54)
# --------------------------------------------------------------------------------------------
class ReplayBufferOracle:
    """SoA ring buffer semantics of the reference: `store` writes at ptr, ptr wraps, size saturates
    (buffer/replay_buffer.py:29-43); `gather(idx)` is `sample` (:45-54) with the index draw injected
    (the reference draws `random.sample(range(size), B)`, without replacement, :46)."""

    def __init__(self, size, obs_dim, act_dim):
        self.state = np.zeros((size, obs_dim), np.float32)
        self.actions = np.zeros((size, act_dim), np.float32)
        self.rewards = np.zeros(size, np.float32)
        self.next_state = np.zeros((size, obs_dim), np.float32)
        self.done = np.zeros(size, np.bool_)
        self.ptr, self.size, self.max_size = 0, 0, size

    def store(self, obs, act, rew, next_obs, done):
        i = self.ptr
        self.state[i], self.actions[i], self.rewards[i] = obs, act, rew
        self.next_state[i], self.done[i] = next_obs, done
        self.ptr = (self.ptr + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def gather(self, idx, dtype=np.float32):
        idx = np.asarray(idx)
        if len(idx) > self.size:
            raise ValueError("Sample larger than population or is negative")
        return dict(states=self.state[idx].astype(dtype), actions=self.actions[idx].astype(dtype),
                    rewards=self.rewards[idx].astype(dtype), next_states=self.next_state[idx].astype(dtype),
                    done=self.done[idx].astype(dtype))


def cast_params(p, dtype):
    return {k: v.astype(dtype).copy() for k, v in p.items()}
