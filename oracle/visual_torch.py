"""torch-CPU oracle of the VISUAL SAC update -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates `networks/convolutional.py:14-183` (simple_cnn, VisualActor, VisualCritic, VisualDoubleCritic) and the update
`sac/algorithm.py:30-81,115-162` on `VisualBatch`es (`buffer/visual_replay_buffer.py:12-66`) in functional form over plain
parameter dicts, in float32 or float64, with torch autograd and a restated Adam.  Pinned against golden vectors produced by
the unmodified reference (`oracle/make_golden_visual.py` -> `tests/golden/visual_*.npz`, checked by
`tests/test_oracle_visual.py`); the CUDA visual engine is compared against it on the GPU box.

Reference behaviours reproduced:
  * simple_cnn: Conv(k8,s4)+ReLU, Conv(k4,s2)+ReLU, Conv(k3,s1)+ReLU, Flatten (NCHW order), Linear(flat, 512), Linear(512, 1)
    with NO activation between the two linear layers (`convolutional.py:46-49`); frames are raw 0..255 floats.
  * VisualActor: heads are Linear(H2 + 1, A) on cat([trunk, cnn_scalar]) (`:78-79,103-106`); log-prob as written (`:118-119`).
  * VisualCritic: ReLU after EVERY MLP layer including the 1-wide output (`:156-158`); q = Linear(2, 1)(cat([relu_mlp, cnn]))
    (`:142,160-162`).
  * the actor loss evaluates the policy at next_states and Q at (states, pi) (`sac/algorithm.py:37-38`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

FILTERS, KERNELS, STRIDES, DENSE = (32, 64, 64), (8, 4, 3), (4, 2, 1), 512     # main.py:65-67, convolutional.py:17-22


@dataclass
class VisualHyperParams:
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
    act_limit: float = 1.0
    logprob_reference_form: bool = True


def conv_sizes(vis_dim):
    """Spatial sizes after each conv layer and the flattened size (convolutional.py:14-29)."""
    c, h, w = vis_dim
    sizes = []
    for f, k, s in zip(FILTERS, KERNELS, STRIDES):
        h, w = (h - k) // s + 1, (w - k) // s + 1
        sizes.append((f, h, w))
    return sizes, FILTERS[-1] * h * w


def cnn_keys(prefix="visual_network"):
    ks = []
    for i in range(3):
        ks += [f"{prefix}.conv_{i}.weight", f"{prefix}.conv_{i}.bias"]
    return ks + [f"{prefix}.linear.weight", f"{prefix}.linear.bias", f"{prefix}.final.weight", f"{prefix}.final.bias"]


def actor_keys():
    """named_parameters() order of the reference VisualActor (layers, visual_network, mu_layer, log_std_layer)."""
    return (["layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias"] + cnn_keys() +
            ["mu_layer.weight", "mu_layer.bias", "log_std_layer.weight", "log_std_layer.bias"])


def critic_keys():
    """named_parameters() order of the reference VisualDoubleCritic (q1 then q2; layers, final, visual_network)."""
    ks = []
    for q in ("q1", "q2"):
        ks += [f"{q}.layers.{i}.{wb}" for i in range(3) for wb in ("weight", "bias")]
        ks += [f"{q}.final.weight", f"{q}.final.bias"] + cnn_keys(f"{q}.visual_network")
    return ks


def _lin(rng, fan_out, fan_in, dtype):
    k = 1.0 / math.sqrt(fan_in)
    return rng.uniform(-k, k, (fan_out, fan_in)).astype(dtype), rng.uniform(-k, k, fan_out).astype(dtype)


def _cnn_params(rng, prefix, vis_dim, dtype, frame_scale):
    """torch-default-like init U(+-1/sqrt(fan_in)); conv_0 is divided by `frame_scale` so that raw 0..255 frames do not
    saturate everything downstream in the synthetic fixtures (the reference feeds raw pixels too, convolutional.py:99)."""
    p = {}
    cin = vis_dim[0]
    for i, (f, k) in enumerate(zip(FILTERS, KERNELS)):
        fan = cin * k * k
        b = 1.0 / math.sqrt(fan)
        p[f"{prefix}.conv_{i}.weight"] = (rng.uniform(-b, b, (f, cin, k, k)) / (frame_scale if i == 0 else 1.0)).astype(dtype)
        p[f"{prefix}.conv_{i}.bias"] = rng.uniform(-b, b, f).astype(dtype)
        cin = f
    _, flat = conv_sizes(vis_dim)
    p[f"{prefix}.linear.weight"], p[f"{prefix}.linear.bias"] = _lin(rng, DENSE, flat, dtype)
    p[f"{prefix}.final.weight"], p[f"{prefix}.final.bias"] = _lin(rng, 1, DENSE, dtype)
    return p


def init_visual_params(obs_dim, act_dim, vis_dim, hidden, rng, dtype=np.float32, frame_scale=64.0):
    """(actor, critic) parameter dicts keyed like the reference modules' state_dict()s, drawn from a numpy RNG so that
    fixtures can be regenerated without shipping 30 MB of weights."""
    h1, h2 = hidden
    a = {}
    a["layers.0.weight"], a["layers.0.bias"] = _lin(rng, h1, obs_dim, dtype)
    a["layers.1.weight"], a["layers.1.bias"] = _lin(rng, h2, h1, dtype)
    a.update(_cnn_params(rng, "visual_network", vis_dim, dtype, frame_scale))
    a["mu_layer.weight"], a["mu_layer.bias"] = _lin(rng, act_dim, h2 + 1, dtype)
    a["log_std_layer.weight"], a["log_std_layer.bias"] = _lin(rng, act_dim, h2 + 1, dtype)
    c = {}
    for q in ("q1", "q2"):
        c[f"{q}.layers.0.weight"], c[f"{q}.layers.0.bias"] = _lin(rng, h1, obs_dim + act_dim, dtype)
        c[f"{q}.layers.1.weight"], c[f"{q}.layers.1.bias"] = _lin(rng, h2, h1, dtype)
        c[f"{q}.layers.2.weight"], c[f"{q}.layers.2.bias"] = _lin(rng, 1, h2, dtype)
        c[f"{q}.final.weight"], c[f"{q}.final.bias"] = _lin(rng, 1, 2, dtype)
        c.update(_cnn_params(rng, f"{q}.visual_network", vis_dim, dtype, frame_scale))
    return {k: a[k] for k in actor_keys()}, {k: c[k] for k in critic_keys()}


def make_visual_arrays(rng, n, obs_dim, act_dim, vis_dim, act_limit, p_done=0.05):
    """Synthetic visual transitions: features ~ N(0,1), frames = uniform integers 0..255 (stored u8), SURVEY.md section 8d."""
    return dict(features=rng.standard_normal((n, obs_dim)).astype(np.float32),
                frames=rng.integers(0, 256, (n, *vis_dim), dtype=np.uint8),
                actions=rng.uniform(-act_limit, act_limit, (n, act_dim)).astype(np.float32),
                rewards=rng.standard_normal(n).astype(np.float32),
                next_features=rng.standard_normal((n, obs_dim)).astype(np.float32),
                next_frames=rng.integers(0, 256, (n, *vis_dim), dtype=np.uint8),
                done=rng.random(n) < p_done)


# ----------------------------------------------------------------------------------------------------------------
# forward passes
# ----------------------------------------------------------------------------------------------------------------
def cnn_forward(p, prefix, frame):
    """simple_cnn, convolutional.py:32-51.  frame [B,C,H,W] -> [B,1]."""
    x = frame
    for i, s in enumerate(STRIDES):
        x = F.relu(F.conv2d(x, p[f"{prefix}.conv_{i}.weight"], p[f"{prefix}.conv_{i}.bias"], stride=s))
    x = x.flatten(1)
    x = F.linear(x, p[f"{prefix}.linear.weight"], p[f"{prefix}.linear.bias"])
    return F.linear(x, p[f"{prefix}.final.weight"], p[f"{prefix}.final.bias"])


def actor_forward(p, feats, frame, hp: VisualHyperParams, eps=None, deterministic=False):
    """VisualActor.forward, convolutional.py:85-121 (batched)."""
    h = F.relu(F.linear(feats, p["layers.0.weight"], p["layers.0.bias"]))
    h = F.relu(F.linear(h, p["layers.1.weight"], p["layers.1.bias"]))
    x = torch.cat([h, cnn_forward(p, "visual_network", frame)], dim=1)
    mu = F.linear(x, p["mu_layer.weight"], p["mu_layer.bias"])
    log_std = torch.clip(F.linear(x, p["log_std_layer.weight"], p["log_std_layer.bias"]), hp.log_min_std, hp.log_max_std)
    std = torch.exp(log_std)
    u = mu if deterministic else mu + eps * std
    pi = torch.tanh(u) * hp.act_limit
    logp = (-((u - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)
    if hp.logprob_reference_form:
        logp = logp - (2 * math.log(2) - u - F.softplus(-2 * u)).sum(dim=-1)
    else:
        logp = logp - (2 * (math.log(2) - u - F.softplus(-2 * u))).sum(dim=-1)
    return pi, logp


def critic_forward(p, q, feats, frame, act):
    """VisualCritic.forward, convolutional.py:144-164: ReLU on all three MLP layers, then Linear(2,1) on [mlp, cnn]."""
    c = cnn_forward(p, f"{q}.visual_network", frame)
    x = torch.cat([feats, act], dim=-1)
    for i in range(3):
        x = F.relu(F.linear(x, p[f"{q}.layers.{i}.weight"], p[f"{q}.layers.{i}.bias"]))
    x = torch.cat([x, c], dim=1)
    return F.linear(x, p[f"{q}.final.weight"], p[f"{q}.final.bias"]).squeeze(-1)


def q_loss(actor, critic, target, b, hp, eps):
    """eval_q_loss, sac/algorithm.py:46-74."""
    with torch.no_grad():
        a2, logp = actor_forward(actor, b["next_features"], b["next_frames"], hp, eps)
        q1t = critic_forward(target, "q1", b["next_features"], b["next_frames"], a2)
        q2t = critic_forward(target, "q2", b["next_features"], b["next_frames"], a2)
        backup = hp.reward_scale * b["rewards"] + hp.gamma * (1 - b["done"]) * (torch.min(q1t, q2t) - hp.alpha * logp)
    q1 = critic_forward(critic, "q1", b["features"], b["frames"], b["actions"])
    q2 = critic_forward(critic, "q2", b["features"], b["frames"], b["actions"])
    loss = ((q1 - backup) ** 2).mean() + ((q2 - backup) ** 2).mean()
    return loss, dict(a2=a2, logp_next=logp, q1_target=q1t, q2_target=q2t, backup=backup, q1=q1, q2=q2)


def pi_loss(actor, critic, b, hp, eps):
    """eval_pi_loss, sac/algorithm.py:30-43: pi from next_states, Q at (states, pi)."""
    pi, logp = actor_forward(actor, b["next_features"], b["next_frames"], hp, eps)
    q1 = critic_forward(critic, "q1", b["features"], b["frames"], pi)
    q2 = critic_forward(critic, "q2", b["features"], b["frames"], pi)
    return (hp.alpha * logp - torch.min(q1, q2)).mean(), dict(pi=pi, logp_pi=logp, q1_pi=q1, q2_pi=q2)


# ----------------------------------------------------------------------------------------------------------------
# one update
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class AdamState:
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)
    t: int = 0


def adam_step(params, grads, st: AdamState, hp, keys):
    """torch.optim.adam._single_tensor_adam (non-capturable branch), in the dtype of the parameters."""
    st.t += 1
    bc1, bc2 = 1 - hp.beta1 ** st.t, 1 - hp.beta2 ** st.t
    step_size, bc2_sqrt = hp.lr / bc1, bc2 ** 0.5
    with torch.no_grad():
        for k in keys:
            g = grads[k]
            if k not in st.m:
                st.m[k], st.v[k] = torch.zeros_like(params[k]), torch.zeros_like(params[k])
            st.m[k].lerp_(g, 1 - hp.beta1)
            st.v[k].mul_(hp.beta2).addcmul_(g, g, value=1 - hp.beta2)
            denom = (st.v[k].sqrt() / bc2_sqrt).add_(hp.adam_eps)
            params[k].addcdiv_(st.m[k], denom, value=-step_size)


def to_torch(params, dtype=torch.float64):
    return {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in params.items()}


def torch_batch(arrays, idx, dtype=torch.float64):
    idx = np.asarray(idx)
    t = lambda a: torch.tensor(np.asarray(a[idx]), dtype=dtype)          # noqa: E731
    return dict(features=t(arrays["features"]), frames=t(arrays["frames"]), actions=t(arrays["actions"]),
                rewards=t(arrays["rewards"]), next_features=t(arrays["next_features"]), next_frames=t(arrays["next_frames"]),
                done=t(arrays["done"]))


def visual_sac_update(actor, critic, target, opt_pi: AdamState, opt_q: AdamState, batch, eps_q, eps_pi, hp):
    """update_critic -> update_policy -> update_targets (sac/algorithm.py:275-278) on one VisualBatch.  `actor`, `critic`,
    `target` are dicts of torch tensors (mutated in place); returns every intermediate the parity tests compare."""
    ak, ck = actor_keys(), critic_keys()
    for k in ck:
        critic[k].requires_grad_(True)
    loss_q, aux_q = q_loss(actor, critic, target, batch, hp, eps_q)
    gq = dict(zip(ck, torch.autograd.grad(loss_q, [critic[k] for k in ck])))
    for k in ck:
        critic[k].requires_grad_(False)
    adam_step(critic, gq, opt_q, hp, ck)
    for k in ak:
        actor[k].requires_grad_(True)
    loss_pi, aux_pi = pi_loss(actor, critic, batch, hp, eps_pi)
    gpi = dict(zip(ak, torch.autograd.grad(loss_pi, [actor[k] for k in ak])))
    for k in ak:
        actor[k].requires_grad_(False)
    adam_step(actor, gpi, opt_pi, hp, ak)
    with torch.no_grad():
        for k in ck:                                                     # update_targets, sac/algorithm.py:77-81
            target[k].mul_(hp.polyak)
            target[k].add_((1 - hp.polyak) * critic[k])
    out = dict(loss_q=float(loss_q.detach()), loss_pi=float(loss_pi.detach()), critic_grads=gq, actor_grads=gpi)
    out.update({k: v.detach() for k, v in aux_q.items()})
    out.update({k: v.detach() for k, v in aux_pi.items()})
    return out
