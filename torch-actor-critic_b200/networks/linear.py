"""Actor / Critic / DoubleCritic with the reference's constructor and forward contracts (networks/linear.py:13-79),
whose parameters are views into flat HBM arenas and whose forward passes are libsacb200 kernels.

Differences from the reference that a caller can observe: the modules live on the GPU (outputs are CUDA tensors;
CPU inputs are copied over), forward builds no autograd graph (training goes through `sac.algorithm.SAC`, which runs
the hand-derived backward on the device), and exactly two hidden layers are supported.
"""
from __future__ import annotations

import copy
import typing as t

import torch
import torch.nn as nn

from ..engine import Engine, actor_views, critic_views, layout_for
from .core import ArenaLinear

_FWD_ROWS = 256     # rows per kernel pass of the private forward engines


def _device_index(device):
    if device is None:
        return torch.cuda.current_device()
    return torch.device(device).index if not isinstance(device, int) else device


def _rebuild(cls, kwargs, state, requires_grad):
    """Unpickling hook (mlflow.pytorch.load_model / torch.load of a whole module, main.py:33-34, run_agent.py): re-create the
    arena-backed module on the current CUDA device and load the saved parameters.  Engines are re-created lazily."""
    m = cls(**kwargs)
    m.load_state_dict(state)
    for p, rg in zip(m.parameters(), requires_grad):
        p.requires_grad = rg
    return m


class _Picklable:
    def _ctor_kwargs(self):
        raise NotImplementedError

    def __reduce__(self):
        state = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        return (_rebuild, (type(self), self._ctor_kwargs(), state, [p.requires_grad for p in self.parameters()]))


class Actor(_Picklable, nn.Module):
    """tanh-Gaussian policy, reference networks/linear.py:13-53.

    forward(x, deterministic=False, with_logprob=True) -> (pi_action, logprob | None); x may be [B,O] or [O]."""

    def __init__(self, obs_dim: int, act_dim: int, hidden_sizes: t.List[int], log_min_std: float = -20,
                 log_max_std: float = 2, act_limit: float = 10, device=None):
        super().__init__()
        hidden_sizes = list(hidden_sizes)
        self.obs_dim, self.act_dim, self.hidden_sizes = obs_dim, act_dim, hidden_sizes
        self.log_min_std, self.log_max_std, self.act_limit = log_min_std, log_max_std, float(act_limit)
        self._dev = _device_index(device)
        self._layout = layout_for(obs_dim, act_dim, hidden_sizes)
        self.arena = torch.zeros(self._layout.actor_floats, dtype=torch.float32, device=torch.device("cuda", self._dev))
        v = actor_views(self._layout, self.arena, obs_dim, act_dim, hidden_sizes)
        h1, h2 = hidden_sizes
        self.layers = nn.ModuleList([ArenaLinear(obs_dim, h1, v["layers.0.weight"], v["layers.0.bias"]),
                                     ArenaLinear(h1, h2, v["layers.1.weight"], v["layers.1.bias"])])
        self.mu_layer = ArenaLinear(h2, act_dim, v["mu_layer.weight"], v["mu_layer.bias"])
        self.log_std_layer = ArenaLinear(h2, act_dim, v["log_std_layer.weight"], v["log_std_layer.bias"])
        self._engine = None

    def _ctor_kwargs(self):
        return dict(obs_dim=self.obs_dim, act_dim=self.act_dim, hidden_sizes=self.hidden_sizes, log_min_std=self.log_min_std,
                    log_max_std=self.log_max_std, act_limit=self.act_limit)

    def _fwd_engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.obs_dim, self.act_dim, self.hidden_sizes, batch_size=_FWD_ROWS, max_burst=1,
                                  device=self._dev, log_min_std=self.log_min_std, log_max_std=self.log_max_std,
                                  act_limit=self.act_limit, seed=torch.initial_seed())
            self._engine.bind_params(actor=self.arena)
        return self._engine

    def forward(self, x, deterministic: bool = False, with_logprob: bool = True, eps=None):
        x = torch.as_tensor(x, dtype=torch.float32, device=self.arena.device)
        squeeze = x.dim() == 1
        if squeeze:
            x = x.unsqueeze(0)
        lead = x.shape[:-1]
        pi, logp = self._fwd_engine().actor_forward(x.reshape(-1, self.obs_dim), eps=eps, deterministic=deterministic,
                                                    with_logprob=with_logprob)
        pi = pi.reshape(*lead, self.act_dim)
        if logp is not None:
            logp = logp.reshape(*lead)
        if squeeze:
            pi = pi[0]
            logp = logp[0] if logp is not None else None
        return pi, logp

    def __deepcopy__(self, memo):
        new = Actor(self.obs_dim, self.act_dim, self.hidden_sizes, self.log_min_std, self.log_max_std, self.act_limit,
                    device=self._dev)
        new.arena.copy_(self.arena)
        for p, q in zip(new.parameters(), self.parameters()):
            p.requires_grad = q.requires_grad
        return new


class Critic(nn.Module):
    """Q(s,a) MLP, reference networks/linear.py:56-69.  Inside a DoubleCritic it owns slot q1 or q2 of the twin arena;
    stand-alone it owns a private twin arena and uses slot 0."""

    def __init__(self, obs_dim: int, act_dim: int, hidden_sizes: t.List[int], device=None, _arena=None, _slot=0,
                 _layout=None):
        super().__init__()
        hidden_sizes = list(hidden_sizes)
        self.obs_dim, self.act_dim, self.hidden_sizes = obs_dim, act_dim, hidden_sizes
        self._dev = _device_index(device)
        self._slot = _slot
        self._layout = _layout or layout_for(obs_dim, act_dim, hidden_sizes)
        self._owns = _arena is None
        arena = _arena if _arena is not None else torch.zeros(self._layout.critic_floats, dtype=torch.float32,
                                                              device=torch.device("cuda", self._dev))
        object.__setattr__(self, "_arena_ref", arena)
        v = critic_views(self._layout, arena, obs_dim, act_dim, hidden_sizes)
        q = ("q1", "q2")[_slot]
        h1, h2 = hidden_sizes
        self.layers = nn.ModuleList([
            ArenaLinear(obs_dim + act_dim, h1, v[f"{q}.layers.0.weight"], v[f"{q}.layers.0.bias"]),
            ArenaLinear(h1, h2, v[f"{q}.layers.1.weight"], v[f"{q}.layers.1.bias"]),
            ArenaLinear(h2, 1, v[f"{q}.layers.2.weight"], v[f"{q}.layers.2.bias"]),
        ])
        self._engine = None

    def forward(self, state, action):
        if self._engine is None:
            self._engine = Engine(self.obs_dim, self.act_dim, self.hidden_sizes, batch_size=_FWD_ROWS, max_burst=1,
                                  device=self._dev)
            self._engine.bind_params(critic=self._arena_ref)
        state = torch.as_tensor(state, dtype=torch.float32, device=self._arena_ref.device)
        action = torch.as_tensor(action, dtype=torch.float32, device=self._arena_ref.device)
        lead = state.shape[:-1]
        qs = self._engine.critic_forward(state.reshape(-1, self.obs_dim), action.reshape(-1, self.act_dim))
        return qs[self._slot].reshape(*lead)


class DoubleCritic(_Picklable, nn.Module):
    """Twin Q networks, reference networks/linear.py:72-79.  forward(state, action) -> (q1, q2)."""

    def __init__(self, obs_dim: int, act_dim: int, hidden_sizes: t.List[int], device=None):
        super().__init__()
        hidden_sizes = list(hidden_sizes)
        self.obs_dim, self.act_dim, self.hidden_sizes = obs_dim, act_dim, hidden_sizes
        self._dev = _device_index(device)
        self._layout = layout_for(obs_dim, act_dim, hidden_sizes)
        self.arena = torch.zeros(self._layout.critic_floats, dtype=torch.float32, device=torch.device("cuda", self._dev))
        self.q1 = Critic(obs_dim, act_dim, hidden_sizes, device=self._dev, _arena=self.arena, _slot=0, _layout=self._layout)
        self.q2 = Critic(obs_dim, act_dim, hidden_sizes, device=self._dev, _arena=self.arena, _slot=1, _layout=self._layout)
        self._engine = None

    def _ctor_kwargs(self):
        return dict(obs_dim=self.obs_dim, act_dim=self.act_dim, hidden_sizes=self.hidden_sizes)

    def _fwd_engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.obs_dim, self.act_dim, self.hidden_sizes, batch_size=_FWD_ROWS, max_burst=1,
                                  device=self._dev)
            self._engine.bind_params(critic=self.arena)
        return self._engine

    def forward(self, state, action) -> t.Tuple[torch.Tensor, torch.Tensor]:
        state = torch.as_tensor(state, dtype=torch.float32, device=self.arena.device)
        action = torch.as_tensor(action, dtype=torch.float32, device=self.arena.device)
        lead = state.shape[:-1]
        q1, q2 = self._fwd_engine().critic_forward(state.reshape(-1, self.obs_dim), action.reshape(-1, self.act_dim))
        return q1.reshape(*lead), q2.reshape(*lead)

    def __deepcopy__(self, memo):
        """`deepcopy(critic)` is how the reference creates the target network (sac/algorithm.py:194): keep the copy
        arena-backed."""
        new = DoubleCritic(self.obs_dim, self.act_dim, self.hidden_sizes, device=self._dev)
        new.arena.copy_(self.arena)
        for p, q in zip(new.parameters(), self.parameters()):
            p.requires_grad = q.requires_grad
        return new
