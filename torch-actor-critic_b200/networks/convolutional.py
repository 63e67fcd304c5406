"""VisualActor / VisualCritic / VisualDoubleCritic / simple_cnn with the reference's constructor and forward contracts
(networks/convolutional.py:14-183), whose parameters are views into flat HBM arenas and whose forward passes are
libsacb200 kernels (implicit-GEMM convolutions on the tensor cores, `sacv_*` in include/sacb200.h).

As with the MLP modules: outputs are CUDA tensors, forward builds no autograd graph (training goes through
`sac.algorithm.SAC`, which runs the hand-derived backward on the device), two hidden layers / three conv layers.
"""
from __future__ import annotations

import typing as t

import numpy as np
import torch
import torch.nn as nn

from ..environments.wall_runner import MultiObservation
from ..visual_engine import DENSE, VisualEngine, vactor_views, vcritic_views, vlayout_for
from .core import ArenaLinear

_FWD_ROWS = 64      # rows per kernel pass of the private forward engines


def calculate_size(image_shape, filters, kernel_sizes, strides) -> int:
    """Flattened size of the conv stack's output (reference networks/convolutional.py:14-29)."""
    image_shape = list(image_shape)
    for f, k, s in zip(filters, kernel_sizes, strides):
        image_shape[0] = f
        image_shape[1] = int(np.floor((image_shape[1] - k) / s + 1))
        image_shape[2] = int(np.floor((image_shape[2] - k) / s + 1))
    return int(np.prod(image_shape))


def _device_index(device):
    if device is None:
        return torch.cuda.current_device()
    return torch.device(device).index if not isinstance(device, int) else device


class ArenaParams(nn.Module):
    """Stands in for nn.Conv2d / nn.Linear: `weight` / `bias` parameters (default torch init U(+-1/sqrt(fan_in))) that are
    views into a flat HBM arena."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        self.weight, self.bias = nn.Parameter(weight), nn.Parameter(bias)
        k = 1.0 / float(np.sqrt(np.prod(weight.shape[1:])))
        with torch.no_grad():
            self.weight.uniform_(-k, k)
            self.bias.uniform_(-k, k)


class _CnnParams(nn.Module):
    """Parameter container with the names of the reference's `simple_cnn` Sequential (convolutional.py:32-51):
    conv_0..2, linear, final."""

    def __init__(self, views, prefix):
        super().__init__()
        for i in range(3):
            setattr(self, f"conv_{i}", ArenaParams(views[f"{prefix}conv_{i}.weight"], views[f"{prefix}conv_{i}.bias"]))
        self.linear = ArenaParams(views[f"{prefix}linear.weight"], views[f"{prefix}linear.bias"])
        self.final = ArenaParams(views[f"{prefix}final.weight"], views[f"{prefix}final.bias"])


def _rebuild(cls, kwargs, state, requires_grad):
    m = cls(**kwargs)
    m.load_state_dict(state)
    for p, rg in zip(m.parameters(), requires_grad):
        p.requires_grad = rg
    return m


class _Picklable:
    def __reduce__(self):
        state = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        return (_rebuild, (type(self), self._ctor_kwargs(), state, [p.requires_grad for p in self.parameters()]))


def _obs(x, obs_dim, vis_dim, dev):
    f = torch.as_tensor(x.features, dtype=torch.float32, device=dev)
    fr = torch.as_tensor(x.frame, dtype=torch.float32, device=dev)
    single = f.dim() == 1
    return f.reshape(-1, obs_dim), fr.reshape(-1, *vis_dim), single


class VisualActor(_Picklable, nn.Module):
    """Reference networks/convolutional.py:54-121.  forward(x: MultiObservation, deterministic=False, with_logprob=True) ->
    (squeeze(pi_action), squeeze(logprob))."""

    def __init__(self, obs_dim: int, act_dim: int, vis_dim: t.Tuple[int, int, int], hidden_sizes: t.List[int] = [256, 256],
                 act_limit: float = 10, filters: t.List[int] = [32, 64, 64], kernel_sizes: t.List[int] = [8, 4, 3],
                 strides: t.List[int] = [4, 2, 1], log_min_std: float = -20, log_max_std: float = 2, device=None):
        super().__init__()
        hidden_sizes, filters, kernel_sizes, strides = list(hidden_sizes), list(filters), list(kernel_sizes), list(strides)
        self.obs_dim, self.act_dim, self.vis_dim, self.hidden_sizes = obs_dim, act_dim, tuple(vis_dim), hidden_sizes
        self.filters, self.kernel_sizes, self.strides = filters, kernel_sizes, strides
        self.log_min_std, self.log_max_std, self.act_limit = log_min_std, log_max_std, float(act_limit)
        self._dev = _device_index(device)
        self._layout = vlayout_for(obs_dim, act_dim, vis_dim, hidden_sizes, filters=filters, kernel_sizes=kernel_sizes, strides=strides)
        self.arena = torch.zeros(self._layout.actor_floats, dtype=torch.float32, device=torch.device("cuda", self._dev))
        v = vactor_views(self._layout, self.arena, obs_dim, act_dim, self.vis_dim[0], hidden_sizes, filters, kernel_sizes, DENSE)
        h1, h2 = hidden_sizes
        self.layers = nn.ModuleList([ArenaLinear(obs_dim, h1, v["layers.0.weight"], v["layers.0.bias"]),
                                     ArenaLinear(h1, h2, v["layers.1.weight"], v["layers.1.bias"])])
        self.visual_network = _CnnParams(v, "visual_network.")
        self.mu_layer = ArenaLinear(h2 + 1, act_dim, v["mu_layer.weight"], v["mu_layer.bias"])
        self.log_std_layer = ArenaLinear(h2 + 1, act_dim, v["log_std_layer.weight"], v["log_std_layer.bias"])
        self._engine = None

    def _ctor_kwargs(self):
        return dict(obs_dim=self.obs_dim, act_dim=self.act_dim, vis_dim=self.vis_dim, hidden_sizes=self.hidden_sizes, act_limit=self.act_limit,
                    filters=self.filters, kernel_sizes=self.kernel_sizes, strides=self.strides, log_min_std=self.log_min_std,
                    log_max_std=self.log_max_std)

    def _engine_kwargs(self):
        return dict(filters=self.filters, kernel_sizes=self.kernel_sizes, strides=self.strides, log_min_std=self.log_min_std,
                    log_max_std=self.log_max_std, act_limit=self.act_limit)

    def _fwd_engine(self) -> VisualEngine:
        if self._engine is None:
            self._engine = VisualEngine(self.obs_dim, self.act_dim, self.vis_dim, self.hidden_sizes, batch_size=_FWD_ROWS, device=self._dev,
                                        seed=torch.initial_seed(), **self._engine_kwargs())
            self._engine.bind_params(actor=self.arena)
        return self._engine

    def forward(self, x: MultiObservation, deterministic: bool = False, with_logprob: bool = True, eps=None):
        f, fr, _ = _obs(x, self.obs_dim, self.vis_dim, self.arena.device)
        pi, logp = self._fwd_engine().actor_forward(f, fr, eps=eps, deterministic=deterministic, with_logprob=with_logprob)
        return torch.squeeze(pi), (torch.squeeze(logp) if logp is not None else None)      # convolutional.py:121

    def __deepcopy__(self, memo):
        new = type(self)(device=self._dev, **self._ctor_kwargs())
        new.arena.copy_(self.arena)
        for p, q in zip(new.parameters(), self.parameters()):
            p.requires_grad = q.requires_grad
        return new


class VisualCritic(nn.Module):
    """Parameter view of one Q network (reference convolutional.py:124-164): MLP [O+A, H1, H2, 1] with ReLU after every layer,
    `final` = Linear(2, 1) over [mlp, cnn], and its own simple_cnn.  Owns slot q1 / q2 of a twin arena."""

    def __init__(self, views, q, obs_dim, act_dim, hidden_sizes):
        super().__init__()
        h1, h2 = hidden_sizes
        self.layers = nn.ModuleList([ArenaLinear(obs_dim + act_dim, h1, views[f"{q}.layers.0.weight"], views[f"{q}.layers.0.bias"]),
                                     ArenaLinear(h1, h2, views[f"{q}.layers.1.weight"], views[f"{q}.layers.1.bias"]),
                                     ArenaLinear(h2, 1, views[f"{q}.layers.2.weight"], views[f"{q}.layers.2.bias"])])
        self.final = ArenaLinear(2, 1, views[f"{q}.final.weight"], views[f"{q}.final.bias"])
        self.visual_network = _CnnParams(views, f"{q}.visual_network.")


class VisualDoubleCritic(_Picklable, nn.Module):
    """Reference convolutional.py:167-183.  forward(state: MultiObservation, action) -> (q1, q2)."""

    def __init__(self, obs_dim: int, act_dim: int, vis_dim: t.Tuple[int, int, int], hidden_sizes: t.List[int] = [256, 256],
                 filters: t.List[int] = [32, 64, 64], kernel_sizes: t.List[int] = [8, 4, 3], strides: t.List[int] = [4, 2, 1], device=None):
        super().__init__()
        hidden_sizes, filters, kernel_sizes, strides = list(hidden_sizes), list(filters), list(kernel_sizes), list(strides)
        self.obs_dim, self.act_dim, self.vis_dim, self.hidden_sizes = obs_dim, act_dim, tuple(vis_dim), hidden_sizes
        self.filters, self.kernel_sizes, self.strides = filters, kernel_sizes, strides
        self._dev = _device_index(device)
        self._layout = vlayout_for(obs_dim, act_dim, vis_dim, hidden_sizes, filters=filters, kernel_sizes=kernel_sizes, strides=strides)
        self.arena = torch.zeros(self._layout.critic_floats, dtype=torch.float32, device=torch.device("cuda", self._dev))
        v = vcritic_views(self._layout, self.arena, obs_dim, act_dim, self.vis_dim[0], hidden_sizes, filters, kernel_sizes, DENSE)
        self.q1 = VisualCritic(v, "q1", obs_dim, act_dim, hidden_sizes)
        self.q2 = VisualCritic(v, "q2", obs_dim, act_dim, hidden_sizes)
        self._engine = None

    def _ctor_kwargs(self):
        return dict(obs_dim=self.obs_dim, act_dim=self.act_dim, vis_dim=self.vis_dim, hidden_sizes=self.hidden_sizes, filters=self.filters,
                    kernel_sizes=self.kernel_sizes, strides=self.strides)

    def _engine_kwargs(self):
        return dict(filters=self.filters, kernel_sizes=self.kernel_sizes, strides=self.strides)

    def _fwd_engine(self) -> VisualEngine:
        if self._engine is None:
            self._engine = VisualEngine(self.obs_dim, self.act_dim, self.vis_dim, self.hidden_sizes, batch_size=_FWD_ROWS, device=self._dev,
                                        **self._engine_kwargs())
            self._engine.bind_params(critic=self.arena)
        return self._engine

    def forward(self, state: MultiObservation, action):
        f, fr, single = _obs(state, self.obs_dim, self.vis_dim, self.arena.device)
        a = torch.as_tensor(action, dtype=torch.float32, device=self.arena.device).reshape(-1, self.act_dim)
        q1, q2 = self._fwd_engine().critic_forward(f, fr, a)
        return q1, q2

    def __deepcopy__(self, memo):
        """`deepcopy(critic)` is how the reference creates the target network (sac/algorithm.py:194)."""
        new = type(self)(device=self._dev, **self._ctor_kwargs())
        new.arena.copy_(self.arena)
        for p, q in zip(new.parameters(), self.parameters()):
            p.requires_grad = q.requires_grad
        return new
