"""Python handle around one `sacb_handle` (include/sacb200.h).  PyTorch supplies device memory and the stream;
every computation happens inside libsacb200.so."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import SacbConfig, SacbLayout, check

ACTOR_PARAM_NAMES = ("layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias",
                     "mu_layer.weight", "mu_layer.bias", "log_std_layer.weight", "log_std_layer.bias")
CRITIC_PARAM_NAMES = tuple(f"{q}.layers.{i}.{wb}" for q in ("q1", "q2") for i in (0, 1, 2) for wb in ("weight", "bias"))


MAX_BURST = 64          # updates per launch (kMaxBurst in csrc/sacb_device.cuh); Engine.update splits longer bursts


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _check_tensor(t, name, dtype=torch.float32, device=None):
    if t is None:
        return
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if device is not None and t.device.index != device:
        raise ValueError(f"{name} lives on cuda:{t.device.index}, engine on cuda:{device}")


@dataclass
class Dims:
    obs_dim: int
    act_dim: int
    hidden1: int
    hidden2: int


def make_config(obs_dim, act_dim, hidden_sizes, batch_size=256, max_burst=64, device=0, alpha=0.2, gamma=0.99,
                polyak=0.995, reward_scale=1.0, log_min_std=-20.0, log_max_std=2.0, act_limit=10.0, lr_pi=3e-4,
                lr_q=3e-4, beta1=0.9, beta2=0.999, adam_eps=1e-8, flags=0, seed=0, rank=0, world=1) -> SacbConfig:
    hidden_sizes = list(hidden_sizes)
    if len(hidden_sizes) != 2:
        raise NotImplementedError("the B200 engine implements the reference's two-hidden-layer MLPs "
                                  f"(main.py:61 uses [256, 256]); got hidden_sizes={hidden_sizes}")
    cfg = SacbConfig()
    cfg.struct_size = C.sizeof(SacbConfig)
    cfg.device, cfg.obs_dim, cfg.act_dim = int(device), int(obs_dim), int(act_dim)
    cfg.hidden1, cfg.hidden2 = int(hidden_sizes[0]), int(hidden_sizes[1])
    cfg.batch_size, cfg.max_burst = int(batch_size), int(max_burst)
    cfg.alpha, cfg.gamma, cfg.polyak, cfg.reward_scale = alpha, gamma, polyak, reward_scale
    cfg.log_min_std, cfg.log_max_std, cfg.act_limit = log_min_std, log_max_std, act_limit
    cfg.lr_pi, cfg.lr_q, cfg.beta1, cfg.beta2, cfg.adam_eps = lr_pi, lr_q, beta1, beta2, adam_eps
    cfg.flags, cfg.seed, cfg.rank, cfg.world = int(flags), int(seed) & (2**64 - 1), int(rank), int(world)
    return cfg


def layout_for(obs_dim, act_dim, hidden_sizes) -> SacbLayout:
    cfg = make_config(obs_dim, act_dim, hidden_sizes)
    lay = SacbLayout()
    check(_lib.load().sacb_layout(C.byref(cfg), C.byref(lay)))
    return lay


def actor_views(lay: SacbLayout, arena: torch.Tensor, obs_dim, act_dim, hidden_sizes):
    """name -> view into the flat actor arena, shapes as in networks/linear.py:23-27."""
    h1, h2 = hidden_sizes
    spec = {
        "layers.0.weight": (lay.actor_w0, (h1, obs_dim)), "layers.0.bias": (lay.actor_b0, (h1,)),
        "layers.1.weight": (lay.actor_w1, (h2, h1)), "layers.1.bias": (lay.actor_b1, (h2,)),
        "mu_layer.weight": (lay.actor_w_mu, (act_dim, h2)), "mu_layer.bias": (lay.actor_b_mu, (act_dim,)),
        "log_std_layer.weight": (lay.actor_w_ls, (act_dim, h2)), "log_std_layer.bias": (lay.actor_b_ls, (act_dim,)),
    }
    return {k: arena[o:o + int(np.prod(s))].view(*s) for k, (o, s) in spec.items()}


def critic_views(lay: SacbLayout, arena: torch.Tensor, obs_dim, act_dim, hidden_sizes):
    """name -> view into the flat twin-critic arena, shapes as in networks/linear.py:59,75-76."""
    h1, h2 = hidden_sizes
    out = {}
    for z, q in enumerate(("q1", "q2")):
        spec = {
            f"{q}.layers.0.weight": (lay.critic_w0[z], (h1, obs_dim + act_dim)), f"{q}.layers.0.bias": (lay.critic_b0[z], (h1,)),
            f"{q}.layers.1.weight": (lay.critic_w1[z], (h2, h1)), f"{q}.layers.1.bias": (lay.critic_b1[z], (h2,)),
            f"{q}.layers.2.weight": (lay.critic_w2[z], (1, h2)), f"{q}.layers.2.bias": (lay.critic_b2[z], (1,)),
        }
        out.update({k: arena[o:o + int(np.prod(s))].view(*s) for k, (o, s) in spec.items()})
    return out


class Engine:
    """One engine = one `sacb_handle` on one GPU."""

    def __init__(self, obs_dim, act_dim, hidden_sizes, batch_size=256, max_burst=64, device=None, **scalars):
        self._h = None
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("torch-actor-critic_b200 needs a CUDA device (B200); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device(device).index if not isinstance(device, int) else device
        self.device = int(device)
        self.obs_dim, self.act_dim, self.hidden_sizes = int(obs_dim), int(act_dim), [int(h) for h in hidden_sizes]
        self.batch_size, self.max_burst = int(batch_size), max(1, min(int(max_burst), MAX_BURST))
        max_burst = self.max_burst
        self.cfg = make_config(obs_dim, act_dim, hidden_sizes, batch_size, max_burst, self.device, **scalars)
        self.layout = SacbLayout()
        check(lib.sacb_layout(C.byref(self.cfg), C.byref(self.layout)))
        h = C.c_void_p()
        check(lib.sacb_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self._lib = lib
        self._keep = {}
        self.world = 1

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.sacb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def tdev(self):
        return torch.device("cuda", self.device)

    # ------------------------------------------------------------------ binding
    def new_actor_arena(self):
        return torch.zeros(self.layout.actor_floats, dtype=torch.float32, device=self.tdev())

    def new_critic_arena(self):
        return torch.zeros(self.layout.critic_floats, dtype=torch.float32, device=self.tdev())

    def bind_params(self, actor=None, critic=None, target=None):
        for t, n, sz in ((actor, "actor arena", self.layout.actor_floats), (critic, "critic arena", self.layout.critic_floats),
                         (target, "target arena", self.layout.critic_floats)):
            _check_tensor(t, n, device=self.device)
            if t is not None and t.numel() != sz:
                raise ValueError(f"{n} must have {sz} floats, got {t.numel()}")
        check(self._lib.sacb_bind_params(self._h, _ptr(actor), _ptr(critic), _ptr(target)))
        self._keep.update(actor=actor, critic=critic, target=target)

    def bind_adam(self, m_actor=None, v_actor=None, step_actor=0, m_critic=None, v_critic=None, step_critic=0):
        for t, n in ((m_actor, "m_actor"), (v_actor, "v_actor"), (m_critic, "m_critic"), (v_critic, "v_critic")):
            _check_tensor(t, n, device=self.device)
        check(self._lib.sacb_bind_adam(self._h, _ptr(m_actor), _ptr(v_actor), int(step_actor), _ptr(m_critic), _ptr(v_critic),
                                       int(step_critic)))
        self._keep.update(m_actor=m_actor, v_actor=v_actor, m_critic=m_critic, v_critic=v_critic)

    def adam_steps(self):
        a, c = C.c_int64(), C.c_int64()
        check(self._lib.sacb_get_adam_steps(self._h, C.byref(a), C.byref(c)))
        return a.value, c.value

    def set_scalar(self, name, value):
        check(self._lib.sacb_set_scalar(self._h, name.encode(), float(value)))

    def bind_buffer(self, state, actions, rewards, next_state, done):
        for t, n in ((state, "state"), (actions, "actions"), (rewards, "rewards"), (next_state, "next_state")):
            _check_tensor(t, n, device=self.device)
        if done.dtype not in (torch.uint8, torch.bool):
            raise ValueError("done must be uint8/bool")
        _check_tensor(done, "done", dtype=done.dtype, device=self.device)
        cap = rewards.shape[0]
        check(self._lib.sacb_bind_buffer(self._h, _ptr(state), _ptr(actions), _ptr(rewards), _ptr(next_state), _ptr(done), cap))
        self._keep.update(buf=(state, actions, rewards, next_state, done))

    def bind_buffer_packed(self, records):
        """Packed ring [N, pitch] float32: [state | action | reward | next_state | done | pad] per row (include/sacb200.h)."""
        _check_tensor(records, "records", device=self.device)
        check(self._lib.sacb_bind_buffer_packed(self._h, _ptr(records), int(records.shape[1]), int(records.shape[0])))
        self._keep.update(buf=(records,))

    def buffer_state(self):
        p, s, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self._lib.sacb_buffer_state(self._h, C.byref(p), C.byref(s), C.byref(c)))
        return p.value, s.value, c.value

    def set_buffer_state(self, ptr, size):
        check(self._lib.sacb_set_buffer_state(self._h, int(ptr), int(size)))

    # ------------------------------------------------------------------ replay ring
    def store(self, obs, act, rew, next_obs, done):
        """Append host rows (numpy) at the ring pointer: ReplayBuffer.store, buffer/replay_buffer.py:29-43."""
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, self.obs_dim)
        n = obs.shape[0]
        act = np.ascontiguousarray(act, np.float32).reshape(n, self.act_dim)
        rew = np.ascontiguousarray(rew, np.float32).reshape(n)
        next_obs = np.ascontiguousarray(next_obs, np.float32).reshape(n, self.obs_dim)
        done = np.ascontiguousarray(np.asarray(done).astype(np.uint8)).reshape(n)
        check(self._lib.sacb_store(self._h, obs.ctypes.data, act.ctypes.data, rew.ctypes.data, next_obs.ctypes.data,
                                   done.ctypes.data, n, self._stream()))

    def sample(self, batch, idx=None, return_idx=False):
        """ReplayBuffer.sample (buffer/replay_buffer.py:45-54): duplicate-free device draw unless idx is given."""
        d = self.tdev()
        batch = int(batch)
        if idx is not None:
            idx = torch.as_tensor(idx, dtype=torch.int64, device=d).contiguous()
        n = max(batch, 0)
        s = torch.empty(n, self.obs_dim, device=d); a = torch.empty(n, self.act_dim, device=d)
        r = torch.empty(n, device=d); s2 = torch.empty(n, self.obs_dim, device=d); dn = torch.empty(n, device=d)
        io = torch.empty(n, dtype=torch.int64, device=d) if return_idx else None
        check(self._lib.sacb_sample(self._h, batch, _ptr(idx), _ptr(s), _ptr(a), _ptr(r), _ptr(s2), _ptr(dn), _ptr(io), self._stream()))
        return (s, a, r, s2, dn, io) if return_idx else (s, a, r, s2, dn)

    # ------------------------------------------------------------------ training
    def update(self, n_updates, idx=None, eps=None):
        """n_updates x (sample -> update_critic -> update_policy -> update_targets): sac/algorithm.py:274-281.
        Returns device tensors (loss_q[n], loss_pi[n]); nothing is synchronised."""
        d = self.tdev()
        n = int(n_updates)
        if idx is not None:
            idx = torch.as_tensor(idx, dtype=torch.int64, device=d).contiguous()
            if tuple(idx.shape) != (n, self.batch_size):
                raise ValueError(f"idx must be [{n},{self.batch_size}]")
        if eps is not None:
            eps = torch.as_tensor(eps, dtype=torch.float32, device=d).contiguous()
            if tuple(eps.shape) != (n, 2, self.batch_size, self.act_dim):
                raise ValueError(f"eps must be [{n},2,{self.batch_size},{self.act_dim}]")
        lq = torch.empty(n, device=d); lp = torch.empty(n, device=d)
        # one launch per <= max_burst updates (the kernel carries a 64-entry table of Adam step scalars); a longer burst
        # (e.g. update_every = 100) is a sequence of launches -- the Adam step counters carry across calls
        for k0 in range(0, n, self.max_burst):
            m = min(self.max_burst, n - k0)
            check(self._lib.sacb_update(self._h, m, _ptr(idx[k0:k0 + m]) if idx is not None else 0,
                                        _ptr(eps[k0:k0 + m]) if eps is not None else 0, _ptr(lq[k0:]), _ptr(lp[k0:]), self._stream()))
        return lq, lp

    def load_batch(self, states, actions, rewards, next_states, done):
        B = self.batch_size
        ts = []
        for t, shape, n in ((states, (B, self.obs_dim), "states"), (actions, (B, self.act_dim), "actions"), (rewards, (B,), "rewards"),
                            (next_states, (B, self.obs_dim), "next_states"), (done, (B,), "done")):
            t = torch.as_tensor(t, dtype=torch.float32, device=self.tdev()).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"{n} must have shape {shape}, got {tuple(t.shape)} (engine batch_size={B})")
            ts.append(t)
        check(self._lib.sacb_load_batch(self._h, *[_ptr(t) for t in ts], self._stream()))
        self._keep["batch"] = ts

    def _eps(self, eps):
        if eps is None:
            return None
        eps = torch.as_tensor(eps, dtype=torch.float32, device=self.tdev()).contiguous()
        if tuple(eps.shape) != (self.batch_size, self.act_dim):
            raise ValueError(f"eps must be [{self.batch_size},{self.act_dim}]")
        return eps

    def update_critic(self, eps=None):
        eps = self._eps(eps)
        out = torch.empty(1, device=self.tdev())
        check(self._lib.sacb_update_critic(self._h, _ptr(eps), _ptr(out), self._stream()))
        return out[0]

    def update_policy(self, eps=None):
        eps = self._eps(eps)
        out = torch.empty(1, device=self.tdev())
        check(self._lib.sacb_update_policy(self._h, _ptr(eps), _ptr(out), self._stream()))
        return out[0]

    def update_targets(self):
        check(self._lib.sacb_update_targets(self._h, self._stream()))

    # ------------------------------------------------------------------ forward passes
    def actor_forward(self, x, eps=None, deterministic=False, with_logprob=True):
        d = self.tdev()
        x = torch.as_tensor(x, dtype=torch.float32, device=d).contiguous()
        n = x.shape[0]
        if eps is not None:
            eps = torch.as_tensor(eps, dtype=torch.float32, device=d).contiguous()
        pi = torch.empty(n, self.act_dim, device=d)
        logp = torch.empty(n, device=d) if with_logprob else None
        check(self._lib.sacb_actor_forward(self._h, _ptr(x), n, _ptr(eps), 1 if deterministic else 0, _ptr(pi), _ptr(logp), self._stream()))
        return pi, logp

    def critic_forward(self, state, action, which=0):
        d = self.tdev()
        state = torch.as_tensor(state, dtype=torch.float32, device=d).contiguous()
        action = torch.as_tensor(action, dtype=torch.float32, device=d).contiguous()
        n = state.shape[0]
        q1 = torch.empty(n, device=d); q2 = torch.empty(n, device=d)
        check(self._lib.sacb_critic_forward(self._h, int(which), _ptr(state), _ptr(action), n, _ptr(q1), _ptr(q2), self._stream()))
        return q1, q2

    # ------------------------------------------------------------------ data parallel
    def comm_init_from_torch(self):
        """Create the engine's NCCL communicator from the default torch.distributed group (replaces sac/mpi.py)."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        path = _lib.nccl_library_path()
        pb = path.encode() if path else None
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            check(self._lib.sacb_comm_unique_id(pb, uid))
        obj = [bytes(uid)]
        dist.broadcast_object_list(obj, src=0)
        uid = (C.c_uint8 * 128).from_buffer_copy(obj[0])
        check(self._lib.sacb_comm_init(self._h, pb, uid, rank, world))
        self.world = world

    def peer_init_from_torch(self):
        """Fused in-kernel NVLink gradient exchange (include/sacb200.h, sacb_peer_bind): map every rank's engine workspace
        into every other rank via CUDA IPC; handles travel over the default torch.distributed group."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        hnd = (C.c_uint8 * 64)()
        n = C.c_int64()
        check(self._lib.sacb_peer_export(self._h, hnd, C.byref(n)))
        gathered = [None] * world
        dist.all_gather_object(gathered, (bytes(hnd), n.value))
        if len({g[1] for g in gathered}) != 1:
            raise RuntimeError("peer mode needs identical engine configurations on all ranks")
        blob = b"".join(g[0] for g in gathered)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(self._lib.sacb_peer_bind(self._h, buf, rank, world))
        self.world = world
        dist.barrier()

    def sync_params(self):
        check(self._lib.sacb_sync_params(self._h, self._stream()))

    def peer_status(self):
        """Raises SacbError(SACB_ERR_PEER) when the in-kernel gradient exchange timed out on a dead / stalled rank."""
        check(self._lib.sacb_peer_status(self._h, self._stream()))

    # ------------------------------------------------------------------ debug
    def debug_read(self, name):
        cap = max(self.layout.actor_floats, self.layout.critic_floats, 16 * self.batch_size * max(self.act_dim, 1))
        buf = np.empty(cap, np.float32)
        n = C.c_int64()
        check(self._lib.sacb_debug_read(self._h, name.encode(), buf.ctypes.data, cap, C.byref(n)))
        out = buf[: n.value].copy()
        if name in ("a2", "pi"):
            out = out.reshape(self.batch_size, self.act_dim)
        return out

    def profile_phases(self, n_updates):
        """[n_updates, n_phases, n_ctas, 2] globaltimer stamps (ns): after each phase's work / after its grid barrier."""
        cap = int(n_updates) * 64 * 1024 * 2
        buf = np.zeros(cap, np.uint64)
        nph, ncta = C.c_int32(), C.c_int32()
        check(self._lib.sacb_debug_profile(self._h, int(n_updates), buf.ctypes.data, cap, C.byref(nph), C.byref(ncta), self._stream()))
        n = int(n_updates) * nph.value * ncta.value * 2
        return buf[:n].reshape(int(n_updates), nph.value, ncta.value, 2)

    def launch_count(self):
        return int(self._lib.sacb_launch_count(self._h))

    def set_mode(self, mode):
        check(self._lib.sacb_set_mode(self._h, int(mode)))

    def set_precision(self, tf32x3):
        """1 (default): mma.sync tensor-core GEMMs with 3xTF32 error compensation; 0: plain fp32 FFMA; 2: 3xTF32 on
        tcgen05/TMEM; 3: tcgen05 with bf16 operands, fp32 accumulation (Humanoid-shaped "bf16 MLP" configuration)."""
        check(self._lib.sacb_set_precision(self._h, int(tf32x3)))
