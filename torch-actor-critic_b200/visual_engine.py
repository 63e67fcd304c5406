"""Python handle around one `sacv_handle` (include/sacb200.h): the visual SAC engine.  PyTorch supplies device memory and
the stream; every computation happens inside libsacb200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SacvConfig, SacvLayout, checkv

FILTERS, KERNELS, STRIDES, DENSE = (32, 64, 64), (8, 4, 3), (4, 2, 1), 512      # reference defaults, convolutional.py:17-22


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def cnn_param_names(prefix):
    names = []
    for i in range(3):
        names += [f"{prefix}conv_{i}.weight", f"{prefix}conv_{i}.bias"]
    return names + [f"{prefix}linear.weight", f"{prefix}linear.bias", f"{prefix}final.weight", f"{prefix}final.bias"]


def actor_param_names():
    """named_parameters() order of the reference VisualActor (convolutional.py:66-83)."""
    return (["layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias"] + cnn_param_names("visual_network.") +
            ["mu_layer.weight", "mu_layer.bias", "log_std_layer.weight", "log_std_layer.bias"])


def critic_param_names():
    """named_parameters() order of the reference VisualDoubleCritic (convolutional.py:140-143,179-180)."""
    names = []
    for q in ("q1", "q2"):
        names += [f"{q}.layers.{i}.{wb}" for i in range(3) for wb in ("weight", "bias")]
        names += [f"{q}.final.weight", f"{q}.final.bias"] + cnn_param_names(f"{q}.visual_network.")
    return names


def make_vconfig(obs_dim, act_dim, vis_dim, hidden_sizes, batch_size=256, device=0, filters=FILTERS, kernel_sizes=KERNELS,
                 strides=STRIDES, dense=DENSE, alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, log_min_std=-20.0,
                 log_max_std=2.0, act_limit=10.0, lr_pi=3e-4, lr_q=3e-4, beta1=0.9, beta2=0.999, adam_eps=1e-8, flags=0, seed=0,
                 precision=1) -> SacvConfig:
    hidden_sizes = list(hidden_sizes)
    if len(hidden_sizes) != 2 or len(filters) != 3 or len(kernel_sizes) != 3 or len(strides) != 3:
        raise NotImplementedError("the visual engine implements the reference architecture: two hidden layers, three conv "
                                  "layers (main.py:61-67)")
    cfg = SacvConfig()
    cfg.struct_size = C.sizeof(SacvConfig)
    cfg.device, cfg.obs_dim, cfg.act_dim = int(device), int(obs_dim), int(act_dim)
    cfg.hidden1, cfg.hidden2, cfg.batch_size = int(hidden_sizes[0]), int(hidden_sizes[1]), int(batch_size)
    cfg.vis_c, cfg.vis_h, cfg.vis_w = (int(v) for v in vis_dim)
    for i in range(3):
        cfg.filters[i], cfg.kernels[i], cfg.strides[i] = int(filters[i]), int(kernel_sizes[i]), int(strides[i])
    cfg.dense = int(dense)
    cfg.log_min_std, cfg.log_max_std, cfg.act_limit, cfg.precision = log_min_std, log_max_std, act_limit, int(precision)
    cfg.alpha, cfg.gamma, cfg.polyak, cfg.reward_scale = alpha, gamma, polyak, reward_scale
    cfg.lr_pi, cfg.lr_q, cfg.beta1, cfg.beta2, cfg.adam_eps = lr_pi, lr_q, beta1, beta2, adam_eps
    cfg.flags, cfg.seed = int(flags), int(seed) & (2**64 - 1)
    return cfg


def vlayout_for(obs_dim, act_dim, vis_dim, hidden_sizes, **kw) -> SacvLayout:
    cfg = make_vconfig(obs_dim, act_dim, vis_dim, hidden_sizes, **kw)
    lay = SacvLayout()
    checkv(_lib.load().sacv_layout(C.byref(cfg), C.byref(lay)))
    return lay


def _cnn_spec(lay, base, prefix, vis_c, filters, kernels, dense):
    spec, cin = {}, vis_c
    for i in range(3):
        # conv 0 is stored [F, C, KH, KW] like the reference; conv 1 / 2 are stored [F, KH, KW, C] (channels innermost: the
        # im2col of a [pixels, channels] activation is then contiguous copies) and exposed as a permuted view of that shape
        spec[f"{prefix}conv_{i}.weight"] = (base + lay.cnn_w[i], (filters[i], cin, kernels[i], kernels[i]) if i == 0 else
                                            ("nhwc", filters[i], cin, kernels[i], kernels[i]))
        spec[f"{prefix}conv_{i}.bias"] = (base + lay.cnn_b[i], (filters[i],))
        cin = filters[i]
    spec[f"{prefix}linear.weight"] = (base + lay.cnn_lw, (dense, lay.flat)); spec[f"{prefix}linear.bias"] = (base + lay.cnn_lb, (dense,))
    spec[f"{prefix}final.weight"] = (base + lay.cnn_fw, (1, dense)); spec[f"{prefix}final.bias"] = (base + lay.cnn_fb, (1,))
    return spec


def _view(arena, o, s):
    if s[0] == "nhwc":
        _, f, c, kh, kw = s
        return arena[o:o + f * c * kh * kw].view(f, kh, kw, c).permute(0, 3, 1, 2)
    return arena[o:o + int(np.prod(s))].view(*s)


def vactor_views(lay, arena, obs_dim, act_dim, vis_c, hidden, filters=FILTERS, kernels=KERNELS, dense=DENSE):
    """name -> view into the flat actor arena, shapes as the reference VisualActor's state_dict."""
    h1, h2 = hidden
    spec = {"layers.0.weight": (lay.a_l0w, (h1, obs_dim)), "layers.0.bias": (lay.a_l0b, (h1,)),
            "layers.1.weight": (lay.a_l1w, (h2, h1)), "layers.1.bias": (lay.a_l1b, (h2,))}
    spec.update(_cnn_spec(lay, lay.a_cnn, "visual_network.", vis_c, filters, kernels, dense))
    spec.update({"mu_layer.weight": (lay.a_muw, (act_dim, h2 + 1)), "mu_layer.bias": (lay.a_mub, (act_dim,)),
                 "log_std_layer.weight": (lay.a_lsw, (act_dim, h2 + 1)), "log_std_layer.bias": (lay.a_lsb, (act_dim,))})
    return {k: _view(arena, o, s) for k, (o, s) in spec.items()}


def vcritic_views(lay, arena, obs_dim, act_dim, vis_c, hidden, filters=FILTERS, kernels=KERNELS, dense=DENSE):
    h1, h2 = hidden
    out = {}
    for z, q in enumerate(("q1", "q2")):
        b = z * lay.critic_net_floats
        spec = {f"{q}.layers.0.weight": (b + lay.c_l0w, (h1, obs_dim + act_dim)), f"{q}.layers.0.bias": (b + lay.c_l0b, (h1,)),
                f"{q}.layers.1.weight": (b + lay.c_l1w, (h2, h1)), f"{q}.layers.1.bias": (b + lay.c_l1b, (h2,)),
                f"{q}.layers.2.weight": (b + lay.c_l2w, (1, h2)), f"{q}.layers.2.bias": (b + lay.c_l2b, (1,)),
                f"{q}.final.weight": (b + lay.c_fw, (1, 2)), f"{q}.final.bias": (b + lay.c_fb, (1,))}
        spec.update(_cnn_spec(lay, b + lay.c_cnn, f"{q}.visual_network.", vis_c, filters, kernels, dense))
        out.update({k: _view(arena, o, s) for k, (o, s) in spec.items()})
    return out


class VisualEngine:
    """One engine = one `sacv_handle` on one GPU, for a fixed batch size."""

    def __init__(self, obs_dim, act_dim, vis_dim, hidden_sizes, batch_size=256, device=None, **kw):
        self._h = None
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("torch-actor-critic_b200 needs a CUDA device (B200); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device(device).index if not isinstance(device, int) else device
        self.device = int(device)
        self.obs_dim, self.act_dim, self.vis_dim = int(obs_dim), int(act_dim), tuple(int(v) for v in vis_dim)
        self.hidden_sizes, self.batch_size = [int(h) for h in hidden_sizes], int(batch_size)
        self.cfg = make_vconfig(obs_dim, act_dim, vis_dim, hidden_sizes, batch_size, self.device, **kw)
        self.layout = SacvLayout()
        checkv(lib.sacv_layout(C.byref(self.cfg), C.byref(self.layout)))
        h = C.c_void_p()
        checkv(lib.sacv_create(C.byref(self.cfg), C.byref(h)))
        self._h, self._lib, self._keep = h, lib, {}
        self.frame_floats = int(np.prod(self.vis_dim))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sacv_destroy(self._h)
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

    def new_actor_arena(self):
        return torch.zeros(self.layout.actor_floats, dtype=torch.float32, device=self.tdev())

    def new_critic_arena(self):
        return torch.zeros(self.layout.critic_floats, dtype=torch.float32, device=self.tdev())

    def bind_params(self, actor=None, critic=None, target=None):
        checkv(self._lib.sacv_bind_params(self._h, _ptr(actor), _ptr(critic), _ptr(target)))
        self._keep.update(actor=actor, critic=critic, target=target)

    def bind_adam(self, m_actor=None, v_actor=None, step_actor=0, m_critic=None, v_critic=None, step_critic=0):
        checkv(self._lib.sacv_bind_adam(self._h, _ptr(m_actor), _ptr(v_actor), int(step_actor), _ptr(m_critic), _ptr(v_critic), int(step_critic)))
        self._keep.update(m_actor=m_actor, v_actor=v_actor, m_critic=m_critic, v_critic=v_critic)

    def adam_steps(self):
        a, c = C.c_int64(), C.c_int64()
        checkv(self._lib.sacv_get_adam_steps(self._h, C.byref(a), C.byref(c)))
        return a.value, c.value

    def set_scalar(self, name, value):
        checkv(self._lib.sacv_set_scalar(self._h, name.encode(), float(value)))

    def bind_buffer(self, features, frames, actions, rewards, next_features, next_frames, done):
        checkv(self._lib.sacv_bind_buffer(self._h, _ptr(features), _ptr(frames), _ptr(actions), _ptr(rewards), _ptr(next_features),
                                          _ptr(next_frames), _ptr(done), rewards.shape[0]))
        self._keep.update(buf=(features, frames, actions, rewards, next_features, next_frames, done))

    def buffer_state(self):
        p, s, c = C.c_int64(), C.c_int64(), C.c_int64()
        checkv(self._lib.sacv_buffer_state(self._h, C.byref(p), C.byref(s), C.byref(c)))
        return p.value, s.value, c.value

    def set_buffer_state(self, ptr, size):
        checkv(self._lib.sacv_set_buffer_state(self._h, int(ptr), int(size)))

    def store(self, features, frames, actions, rewards, next_features, next_frames, done):
        """Append host rows at the ring pointer (VisualReplayBuffer.store, visual_replay_buffer.py:33-47); frames uint8."""
        f = np.ascontiguousarray(features, np.float32).reshape(-1, self.obs_dim)
        n = f.shape[0]
        fr = np.ascontiguousarray(frames, np.uint8).reshape(n, self.frame_floats)
        a = np.ascontiguousarray(actions, np.float32).reshape(n, self.act_dim)
        r = np.ascontiguousarray(rewards, np.float32).reshape(n)
        nf = np.ascontiguousarray(next_features, np.float32).reshape(n, self.obs_dim)
        nfr = np.ascontiguousarray(next_frames, np.uint8).reshape(n, self.frame_floats)
        d = np.ascontiguousarray(np.asarray(done).astype(np.uint8)).reshape(n)
        checkv(self._lib.sacv_store(self._h, f.ctypes.data, fr.ctypes.data, a.ctypes.data, r.ctypes.data, nf.ctypes.data, nfr.ctypes.data,
                                    d.ctypes.data, n, self._stream()))

    def sample(self, batch, idx=None):
        d = self.tdev()
        batch = int(batch)
        if idx is not None:
            idx = torch.as_tensor(idx, dtype=torch.int64, device=d).contiguous()
        n = max(batch, 0)
        f = torch.empty(n, self.obs_dim, device=d); fr = torch.empty(n, *self.vis_dim, device=d)
        a = torch.empty(n, self.act_dim, device=d); r = torch.empty(n, device=d)
        nf = torch.empty(n, self.obs_dim, device=d); nfr = torch.empty(n, *self.vis_dim, device=d); dn = torch.empty(n, device=d)
        checkv(self._lib.sacv_sample(self._h, batch, _ptr(idx), _ptr(f), _ptr(fr), _ptr(a), _ptr(r), _ptr(nf), _ptr(nfr), _ptr(dn), 0, self._stream()))
        return f, fr, a, r, nf, nfr, dn

    def update(self, n_updates, idx=None, eps=None):
        d, n = self.tdev(), int(n_updates)
        if idx is not None:
            idx = torch.as_tensor(idx, dtype=torch.int64, device=d).contiguous()
            if tuple(idx.shape) != (n, self.batch_size):
                raise ValueError(f"idx must be [{n},{self.batch_size}]")
        if eps is not None:
            eps = torch.as_tensor(eps, dtype=torch.float32, device=d).contiguous()
            if tuple(eps.shape) != (n, 2, self.batch_size, self.act_dim):
                raise ValueError(f"eps must be [{n},2,{self.batch_size},{self.act_dim}]")
        lq = torch.empty(n, device=d); lp = torch.empty(n, device=d)
        checkv(self._lib.sacv_update(self._h, n, _ptr(idx), _ptr(eps), _ptr(lq), _ptr(lp), self._stream()))
        return lq, lp

    def load_batch(self, features, frames, actions, rewards, next_features, next_frames, done):
        B, d = self.batch_size, self.tdev()
        ts = []
        for t, shape, nme in ((features, (B, self.obs_dim), "features"), (frames, (B, *self.vis_dim), "frames"), (actions, (B, self.act_dim), "actions"),
                              (rewards, (B,), "rewards"), (next_features, (B, self.obs_dim), "next_features"),
                              (next_frames, (B, *self.vis_dim), "next_frames"), (done, (B,), "done")):
            t = torch.as_tensor(t, dtype=torch.float32, device=d).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"{nme} must have shape {shape}, got {tuple(t.shape)} (engine batch_size={B})")
            ts.append(t)
        checkv(self._lib.sacv_load_batch(self._h, *[_ptr(t) for t in ts], self._stream()))
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
        checkv(self._lib.sacv_update_critic(self._h, _ptr(eps), _ptr(out), self._stream()))
        return out[0]

    def update_policy(self, eps=None):
        eps = self._eps(eps)
        out = torch.empty(1, device=self.tdev())
        checkv(self._lib.sacv_update_policy(self._h, _ptr(eps), _ptr(out), self._stream()))
        return out[0]

    def update_targets(self):
        checkv(self._lib.sacv_update_targets(self._h, self._stream()))

    def actor_forward(self, features, frames, eps=None, deterministic=False, with_logprob=True):
        d = self.tdev()
        f = torch.as_tensor(features, dtype=torch.float32, device=d).contiguous().reshape(-1, self.obs_dim)
        fr = torch.as_tensor(frames, dtype=torch.float32, device=d).contiguous().reshape(-1, *self.vis_dim)
        n = f.shape[0]
        pis, lps = [], []
        for r0 in range(0, n, self.batch_size):
            m = min(self.batch_size, n - r0)
            e = None if eps is None else torch.as_tensor(eps, dtype=torch.float32, device=d).contiguous()[r0:r0 + m].contiguous()
            pi = torch.empty(m, self.act_dim, device=d)
            lp = torch.empty(m, device=d) if with_logprob else None
            checkv(self._lib.sacv_actor_forward(self._h, _ptr(f[r0:r0 + m]), _ptr(fr[r0:r0 + m]), m, _ptr(e), 1 if deterministic else 0, _ptr(pi),
                                                _ptr(lp), self._stream()))
            pis.append(pi); lps.append(lp)
        return torch.cat(pis), (torch.cat(lps) if with_logprob else None)

    def critic_forward(self, features, frames, actions, which=0):
        d = self.tdev()
        f = torch.as_tensor(features, dtype=torch.float32, device=d).contiguous().reshape(-1, self.obs_dim)
        fr = torch.as_tensor(frames, dtype=torch.float32, device=d).contiguous().reshape(-1, *self.vis_dim)
        a = torch.as_tensor(actions, dtype=torch.float32, device=d).contiguous().reshape(-1, self.act_dim)
        n = f.shape[0]
        q1s, q2s = [], []
        for r0 in range(0, n, self.batch_size):
            m = min(self.batch_size, n - r0)
            q1 = torch.empty(m, device=d); q2 = torch.empty(m, device=d)
            checkv(self._lib.sacv_critic_forward(self._h, int(which), _ptr(f[r0:r0 + m]), _ptr(fr[r0:r0 + m]), _ptr(a[r0:r0 + m]), m, _ptr(q1), _ptr(q2),
                                                 self._stream()))
            q1s.append(q1); q2s.append(q2)
        return torch.cat(q1s), torch.cat(q2s)

    def debug_read(self, name):
        cap = max(self.layout.actor_floats, self.layout.critic_floats, 16 * self.batch_size * max(self.act_dim, self.obs_dim, 1))
        buf = np.empty(cap, np.float32)
        n = C.c_int64()
        checkv(self._lib.sacv_debug_read(self._h, name.encode(), buf.ctypes.data, cap, C.byref(n)))
        out = buf[: n.value].copy()
        if name in ("a2", "pi"):
            out = out.reshape(self.batch_size, self.act_dim)
        return out

    def launch_count(self):
        return int(self._lib.sacv_launch_count(self._h))
