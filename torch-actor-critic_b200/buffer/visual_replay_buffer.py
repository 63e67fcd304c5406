"""HBM-resident visual replay ring with the reference's VisualReplayBuffer / VisualBatch surface
(buffer/visual_replay_buffer.py:12-66).

The reference keeps an object array of `MultiObservation(features, frame)` per slot and stacks B of them in Python on
every `sample`.  Here the ring is a struct of arrays in HBM -- `features [N,O]` f32, `frames [N,C,H,W]` **uint8** (the
frames are raw 0..255 camera pixels, environments/wall_runner.py:54-58: integral values are stored losslessly in a quarter
of the bytes), `actions`, `rewards`, `next_features`, `next_frames`, `done` -- that the CUDA engine gathers from directly
(`sacv_sample`; inside a training burst the uint8 frames go straight into im2col form).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .. import _lib
from ..environments.wall_runner import MultiObservation
from ..visual_engine import VisualEngine


@dataclass(frozen=True)
class VisualBatch:
    """Same fields as the reference VisualBatch (visual_replay_buffer.py:12-18)."""
    states: MultiObservation
    actions: torch.Tensor
    rewards: torch.Tensor
    next_states: MultiObservation
    done: torch.Tensor


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class VisualReplayBuffer:
    """VisualReplayBuffer(size, act_dim) as in the reference; the observation shapes are taken from the first `store`
    (the reference's object array does not need them) or given up front with obs_dim / vis_dim."""
    _STAGE_ROWS = 64

    def __init__(self, size: int, act_dim: int, obs_dim: int = None, vis_dim=None, device=None, seed: int = 0):
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self._dev = torch.device(device).index if not isinstance(device, int) else device
        self.max_size, self.act_dim, self._seed = int(size), int(act_dim), seed
        self.obs_dim, self.vis_dim = obs_dim, tuple(vis_dim) if vis_dim is not None else None
        self._engine = None
        self._pending = 0
        if obs_dim is not None and vis_dim is not None:
            self._allocate()

    def _allocate(self):
        O, A, V, N = self.obs_dim, self.act_dim, self.vis_dim, self.max_size
        # the private engine only serves store/sample; network dims are irrelevant to it
        self._engine = VisualEngine(O, A, V, [4, 4], batch_size=1, device=self._dev, seed=self._seed,
                                    kernel_sizes=(1, 1, 1), strides=(1, 1, 1), filters=(1, 1, 1), dense=1)
        d = self._engine.tdev()
        self._features = torch.zeros(N, O, dtype=torch.float32, device=d)
        self._frames = torch.zeros(N, *V, dtype=torch.uint8, device=d)
        self._actions = torch.zeros(N, A, dtype=torch.float32, device=d)
        self._rewards = torch.zeros(N, dtype=torch.float32, device=d)
        self._next_features = torch.zeros(N, O, dtype=torch.float32, device=d)
        self._next_frames = torch.zeros(N, *V, dtype=torch.uint8, device=d)
        self._done = torch.zeros(N, dtype=torch.bool, device=d)
        self._engine.bind_buffer(self._features, self._frames, self._actions, self._rewards, self._next_features, self._next_frames, self._done)
        n = min(self._STAGE_ROWS, N)
        self._h = dict(f=np.empty((n, O), np.float32), fr=np.empty((n, *V), np.uint8), a=np.empty((n, A), np.float32),
                       r=np.empty(n, np.float32), nf=np.empty((n, O), np.float32), nfr=np.empty((n, *V), np.uint8), d=np.empty(n, np.uint8))

    # ring state, as in the reference (ptr / size / max_size); staged rows count as stored
    @property
    def ptr(self):
        if self._engine is None:
            return 0
        p, _, c = self._engine.buffer_state()
        return (p + self._pending) % c

    @property
    def size(self):
        if self._engine is None:
            return 0
        _, s, c = self._engine.buffer_state()
        return min(s + self._pending, c)

    @staticmethod
    def _frame_u8(frame):
        fr = _np(frame)
        if fr.dtype != np.uint8:
            if not np.all((fr >= 0) & (fr <= 255) & (fr == np.rint(fr))):
                raise ValueError("frames must hold integral pixel values in 0..255 (environments/wall_runner.py:54-58 casts camera "
                                 "pixels to float); the HBM ring stores them as uint8")
            fr = fr.astype(np.uint8)
        return fr

    def store(self, obs: MultiObservation, act, rew, next_obs: MultiObservation, done):
        """One transition at `ptr`; ptr wraps, size saturates (visual_replay_buffer.py:33-47)."""
        if self._engine is None:
            self.obs_dim, self.vis_dim = int(_np(obs.features).shape[-1]), tuple(_np(obs.frame).shape[-3:])
            self._allocate()
        i, h = self._pending, self._h
        h["f"][i], h["fr"][i], h["a"][i], h["r"][i] = _np(obs.features), self._frame_u8(obs.frame), _np(act), float(rew)
        h["nf"][i], h["nfr"][i], h["d"][i] = _np(next_obs.features), self._frame_u8(next_obs.frame), bool(done)
        self._pending = i + 1
        if self._pending == len(h["r"]):
            self.flush()

    def store_many(self, features, frames, actions, rewards, next_features, next_frames, done):
        """Vectorised store of n transitions (same ring semantics as n calls of `store`)."""
        if self._engine is None:
            self.obs_dim, self.vis_dim = int(np.asarray(features).shape[-1]), tuple(np.asarray(frames).shape[-3:])
            self._allocate()
        self.flush()
        n, chunk = len(rewards), 2048
        for r0 in range(0, n, chunk):
            s = slice(r0, r0 + chunk)
            self._engine.store(features[s], self._frame_u8(frames[s]), actions[s], rewards[s], next_features[s], self._frame_u8(next_frames[s]), done[s])

    def flush(self):
        n = self._pending
        if n:
            self._pending = 0
            h = self._h
            self._engine.store(h["f"][:n], h["fr"][:n], h["a"][:n], h["r"][:n], h["nf"][:n], h["nfr"][:n], h["d"][:n])

    def sample(self, batch_size, idx=None) -> VisualBatch:
        """Uniform batch without replacement, stacked like the reference's `sample` (visual_replay_buffer.py:49-66): frames
        come back as float32 tensors of raw pixel values.  Raises ValueError when size < batch_size."""
        if self._engine is None:
            raise ValueError("Sample larger than population or is negative")
        self.flush()
        try:
            f, fr, a, r, nf, nfr, d = self._engine.sample(batch_size, idx=idx)
        except _lib.SacbError as e:
            if e.code == _lib.SACB_ERR_SAMPLE:
                raise ValueError("Sample larger than population or is negative") from None
            raise
        return VisualBatch(MultiObservation(f, fr), a, r, MultiObservation(nf, nfr), d)
