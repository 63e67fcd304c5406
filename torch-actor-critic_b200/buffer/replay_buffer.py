"""HBM-resident replay ring with the reference's ReplayBuffer / Batch surface (buffer/replay_buffer.py:8-54).

The ring is ONE packed torch CUDA tensor `[N, pitch]` of float32 records `[state(O) | action(A) | reward | next_state(O) |
done | pad]`, pitch a multiple of 64 bytes, so that a sampled transition is a single aligned, contiguous read (inside the
update kernel: one TMA bulk copy per row); `state`, `actions`, `rewards`, `next_state`, `done` -- the reference's five
attributes -- are strided views of it.  `store` stages rows in host memory and flushes them with one H2D copy per ring
segment (`sacb_store`); `sample` draws a duplicate-free index set on the device (cuRAND Philox) and gathers with a CUDA
kernel (`sacb_sample`).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .. import _lib
from ..engine import Engine


@dataclass(frozen=True)
class Batch:
    """Same fields as the reference Batch (buffer/replay_buffer.py:8-14); float32 CUDA tensors."""
    states: torch.Tensor
    actions: torch.Tensor
    rewards: torch.Tensor
    next_states: torch.Tensor
    done: torch.Tensor


class ReplayBuffer:
    _STAGE_ROWS = 512

    def __init__(self, size: int, obs_dim: int, act_dim: int, device=None, seed: int = 0):
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self._dev = torch.device(device).index if not isinstance(device, int) else device
        # the private engine only serves store/sample; network dims are irrelevant to it
        self._engine = Engine(obs_dim, act_dim, [4, 4], batch_size=256, max_burst=1, device=self._dev, seed=seed)
        d = self._engine.tdev()
        self.obs_dim, self.act_dim = obs_dim, act_dim
        O, A = obs_dim, act_dim
        self.pitch = (2 * O + A + 2 + 15) // 16 * 16                    # floats; records start on 64-byte boundaries
        self._rec = torch.zeros(size, self.pitch, dtype=torch.float32, device=d)
        self._state, self._actions = self._rec[:, :O], self._rec[:, O:O + A]
        self._rewards, self._next_state = self._rec[:, O + A], self._rec[:, O + A + 1:2 * O + A + 1]
        self._done_f = self._rec[:, 2 * O + A + 1]
        self._engine.bind_buffer_packed(self._rec)
        self.max_size = size
        # host staging (flushed lazily: before any sample/update/array access)
        n = min(self._STAGE_ROWS, size)
        self._h_obs = np.empty((n, obs_dim), np.float32)
        self._h_act = np.empty((n, act_dim), np.float32)
        self._h_rew = np.empty(n, np.float32)
        self._h_next = np.empty((n, obs_dim), np.float32)
        self._h_done = np.empty(n, np.uint8)
        self._pending = 0

    # ring state, as in the reference (ptr / size / max_size); pending staged rows count as stored
    @property
    def ptr(self):
        p, _, c = self._engine.buffer_state()
        return (p + self._pending) % c

    @property
    def size(self):
        _, s, c = self._engine.buffer_state()
        return min(s + self._pending, c)

    @size.setter
    def size(self, value):
        self.flush()
        p, _, _ = self._engine.buffer_state()
        self._engine.set_buffer_state(p, int(value))

    def set_ring_state(self, ptr, size):
        self.flush()
        self._engine.set_buffer_state(int(ptr), int(size))

    # the five arrays (device tensors); reading them flushes staged rows first
    state = property(lambda self: (self.flush(), self._state)[1])
    actions = property(lambda self: (self.flush(), self._actions)[1])
    rewards = property(lambda self: (self.flush(), self._rewards)[1])
    next_state = property(lambda self: (self.flush(), self._next_state)[1])
    done = property(lambda self: (self.flush(), self._done_f != 0)[1])

    def store(self, obs, act, rew, next_obs, done):
        """One transition at `ptr`; ptr wraps, size saturates (buffer/replay_buffer.py:29-43)."""
        i = self._pending
        self._h_obs[i], self._h_act[i], self._h_rew[i], self._h_next[i], self._h_done[i] = obs, act, rew, next_obs, bool(done)
        self._pending = i + 1
        if self._pending == len(self._h_rew):
            self.flush()

    def store_many(self, obs, act, rew, next_obs, done):
        """Vectorised store of n transitions (same ring semantics as n calls of `store`)."""
        self.flush()
        self._engine.store(obs, act, rew, next_obs, done)

    def flush(self):
        n = self._pending
        if n:
            self._pending = 0
            self._engine.store(self._h_obs[:n], self._h_act[:n], self._h_rew[:n], self._h_next[:n], self._h_done[:n])

    def sample(self, batch_size, idx=None) -> Batch:
        """Uniform batch without replacement (buffer/replay_buffer.py:45-54).  Raises ValueError when
        size < batch_size, as `random.sample` does in the reference."""
        self.flush()
        try:
            s, a, r, s2, d = self._engine.sample(batch_size, idx=idx)
        except _lib.SacbError as e:
            if e.code == _lib.SACB_ERR_SAMPLE:
                raise ValueError("Sample larger than population or is negative") from None
            raise
        return Batch(s, a, r, s2, d)
