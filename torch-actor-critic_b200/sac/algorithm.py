"""Soft Actor-Critic trainer with the reference's `sac.algorithm` surface (sac/algorithm.py:30-307), driving the
B200 engine.

`SAC.update_critic`, `SAC.update_policy` and `update_targets` keep their signatures and update the given modules /
optimizers in place, but each is a handful of CUDA kernel launches on HBM-resident state instead of ~45 torch CPU ops
and two autograd graphs.  `SAC.train` issues each burst of `update_every` gradient updates (sac/algorithm.py:273-281)
as ONE persistent-kernel launch (`Engine.update`) and reads the losses back once per burst.
"""
from __future__ import annotations

import logging
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from ..buffer.replay_buffer import Batch, ReplayBuffer
from ..engine import ACTOR_PARAM_NAMES, CRITIC_PARAM_NAMES, Engine, actor_views, critic_views
from ..buffer.visual_replay_buffer import VisualBatch, VisualReplayBuffer
from ..networks.convolutional import VisualActor, VisualDoubleCritic
from ..networks.linear import Actor, DoubleCritic
from ..visual_engine import VisualEngine, actor_param_names, critic_param_names, vactor_views, vcritic_views
from .mpi import num_procs, proc_id, setup_pytorch_for_mpi, sync_params

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)


# ----------------------------------------------------------------------------------------------------------------
# loss evaluation (diagnostic, forward only).  Training does not call these: the engine fuses them with backward.
# ----------------------------------------------------------------------------------------------------------------
def eval_pi_loss(actor: nn.Module, critic: nn.Module, state, next_state, alpha: float) -> torch.Tensor:
    """loss_pi = mean(alpha*logp - min(q1,q2)) with the policy evaluated at `next_state` and Q at
    (state, pi) -- exactly as the reference does (sac/algorithm.py:30-43).  No autograd graph is built."""
    pi, logp_pi = actor(next_state)
    q1, q2 = critic(state, pi)
    return (alpha * logp_pi - torch.min(q1, q2)).mean()


def eval_q_loss(actor, critic, target_critic, states, actions, rewards, next_states, done, alpha: float, gamma: float,
                reward_scale: float) -> torch.Tensor:
    """Twin MSE against the soft Bellman backup (sac/algorithm.py:46-74).  Forward only."""
    a2, logp = actor(next_states)
    q1t, q2t = target_critic(next_states, a2)
    dev = q1t.device
    rewards = torch.as_tensor(rewards, dtype=torch.float32, device=dev)
    done = torch.as_tensor(done, dtype=torch.float32, device=dev)
    backup = reward_scale * rewards + gamma * (1 - done) * (torch.min(q1t, q2t) - alpha * logp)
    q1, q2 = critic(states, actions)
    return ((q1 - backup) ** 2).mean() + ((q2 - backup) ** 2).mean()


def _polyak_engine(source: DoubleCritic, target: DoubleCritic) -> Engine:
    eng = getattr(target, "_polyak_engine", None)
    if eng is None or eng._keep.get("critic") is not source.arena:
        eng = Engine(source.obs_dim, source.act_dim, source.hidden_sizes, batch_size=8, max_burst=1, device=source._dev)
        eng.bind_params(critic=source.arena, target=target.arena)
        object.__setattr__(target, "_polyak_engine", eng)
    return eng


def _polyak_vengine(source: VisualDoubleCritic, target: VisualDoubleCritic) -> VisualEngine:
    eng = getattr(target, "_polyak_engine", None)
    if eng is None or eng._keep.get("critic") is not source.arena:
        eng = VisualEngine(source.obs_dim, source.act_dim, source.vis_dim, source.hidden_sizes, batch_size=1, device=source._dev,
                           **source._engine_kwargs())
        eng.bind_params(critic=source.arena, target=target.arena)
        object.__setattr__(target, "_polyak_engine", eng)
    return eng


def update_targets(source: nn.Module, target: nn.Module, polyak: float):
    """targ <- polyak*targ + (1-polyak)*src for every parameter (sac/algorithm.py:77-81), one fused kernel."""
    if isinstance(source, VisualDoubleCritic) and isinstance(target, VisualDoubleCritic):
        eng = _polyak_vengine(source, target)
        eng.set_scalar("polyak", polyak)
        eng.update_targets()
        return
    if not (isinstance(source, DoubleCritic) and isinstance(target, DoubleCritic)):
        raise TypeError("update_targets expects arena-backed DoubleCritic modules")
    eng = _polyak_engine(source, target)
    eng.set_scalar("polyak", polyak)
    eng.update_targets()


# ----------------------------------------------------------------------------------------------------------------
# optimizer <-> engine plumbing
# ----------------------------------------------------------------------------------------------------------------
def _adam_hparams(opt: optim.Optimizer):
    if not isinstance(opt, optim.Adam):
        raise TypeError(f"the engine implements torch.optim.Adam (main.py:93-95); got {type(opt).__name__}")
    if len(opt.param_groups) != 1:
        raise NotImplementedError("exactly one param group is supported")
    g = opt.param_groups[0]
    if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
        raise NotImplementedError("weight_decay / amsgrad / maximize are not implemented in the engine's Adam")
    return float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"])


class _AdamBinding:
    """Flat exp_avg / exp_avg_sq arenas for one optimizer, exposed to torch as views in `opt.state` so that
    `opt.state_dict()` / `load_state_dict` (checkpointing, sac/algorithm.py:176-178, main.py:38-42) keep working."""

    def __init__(self, opt, module, names, views_fn, n_floats, device):
        self.opt = opt
        self.m = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.v = torch.zeros(n_floats, dtype=torch.float32, device=device)
        mv, vv = views_fn(self.m), views_fn(self.v)
        params = dict(module.named_parameters())
        group = opt.param_groups[0]["params"]
        if len(group) != len(names) or any(params[n] is not p for n, p in zip(names, group)):
            raise ValueError("optimizer must have been built from module.parameters() (main.py:94-95)")
        self.steps = 0
        self.step_tensor = torch.tensor(0.0, dtype=torch.float32)
        for n in names:
            p = params[n]
            st = opt.state.get(p, {})
            if st:                                    # resumed optimizer: adopt its moments
                mv[n].copy_(st["exp_avg"].to(device))
                vv[n].copy_(st["exp_avg_sq"].to(device))
                self.steps = int(float(st["step"]))
            opt.state[p] = {"step": self.step_tensor, "exp_avg": mv[n], "exp_avg_sq": vv[n]}
        self.step_tensor.fill_(float(self.steps))
        self._probe = (params[names[0]], mv[names[0]].data_ptr())

    def still_bound(self) -> bool:
        """False once `opt.load_state_dict()` (or anything else) replaced the moment tensors the engine steps: the
        binding must then be rebuilt so that the loaded state is adopted instead of silently ignored."""
        p, ptr = self._probe
        st = self.opt.state.get(p)
        return bool(st) and st.get("exp_avg") is not None and st["exp_avg"].data_ptr() == ptr and st.get("step") is self.step_tensor

    def advance(self, n):
        self.steps += n
        self.step_tensor.fill_(float(self.steps))


class SAC(object):
    def __init__(self, alpha: float, gamma: float, polyak: float, reward_scale: float, epochs: int, batch_size: int,
                 start_steps: int, steps_per_epoch: int, max_ep_len: int, update_after: int, update_every: int,
                 save_every: int):
        self.alpha = alpha
        self.gamma = gamma
        self.polyak = polyak
        self.reward_scale = reward_scale
        self.epochs = epochs
        self.batch_size = batch_size
        self.start_steps = start_steps
        self.steps_per_epoch = steps_per_epoch
        self.max_ep_len = max_ep_len
        self.update_after = update_after
        self.update_every = update_every
        self.save_every = save_every
        self._engines = {}
        self._bindings = {}          # id(optimizer) -> _AdamBinding, shared by every engine that steps that optimizer
        self.seed = 0
        setup_pytorch_for_mpi()

    # ------------------------------------------------------------------------------------------------------------
    def _engine(self, actor: Actor, critic: DoubleCritic, batch: int, target: DoubleCritic = None, pi_opt=None, q_opt=None) -> Engine:
        if not isinstance(actor, Actor) or not isinstance(critic, DoubleCritic):
            raise TypeError("SAC drives arena-backed Actor / DoubleCritic modules (networks/linear.py of this package)")
        key = (id(actor), id(critic), int(batch))
        ent = self._engines.get(key)
        if ent is None:
            eng = Engine(actor.obs_dim, actor.act_dim, actor.hidden_sizes, batch_size=batch, max_burst=max(int(self.update_every), 1),   # clamped to 64 per launch; longer bursts are split (Engine.update)
                         device=actor._dev, alpha=self.alpha, gamma=self.gamma, polyak=self.polyak,
                         reward_scale=self.reward_scale, log_min_std=actor.log_min_std, log_max_std=actor.log_max_std,
                         act_limit=actor.act_limit, seed=self.seed, rank=proc_id(), world=num_procs())
            ent = dict(engine=eng, target=None, pi=None, q=None)
            eng.bind_params(actor=actor.arena, critic=critic.arena, target=None)
            if num_procs() > 1:
                eng.comm_init_from_torch()                 # NCCL communicator (parameter broadcast, fallback exchange)
                import os
                if os.environ.get("SACB_COMM", "peer") == "peer":
                    eng.peer_init_from_torch()             # gradient exchange fused into the update kernel (NVLink)
            self._engines[key] = ent
        eng = ent["engine"]
        for name, val in (("alpha", self.alpha), ("gamma", self.gamma), ("polyak", self.polyak), ("reward_scale", self.reward_scale)):
            eng.set_scalar(name, val)
        if target is not None and ent["target"] is not target:
            eng.bind_params(actor=actor.arena, critic=critic.arena, target=target.arena)
            ent["target"] = target
        rebind = False
        if pi_opt is not None:
            b = self._binding(pi_opt, actor, True, eng)
            if ent["pi"] is not b:
                ent["pi"], rebind = b, True
        if q_opt is not None:
            b = self._binding(q_opt, critic, False, eng)
            if ent["q"] is not b:
                ent["q"], rebind = b, True
        if not rebind and (ent["pi"] is not None or ent["q"] is not None):
            # the same optimizer may have been stepped through another engine (another batch size): adopt its step count
            sa, sc = eng.adam_steps()
            rebind = (ent["pi"] is not None and ent["pi"].steps != sa) or (ent["q"] is not None and ent["q"].steps != sc)
        if rebind:
            pi, q = ent["pi"], ent["q"]
            eng.bind_adam(pi.m if pi else None, pi.v if pi else None, pi.steps if pi else 0,
                          q.m if q else None, q.v if q else None, q.steps if q else 0)
        for b, which in ((ent["pi"], "pi"), (ent["q"], "q")):
            if b is not None:
                lr, b1, b2, eps = _adam_hparams(b.opt)
                eng.set_scalar(f"lr_{which}", lr)
                eng.set_scalar("beta1", b1); eng.set_scalar("beta2", b2); eng.set_scalar("adam_eps", eps)
        return eng

    def _binding(self, opt, module, is_actor, eng: Engine):
        """One _AdamBinding per optimizer (shared across engines); rebuilt when the optimizer's state tensors were replaced
        (`load_state_dict` after the first update), so the loaded moments / step are adopted."""
        b = self._bindings.get(id(opt))
        if b is None or b.opt is not opt or not b.still_bound():
            b = self._bind_opt(opt, module, is_actor, eng)
            self._bindings[id(opt)] = b
        return b

    @staticmethod
    def _bind_opt(opt, module, is_actor, eng: Engine):
        _adam_hparams(opt)
        lay, dev = eng.layout, eng.tdev()
        if is_actor:
            return _AdamBinding(opt, module, ACTOR_PARAM_NAMES,
                                lambda a: actor_views(lay, a, module.obs_dim, module.act_dim, module.hidden_sizes),
                                lay.actor_floats, dev)
        return _AdamBinding(opt, module, CRITIC_PARAM_NAMES,
                            lambda a: critic_views(lay, a, module.obs_dim, module.act_dim, module.hidden_sizes),
                            lay.critic_floats, dev)

    def _entry(self, actor, critic, batch):
        return self._engines[(id(actor), id(critic), int(batch))]

    # ---- visual path (networks/convolutional.py modules, buffer/visual_replay_buffer.py) --------------------------------
    def _vengine(self, actor: VisualActor, critic: VisualDoubleCritic, batch: int, target=None, pi_opt=None, q_opt=None) -> VisualEngine:
        if not isinstance(critic, VisualDoubleCritic):
            raise TypeError("a VisualActor trains against a VisualDoubleCritic (main.py:69-87)")
        key = (id(actor), id(critic), int(batch))
        ent = self._engines.get(key)
        if ent is None:
            eng = VisualEngine(actor.obs_dim, actor.act_dim, actor.vis_dim, actor.hidden_sizes, batch_size=batch, device=actor._dev,
                               alpha=self.alpha, gamma=self.gamma, polyak=self.polyak, reward_scale=self.reward_scale, seed=self.seed,
                               **actor._engine_kwargs())
            ent = dict(engine=eng, target=None, pi=None, q=None)
            eng.bind_params(actor=actor.arena, critic=critic.arena, target=None)
            self._engines[key] = ent
        eng = ent["engine"]
        for name, val in (("alpha", self.alpha), ("gamma", self.gamma), ("polyak", self.polyak), ("reward_scale", self.reward_scale)):
            eng.set_scalar(name, val)
        if target is not None and ent["target"] is not target:
            eng.bind_params(actor=actor.arena, critic=critic.arena, target=target.arena)
            ent["target"] = target
        lay, dev = eng.layout, eng.tdev()
        rebind = False
        for which, opt, module in (("pi", pi_opt, actor), ("q", q_opt, critic)):
            if opt is None:
                continue
            b = self._bindings.get(id(opt))
            if b is None or b.opt is not opt or not b.still_bound():
                _adam_hparams(opt)
                if which == "pi":
                    b = _AdamBinding(opt, module, actor_param_names(),
                                     lambda a: vactor_views(lay, a, actor.obs_dim, actor.act_dim, actor.vis_dim[0], actor.hidden_sizes,
                                                            actor.filters, actor.kernel_sizes), lay.actor_floats, dev)
                else:
                    b = _AdamBinding(opt, module, critic_param_names(),
                                     lambda a: vcritic_views(lay, a, critic.obs_dim, critic.act_dim, critic.vis_dim[0], critic.hidden_sizes,
                                                             critic.filters, critic.kernel_sizes), lay.critic_floats, dev)
                self._bindings[id(opt)] = b
            if ent[which] is not b:
                ent[which], rebind = b, True
        if not rebind and (ent["pi"] is not None or ent["q"] is not None):
            sa, sc = eng.adam_steps()
            rebind = (ent["pi"] is not None and ent["pi"].steps != sa) or (ent["q"] is not None and ent["q"].steps != sc)
        if rebind:
            pi, q = ent["pi"], ent["q"]
            eng.bind_adam(pi.m if pi else None, pi.v if pi else None, pi.steps if pi else 0, q.m if q else None, q.v if q else None,
                          q.steps if q else 0)
        for b, which in ((ent["pi"], "pi"), (ent["q"], "q")):
            if b is not None:
                lr, b1, b2, eps = _adam_hparams(b.opt)
                eng.set_scalar(f"lr_{which}", lr)
                eng.set_scalar("beta1", b1); eng.set_scalar("beta2", b2); eng.set_scalar("adam_eps", eps)
        return eng

    @staticmethod
    def _load_visual(eng: VisualEngine, samples: VisualBatch):
        eng.load_batch(samples.states.features, samples.states.frame, samples.actions, samples.rewards, samples.next_states.features,
                       samples.next_states.frame, samples.done)

    # ------------------------------------------------------------------------------------------------------------
    def update_critic(self, q_opt: optim.Optimizer, actor: nn.Module, critic: nn.Module, target_critic: nn.Module,
                      samples: Batch, eps=None) -> torch.Tensor:
        """One critic step on `samples` (sac/algorithm.py:115-141): soft Bellman target, twin MSE, hand-derived
        backward, (multi-GPU: NCCL gradient average,) Adam.  Returns loss_q as a 0-dim CUDA tensor."""
        if isinstance(actor, VisualActor):                              # VisualBatch (visual_replay_buffer.py:12-18)
            B = samples.actions.shape[0]
            eng = self._vengine(actor, critic, B, target=target_critic, q_opt=q_opt)
            self._load_visual(eng, samples)
            loss_q = eng.update_critic(eps)
            self._entry(actor, critic, B)["q"].advance(1)
            return loss_q
        B = samples.states.shape[0]
        eng = self._engine(actor, critic, B, target=target_critic, q_opt=q_opt)
        eng.load_batch(samples.states, samples.actions, samples.rewards, samples.next_states, samples.done)
        loss_q = eng.update_critic(eps)
        self._entry(actor, critic, B)["q"].advance(1)
        return loss_q

    def update_policy(self, pi_opt: optim.Optimizer, actor: nn.Module, critic: nn.Module, samples: Batch, eps=None) -> torch.Tensor:
        """One actor step on `samples` (sac/algorithm.py:143-162); the critic is only read."""
        if isinstance(actor, VisualActor):
            B = samples.actions.shape[0]
            eng = self._vengine(actor, critic, B, pi_opt=pi_opt)
            self._load_visual(eng, samples)
            loss_pi = eng.update_policy(eps)
            self._entry(actor, critic, B)["pi"].advance(1)
            return loss_pi
        B = samples.states.shape[0]
        eng = self._engine(actor, critic, B, pi_opt=pi_opt)
        eng.load_batch(samples.states, samples.actions, samples.rewards, samples.next_states, samples.done)
        loss_pi = eng.update_policy(eps)
        self._entry(actor, critic, B)["pi"].advance(1)
        return loss_pi

    def update_burst(self, n_updates, actor, critic, target_critic, buffer: ReplayBuffer, pi_opt, q_opt, idx=None, eps=None):
        """`n_updates` x (sample, update_critic, update_policy, update_targets) = the inner loop
        sac/algorithm.py:274-278, as one persistent-kernel launch.  Returns (loss_q[n], loss_pi[n]) device tensors."""
        if isinstance(actor, VisualActor):
            if not isinstance(buffer, VisualReplayBuffer):
                raise TypeError("a visual update_burst needs the HBM VisualReplayBuffer of this package")
            eng = self._vengine(actor, critic, self.batch_size, target=target_critic, pi_opt=pi_opt, q_opt=q_opt)
            buffer.flush()
            ent = self._entry(actor, critic, self.batch_size)
            if ent.get("buffer") is not buffer:
                eng.bind_buffer(buffer._features, buffer._frames, buffer._actions, buffer._rewards, buffer._next_features, buffer._next_frames,
                                buffer._done)
                ent["buffer"] = buffer
            if buffer.size < self.batch_size:
                raise ValueError("Sample larger than population or is negative")
            eng.set_buffer_state(buffer.ptr, buffer.size)
            lq, lp = eng.update(n_updates, idx=idx, eps=eps)
            ent["pi"].advance(n_updates)
            ent["q"].advance(n_updates)
            return lq, lp
        if not isinstance(buffer, ReplayBuffer):
            raise TypeError("update_burst needs the HBM ReplayBuffer of this package")
        eng = self._engine(actor, critic, self.batch_size, target=target_critic, pi_opt=pi_opt, q_opt=q_opt)
        buffer.flush()
        ent = self._entry(actor, critic, self.batch_size)
        if ent.get("buffer") is not buffer:
            eng.bind_buffer_packed(buffer._rec)
            ent["buffer"] = buffer
        if buffer.size < self.batch_size:
            raise ValueError("Sample larger than population or is negative")
        eng.set_buffer_state(buffer.ptr, buffer.size)
        lq, lp = eng.update(n_updates, idx=idx, eps=eps)
        ent["pi"].advance(n_updates)
        ent["q"].advance(n_updates)
        return lq, lp

    # ------------------------------------------------------------------------------------------------------------
    def save_model(self, actor, critic, pi_opt, q_opt, epoch: int):
        """Checkpoint (sac/algorithm.py:164-180).  With mlflow installed the reference's artifacts are written;
        otherwise a plain torch file with the same content."""
        aux = {"pi_opt": pi_opt.state_dict(), "q_opt": q_opt.state_dict(), "epoch": epoch}
        try:
            import mlflow
            # the reference's artifact layout (sac/algorithm.py:172-180): main.py:load_session and run_agent.py read
            # artifacts/actor and artifacts/critic back with mlflow.pytorch.load_model, artifacts/auxiliaries with
            # load_state_dict.  The modules pickle as plain state (arena + dims) and re-create their engine on load.
            mlflow.pytorch.log_model(actor, "actor")
            mlflow.pytorch.log_model(critic, "critic")
            mlflow.pytorch.log_state_dict(aux, artifact_path="auxiliaries")
        except ImportError:
            import os
            d = os.path.join("sacb200_checkpoints", f"epoch_{epoch}")
            os.makedirs(d, exist_ok=True)
            torch.save(actor, os.path.join(d, "actor.pt"))            # same three artifacts, torch.save instead of mlflow
            torch.save(critic, os.path.join(d, "critic.pt"))
            torch.save(aux, os.path.join(d, "auxiliaries.pt"))

    def train(self, start_epoch: int, env, actor: nn.Module, critic: nn.Module, buffer: ReplayBuffer,
              pi_opt: optim.Optimizer, q_opt: optim.Optimizer, render: bool = True, logging: bool = True):
        """Environment loop + update schedule of the reference (sac/algorithm.py:182-307).  The simulator runs on the
        host; transitions are staged into the HBM ring; every `update_every` steps one burst runs on the GPU."""
        target_critic = deepcopy(critic)
        for p in target_critic.parameters():
            p.requires_grad = False

        sync_params(actor)
        sync_params(critic)
        sync_params(target_critic)

        seed = 10000 * proc_id()
        self.seed = seed
        torch.manual_seed(seed)
        np.random.seed(seed)

        state = env.reset()
        step, ep_ret, ep_len = 0, 0, 0
        try:
            import tqdm
            pbar = tqdm.trange(start_epoch, start_epoch + self.epochs, ncols=0, initial=start_epoch)
        except ImportError:
            pbar = range(start_epoch, start_epoch + self.epochs)
        metrics = {}
        for e in pbar:
            episode_rewards, episode_lengths = [], []
            burst_losses = []                               # device tensors; read back once per epoch

            for t in range(self.steps_per_epoch):
                if step < self.start_steps:
                    action = env.action_space.sample()
                else:
                    action, _ = actor(state)
                    action = action.cpu().detach().numpy()

                next_state, reward, done, _ = env.step(action)
                done = False if ep_len == self.max_ep_len else done     # as the reference, sac/algorithm.py:241
                if render and proc_id() == 0:
                    env.render()
                ep_len += 1
                ep_ret += reward
                buffer.store(state, action, reward, next_state, done)
                state = next_state

                if done or t == self.steps_per_epoch - 1:
                    episode_rewards.append(ep_ret)
                    episode_lengths.append(ep_len)
                    state = env.reset()
                    ep_ret, ep_len = 0, 0

                if step > self.update_after and step % self.update_every == 0:
                    burst_losses.append(self.update_burst(self.update_every, actor, critic, target_critic, buffer, pi_opt, q_opt))
                step += 1

            if num_procs() > 1:                              # one gather per epoch instead of a blocking recv per step
                import torch.distributed as dist
                gathered = [None] * num_procs()
                dist.all_gather_object(gathered, (episode_rewards, episode_lengths))
                episode_rewards = [r for g in gathered for r in g[0]]
                episode_lengths = [l for g in gathered for l in g[1]]
            if burst_losses:
                lq = torch.cat([b[0] for b in burst_losses]).mean().item()
                lp = torch.cat([b[1] for b in burst_losses]).mean().item()
            else:
                lq = lp = float("nan")
            metrics = {"episode_length": float(np.mean(episode_lengths)), "reward": float(np.mean(episode_rewards)),
                       "loss_q": lq, "loss_pi": lp}
            if proc_id() == 0 and logging:
                if e % self.save_every == 0:
                    self.save_model(actor, critic, pi_opt, q_opt, e)
                try:
                    import mlflow
                    mlflow.log_metrics(metrics, e)
                except ImportError:
                    logger.info("epoch %d: %s", e, metrics)
            metrics["step"] = step
            if hasattr(pbar, "set_postfix"):
                pbar.set_postfix(metrics)
            state = env.reset()
            ep_ret, ep_len = 0, 0
        self.target_critic = target_critic
        return metrics
