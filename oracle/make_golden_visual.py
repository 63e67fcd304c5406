"""Generate tests/golden/visual_*.npz by executing the UNMODIFIED reference's visual path (/root/reference) on the CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_visual.py

Executes `buffer/visual_replay_buffer.py` (VisualReplayBuffer.store / sample, VisualBatch), `networks/convolutional.py`
(VisualActor, VisualDoubleCritic) and `sac/algorithm.py` (SAC.update_critic, SAC.update_policy, update_targets) exactly as
`main.py:63-87` wires them, under the import stubs of oracle/ref_stubs.py.  Determinism recipe as in oracle/make_golden.py:
parameters from a numpy RNG via load_state_dict (oracle.visual_torch.init_visual_params), the batch drawn by the
reference's own `sample` under `random.seed`, the rsample noise reproduced with `torch.manual_seed` before each update call.
"""
import os
import random
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle.ref_stubs import import_reference_visual
from oracle.visual_torch import actor_keys, critic_keys, init_visual_params, make_visual_arrays

REF = "/root/reference"


def sd_to_np(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def summarize(out, prefix, d, full):
    for n, v in d.items():
        v = np.asarray(v)
        if full or v.size <= 4096:
            out[f"{prefix}.{n}"] = v
        else:                                   # summaries: sum, abs-sum, strided subsample
            flat = v.reshape(-1).astype(np.float64)
            out[f"{prefix}.{n}.sum"] = flat.sum()
            out[f"{prefix}.{n}.abssum"] = np.abs(flat).sum()
            out[f"{prefix}.{n}.sub"] = v.reshape(-1)[:: max(1, v.size // 512)][:512].copy()


def run_case(name, obs_dim, act_dim, vis_dim, hidden, batch, n_buf, n_updates, act_limit, seed, frame_scale, full):
    (alg, VisualBatch, VisualReplayBuffer, MultiObservation, VisualActor, VisualCritic, VisualDoubleCritic,
     calculate_size, simple_cnn) = import_reference_visual(REF)
    from copy import deepcopy
    rng = np.random.default_rng(seed)
    actor_np, critic_np = init_visual_params(obs_dim, act_dim, vis_dim, hidden, rng, frame_scale=frame_scale)
    arrays = make_visual_arrays(rng, n_buf, obs_dim, act_dim, vis_dim, act_limit)
    torch.set_num_threads(4)
    actor = VisualActor(obs_dim, act_dim, vis_dim, list(hidden), act_limit)          # main.py:69-78
    critic = VisualDoubleCritic(obs_dim, act_dim, vis_dim, list(hidden))             # main.py:79-87
    assert list(dict(actor.named_parameters())) == actor_keys() and list(dict(critic.named_parameters())) == critic_keys()
    actor.load_state_dict({k: torch.tensor(v) for k, v in actor_np.items()})
    critic.load_state_dict({k: torch.tensor(v) for k, v in critic_np.items()})
    target = deepcopy(critic)                                                        # sac/algorithm.py:194-196
    for p in target.parameters():
        p.requires_grad = False
    pi_opt = torch.optim.Adam(actor.parameters(), lr=3e-4)                           # main.py:93-95
    q_opt = torch.optim.Adam(critic.parameters(), lr=3e-4)
    sac = alg.SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=batch, start_steps=0,
                  steps_per_epoch=1, max_ep_len=1, update_after=0, update_every=1, save_every=1)

    # VisualReplayBuffer.store, transition by transition, with MultiObservation(FloatTensor, FloatTensor) as
    # environments/wall_runner.py:54-58 builds them (frames are raw 0..255 floats)
    buf = VisualReplayBuffer(n_buf, act_dim)
    for i in range(n_buf):
        obs = MultiObservation(torch.FloatTensor(arrays["features"][i]), torch.FloatTensor(arrays["frames"][i].astype(np.float32)))
        nobs = MultiObservation(torch.FloatTensor(arrays["next_features"][i]), torch.FloatTensor(arrays["next_frames"][i].astype(np.float32)))
        buf.store(obs, arrays["actions"][i], arrays["rewards"][i], nobs, arrays["done"][i])
    assert buf.size == n_buf and buf.ptr == 0

    out = dict(obs_dim=obs_dim, act_dim=act_dim, vis_dim=np.array(vis_dim), hidden=np.array(hidden), batch=batch, n_buf=n_buf,
               n_updates=n_updates, act_limit=act_limit, seed=seed, frame_scale=frame_scale)
    # the reference's own shape contracts (tests/test_convolutional.py): calculate_size and a single unbatched observation,
    # with the INITIAL parameters
    out["calculate_size"] = int(calculate_size(vis_dim, [32, 64, 64], [8, 4, 3], [4, 2, 1]))
    with torch.no_grad():
        one = MultiObservation(torch.FloatTensor(arrays["features"][0]), torch.FloatTensor(arrays["frames"][0].astype(np.float32)))
        p1, l1 = actor(one, deterministic=True)
        out["fwd.pi_unbatched"], out["fwd.logp_unbatched"] = p1.numpy(), np.asarray(l1.numpy())
        qa, qb = critic(one, torch.FloatTensor(arrays["actions"][0]))
        out["fwd.q1_unbatched"], out["fwd.q2_unbatched"] = np.asarray(qa.numpy()), np.asarray(qb.numpy())
    captured = {}

    def hook(tag):
        def fn(_m, _inp, outp):
            captured.setdefault(tag, []).append(tuple(o.detach().numpy().copy() for o in outp))
        return fn

    actor.register_forward_hook(hook("actor"))
    critic.register_forward_hook(hook("critic"))
    target.register_forward_hook(hook("target"))

    idx_all = np.zeros((n_updates, batch), np.int64)
    eps_all = np.zeros((n_updates, 2, batch, act_dim), np.float32)
    steps = []
    for k in range(n_updates):
        random.seed(500 + k)
        idx = random.sample(range(buf.size), batch)                                  # visual_replay_buffer.py:50
        idx_all[k] = idx
        random.seed(500 + k)
        samples = buf.sample(batch)                                                  # the reference's own sampling + stacking
        assert isinstance(samples, VisualBatch)
        if k == 0:
            out["batch0.features"] = samples.states.features.numpy().copy()
            out["batch0.frames_sum"] = samples.states.frame.numpy().astype(np.float64).sum(axis=(1, 2, 3))
            out["batch0.next_frames_sum"] = samples.next_states.frame.numpy().astype(np.float64).sum(axis=(1, 2, 3))
            out["batch0.actions"], out["batch0.rewards"] = samples.actions.numpy().copy(), samples.rewards.numpy().copy()
            out["batch0.done"] = samples.done.numpy().copy()
        captured.clear()
        sq, sp = 2000 + 2 * k, 2001 + 2 * k
        torch.manual_seed(sq)
        eps_all[k, 0] = torch.empty(batch, act_dim).normal_().numpy()
        torch.manual_seed(sp)
        eps_all[k, 1] = torch.empty(batch, act_dim).normal_().numpy()
        torch.manual_seed(sq)
        loss_q = sac.update_critic(q_opt, actor, critic, target, samples)
        cg = {n: p.grad.numpy().copy() for n, p in critic.named_parameters()}
        a2, logp_next = captured["actor"][0]
        q1t, q2t = captured["target"][0]
        q1, q2 = captured["critic"][0]
        torch.manual_seed(sp)
        loss_pi = sac.update_policy(pi_opt, actor, critic, samples)
        ag = {n: p.grad.numpy().copy() for n, p in actor.named_parameters()}
        pi, logp_pi = captured["actor"][1]
        q1pi, q2pi = captured["critic"][1]
        alg.update_targets(critic, target, sac.polyak)
        steps.append(dict(loss_q=float(loss_q), loss_pi=float(loss_pi), a2=a2, logp_next=logp_next, q1_target=q1t, q2_target=q2t,
                          q1=q1, q2=q2, pi=pi, logp_pi=logp_pi, q1_pi=q1pi, q2_pi=q2pi, critic_grads=cg, actor_grads=ag,
                          actor=sd_to_np(actor), critic=sd_to_np(critic), target=sd_to_np(target)))
    out["idx"], out["eps"] = idx_all, eps_all
    out["loss_q"] = np.array([s["loss_q"] for s in steps], np.float32)
    out["loss_pi"] = np.array([s["loss_pi"] for s in steps], np.float32)
    for key in ("a2", "logp_next", "q1_target", "q2_target", "q1", "q2", "pi", "logp_pi", "q1_pi", "q2_pi"):
        out[f"step0.{key}"] = steps[0][key]
    summarize(out, "step0.critic_grads", steps[0]["critic_grads"], full)
    summarize(out, "step0.actor_grads", steps[0]["actor_grads"], full)
    for tag in ("actor", "critic", "target"):
        summarize(out, f"step0.{tag}", steps[0][tag], full)
        summarize(out, f"final.{tag}", steps[-1][tag], full)
    if full:
        for n, v in actor_np.items():
            out[f"init.actor.{n}"] = v
        for n, v in critic_np.items():
            out[f"init.critic.{n}"] = v
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.1f} KiB, loss_q={out['loss_q']}, loss_pi={out['loss_pi']}")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    which = sys.argv[1:] or ["visual_small", "visual_raw", "visual_c5"]
    if "visual_small" in which:
        run_case("visual_small", obs_dim=7, act_dim=3, vis_dim=(3, 36, 36), hidden=(32, 32), batch=8, n_buf=24, n_updates=3,
                 act_limit=1.0, seed=21, frame_scale=64.0, full=False)
    if "visual_raw" in which:          # default-scale first conv on raw 0..255 pixels, as the reference runs it
        run_case("visual_raw", obs_dim=5, act_dim=2, vis_dim=(3, 44, 36), hidden=(16, 24), batch=5, n_buf=12, n_updates=2,
                 act_limit=2.0, seed=22, frame_scale=1.0, full=False)
    if "visual_c5" in which:           # BASELINE configs[4] shapes: 3x84x84 frames, features 168, actions 56, batch 256
        run_case("visual_c5", obs_dim=168, act_dim=56, vis_dim=(3, 84, 84), hidden=(256, 256), batch=256, n_buf=320, n_updates=1,
                 act_limit=1.0, seed=23, frame_scale=64.0, full=False)
