"""CPU: the visual oracle (oracle/visual_torch.py) against golden vectors produced by the UNMODIFIED reference's visual
path -- VisualReplayBuffer.sample, VisualActor / VisualDoubleCritic, SAC.update_critic / update_policy / update_targets
(`buffer/visual_replay_buffer.py`, `networks/convolutional.py`, `sac/algorithm.py`)."""
import numpy as np
import pytest
import torch

from golden_helpers import assert_close
from oracle import visual_torch as vt
from visual_helpers import check_summary, load_visual_case


@pytest.mark.parametrize("name", ["visual_small", "visual_raw"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_visual_oracle_matches_reference_golden(name, dtype):
    g, actor_np, critic_np, arrays, hp = load_visual_case(name)
    tight = dtype == torch.float64
    actor, critic = vt.to_torch(actor_np, dtype), vt.to_torch(critic_np, dtype)
    target = {k: v.clone() for k, v in critic.items()}
    opt_pi, opt_q = vt.AdamState(), vt.AdamState()
    K = int(g["n_updates"])
    for k in range(K):
        b = vt.torch_batch(arrays, g["idx"][k], dtype)
        o = vt.visual_sac_update(actor, critic, target, opt_pi, opt_q, b, torch.tensor(g["eps"][k, 0], dtype=dtype),
                                 torch.tensor(g["eps"][k, 1], dtype=dtype), hp)
        assert_close([o["loss_q"]], g["loss_q"][k:k + 1], 2e-4, 1e-5, f"loss_q[{k}]")
        assert_close([o["loss_pi"]], g["loss_pi"][k:k + 1], 2e-4, 1e-5, f"loss_pi[{k}]")
        if k == 0:
            for key in ("a2", "logp_next", "q1_target", "q2_target", "q1", "q2", "pi", "logp_pi", "q1_pi", "q2_pi"):
                assert_close(o[key].numpy(), g[f"step0.{key}"], 2e-4, 5e-5, key)
            for n in vt.critic_keys():
                check_summary(g, "step0.critic_grads", n, o["critic_grads"][n].numpy(), 2e-3 if not tight else 1e-3, 2e-6, assert_close)
            for n in vt.actor_keys():
                check_summary(g, "step0.actor_grads", n, o["actor_grads"][n].numpy(), 2e-3 if not tight else 1e-3, 2e-6, assert_close)
            for n in vt.critic_keys():
                # one Adam step moves a parameter by lr * g / (|g| + eps'): entries whose gradient is ~1e-8 (first-conv weights
                # behind dead ReLUs) are decided by rounding noise, so the bound is a fraction of lr = 3e-4, not an ulp
                check_summary(g, "step0.critic", n, critic[n].numpy(), 1e-6, 2e-5, assert_close, frac=0.02, atol_out=6.1e-4)
                check_summary(g, "step0.target", n, target[n].numpy(), 1e-6, 1e-6, assert_close, frac=0.02, atol_out=4e-6)
            for n in vt.actor_keys():
                check_summary(g, "step0.actor", n, actor[n].numpy(), 1e-6, 2e-5, assert_close, frac=0.02, atol_out=6.1e-4)
    for n in vt.critic_keys():
        check_summary(g, "final.critic", n, critic[n].numpy(), 1e-5, 3e-4 * K, assert_close)
    for n in vt.actor_keys():
        check_summary(g, "final.actor", n, actor[n].numpy(), 1e-5, 3e-4 * K, assert_close)


def test_calculate_size_and_unbatched_contract():
    """tests/test_convolutional.py of the reference: calculate_size and a single (unbatched) MultiObservation."""
    g, actor_np, critic_np, arrays, hp = load_visual_case("visual_small")
    vis = tuple(int(v) for v in g["vis_dim"])
    sizes, flat = vt.conv_sizes(vis)
    assert flat == int(g["calculate_size"])
    assert vt.conv_sizes((3, 84, 84))[0] == [(32, 20, 20), (64, 9, 9), (64, 7, 7)] and vt.conv_sizes((3, 84, 84))[1] == 3136
    assert vt.conv_sizes((3, 64, 64))[1] == 1024                                  # the reference default, main.py:68
    a = vt.to_torch(actor_np, torch.float32)
    f = torch.tensor(arrays["features"][:1]); fr = torch.tensor(arrays["frames"][:1].astype(np.float32))
    pi, logp = vt.actor_forward(a, f, fr, hp, deterministic=True)
    assert_close(pi[0].numpy(), g["fwd.pi_unbatched"], 2e-4, 2e-5, "pi unbatched")
    assert_close(logp[0].numpy(), g["fwd.logp_unbatched"], 2e-4, 5e-5, "logp unbatched")
