"""Pin both oracle flavours against outputs of the unmodified reference (tests/golden/*.npz).

The reference's own tests hold no numerical vectors for this path, so the golden files made by
oracle/make_golden.py ARE the pin.  Tolerances: float32 reassociation only (the oracle does the same fp32
arithmetic in a different summation order); the float64 oracle must agree to the fp32 rounding of the reference.
"""
import numpy as np
import pytest

from oracle import sac_numpy as onp
from oracle.sac_numpy import ACTOR_KEYS, CRITIC_KEYS, AdamState, cast_params
from golden_helpers import assert_close, batch_from, load_case, summary

CASES = ["small", "ragged", "c1"]


def _run_numpy(name, dtype):
    g, actor, critic, arrays, hp = load_case(name)
    actor, critic = cast_params(actor, dtype), cast_params(critic, dtype)
    target = cast_params(critic, dtype)
    opt_pi, opt_q = AdamState.zeros_like(actor), AdamState.zeros_like(critic)
    outs = []
    for k in range(int(g["n_updates"])):
        batch = batch_from(arrays, g["idx"][k], dtype)
        eps = g["eps"][k].astype(dtype)
        o = onp.sac_update(actor, critic, target, opt_pi, opt_q, batch, eps[0], eps[1], hp)
        if k == 0:
            o["snap"] = dict(actor=cast_params(actor, dtype), critic=cast_params(critic, dtype),
                             target=cast_params(target, dtype))
        outs.append(o)
    return g, outs, actor, critic, target, opt_pi, opt_q


def _check_tensor(g, prefix, name, value, rtol, atol):
    if f"{prefix}.{name}" in g:
        assert_close(value, g[f"{prefix}.{name}"], rtol, atol, f"{prefix}.{name}")
    else:
        s, a, sub = summary(value)
        assert_close(sub, g[f"{prefix}.{name}.sub"], rtol, atol, f"{prefix}.{name}.sub")
        assert abs(s - float(g[f"{prefix}.{name}.sum"])) <= atol * value.size + rtol * float(g[f"{prefix}.{name}.abssum"])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_numpy_oracle_matches_reference(name, dtype):
    g, outs, actor, critic, target, opt_pi, opt_q = _run_numpy(name, dtype)
    assert_close([o["loss_q"] for o in outs], g["loss_q"], 2e-5, 1e-6, "loss_q")
    assert_close([o["loss_pi"] for o in outs], g["loss_pi"], 2e-5, 1e-6, "loss_pi")
    o = outs[0]
    for key in ("a2", "logp_next", "q1_target", "q2_target", "q1", "q2", "pi", "logp_pi", "q1_pi", "q2_pi", "backup"):
        assert_close(o[key], g[f"step0.{key}"], 2e-5, 2e-5, key)
    for k in CRITIC_KEYS:
        _check_tensor(g, "step0.critic_grads", k, o["critic_grads"][k], 1e-4, 2e-6)
    for k in ACTOR_KEYS:
        _check_tensor(g, "step0.actor_grads", k, o["actor_grads"][k], 1e-4, 2e-6)
    # after ONE Adam step |dp| = lr for every element whose grad sign agrees; compare params tightly
    for k in ACTOR_KEYS:
        _check_tensor(g, "step0.actor", k, o["snap"]["actor"][k], 1e-6, 1e-6)
    for k in CRITIC_KEYS:
        _check_tensor(g, "step0.critic", k, o["snap"]["critic"][k], 1e-6, 1e-6)
        _check_tensor(g, "step0.target", k, o["snap"]["target"][k], 1e-6, 1e-7)
    # short trajectory: Adam's g/sqrt(v) amplifies rounding of near-zero grads, so the
    # horizon tolerance is a few lr
    n = int(g["n_updates"])
    for k in ACTOR_KEYS:
        _check_tensor(g, "final.actor", k, actor[k], 1e-5, 3e-4 * n)
    for k in CRITIC_KEYS:
        _check_tensor(g, "final.critic", k, critic[k], 1e-5, 3e-4 * n)
        _check_tensor(g, "final.target", k, target[k], 1e-5, 3e-4 * n * 0.005 * n + 1e-6)


def test_adam_state_matches_reference():
    g, outs, actor, critic, target, opt_pi, opt_q = _run_numpy("small", np.float32)
    for k in ACTOR_KEYS:
        assert_close(opt_pi.m[k], g[f"final.pi_opt.exp_avg.{k}"], 1e-3, 1e-6, k)
        assert_close(opt_pi.v[k], g[f"final.pi_opt.exp_avg_sq.{k}"], 1e-3, 1e-9, k)
        assert float(g[f"final.pi_opt.step.{k}"]) == opt_pi.step
    for k in CRITIC_KEYS:
        assert_close(opt_q.m[k], g[f"final.q_opt.exp_avg.{k}"], 1e-3, 1e-6, k)
        assert float(g[f"final.q_opt.step.{k}"]) == opt_q.step


def test_forward_contracts():
    """deterministic / unbatched forward of the reference Actor and DoubleCritic (tests/test_linear.py:7-41)."""
    g, actor, critic, arrays, hp = load_case("small")
    # the fwd.* vectors were taken AFTER the 4 updates: rebuild those params from the golden file
    actor = {k: g[f"final.actor.{k}"] for k in ACTOR_KEYS}
    critic = {k: g[f"final.critic.{k}"] for k in CRITIC_KEYS}
    x = g["fwd.x"]
    pi, logp, _ = onp.actor_forward(actor, x, None, hp, deterministic=True)
    assert_close(pi, g["fwd.pi_det"], 1e-5, 1e-6)
    assert_close(logp, g["fwd.logp_det"], 1e-5, 1e-5)
    pi0, logp0, _ = onp.actor_forward(actor, x[0], None, hp, deterministic=True)
    assert pi0.shape == (int(g["act_dim"]),) and np.shape(logp0) == ()
    assert_close(pi0, g["fwd.pi_unbatched"], 1e-5, 1e-6)
    q1, q2, _, _ = onp.double_critic_forward(critic, x, g["fwd.act"])
    assert_close(q1, g["fwd.q1"], 1e-5, 1e-6)
    assert_close(q2, g["fwd.q2"], 1e-5, 1e-6)


@pytest.mark.parametrize("name", ["small", "ragged"])
def test_torch_cpu_port_matches_reference(name):
    import torch
    from oracle.sac_torch_cpu import TorchCpuSAC
    torch.set_num_threads(1)
    g, actor, critic, arrays, hp = load_case(name)
    sac = TorchCpuSAC(actor, critic, hp)
    for k in range(int(g["n_updates"])):
        b = {kk: torch.tensor(v) for kk, v in batch_from(arrays, g["idx"][k]).items()}
        o = sac.update(b, torch.tensor(g["eps"][k, 0]), torch.tensor(g["eps"][k, 1]))
        assert_close(float(o["loss_q"]), g["loss_q"][k], 1e-6, 1e-7, "loss_q")
        assert_close(float(o["loss_pi"]), g["loss_pi"][k], 1e-6, 1e-7, "loss_pi")
        if k == 0:
            for n_ in CRITIC_KEYS:
                assert_close(o["critic_grads"][n_].numpy(), g[f"step0.critic_grads.{n_}"], 1e-5, 1e-7, n_)
            for n_ in ACTOR_KEYS:
                assert_close(o["actor_grads"][n_].numpy(), g[f"step0.actor_grads.{n_}"], 1e-5, 1e-7, n_)
    for n_ in ACTOR_KEYS:
        assert_close(sac.actor[n_].detach().numpy(), g[f"final.actor.{n_}"], 1e-6, 1e-6, n_)
    for n_ in CRITIC_KEYS:
        assert_close(sac.critic[n_].detach().numpy(), g[f"final.critic.{n_}"], 1e-6, 1e-6, n_)
        assert_close(sac.target[n_].detach().numpy(), g[f"final.target.{n_}"], 1e-6, 1e-7, n_)


def test_replay_buffer_oracle_matches_reference():
    g = dict(np.load("tests/golden/replay_buffer.npz")) if False else None
    import os
    from golden_helpers import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "replay_buffer.npz")))
    buf = onp.ReplayBufferOracle(int(g["n"]), int(g["obs_dim"]), int(g["act_dim"]))
    for i in range(int(g["n_store"])):
        buf.store(g["in.state"][i], g["in.actions"][i], g["in.rewards"][i], g["in.next_state"][i], g["in.done"][i])
        assert buf.size == g["sizes"][i] and buf.ptr == g["ptrs"][i]
    for k in ("state", "actions", "rewards", "next_state", "done"):
        assert np.array_equal(getattr(buf, k), g[f"final.{k}"])
    b = buf.gather(g["idx"])
    for k in ("states", "actions", "rewards", "next_states", "done"):
        assert np.array_equal(b[k], g[f"batch.{k}"]) and b[k].dtype == np.float32
    assert str(g["empty_sample_error"]) == "ValueError"
    with pytest.raises(ValueError):
        onp.ReplayBufferOracle(4, 3, 2).gather([0])


# ---------------------------------------------------------------------------------------------------------------
# bf16 operand-rounding mode of the oracle (the checker of the engine's precision-3 path)
# ---------------------------------------------------------------------------------------------------------------
def test_round_bf16_is_torch_bfloat16():
    import torch
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200_000).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, 30, 200_000).astype(np.float32),
                        np.array([0.0, -0.0, np.inf, -np.inf, 1 + 2 ** -8, 1 + 3 * 2 ** -8, 1 + 2 ** -9, 3.3895314e38, 1e-40, -1e-45], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    got = onp.round_bf16(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))            # bit for bit, ties to even, overflow to inf
    assert np.isnan(onp.round_bf16(np.array([np.nan], np.float32))[0])
    with onp.operand_rounding("bf16"):                                          # products of rounded operands, fp accumulation
        a, b = rng.standard_normal((5, 7)).astype(np.float32), rng.standard_normal((7, 3)).astype(np.float32)
        assert np.array_equal(onp._mm(a, b), onp.round_bf16(a) @ onp.round_bf16(b))
    assert np.array_equal(onp._mm(a, b), a @ b)                                 # and the mode is off again outside


def test_bf16_oracle_matches_torch_autograd_with_rounded_operands(monkeypatch):
    """Independent restatement of the bf16 arithmetic: torch autograd with a custom Linear whose forward rounds both
    operands to bfloat16 and whose backward rounds the incoming gradient and uses the rounded saved operands -- i.e. every
    product of the hidden / head layers has bf16 operands, everything else (relu, tanh-Gaussian, losses, the scalar output
    layer of the critics) is plain autograd.  The hand-derived numpy backward in operand_rounding("bf16") mode must agree."""
    import torch
    import torch.nn.functional as F
    import oracle.sac_torch_cpu as otc

    class Bf16Linear(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b):
            xr, wr = x.bfloat16().to(x.dtype), w.bfloat16().to(w.dtype)
            ctx.save_for_backward(xr, wr)
            return xr @ wr.T + b

        @staticmethod
        def backward(ctx, g):
            xr, wr = ctx.saved_tensors
            gr = g.bfloat16().to(g.dtype)
            return gr @ wr, gr.T @ xr, g.sum(0)

    class FShim:
        relu, softplus = staticmethod(F.relu), staticmethod(F.softplus)

        @staticmethod
        def linear(x, w, b):
            return F.linear(x, w, b) if w.shape[0] == 1 else Bf16Linear.apply(x, w, b)

    monkeypatch.setattr(otc, "F", FShim)
    torch.set_num_threads(1)
    O, A, H, B, N = 11, 3, [72, 40], 96, 400
    rng = np.random.default_rng(7)
    actor, critic = onp.init_params(O, A, H, rng)
    arrays = dict(state=rng.standard_normal((N, O)).astype(np.float32), actions=rng.uniform(-1, 1, (N, A)).astype(np.float32),
                  rewards=rng.standard_normal(N).astype(np.float32), next_state=rng.standard_normal((N, O)).astype(np.float32),
                  done=rng.random(N) < 0.05)
    idx = rng.choice(N, B, replace=False)
    eps = rng.standard_normal((2, B, A)).astype(np.float32)
    hp = onp.HyperParams(act_limit=1.0)
    sac = otc.TorchCpuSAC(actor, critic, hp)
    ot = sac.update({k: torch.tensor(v) for k, v in batch_from(arrays, idx).items()}, torch.tensor(eps[0]), torch.tensor(eps[1]))
    a32, c32, t32 = ({k: v.copy() for k, v in p.items()} for p in (actor, critic, critic))
    with onp.operand_rounding("bf16"):
        on = onp.sac_update(a32, c32, t32, onp.AdamState.zeros_like(a32), onp.AdamState.zeros_like(c32),
                            batch_from(arrays, idx), eps[0], eps[1], hp)
    on32 = onp.sac_update(*({k: v.copy() for k, v in p.items()} for p in (actor, critic, critic)), onp.AdamState.zeros_like(actor),
                          onp.AdamState.zeros_like(critic), batch_from(arrays, idx), eps[0], eps[1], hp)
    assert abs(float(on32["loss_q"]) - float(on["loss_q"])) > 1e-5                        # the mode changes the numbers ...
    assert_close(float(ot["loss_q"]), float(on["loss_q"]), 2e-6, 1e-6, "loss_q")           # ... and both restatements agree on them
    assert_close(float(ot["loss_pi"]), float(on["loss_pi"]), 2e-5, 2e-6, "loss_pi")        # (evaluated after the critic's Adam step)

    def close(got, want, name):                           # fp32 summation order can flip a bf16 rounding downstream: allow a few
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        err = np.abs(got - want)
        bad = err > 1e-6 + 1e-4 * np.abs(want)
        assert bad.mean() <= 0.02 and err.max() <= 1e-2 * np.abs(want).max() + 1e-7, f"{name}: {bad.sum()} / {bad.size} off, max {err.max():.2e}"
    for n_ in CRITIC_KEYS:
        close(ot["critic_grads"][n_].numpy(), on["critic_grads"][n_], n_)
    for n_ in ACTOR_KEYS:
        close(ot["actor_grads"][n_].numpy(), on["actor_grads"][n_], n_)
