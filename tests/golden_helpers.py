"""Shared loaders for the committed golden fixtures (tests/golden/, made by oracle/make_golden.py)."""
import os

import numpy as np

from oracle.sac_numpy import ACTOR_KEYS, CRITIC_KEYS, HyperParams, init_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    """Returns (golden npz dict, actor_np, critic_np, buffer arrays, hp).  For summary-only cases ("c1") the
    inputs are regenerated from the recorded seed exactly as oracle/make_golden.py drew them."""
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    O, A, B = int(g["obs_dim"]), int(g["act_dim"]), int(g["batch"])
    hidden = tuple(int(h) for h in g["hidden"])
    rng = np.random.default_rng(int(g["seed"]))
    actor, critic = init_params(O, A, hidden, rng)
    n = int(g["n_buf"])
    lim = float(g["act_limit"])
    arrays = dict(
        state=rng.standard_normal((n, O)).astype(np.float32),
        actions=rng.uniform(-lim, lim, (n, A)).astype(np.float32),
        rewards=rng.standard_normal(n).astype(np.float32),
        next_state=rng.standard_normal((n, O)).astype(np.float32),
        done=(rng.random(n) < 0.05),
    )
    idx = np.stack([rng.choice(n, size=B, replace=False) for _ in range(int(g["n_updates"]))]).astype(np.int64)
    assert np.array_equal(idx, g["idx"])
    if "init.actor.layers.0.weight" in g:       # full case: regenerated inputs must equal the stored ones
        for k in ACTOR_KEYS:
            assert np.array_equal(actor[k], g[f"init.actor.{k}"])
        for k in CRITIC_KEYS:
            assert np.array_equal(critic[k], g[f"init.critic.{k}"])
        for k, v in arrays.items():
            assert np.array_equal(v, g[f"buffer.{k}"])
    hp = HyperParams(act_limit=lim)
    return g, actor, critic, arrays, hp


def batch_from(arrays, idx, dtype=np.float32):
    return dict(states=arrays["state"][idx].astype(dtype), actions=arrays["actions"][idx].astype(dtype),
                rewards=arrays["rewards"][idx].astype(dtype), next_states=arrays["next_state"][idx].astype(dtype),
                done=arrays["done"][idx].astype(dtype))


def summary(v):
    flat = np.asarray(v).reshape(-1).astype(np.float64)
    return flat.sum(), np.abs(flat).sum(), np.asarray(v).reshape(-1)[:: max(1, v.size // 512)][:512]


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert np.all(err <= tol), f"{what}: max err {err.max():.3e} (tol {tol.reshape(-1)[err.argmax()]:.3e}) at {err.argmax()}"
