"""Loaders for the visual golden fixtures (tests/golden/visual_*.npz, made by oracle/make_golden_visual.py)."""
import os

import numpy as np

from oracle.visual_torch import VisualHyperParams, init_visual_params, make_visual_arrays

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_visual_case(name):
    """(golden dict, actor_np, critic_np, arrays, hp): inputs are regenerated from the recorded seed exactly as
    oracle/make_golden_visual.py drew them (parameters, then buffer arrays, from one numpy Generator)."""
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    O, A, B = int(g["obs_dim"]), int(g["act_dim"]), int(g["batch"])
    vis = tuple(int(v) for v in g["vis_dim"])
    hidden = tuple(int(h) for h in g["hidden"])
    rng = np.random.default_rng(int(g["seed"]))
    actor, critic = init_visual_params(O, A, vis, hidden, rng, frame_scale=float(g["frame_scale"]))
    arrays = make_visual_arrays(rng, int(g["n_buf"]), O, A, vis, float(g["act_limit"]))
    hp = VisualHyperParams(act_limit=float(g["act_limit"]))
    # the regenerated inputs must be the ones the reference saw
    idx0 = g["idx"][0]
    assert np.array_equal(arrays["features"][idx0], g["batch0.features"])
    assert np.array_equal(arrays["frames"][idx0].astype(np.float64).sum(axis=(1, 2, 3)), g["batch0.frames_sum"])
    assert np.array_equal(arrays["next_frames"][idx0].astype(np.float64).sum(axis=(1, 2, 3)), g["batch0.next_frames_sum"])
    assert np.array_equal(arrays["actions"][idx0], g["batch0.actions"]) and np.array_equal(arrays["rewards"][idx0], g["batch0.rewards"])
    assert np.array_equal(arrays["done"][idx0].astype(np.float32), g["batch0.done"])
    return g, actor, critic, arrays, hp


def close_frac(a, b, rtol, atol, what, frac=0.0, atol_out=0.0):
    """|a - b| <= atol + rtol |b| for all but a fraction `frac` of the entries, which must still be within `atol_out`.
    (After ONE Adam step a parameter moved by lr * g / (|g| + eps'): an entry whose gradient is ~1e-8 -- first-conv weights
    behind mostly-dead ReLUs -- is decided by rounding noise, so a few entries may be off by up to lr = 3e-4.)"""
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    err = np.abs(a - b)
    bad = err > atol + rtol * np.abs(b)
    assert bad.mean() <= frac + 1e-12, f"{what}: {bad.sum()} of {bad.size} entries beyond tolerance (max err {err.max():.3e})"
    if bad.any():
        assert err[bad].max() <= atol_out, f"{what}: outlier {err[bad].max():.3e} > {atol_out:.1e}"


def check_summary(g, prefix, name, value, rtol, atol, assert_close, frac=0.0, atol_out=0.0):
    """Compare `value` with a stored array or with its (sum, abs-sum, strided subsample) summary."""
    value = np.asarray(value)
    if f"{prefix}.{name}" in g:
        close_frac(value, g[f"{prefix}.{name}"], rtol, atol, f"{prefix}.{name}", frac, atol_out)
    else:
        sub = value.reshape(-1)[:: max(1, value.size // 512)][:512]
        close_frac(sub, g[f"{prefix}.{name}.sub"], rtol, atol, f"{prefix}.{name}.sub", frac, atol_out)
        s = value.reshape(-1).astype(np.float64).sum()
        assert abs(s - float(g[f"{prefix}.{name}.sum"])) <= atol * value.size + rtol * float(g[f"{prefix}.{name}.abssum"]), f"{prefix}.{name}.sum"
