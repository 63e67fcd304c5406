"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol include/sacb200.h
declares, computes the arena layout, and fails loudly without a GPU (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from torch_actor_critic_b200 import _lib
    return _lib


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sacb200.h")).read()
    declared = set(re.findall(r"\b(sac[bv]_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), f"{name} declared in include/sacb200.h but not exported"
    assert declared == set(lib.EXPORTED_SYMBOLS), declared ^ set(lib.EXPORTED_SYMBOLS)


def test_config_struct_matches_c_abi(lib):
    # struct_size is validated by the library itself: a mismatching ctypes mirror is rejected
    from torch_actor_critic_b200.engine import make_config
    cfg = make_config(17, 6, [256, 256])
    lay = lib.SacbLayout()
    assert lib.load().sacb_layout(ctypes.byref(cfg), ctypes.byref(lay)) == 0
    cfg.struct_size = 4
    assert lib.load().sacb_layout(ctypes.byref(cfg), ctypes.byref(lay)) == lib.SACB_ERR_INVALID
    assert b"ABI" in lib.load().sacb_last_error()


@pytest.mark.parametrize("O,A,H", [(17, 6, (256, 256)), (376, 17, (512, 512)), (5, 3, (32, 32)), (11, 3, (48, 24)), (3, 1, (7, 5))])
def test_layout_covers_reference_parameters(lib, O, A, H):
    """Arena layout holds exactly the reference's tensors (SURVEY.md a5/a7 parameter counts), 16-byte aligned groups,
    twin matrices adjacent so that both Q networks run as one GEMM."""
    import torch
    from torch_actor_critic_b200.engine import actor_views, critic_views, layout_for
    lay = layout_for(O, A, H)
    h1, h2 = H
    n_actor = h1 * O + h1 + h2 * h1 + h2 + 2 * (A * h2 + A)
    n_critic = 2 * (h1 * (O + A) + h1 + h2 * h1 + h2 + h2 + 1)
    if (O, A, H) == (17, 6, (256, 256)):
        assert (n_actor, n_critic) == (73484, 144386)
    av = actor_views(lay, torch.zeros(lay.actor_floats), O, A, list(H))
    cv = critic_views(lay, torch.zeros(lay.critic_floats), O, A, list(H))
    assert sum(v.numel() for v in av.values()) == n_actor and lay.actor_floats >= n_actor
    assert sum(v.numel() for v in cv.values()) == n_critic and lay.critic_floats >= n_critic
    assert lay.actor_floats % 4 == 0 and lay.critic_floats % 4 == 0
    for off in (lay.actor_w0, lay.actor_b0, lay.actor_w1, lay.actor_b1, lay.actor_w_mu, lay.actor_b_mu,
                lay.critic_w0[0], lay.critic_b0[0], lay.critic_w1[0], lay.critic_b1[0], lay.critic_w2[0], lay.critic_b2[0]):
        assert off % 4 == 0
    assert lay.actor_w_ls == lay.actor_w_mu + A * h2 and lay.actor_b_ls == lay.actor_b_mu + A
    assert lay.critic_w0[1] == lay.critic_w0[0] + h1 * (O + A) and lay.critic_b0[1] == lay.critic_b0[0] + h1
    assert lay.critic_w2[1] == lay.critic_w2[0] + h2 and lay.critic_b2[1] == lay.critic_b2[0] + 1
    # no two tensors overlap
    mark = np.zeros(lay.critic_floats, int)
    base = torch.zeros(lay.critic_floats)
    for v in critic_views(lay, base, O, A, list(H)).values():
        o = v.storage_offset()
        mark[o:o + v.numel()] += 1
    assert mark.max() == 1


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from torch_actor_critic_b200.engine import Engine, make_config
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(17, 6, [256, 256])
    cfg = make_config(17, 6, [256, 256])
    h = ctypes.c_void_p()
    rc = lib.load().sacb_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == lib.SACB_ERR_CUDA and b"no CPU fallback" in lib.load().sacb_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "torch-actor-critic_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_hidden_layer_count_is_checked(lib):
    from torch_actor_critic_b200.engine import make_config
    with pytest.raises(NotImplementedError):
        make_config(3, 2, [8, 8, 8])
