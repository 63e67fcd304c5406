"""CPU oracle for the SAC update hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference's hot path
(`/root/reference/sac/algorithm.py:30-162,274-281`, `networks/linear.py:13-79`,
`buffer/replay_buffer.py:17-54`, and torch's `_single_tensor_adam`).  It is the
checker the CUDA engine is compared against.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it; the product package `torch-actor-critic_b200/` never does
and fails loudly when its CUDA library is missing.

Parity pinning: the reference's own tests assert shapes only,
so there are no reference-owned golden vectors for this path.  The oracle is
pinned instead against outputs of the UNMODIFIED reference executed in the
build container through import stubs (`oracle/make_golden.py`); those outputs
are committed under `tests/golden/` and `tests/test_oracle_golden.py` checks
both oracle flavours against them.

Two flavours:
  * `oracle.sac_numpy`  -- numpy, hand-derived backward (no autograd), float32
    or float64.  Independent derivation of every gradient the CUDA kernels
    compute.
  * `oracle.sac_torch_cpu` -- torch CPU ops + autograd + torch.optim.Adam, the
    same third-party arithmetic the reference delegates to.  Used as the timed
    CPU baseline ("port") because the reference tree cannot travel to the GPU box.
"""
