"""Per-phase timeline of the persistent SAC kernel from in-kernel %globaltimer stamps (sacb_debug_profile).

    python profiles/phase_timeline.py [n_updates] > gpurun_out/phase_timeline.txt

For every phase of one update: when the slowest CTA finished its tiles (work), how long the grid barrier then took,
how many tiles the phase has.  Not taken under a profiler; the stamps cost two global stores per CTA per phase.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PHASES = ["gather", "A.L0+C.L0", "A.L1+C.L1", "A.heads", "row_heads", "T.L0", "T.L1", "row_backup", "C.dW1+dh1+db1+dW2",
          "C.dW0+db0+lossq", "adam_c+polyak", "Cpi.L0", "Cpi.L1", "row_pi", "Cpi.dh1", "dpi", "row_ghead", "A.dWh+dh2",
          "A.dW1+dh1", "A.dW0+losspi", "adam_a"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    O, A, H, B, N = 17, 6, [256, 256], 256, 1_000_000
    eng = Engine(O, A, H, batch_size=B, max_burst=50, device=0, act_limit=1.0)
    dev = eng.tdev()
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    for views in (actor_views(eng.layout, pa, O, A, H), critic_views(eng.layout, pc, O, A, H)):
        for k, v in views.items():
            fan_in = v.shape[-1] if k.endswith("weight") else views[k.replace("bias", "weight")].shape[-1]
            v.uniform_(-fan_in ** -0.5, fan_in ** -0.5)
    pt = pc.clone()
    eng.bind_params(pa, pc, pt)
    eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
    s = torch.randn(N, O, device=dev); s2 = torch.randn(N, O, device=dev)
    a = torch.rand(N, A, device=dev) * 2 - 1; r = torch.randn(N, device=dev); d = torch.rand(N, device=dev) < 1e-3
    eng.bind_buffer(s, a, r, s2, d)
    eng.set_buffer_state(0, N)
    eng.update(50)
    torch.cuda.synchronize()
    st = eng.profile_phases(n).astype(np.int64)              # [n, P, ctas, 2]
    nP, nC = st.shape[1], st.shape[2]
    names = PHASES if nP == len(PHASES) else [f"p{i}" for i in range(nP)]
    work_end = st[..., 0].max(axis=2)                        # slowest CTA done with its tiles
    bar_end = st[..., 1].max(axis=2)                         # everyone past the barrier
    first_work_end = st[..., 0].min(axis=2)
    start = np.empty_like(bar_end)
    start[:, 1:] = bar_end[:, :-1]
    start[1:, 0] = bar_end[:-1, -1]
    start[0, 0] = first_work_end[0, 0]
    sl = slice(1, None) if n > 1 else slice(0, None)         # skip the first update (cold)
    work = (work_end - start)[sl].mean(axis=0) / 1e3
    barrier = (bar_end - work_end)[sl].mean(axis=0) / 1e3
    fastest = (first_work_end - start)[sl].mean(axis=0) / 1e3
    per_update = (bar_end[sl, -1] - start[sl, 0]).mean() / 1e3
    print(f"# persistent sac_phase_kernel, {nC} CTAs, {nP} phases, mean over {work_end[sl].shape[0]} updates; microseconds")
    print(f"{'phase':<22}{'work(max cta)':>14}{'work(min cta)':>14}{'barrier':>10}{'total':>9}")
    for i in range(nP):
        print(f"{names[i]:<22}{work[i]:>14.2f}{fastest[i]:>14.2f}{barrier[i]:>10.2f}{work[i] + barrier[i]:>9.2f}")
    print(f"{'SUM':<22}{work.sum():>14.2f}{'':>14}{barrier.sum():>10.2f}{(work + barrier).sum():>9.2f}")
    print(f"per update (stamps): {per_update:.2f} us  -> {1e6 / per_update:.0f} updates/s")


if __name__ == "__main__":
    main()
