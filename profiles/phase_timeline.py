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

PHASES = ["gather(first only)", "A.L0+C.L0q1", "A.L1(+heads)+C.L1q1+C.L0q2", "rows_heads+C.L1q2", "T.L0", "T.L1", "row_backup",
          "C.dh1+db1+dW2+lossq", "C.dW1+dW0+db0", "adam_c+polyak", "Cpi.L0", "Cpi.L1", "row_pi", "Cpi.dh1(+dpi)", "row_ghead",
          "A.dWh+dh2", "A.dW1+dh1", "A.dW0+losspi", "adam_a+gather(k+1)"]


# unfused program (2A > 16 or the bf16 path): heads and d-pi are GEMM phases of their own
PHASES_UNFUSED = ["gather(first only)", "A.L0+C.L0q1", "A.L1+C.L1q1+C.L0q2", "A.heads+C.L1q2", "rows_heads", "T.L0", "T.L1",
                  "row_backup", "C.dh1+db1+dW2+lossq", "C.dW1+dW0+db0", "adam_c+polyak", "Cpi.L0", "Cpi.L1", "row_pi", "Cpi.dh1",
                  "dpi", "row_ghead", "A.dh2+losspi", "A.dh1", "A.dWh+dW1+dW0", "adam_a+gather(k+1)"]


# folded program (csrc/sacb_fold.cuh): 2 first-update-only phases + 8 phases per update (one GPU)
PHASES_FOLD = ["gather(first only)", "C.L1[s,a](first only)", "F1 A.L1(+L0,heads pp)", "F2 T.L1(+heads rows,L0)", "F3 C.dh1(+backup rows)",
               "F4 C.dW1+dW0+Adam+loss_q", "F5 Cpi.L1(+pi rows,L0)+gather(k+1)", "F6 Cpi.dh1(+minQ rows,dpi pp)", "F7 A.dh1(+ghead rows)+C.L1(k+1)",
               "F8 A.dW+Adam+C.L1(k+1)"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
    O, A, H, B, N = 17, 6, [256, 256], 256, 1_000_000
    if os.environ.get("SACB_WORKLOAD", "c1") == "c2":           # Humanoid shapes (BASELINE configs[2])
        O, A, H, B, N = 376, 17, [512, 512], 1024, 200_000
    B = int(os.environ.get("SACB_BATCH", B))                    # experiment knob (working-set studies)
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
    eng.set_precision(int(os.environ.get('SACB_PRECISION', '1')))
    if os.environ.get("SACB_RING", "packed") == "packed":       # the product's ring layout (ReplayBuffer): one aligned record per row
        pitch = (2 * O + A + 2 + 15) // 16 * 16
        rec = torch.zeros(N, pitch, device=dev)
        rec[:, :O] = s; rec[:, O:O + A] = a; rec[:, O + A] = r; rec[:, O + A + 1:2 * O + A + 1] = s2; rec[:, 2 * O + A + 1] = d.float()
        eng.bind_buffer_packed(rec)
    else:
        eng.bind_buffer(s, a, r, s2, d)
    eng.set_buffer_state(0, N)
    eng.update(50)
    torch.cuda.synchronize()
    st = eng.profile_phases(n).astype(np.int64)              # [n, P, ctas, 2]
    nP, nC = st.shape[1], st.shape[2]
    names = (PHASES if nP == len(PHASES) else PHASES_UNFUSED if nP == len(PHASES_UNFUSED) else PHASES_FOLD if nP == len(PHASES_FOLD)
             else [f"p{i}" for i in range(nP)])
    work_end = st[..., 0].max(axis=2)                        # slowest CTA done with its tiles
    bar_end = st[..., 1].max(axis=2)                         # everyone past the barrier
    first_work_end = np.where(st[..., 0] > 0, st[..., 0], np.iinfo(np.int64).max).min(axis=2)
    work = np.zeros(nP); barrier = np.zeros(nP); fastest = np.zeros(nP); cnt = np.zeros(nP)
    per_update = []
    prev_end = None
    for k in range(n):
        t_start = None
        for p in range(nP):
            if work_end[k, p] == 0:                          # phase skipped for this update (initial gather)
                continue
            if prev_end is not None and k >= 1:
                work[p] += work_end[k, p] - prev_end
                fastest[p] += first_work_end[k, p] - prev_end
                barrier[p] += max(bar_end[k, p], work_end[k, p]) - work_end[k, p]
                cnt[p] += 1
            if t_start is None:
                t_start = prev_end
            prev_end = max(bar_end[k, p], work_end[k, p])
        if k >= 1 and t_start is not None:
            per_update.append(prev_end - t_start)
    cnt = np.maximum(cnt, 1)
    work, barrier, fastest = work / cnt / 1e3, barrier / cnt / 1e3, fastest / cnt / 1e3
    per_update = float(np.mean(per_update)) / 1e3
    print(f"# persistent sac_phase_kernel, {nC} CTAs, {nP} phases, mean over {n - 1} updates (first skipped); microseconds")
    print(f"{'phase':<24}{'work(max cta)':>14}{'work(min cta)':>14}{'barrier':>10}{'total':>9}")
    for i in range(nP):
        print(f"{names[i]:<24}{work[i]:>14.2f}{fastest[i]:>14.2f}{barrier[i]:>10.2f}{work[i] + barrier[i]:>9.2f}")
    print(f"{'SUM':<24}{work.sum():>14.2f}{'':>14}{barrier.sum():>10.2f}{(work + barrier).sum():>9.2f}")
    print(f"per update (stamps): {per_update:.2f} us  -> {1e6 / per_update:.0f} updates/s")
    # SACB_DUMP_PHASES="8,16": per-CTA time from the previous phase's end to "tiles done" for those phases (update 2),
    # CTA i runs tile i of a one-round phase, so this attributes the phase's critical path to a job
    for p in [int(x) for x in os.environ.get("SACB_DUMP_PHASES", "").split(",") if x]:
        k = min(2, n - 1)
        prev = max(st[k, p - 1, :, 1].max(), st[k, p - 1, :, 0].max())
        d = (st[k, p, :, 0] - prev) / 1e3
        print(f"# phase {p} per-CTA work (us): " + " ".join(f"{x:.1f}" for x in d))


if __name__ == "__main__":
    main()
