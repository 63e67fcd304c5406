"""Dev tool: compare GEMM paths on a synthetic config: python tools/compare_paths_synth.py O A H1 H2 B paths"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sac_numpy as onp
from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
O, A, H1, H2, B = (int(x) for x in sys.argv[1:6])
paths = [int(x) for x in sys.argv[6].split(",")]
H = [H1, H2]; N = 4 * B
rng = np.random.default_rng(3)
actor, critic = onp.init_params(O, A, H, rng)
arrays = dict(state=rng.standard_normal((N, O)).astype(np.float32), actions=rng.uniform(-0.4, 0.4, (N, A)).astype(np.float32),
              rewards=rng.standard_normal(N).astype(np.float32), next_state=rng.standard_normal((N, O)).astype(np.float32),
              done=rng.random(N) < 0.05)
idx = rng.choice(N, B, replace=False)[None]
eps = rng.standard_normal((1, 2, B, A)).astype(np.float32)
res = {}
for pth in paths:
    eng = Engine(O, A, H, batch_size=B, max_burst=2, device=0, act_limit=0.4)
    eng.set_precision(pth)
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    for k, v in actor_views(eng.layout, pa, O, A, H).items(): v.copy_(torch.from_numpy(actor[k]))
    for k, v in critic_views(eng.layout, pc, O, A, H).items(): v.copy_(torch.from_numpy(critic[k]))
    pt = pc.clone()
    eng.bind_params(pa, pc, pt)
    eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}
    eng.bind_buffer(t["state"], t["actions"], t["rewards"], t["next_state"], t["done"])
    eng.set_buffer_state(0, N)
    lq, lp = eng.update(1, idx=idx, eps=eps)
    d = {"loss_q": lq.cpu().numpy(), "loss_pi": lp.cpu().numpy()}
    for key in ("a2", "logp_next", "q1_target", "q2_target", "backup", "q1", "q2", "pi", "logp_pi", "q1_pi", "q2_pi"):
        d[key] = eng.debug_read(key)
    ga = actor_views(eng.layout, torch.from_numpy(eng.debug_read("grad_actor")), O, A, H)
    gc = critic_views(eng.layout, torch.from_numpy(eng.debug_read("grad_critic")), O, A, H)
    for k, v in ga.items(): d["ga." + k] = v.numpy().copy()
    for k, v in gc.items(): d["gc." + k] = v.numpy().copy()
    res[pth] = d
ref = res[paths[0]]
for pth in paths[1:]:
    print(f"== path {pth} vs path {paths[0]}")
    for k in ref:
        a, b = res[pth][k].astype(np.float64), ref[k].astype(np.float64)
        e = np.abs(a - b); err = e.max(); scale = np.abs(b).max() + 1e-30
        flag = "  <<<<" if err / scale > 1e-4 else ""
        where = np.unravel_index(e.argmax(), e.shape) if e.ndim else ()
        print(f"{k:32s} max|diff| {err:.3e}  rel {err/scale:.3e} at {where} n_bad {(e > 1e-4*scale).sum()}{flag}")
