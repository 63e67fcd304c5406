"""Dev tool: compare the GEMM arithmetic paths (0 FFMA, 1 mma.sync, 2 tcgen05) on a golden case, array by array."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_helpers import load_case
from test_gpu_parity import _engine_for
from torch_actor_critic_b200.engine import actor_views, critic_views

name = sys.argv[1] if len(sys.argv) > 1 else "small"
paths = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,2").split(",")]
g, actor, critic, arrays, hp = load_case(name)
O, A, H = int(g["obs_dim"]), int(g["act_dim"]), [int(h) for h in g["hidden"]]
res = {}
for pth in paths:
    eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp)
    eng.set_precision(pth)
    lq, lp = eng.update(1, idx=g["idx"][:1], eps=g["eps"][:1])
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
    print(f"== path {pth} vs path {paths[0]} ({name})")
    for k in ref:
        a, b = res[pth][k].astype(np.float64), ref[k].astype(np.float64)
        err = np.abs(a - b).max()
        scale = np.abs(b).max() + 1e-30
        flag = "  <<<<" if err / scale > 1e-4 else ""
        print(f"{k:32s} max|diff| {err:.3e}  rel {err/scale:.3e}{flag}")
