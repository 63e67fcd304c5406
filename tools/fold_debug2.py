"""Dev script: which intermediate of update k=1 differs between a folded burst of 2 and two single launches."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_helpers import load_case
from test_gpu_parity import _engine_for
name = sys.argv[1] if len(sys.argv) > 1 else "small"
g, actor, critic, arrays, hp = load_case(name)
os.environ["SACB_FOLD"] = "1"
keys = ("rewards", "done", "a2", "logp_next", "q1_target", "q2_target", "backup", "q1", "q2", "pi", "logp_pi", "q1_pi", "q2_pi", "grad_critic", "grad_actor")
out = {}
for tag in ("burst", "single"):
    eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp)
    if tag == "burst":
        eng.update(2, idx=g["idx"][:2], eps=g["eps"][:2])
    else:
        eng.update(1, idx=g["idx"][:1], eps=g["eps"][:1]); eng.update(1, idx=g["idx"][1:2], eps=g["eps"][1:2])
    torch.cuda.synchronize()
    out[tag] = {k: eng.debug_read(k) for k in keys}
for k in keys:
    a, b = out["burst"][k], out["single"][k]
    d = np.abs(a - b)
    bad = np.argwhere(d.reshape(d.shape[0], -1).max(1) > 1e-6).ravel() if d.ndim else []
    print(f"{k:12s} max diff {d.max():.3e}  rows/elems differing: {len(bad)} of {d.shape[0]}  first {bad[:12]}")
