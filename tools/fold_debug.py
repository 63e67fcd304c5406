"""Dev script: folded program, burst of K vs K single-update launches vs the first (16-phase) program."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_helpers import load_case
from test_gpu_parity import _engine_for

name = sys.argv[1] if len(sys.argv) > 1 else "small"
g, actor, critic, arrays, hp = load_case(name)
K = int(g["n_updates"])
res = {}
for tag, fold, mode, single in (("fold-burst", "1", 0, False), ("fold-single", "1", 0, True), ("fold-phased", "1", 1, False), ("old-burst", "0", 0, False)):
    os.environ["SACB_FOLD"] = fold
    eng, av, cv, tv, adam = _engine_for(g, actor, critic, arrays, hp)
    eng.set_mode(mode)
    if single:
        parts = [eng.update(1, idx=g["idx"][k:k + 1], eps=g["eps"][k:k + 1]) for k in range(K)]
        lq = torch.cat([p[0] for p in parts]); lp = torch.cat([p[1] for p in parts])
    else:
        lq, lp = eng.update(K, idx=g["idx"], eps=g["eps"])
    torch.cuda.synchronize()
    res[tag] = (lq.cpu().numpy(), lp.cpu().numpy(), eng._keep["actor"].cpu().numpy(), eng._keep["critic"].cpu().numpy())
    print(tag, "loss_q", res[tag][0], "loss_pi", res[tag][1])
print("golden loss_q", g["loss_q"], "loss_pi", g["loss_pi"])
ref = res["old-burst"]
for tag, r in res.items():
    print(tag, "max|actor - old|", np.abs(r[2] - ref[2]).max(), "max|critic - old|", np.abs(r[3] - ref[3]).max())
