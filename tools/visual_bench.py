"""Dev/profiling script: time the visual SAC update (BASELINE configs[4] shapes) stand-alone.
    python tools/visual_bench.py [n_updates] [precision]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_actor_critic_b200.visual_engine import VisualEngine, vactor_views, vcritic_views
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
O, A, V, H, B, N = 168, 56, (3, 84, 84), [256, 256], 256, 2048
eng = VisualEngine(O, A, V, H, batch_size=B, device=0, act_limit=1.0, precision=prec)
dev = eng.tdev()
pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
g = torch.Generator(device="cpu").manual_seed(0)
for views in (vactor_views(eng.layout, pa, O, A, 3, H), vcritic_views(eng.layout, pc, O, A, 3, H)):
    for k, v in views.items():
        fan = int(np.prod(v.shape[1:])) if k.endswith("weight") else int(np.prod(views[k.replace("bias", "weight")].shape[1:]))
        v.copy_(((torch.rand(v.shape, generator=g) * 2 - 1) * fan ** -0.5 / (64.0 if "conv_0.weight" in k else 1.0)).to(dev))
pt = pc.clone()
eng.bind_params(pa, pc, pt)
eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
ring = dict(f=torch.randn(N, O, device=dev), fr=torch.randint(0, 256, (N, *V), dtype=torch.uint8, device=dev), a=torch.rand(N, A, device=dev) * 2 - 1,
            r=torch.randn(N, device=dev), nf=torch.randn(N, O, device=dev), nfr=torch.randint(0, 256, (N, *V), dtype=torch.uint8, device=dev),
            d=torch.rand(N, device=dev) < 1e-3)
eng.bind_buffer(ring["f"], ring["fr"], ring["a"], ring["r"], ring["nf"], ring["nfr"], ring["d"])
eng.set_buffer_state(0, N)
eng.update(3); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lq, lp = eng.update(n); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"visual update: {ms:.3f} ms/update  {1000/ms:.1f} upd/s  precision={prec}  loss_q {lq[-1].item():.4f} loss_pi {lp[-1].item():.4f}  launches/update {eng.launch_count() // (n + 3)}")
