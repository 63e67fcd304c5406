"""time bursts of n updates (CUDA events), C1 shapes: python tools/burst_time.py"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_actor_critic_b200.engine import Engine, actor_views, critic_views
O, A, H, B, N = 17, 6, [256, 256], 256, 100000
eng = Engine(O, A, H, batch_size=B, max_burst=50, device=0, act_limit=1.0)
dev = eng.tdev()
pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
for views in (actor_views(eng.layout, pa, O, A, H), critic_views(eng.layout, pc, O, A, H)):
    for k, v in views.items():
        v.uniform_(-0.05, 0.05)
pt = pc.clone()
eng.bind_params(pa, pc, pt)
eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
pitch = (2 * O + A + 2 + 15) // 16 * 16
rec = torch.randn(N, pitch, device=dev)
eng.bind_buffer_packed(rec); eng.set_buffer_state(0, N)
for n in (50, 50):
    for _ in range(5): eng.update(n)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.update(n); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"burst {n:3d}: median {np.median(ts):9.1f} us  min {min(ts):9.1f} us  per update {np.median(ts)/n:7.2f}")
