"""Latency of the acting path `actor(state)` (one env step, n rows): cluster kernel vs the batched phase-program path.
    python profiles/act_latency.py > gpurun_out/act_latency.txt"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_actor_critic_b200.networks.linear import Actor

actor = Actor(17, 6, [256, 256], act_limit=1.0)
eng = actor._fwd_engine()
for n in (1, 8, 32):
    x = torch.randn(n, 17, device="cuda")
    for mode, label in ((0, "cluster kernel (1 launch, 8-CTA cluster, DSMEM)"), (1, "phase programs (4 launches + 3 copies)")):
        eng.set_mode(mode)
        for _ in range(20):
            eng.actor_forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 300
        t0 = time.perf_counter(); e0.record()
        for _ in range(iters):
            eng.actor_forward(x)
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / iters * 1e6
        print(f"n={n:3d}  {label:50s} device {e0.elapsed_time(e1) / iters * 1e3:8.2f} us/call   host wall {wall:8.2f} us/call")
