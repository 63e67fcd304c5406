"""precision 3: TMA-fed bf16 tiles vs the staged cp.async bf16 tiles -- same arithmetic, so results must agree (bit-for-bit
when the accumulation order is the same).  python tools/tma_check.py [shape]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_actor_critic_b200.engine import Engine, actor_views, critic_views

SHAPES = {"humanoid": (376, 17, [512, 512], 1024, 4096), "ragged8": (24, 5, [72, 40], 96, 700), "c1": (17, 6, [256, 256], 256, 4096),
          "odd": (11, 3, [72, 40], 96, 700)}

def run(shape, tma, n_upd, mode=0):
    O, A, H, B, N = SHAPES[shape]
    os.environ["SACB_TMA"] = "1" if tma else "0"
    os.environ["SACB_TMA_VERBOSE"] = "1"
    g = torch.Generator(device="cpu").manual_seed(3)
    eng = Engine(O, A, H, batch_size=B, max_burst=8, device=0, act_limit=1.0)
    eng.set_precision(3); eng.set_mode(mode)
    dev = eng.tdev()
    pa, pc = eng.new_actor_arena(), eng.new_critic_arena()
    for views in (actor_views(eng.layout, pa, O, A, H), critic_views(eng.layout, pc, O, A, H)):
        for k, v in views.items():
            fan_in = v.shape[-1] if k.endswith("weight") else views[k.replace("bias", "weight")].shape[-1]
            v.copy_(((torch.rand(v.shape, generator=g) * 2 - 1) * fan_in ** -0.5).to(dev))
    pt = pc.clone()
    eng.bind_params(pa, pc, pt)
    eng.bind_adam(eng.new_actor_arena(), eng.new_actor_arena(), 0, eng.new_critic_arena(), eng.new_critic_arena(), 0)
    s = torch.randn(N, O, generator=g).to(dev); s2 = torch.randn(N, O, generator=g).to(dev)
    a = (torch.rand(N, A, generator=g) * 2 - 1).to(dev); r = torch.randn(N, generator=g).to(dev); d = (torch.rand(N, generator=g) < 0.05).to(dev)
    eng.bind_buffer(s, a, r, s2, d); eng.set_buffer_state(0, N)
    idx = torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(n_upd)]).numpy()
    eps = torch.randn(n_upd, 2, B, A, generator=g).numpy().astype(np.float32)
    lq, lp = eng.update(n_upd, idx=idx, eps=eps)
    torch.cuda.synchronize()
    out = dict(lq=lq.cpu().numpy(), lp=lp.cpu().numpy(), pa=pa.cpu().numpy(), pc=pc.cpu().numpy(), pt=pt.cpu().numpy(),
               ga=eng.debug_read("grad_actor"), gc=eng.debug_read("grad_critic"))
    for k in ("q1", "q2", "backup", "a2", "pi", "q1_pi"):
        out[k] = eng.debug_read(k)
    return out

def main():
    shapes = sys.argv[1:] or ["ragged8", "humanoid"]
    for shape in shapes:
        for n in (1, 3):
            a = run(shape, True, n); b = run(shape, False, n)
            print(f"== {shape} n={n}  loss_q {a['lq']} vs {b['lq']}  loss_pi {a['lp']} vs {b['lp']}")
            for k in a:
                x, y = np.asarray(a[k], np.float64), np.asarray(b[k], np.float64)
                err = np.abs(x - y).max(); sc = np.abs(y).max() + 1e-30
                print(f"   {k:8s} max|diff| {err:.3e}  rel {err / sc:.3e}  {'EXACT' if err == 0 else ''}")
main()
