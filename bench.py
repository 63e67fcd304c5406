#!/usr/bin/env python
"""bench.py -- SAC gradient-updates/sec on the BASELINE.json workload (configs[1]: HalfCheetah-shaped synthetic
buffer, 1e6 transitions in HBM, obs=17, act=6, MLP 256x256, batch=256, fp32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one SAC gradient update = sample -> update_critic -> update_policy -> update_targets
(reference sac/algorithm.py:275-278).  Updates are issued in bursts of update_every=50 (sac/algorithm.py:273-274),
each burst being one persistent-kernel launch.  N>1: one process per GPU under torchrun, every rank owns an
independent replay shard (default_rng(rank)) and the critic / actor gradients are NCCL-averaged inside every update
(weak scaling: batch 256 per GPU).  `value` counts batch-256 gradient updates summed over GPUs (= samples/s / 256);
`config.global_steps_per_s` is the synchronous optimizer-step rate.

`--impl reference` times the CPU arm: the torch-CPU port of the reference's hot loop (oracle/sac_torch_cpu.py; the
reference tree itself is Python under /root/reference and does not exist on the GPU box) on the same synthetic
arrays, on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

O, A, H, B, N_BUF, UPDATE_EVERY = 17, 6, [256, 256], 256, 1_000_000, 50
P_ACTOR, P_CRITIC = 73_484, 144_386
# algorithmic work per update (SURVEY.md section 8d / BASELINE.md section 4)
BYTES_PER_UPDATE = B * (2 * O + A + 2) * 4 + 28 * (P_ACTOR + P_CRITIC) + 12 * P_CRITIC     # 7,876,000
A_F = O * 256 + 256 * 256 + 2 * 256 * A
C_F = (O + A) * 256 + 256 * 256 + 256
FLOPS_PER_UPDATE = 2 * B * (3 * A_F + 10 * C_F + 3 * 256 * 256 + 2 * 256 + 2 * 256 * A)   # 581.6 MFLOP
METRIC = "SAC gradient-updates/sec (batch=256, MLP 256x256)"
UNIT = "updates/s"
WORKLOAD = "HalfCheetah-shaped synthetic replay (obs=17, act=6, 1e6 transitions), MLP 256x256, batch=256, fp32"


def synth_buffer(rank, n=N_BUF):
    """states/next_states ~ N(0,1), actions ~ U(-1,1), rewards ~ N(0,1), done ~ Bernoulli(1e-3)."""
    rng = np.random.default_rng(rank)
    return dict(state=rng.standard_normal((n, O), dtype=np.float32), actions=rng.uniform(-1, 1, (n, A)).astype(np.float32),
                rewards=rng.standard_normal(n, dtype=np.float32), next_state=rng.standard_normal((n, O), dtype=np.float32),
                done=rng.random(n) < 1e-3)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(arrays, threads, n_updates, warmup, max_seconds):
    """updates/s of the torch-CPU port of sac/algorithm.py:274-281 on the host (oracle = checker / baseline only).
    Bounded: stops after `max_seconds` of timed work."""
    import torch
    from oracle.sac_numpy import HyperParams, init_params
    from oracle.sac_torch_cpu import HostReplay, TorchCpuSAC, time_cpu_updates
    torch.set_num_threads(threads)
    actor, critic = init_params(O, A, H, np.random.default_rng(0))
    sac = TorchCpuSAC(actor, critic, HyperParams(act_limit=1.0))
    replay = HostReplay(arrays["state"], arrays["actions"], arrays["rewards"], arrays["next_state"], arrays["done"])
    dt, done = time_cpu_updates(replay, sac, B, n_updates, warmup, UPDATE_EVERY, max_seconds=max_seconds)
    return done / dt, done


def cpu_baseline(arrays, steps, warmup, budget_s):
    """Best thread count among 1, 2 (the reference's own setting, main.py:130), 8, 16, 32 (<= host cores): the update is a
    chain of tiny GEMMs, so more threads stop helping early and oversubscription hurts.  Short probes pick the thread
    count, then the bounded sample runs at it."""
    cores = os.cpu_count() or 1
    cands = sorted({t for t in (1, 2, 8, 16, 32) if t <= cores})
    probe = {}
    for t in cands:
        probe[t] = round(cpu_port_rate(arrays, t, 40, 5, budget_s / (4 * len(cands)))[0], 2)
    best = max(probe, key=probe.get)
    rate, done = cpu_port_rate(arrays, best, steps, warmup, budget_s / 2)
    return rate, best, done, probe, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    arrays = synth_buffer(0)
    rate, thr, done, probe, cores = cpu_baseline(arrays, max(1, args.steps), max(3, min(args.warmup, 50)), budget_s=60.0)
    sample = (f"{done} updates (bursts of {UPDATE_EVERY}) timed at {thr} torch threads after warm-up, torch {torch.__version__} CPU; "
              f"thread probe (updates/s) {probe}; host cpu_count={cores}")
    line = {"impl": "reference", "metric": METRIC, "value": round(rate, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": max(3, min(args.warmup, 50)), "ms_per_step": round(1000.0 / rate, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "update_every": UPDATE_EVERY, "host_cpu_count": cores},
            "cpu_baseline": {"value": round(rate, 2), "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": round(rate, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import torch.optim as optim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from torch_actor_critic_b200.buffer.replay_buffer import ReplayBuffer
    from torch_actor_critic_b200.networks.linear import Actor, DoubleCritic
    from torch_actor_critic_b200.sac.algorithm import SAC
    from torch_actor_critic_b200.sac.mpi import sync_params
    from copy import deepcopy

    arrays = synth_buffer(rank)
    torch.manual_seed(0)
    actor = Actor(O, A, H, act_limit=1.0 if args.workload == 'c1' else 0.4, device=local)
    critic = DoubleCritic(O, A, H, device=local)
    target = deepcopy(critic)
    sync_params(actor); sync_params(critic); sync_params(target)
    pi_opt = optim.Adam(actor.parameters(), lr=3e-4)
    q_opt = optim.Adam(critic.parameters(), lr=3e-4)
    sac = SAC(alpha=0.2, gamma=0.99, polyak=0.995, reward_scale=1.0, epochs=1, batch_size=B, start_steps=0,
              steps_per_epoch=1, max_ep_len=1000, update_after=0, update_every=UPDATE_EVERY, save_every=1)
    sac.seed = 10000 * rank                                        # sac/algorithm.py:203
    buf = ReplayBuffer(N_BUF, O, A, device=local, seed=10000 * rank)
    buf.store_many(arrays["state"], arrays["actions"], arrays["rewards"], arrays["next_state"], arrays["done"])
    torch.cuda.synchronize()

    def burst(n):
        return sac.update_burst(n, actor, critic, target, buf, pi_opt, q_opt)

    def run_updates(total, with_host_io):
        """`total` updates in bursts of UPDATE_EVERY.  with_host_io: the reference training loop's host traffic --
        before each burst the update_every new transitions go host->HBM (ReplayBuffer.store, pinned staging inside
        sacb_store), after it the 2 x n losses come back (sac/algorithm.py:280-281)."""
        done, cursor, last = 0, 0, None
        while done < total:
            n = min(UPDATE_EVERY, total - done)
            if with_host_io:
                for i in range(n):
                    j = (cursor + i) % N_BUF
                    buf.store(arrays["state"][j], arrays["actions"][j], arrays["rewards"][j], arrays["next_state"][j],
                              arrays["done"][j])
                cursor += n
            lq, lp = burst(n)
            if with_host_io:
                last = (lq.cpu(), lp.cpu())
            done += n
        return last

    def timed(total, with_host_io):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        run_updates(total, with_host_io)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), wall * 1e3 if with_host_io else 0.0)     # host I/O arm: wall clock dominates
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    eng = sac._engine(actor, critic, B, target=target, pi_opt=pi_opt, q_opt=q_opt)
    eng.set_mode(args.mode)
    eng.set_precision(args.precision)
    run_updates(args.warmup, False)                                 # W untimed warm-up steps
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ms = timed(args.steps, False)                                   # EXACTLY K timed steps, inputs resident in HBM
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    run_updates(min(args.warmup, 2 * UPDATE_EVERY), True)
    ms_e2e = timed(args.steps, True)

    # dominant kernel (the persistent sac_phase_kernel): CUDA events around single full bursts on the launch stream
    kms = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); burst(UPDATE_EVERY); e1.record(); torch.cuda.synchronize()
        kms.append(e0.elapsed_time(e1))
    k_ms = float(np.mean(kms))

    if rank != 0:
        return
    hbm_peak, tensor_peak, peak_src = peaks()
    upd_per_launch = UPDATE_EVERY
    ach = BYTES_PER_UPDATE * upd_per_launch / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch") if args.workload == "c1" else None
    value = world * args.steps / (ms * 1e-3)
    e2e_v = world * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 6), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == 3 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "update_every": UPDATE_EVERY, "batch_per_gpu": B,
                   "gemm_arithmetic": GEMM_ARITH[args.precision], "global_batch": B * world,
                   "parallelism": f"dp{world}", "global_steps_per_s": round(args.steps / (ms * 1e-3), 1),
                   "samples_per_s": round(value * B, 1),
                   "l2_policy": f"inputs larger than L2: {N_BUF * ((2 * O + A + 1) * 4 + 1) / 1e6:.0f} MB replay ring, random-row gather; the "
                                f"{(P_ACTOR + P_CRITIC) * 12 / 1e6 + P_CRITIC * 4 / 1e6:.1f} MB of weights/Adam state is the algorithm's "
                                "working set and stays L2-resident by design",
                   "value_definition": f"batch-{B} gradient updates summed over GPUs"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_v, 1), "unit": UNIT, "h2d_bytes_per_step": (2 * O + A + 1) * 4 + 1, "d2h_bytes_per_step": 8,
                "path": "ReplayBuffer.store (host rows -> pinned staging -> HBM ring) + SAC.update_burst + loss readback per burst"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(ach, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 5),
                     "traffic": traffic, "peak_source": peak_src, "kernel": "sac_phase_kernel (persistent, 50 updates/launch)",
                     "kernel_ms_per_launch": round(k_ms, 4), "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * upd_per_launch,
                     "flops": {"achieved_tflops": round(FLOPS_PER_UPDATE * upd_per_launch / (k_ms * 1e-3) / 1e12, 3),
                               "note": "latency-bound dependency chain of ~18 grid-synchronised phases per update"}},
    }
    if world == 1 and not args.no_cpu:
        rate, thr, done, probe, cores = cpu_baseline(arrays, 600, 20, budget_s=30.0)
        line["cpu_baseline"] = {"value": round(rate, 2), "unit": UNIT, "cores": thr, "kind": "port",
                                "sample": f"{done} updates at {thr} torch threads on the same synthetic arrays; thread probe "
                                          f"(updates/s) {probe}; host cpu_count={cores}"}
    print(json.dumps(line), flush=True)


GEMM_ARITH = {0: "fp32 FFMA", 1: "tensor-core 3xTF32 via mma.sync (fp32-level accuracy), fp32 accumulate",
              2: "tensor-core 3xTF32 via tcgen05/TMEM (fp32-level accuracy), fp32 accumulate",
              3: "tcgen05 bf16 operands (rounded from fp32 on load), fp32 accumulate in TMEM, fp32 master weights/activations/Adam"}


def set_workload(name):
    global O, A, H, B, N_BUF, P_ACTOR, P_CRITIC, BYTES_PER_UPDATE, FLOPS_PER_UPDATE, WORKLOAD
    if name == "c2":
        O, A, H, B, N_BUF = 376, 17, [512, 512], 1024, 200_000
        P_ACTOR, P_CRITIC = 473_122, 929_794
        BYTES_PER_UPDATE = B * (2 * O + A + 2) * 4 + 28 * (P_ACTOR + P_CRITIC) + 12 * P_CRITIC
        a_f = O * 512 + 512 * 512 + 2 * 512 * A
        c_f = (O + A) * 512 + 512 * 512 + 512
        FLOPS_PER_UPDATE = 2 * B * (3 * a_f + 10 * c_f + 3 * 512 * 512 + 2 * 512 + 2 * 512 * A)
        WORKLOAD = "Humanoid-shaped synthetic replay (obs=376, act=17, 2e5 transitions), MLP 512x512, batch=1024, fp32 storage"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", type=int, default=0, help="profiling only: 1 = one kernel launch per phase")
    ap.add_argument("--no-cpu", action="store_true", help="profiling only: skip the cpu_baseline leg")
    ap.add_argument("--precision", type=int, default=1, help="1 = mma.sync 3xTF32 GEMMs (default), 0 = fp32 FFMA, 2 = tcgen05 3xTF32, 3 = tcgen05 bf16 operands")
    ap.add_argument("--workload", default="c1", choices=["c1", "c2"],
                    help="c1 = BASELINE metric workload (default); c2 = Humanoid-shaped (obs=376, act=17, MLP 512x512, batch=1024), "
                         "fp32 storage -- side measurement only, not the bench line the driver reads")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    set_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
