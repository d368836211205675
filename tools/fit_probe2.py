"""hpe_fit_frames: single calls, pipelined calls (submit / wait, two in flight) and long calls, repeated"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n = 1184
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 1776
hand = hpe_b200.handmodel()
rng = np.random.default_rng(5)
truths = np.clip(D.X0 + rng.normal(size=(n, 26)) * D.STD * 0.5, D.LB, D.UB)
frames = torch.empty(n, D.H, D.W, dtype=torch.float32, device="cuda")
hpe_b200.binding.render_dev(hand, torch.from_numpy(truths).cuda().data_ptr(), n, frames.data_ptr())
host = frames.cpu().pin_memory()
big = host.repeat(4, 1, 1).pin_memory()
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
x0 = torch.from_numpy(truths + rng.normal(size=truths.shape) * D.STD * 0.3).pin_memory()
x0b = x0.repeat(4, 1).pin_memory()
obs = hpe_b200.observedmodel(capacity=cap, num_samples=250)
cf = hpe_b200.costfunc(hand, obs)
def single():
    t0 = time.perf_counter(); pso.fit_frames(cf, host.numpy(), x0.numpy(), 64); return n / (time.perf_counter() - t0)
def long():
    t0 = time.perf_counter(); pso.fit_frames(cf, big.numpy(), x0b.numpy(), 64); return 4 * n / (time.perf_counter() - t0)
def piped(calls=4):
    t0 = time.perf_counter(); jobs = []
    for r in range(calls):
        jobs.append(pso.fit_frames_submit(cf, host.numpy(), x0.numpy(), 64))
        if len(jobs) == 2: jobs.pop(0).wait()
    jobs.pop(0).wait()
    return calls * n / (time.perf_counter() - t0)
single()
print("capacity", cap)
print("single   ", " ".join(f"{single():.0f}" for _ in range(3)))
print("long     ", " ".join(f"{long():.0f}" for _ in range(3)))
piped(2)
print("pipelined", " ".join(f"{piped():.0f}" for _ in range(3)))
print("long     ", " ".join(f"{long():.0f}" for _ in range(3)))
print("single   ", " ".join(f"{single():.0f}" for _ in range(3)))
print("pipelined", " ".join(f"{piped(8):.0f}" for _ in range(2)))
