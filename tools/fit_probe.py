"""frames/s of hpe_fit_frames (host frames in, best poses out) for several observation-buffer capacities"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n = 1184
hand = hpe_b200.handmodel()
rng = np.random.default_rng(5)
truths = np.clip(D.X0 + rng.normal(size=(n, 26)) * D.STD * 0.5, D.LB, D.UB)
frames = torch.empty(n, D.H, D.W, dtype=torch.float32, device="cuda")
hpe_b200.binding.render_dev(hand, torch.from_numpy(truths).cuda().data_ptr(), n, frames.data_ptr())
host = frames.cpu().pin_memory()
big = host.repeat(4, 1, 1).pin_memory()
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
x0 = truths + rng.normal(size=truths.shape) * D.STD * 0.3
x0b = np.tile(x0, (4, 1))
ref = None
for cap in [int(a) for a in sys.argv[1:]] or [888, 1184, 1776, 2368]:
    obs = hpe_b200.observedmodel(capacity=cap, num_samples=250)
    cf = hpe_b200.costfunc(hand, obs)
    pso.fit_frames(cf, host.numpy(), x0, 64)
    t0 = time.perf_counter(); bp, bc = pso.fit_frames(cf, host.numpy(), x0, 64); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); bp4, bc4 = pso.fit_frames(cf, big.numpy(), x0b, 64); dt4 = time.perf_counter() - t0
    if ref is None: ref = bc
    same = np.array_equal(bc, ref) and np.array_equal(bc4[:n], ref) and np.array_equal(bc4[-n:], ref)
    print(f"capacity {cap}: {n} frames {n / dt:.0f} frames/s   {4 * n} frames {4 * n / dt4:.0f} frames/s   identical {same}  checksum {bc.sum():.9f}")
    del obs, cf
