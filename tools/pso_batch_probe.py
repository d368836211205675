"""frames/s of hpe_pso_evolve on a large batch of independent 64 x 30 swarms (device-resident frames)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 40
hand = hpe_b200.handmodel()
rng = np.random.default_rng(5)
truths = np.clip(D.X0 + rng.normal(size=(n, 26)) * D.STD * 0.5, D.LB, D.UB)
frames = torch.empty(n, D.H, D.W, dtype=torch.float32, device="cuda")
th = torch.from_numpy(truths).cuda()
hpe_b200.binding.render_dev(hand, th.data_ptr(), n, frames.data_ptr())
obs = hpe_b200.observedmodel(capacity=n, num_samples=250)
obs.upload_dev(frames.data_ptr(), 0, n)
cf = hpe_b200.costfunc(hand, obs)
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
x0 = truths + rng.normal(size=truths.shape) * D.STD * 0.3
pso.pso_evolve_batch(cf, x0[:2000], 64, want_trace=False)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bp, bc, _ = pso.pso_evolve_batch(cf, x0, 64, want_trace=False)
    dt = time.perf_counter() - t0
print(f"{n} swarms of 64 x 30: {dt * 1e3:.1f} ms  {n / dt:.0f} frames/s  {n * 64 * 30 / dt:.3e} evals/s  checksum {bc.sum():.9f}")
