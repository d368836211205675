"""frames/s of hpe_track_sequence on one synthetic sequence (cluster kernel, latency path)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
hand = hpe_b200.handmodel()
rng = np.random.default_rng(3)
pose = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
walk = np.empty((n, 26))
for k in range(n):
    pose = np.clip(pose + rng.uniform(-1, 1, size=26) * np.r_[[2.0] * 3, [0.2] * 3, [2.0] * 20], D.LB, D.UB)
    walk[k] = pose
seq = torch.empty(n, D.H, D.W, dtype=torch.float32, device="cuda")
hpe_b200.binding.render_dev(hand, torch.from_numpy(walk).cuda().data_ptr(), n, seq.data_ptr())
obs = hpe_b200.observedmodel(capacity=n, num_samples=250)
obs.upload_dev(seq.data_ptr(), 0, n)
cf = hpe_b200.costfunc(hand, obs)
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
x0 = walk[0] + rng.normal(size=26) * D.STD * 0.3
pso.track_sequence(cf, x0, 64, 16)
t0 = time.perf_counter()
poses, costs = pso.track_sequence(cf, x0, 64, n)
dt = time.perf_counter() - t0
print(f"{n} frames: {n / dt:.0f} frames/s  ({dt / n / 30 * 1e6:.2f} us per generation)  checksum {costs.sum():.9f}")
