"""observation kernels (cloud + chamfer) on a batch of device-resident frames"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16
hand = hpe_b200.handmodel()
rng = np.random.default_rng(5)
truths = np.clip(D.X0 + rng.normal(size=(n, 26)) * D.STD * 0.5, D.LB, D.UB)
th = torch.from_numpy(truths).cuda()
obs = hpe_b200.observedmodel(capacity=n, num_samples=250)
lib = hpe_b200.lib()
own = lib.hpe_obs_depth_dev(obs._h)                 # render straight into the observation's frame buffer: no copy
hpe_b200.binding.render_dev(hand, th.data_ptr(), n, own)
torch.cuda.synchronize()
for rep in range(4):
    t0 = time.perf_counter()
    obs.upload_dev(own, 0, n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"{n} frames: {dt * 1e3:.2f} ms  {n / dt:.0f} frames/s  {n * 4 * 307200 / dt / 1e12:.3f} TB/s of 4 x 307 KB per frame")
