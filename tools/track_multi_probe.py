"""tracked frames/s of hpe_track_sequences for several sequence counts (one cluster / CTA per sequence)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
counts = [int(a) for a in sys.argv[2:]] or [1, 8, 9, 18, 36, 37, 74, 148, 296]
hand = hpe_b200.handmodel()
rng = np.random.default_rng(3)
n_max = max(counts)
walk = np.empty((n_max, n_frames, 26))
for q in range(n_max):
    pose = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    for k in range(n_frames):
        pose = np.clip(pose + rng.uniform(-1, 1, size=26) * np.r_[[2.0] * 3, [0.2] * 3, [2.0] * 20], D.LB, D.UB)
        walk[q, k] = pose
seq = torch.empty(n_max * n_frames, D.H, D.W, dtype=torch.float32, device="cuda")
hpe_b200.binding.render_dev(hand, torch.from_numpy(walk.reshape(-1, 26)).cuda().data_ptr(), n_max * n_frames, seq.data_ptr())
obs = hpe_b200.observedmodel(capacity=n_max * n_frames, num_samples=250)
obs.upload_dev(seq.data_ptr(), 0, n_max * n_frames)
cf = hpe_b200.costfunc(hand, obs)
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
x0 = walk[:, 0] + rng.normal(size=(n_max, 26)) * D.STD * 0.3
pso.track_sequences(cf, x0[:2], 64, 2, 8)
for n_seq in counts:
    t0 = time.perf_counter()
    poses, costs = pso.track_sequences(cf, x0[:n_seq], 64, n_seq, n_frames)
    dt = time.perf_counter() - t0
    print(f"{n_seq:4d} sequences x {n_frames} frames: {n_seq * n_frames / dt:9.0f} tracked frames/s  ({dt * 1e3:.1f} ms)  checksum {costs.sum():.6f}")
