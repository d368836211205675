"""Sequential tracking probe (BASELINE configs[2]): one synthetic sequence, 64 x 30 PSO per frame,
frame k+1 starts from frame k's best pose.  Prints frames/s of hpe_track_sequence."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n_fr = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
num_p = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(3)
pose = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
truths = []
for _ in range(n_fr):       # smooth random walk: <= 2 deg / 2 mm per frame
    pose = np.clip(pose + rng.uniform(-1, 1, size=26) * np.r_[[2.0] * 3, [0.2] * 3, [2.0] * 20], D.LB, D.UB)
    truths.append(pose.copy())
truths = np.array(truths)
hand = hpe_b200.handmodel()
t0 = time.perf_counter()
frames = hpe_b200.render_frames(hand, truths)
obs = hpe_b200.observedmodel(capacity=n_fr)
t1 = time.perf_counter()
obs.upload(frames)
t2 = time.perf_counter()
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
cf = hpe_b200.costfunc(hand, obs)
x0 = truths[0] + rng.normal(size=26) * D.STD * 0.3
pso.track_sequence(cf, x0, num_p, 8)
for rep in range(2):
    t3 = time.perf_counter()
    poses, costs = pso.track_sequence(cf, x0, num_p, n_fr)
    dt = time.perf_counter() - t3
    S, J = hand.build_hand_models(poses); St, Jt = hand.build_hand_models(truths)
    err = np.linalg.norm(J - Jt, axis=2).mean(1) * 10
    print(f"{n_fr} frames, {num_p}x30: upload+observe {1e3*(t2-t1):.1f} ms, track {dt*1e3:.1f} ms -> {n_fr/dt:.0f} frames/s "
          f"({n_fr*num_p*30/dt:.3e} evals/s); median cost {np.median(costs):.3f}, mean joint error vs truth {np.median(err):.2f} mm")
