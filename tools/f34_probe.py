"""pso_optimise and bincomp_penalty on synthetic frames (for ncu / timing): N frames x P particles x I iterations
of PSO::pso_optimise, then costfunc::bincomp_penalty of M poses against one frame."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n_fr = int(sys.argv[1]) if len(sys.argv) > 1 else 296
num_p = int(sys.argv[2]) if len(sys.argv) > 2 else 16
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n_bc = int(sys.argv[4]) if len(sys.argv) > 4 else 65536
rng = np.random.default_rng(5)
truths = np.clip(D.X0 + rng.normal(size=(n_fr, 26)) * D.STD * 0.5, D.LB, D.UB)
hand = hpe_b200.handmodel()
frames = hpe_b200.render_frames(hand, truths)
obs = hpe_b200.observedmodel(capacity=n_fr)
obs.upload(frames)
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
x0 = truths + rng.normal(size=truths.shape) * D.STD * 0.3
cf = hpe_b200.costfunc(hand, obs)
pso.pso_optimise_batch(cf, x0[:8], num_p)
for rep in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter()
    bestp, best, ev = pso.pso_optimise_batch(cf, x0, num_p)
    dt = time.perf_counter() - t0
    print(f"pso_optimise {n_fr} frames x {num_p} x {maxiter}: {dt*1e3:.2f} ms  {n_fr/dt:.0f} frames/s  "
          f"{ev.sum()/dt:.3e} evals/s ({ev.mean():.0f} per frame)  median cost {np.median(best):.4f}")
th = truths[0] + rng.normal(size=(n_bc, 26)) * D.STD
cf.bincomp_penalty_batch(th[:64])
for rep in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter()
    pen = cf.bincomp_penalty_batch(th)
    dt = time.perf_counter() - t0
    print(f"bincomp_penalty {n_bc} poses: {dt*1e3:.2f} ms  {n_bc/dt:.3e} poses/s (host poses in, penalties out)  median {np.median(pen):.0f} px")
