"""Small PSO run (for ncu / timing): N frames x 64 particles x 30 generations on synthetic frames."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
n_fr = int(sys.argv[1]) if len(sys.argv) > 1 else 296
num_p = int(sys.argv[2]) if len(sys.argv) > 2 else 64
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 30
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
pso.pso_evolve_batch(cf, x0[:8], num_p, want_trace=False)
for rep in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter()
    bestp, best, _ = pso.pso_evolve_batch(cf, x0, num_p, want_trace=False)
    dt = time.perf_counter() - t0
    print(f"{n_fr} frames x {num_p} x {maxiter}: {dt*1e3:.2f} ms  {n_fr/dt:.0f} frames/s  {n_fr*num_p*maxiter/dt:.3e} evals/s  median cost {np.median(best):.4f}")
