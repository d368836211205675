"""pso_optimise throughput probe: 1,184 swarms of 16 particles x 4 iterations (bench.py's side measurement)"""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, 'hand-pose-estimation_b200')
import hpe_b200
D = hpe_b200.defaults
n_fr = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
rng = np.random.default_rng(7)
truths = np.clip(D.X0 + rng.normal(size=(n_fr, 26)) * D.STD * 0.5, D.LB, D.UB)
hand = hpe_b200.handmodel()
obs = hpe_b200.observedmodel(capacity=n_fr)
obs.upload(hpe_b200.render_frames(hand, truths))
cf = hpe_b200.costfunc(hand, obs)
x0 = truths + rng.normal(size=truths.shape) * D.STD * 0.3
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 4, 1e-8, 1e-8)
pso.pso_optimise_batch(cf, x0[:64], 16)
cnt = C.c_ulonglong()
lib = hpe_b200.lib()
lib.hpe_debug_fallback_points(C.byref(cnt), 1)
t0 = time.perf_counter()
_, bc, ev = pso.pso_optimise_batch(cf, x0, 16)
dt = time.perf_counter() - t0
lib.hpe_debug_fallback_points(C.byref(cnt), 1)
print(f"{n_fr} frames in {dt*1e3:.1f} ms = {n_fr/dt:.0f} frames/s, {ev.sum()/dt:.3g} evals/s, fallback points {cnt.value} ({cnt.value/max(ev.sum(),1):.3f} per eval)")
