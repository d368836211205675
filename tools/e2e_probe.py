"""e2e rate of hpe_cost_batch (65,536 host poses per call) for the host pipelines / upload chunk sizes."""
import os, sys, time, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
N = 65536
hand = hpe_b200.handmodel()
rng = np.random.default_rng(20140528)
truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
obs = hpe_b200.observedmodel(capacity=1, num_samples=256)
obs.upload(hpe_b200.render_frames(hand, truth[None]))
sets = 8
th = torch.from_numpy(truth + np.random.default_rng(2).normal(size=(sets, N, 26)) * D.STD).pin_memory()
out = torch.empty(2, N, dtype=torch.float64).pin_memory()
lib = hpe_b200.lib()
def step(i):
    rc = lib.hpe_cost_batch(hand._h, obs._h, 0, C.c_void_p(th[i % sets].data_ptr()), C.c_int64(N),
                            C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), None)
    assert rc == 0, lib.hpe_last_error_string()
for mode, chunk in ((0, 0), (0, 16384), (0, 0), (0, 16384), (1, 0), (1, 16384)):
    hpe_b200.set_cost_pipeline(mode, chunk)
    for i in range(5): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 100
    for i in range(K): step(i)
    dt = (time.perf_counter() - t0) / K
    print(f"mode {mode} chunk {chunk:6d}: {dt * 1e3:.3f} ms/step  {N / dt:.3e} evals/s")
