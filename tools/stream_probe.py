"""hpe_cost_batch_submit / _wait: end-to-end evaluations/s for chunk sizes x batches in flight (pinned host buffers)."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
N, ROT, STEPS = 65536, 16, int(os.environ.get("STEPS", "300"))
rng = np.random.default_rng(2)
truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
hand = hpe_b200.handmodel()
obs = hpe_b200.observedmodel(capacity=1, num_samples=256)
obs.upload(hpe_b200.render_frames(hand, truth[None]))
th = torch.from_numpy(truth + rng.normal(size=(ROT, N, 26)) * D.STD).pin_memory()
ring = [torch.empty(3, N, dtype=torch.float64).pin_memory() for _ in range(3)]
lib = hpe_b200.lib()
set_bytes = N * 26 * 8


def run(steps, depth):
    pend = []
    for i in range(steps):
        job, buf = C.c_int(-1), ring[i % 3]
        rc = lib.hpe_cost_batch_submit(hand._h, obs._h, 0, C.c_void_p(th.data_ptr() + (i % ROT) * set_bytes), C.c_int64(N),
                                       C.c_void_p(buf[0].data_ptr()), C.c_void_p(buf[1].data_ptr()), None, C.byref(job))
        assert rc == 0, lib.hpe_last_error_string()
        pend.append(job.value)
        if len(pend) == depth:
            assert lib.hpe_cost_batch_wait(pend.pop(0)) == 0
    for j in pend:
        assert lib.hpe_cost_batch_wait(j) == 0


for chunk in (0, 8192, 32768, 65536):
    hpe_b200.set_cost_pipeline(0, chunk)
    for depth in (1, 2, 3):
        run(10, depth)
        t0 = time.perf_counter()
        run(STEPS, depth)
        dt = time.perf_counter() - t0
        print(f"chunk {chunk or 16384:6d} in flight {depth}: {N * STEPS / dt:.4g} evals/s  {dt / STEPS * 1e3:.4f} ms/step  "
              f"{set_bytes * STEPS / dt / 1e9:.1f} GB/s up")
