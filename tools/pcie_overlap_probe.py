"""H2D copy rate with and without the cost kernel running concurrently on another stream."""
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
cf = hpe_b200.costfunc(hand, obs)
th = torch.from_numpy(truth + np.random.default_rng(2).normal(size=(N, 26)) * D.STD)
th_pin = th.pin_memory()
th_dev = th.cuda()
cost_dev = torch.empty(N, dtype=torch.float64, device="cuda")
dst = torch.empty_like(th_dev)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
def copies(nchunk, reps=20):
    per = N // nchunk
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sB):
        e0.record()
        for _ in range(reps):
            for c in range(nchunk):
                dst[c * per:(c + 1) * per].copy_(th_pin[c * per:(c + 1) * per], non_blocking=True)
        e1.record()
    return e0, e1, reps
for nchunk in (1, 4, 8):
    torch.cuda.synchronize()
    e0, e1, reps = copies(nchunk)
    torch.cuda.synchronize()
    alone = e0.elapsed_time(e1) / reps
    torch.cuda.synchronize()
    with torch.cuda.stream(sA):
        for _ in range(40):
            cf.cal_cost_dev(th_dev.data_ptr(), N, cost_dev.data_ptr(), stream=sA.cuda_stream)
    e0, e1, reps = copies(nchunk)
    torch.cuda.synchronize()
    busy = e0.elapsed_time(e1) / reps
    print(f"13.6 MB in {nchunk} copies: alone {alone:.3f} ms ({13.63 / alone:.1f} GB/s)   under the cost kernel {busy:.3f} ms ({13.63 / busy:.1f} GB/s)")
