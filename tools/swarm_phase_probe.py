"""Phase timing of hpe_swarm_run (one-launch sharded swarm) from its globaltimer stamps."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
from hpe_b200 import sharding
D = hpe_b200.defaults
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NP, GEN = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 30
hand = hpe_b200.handmodel()
truth = np.clip(D.X0 + np.random.default_rng(4).normal(size=26) * D.STD * 0.5, D.LB, D.UB)
obs = hpe_b200.observedmodel(capacity=1)
obs.upload(hpe_b200.render_frames(hand, truth[None]))
cf = hpe_b200.costfunc(hand, obs)
pso = hpe_b200.PSO()
pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, GEN, 1e-8, 1e-8)
x0 = truth + np.random.default_rng(44).normal(size=26) * D.STD * 0.3
lo, hi = sharding.shard_range(NP, rank, world)
sw = hpe_b200.Swarm(cf, pso.params, NP, lo, hi, x0, seed=4)
sharding.attach_swarm_peers(sw)
sw.debug_times()
for _ in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    sw.run()
    torch.cuda.synchronize()
t = sw.debug_times().astype(np.int64)
if rank == 0:
    names = ["informant pass", "informant barrier", "step (CTA 0 done)", "step barrier", "post + flag wait", "reduce", "-"]
    d = np.diff(t, axis=1)[3:] / 1e3          # skip the first generations
    print(f"world={world} {NP} particles: mean us per phase over generations 3..{GEN - 1}")
    for n, v in zip(names, d.mean(axis=0)):
        print(f"  {n:20s} {v:8.2f}")
    print(f"  {'generation':20s} {(t[3:, 7] - t[3:, 0]).mean() / 1e3:8.2f}   whole run {(t[-1, 7] - t[0, 0]) / 1e6:.3f} ms")
    bp, bc, tr = sw.best()
    print(f"  best cost {bc!r}")
if world > 1:
    dist.destroy_process_group()
