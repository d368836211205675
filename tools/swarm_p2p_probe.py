"""Sharded 16k swarm: NCCL all-gather vs stepwise peer exchange vs the one-launch kernel (hpe_swarm_run).
Run under torchrun on >= 1 GPUs."""
import os, sys, time
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
res = {}
for mode in ("allgather", "p2p", "run"):
    if mode == "p2p" and world == 1:
        continue
    sw = hpe_b200.Swarm(cf, pso.params, NP, lo, hi, x0, seed=4)
    if mode != "allgather":
        sharding.attach_swarm_peers(sw)
    ts = []
    for rep in range(5):
        barrier()
        e0.record()
        if mode == "allgather":
            out = sharding.run_sharded_swarm(sw, NP)
        elif mode == "p2p":
            out = sharding.run_sharded_swarm_p2p(sw)
        else:
            sw.run()
        e1.record()
        barrier()
        if mode == "run":
            out = sw.best()
        ts.append(sharding.max_over_ranks(e0.elapsed_time(e1), "cuda"))
    res[mode] = (min(ts[1:]), out)
    if rank == 0:
        print(f"{mode:10s} world={world} {NP} x {GEN}: {min(ts[1:]):.3f} ms  best cost {out[1]!r}", flush=True)
if rank == 0:
    a = res["allgather"][1]
    for k, (_, b) in res.items():
        print(k, "identical to allgather:", np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2]))
if world > 1:
    dist.destroy_process_group()
