"""GPU: the peer-memory paths of the sharded swarm across REAL processes (one process per rank, CUDA IPC mappings,
device-side flags): the 2-rank results of hpe_swarm_run and of the stepwise peer path equal the 1-rank result bit for
bit, and a peer that never delivers yields HPE_ERR_STATE instead of a hang or a silently stale minimum.
Ranks use different GPUs when the box has them; on a one-GPU box both processes share the device (time-sliced)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NUM_P, MAXITER, SEED = 512, 8, 4


def _setup(rank, world):
    sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
    import torch
    import hpe_b200
    from hpe_b200 import sharding
    D = hpe_b200.defaults
    torch.cuda.set_device(rank % torch.cuda.device_count())
    rng = np.random.default_rng(4)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    hand, obs = hpe_b200.handmodel(), hpe_b200.observedmodel(capacity=1)
    obs.upload(hpe_b200.render_frames(hand, truth[None]))
    cf, pso = hpe_b200.costfunc(hand, obs), hpe_b200.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, MAXITER, 1e-8, 1e-8)
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    lo, hi = sharding.shard_range(NUM_P, rank, world)
    return hpe_b200, torch, cf, pso, x0, lo, hi


def _worker(rank, world, mode, conns, out):
    try:
        hpe_b200, torch, cf, pso, x0, lo, hi = _setup(rank, world)
        from hpe_b200 import sharding
        sw = hpe_b200.Swarm(cf, pso.params, NUM_P, lo, hi, x0, seed=SEED)
        # exchange the IPC handles over pipes (any host transport will do)
        mine = sw.ipc_handle()
        for c in conns:
            c.send((rank, mine))
        handles = [None] * world
        handles[rank] = mine
        for c in conns:
            r, h = c.recv()
            handles[r] = h
        sw.attach_peers(handles, rank)
        for c in conns:
            c.send("attached")
        for c in conns:
            c.recv()
        if mode == "lost" and rank == 1:
            for c in conns:          # never runs: wait until rank 0 has given up
                c.recv()
            out.put((rank, "idle"))
            return
        if mode == "lost":
            hpe_b200.lib().hpe_swarm_set_wait_limit(1 << 14)
        if mode in ("run", "lost"):
            sw.run()
            torch.cuda.synchronize()
            try:
                res = sw.best()
                out.put((rank, ("ok", res[0], res[1], res[2])))
            except hpe_b200.HpeError as e:
                out.put((rank, ("err", e.code, str(e))))
            if mode == "lost":
                for c in conns:
                    c.send("done")
        else:
            res = sharding.run_sharded_swarm_p2p(sw)
            torch.cuda.synchronize()
            out.put((rank, ("ok", res[0], res[1], res[2])))
            for c in conns:          # keep the tables mapped until every rank has read its result
                c.send("done")
            for c in conns:
                c.recv()
    except Exception as e:           # pragma: no cover
        out.put((rank, ("exc", repr(e))))


def _launch(world, mode):
    ctx = mp.get_context("spawn")
    pipes = {(a, b): ctx.Pipe() for a in range(world) for b in range(a + 1, world)}
    out = ctx.Queue()
    procs = []
    for r in range(world):
        conns = [pipes[(min(r, o), max(r, o))][0 if r < o else 1] for o in range(world) if o != r]
        p = ctx.Process(target=_worker, args=(r, world, mode, conns, out))
        p.start()
        procs.append(p)
    res = {}
    for _ in range(world):
        r, v = out.get(timeout=300)
        res[r] = v
    for p in procs:
        p.join(timeout=60)
    return res


def _single_rank_reference():
    hpe_b200, torch, cf, pso, x0, lo, hi = _setup(0, 1)
    sw = hpe_b200.Swarm(cf, pso.params, NUM_P, 0, NUM_P, x0, seed=SEED)
    sw.run()
    torch.cuda.synchronize()
    return sw.best()


@pytest.mark.parametrize("mode", ["run", "stepwise"])
def test_two_ranks_equal_one_rank(hpe, mode):
    pytest.importorskip("torch")
    ref = _single_rank_reference()
    res = _launch(2, mode)
    for r in (0, 1):
        assert res[r][0] == "ok", res[r]
        _, bp, bc, tr = res[r]
        assert np.array_equal(bp, ref[0]) and bc == ref[1] and np.array_equal(tr, ref[2]), (mode, r)


def test_lost_peer_is_reported(hpe):
    pytest.importorskip("torch")
    res = _launch(2, "lost")
    assert res[1] == "idle"
    assert res[0][0] == "err" and res[0][1] == 4, res[0]       # HPE_ERR_STATE
    assert "peer" in res[0][2]
