"""GPU: the stepwise large-swarm path (hpe_swarm_*) against the one-CTA kernel and the oracle, and
the particle-sharded run (two shards emulated in one process: their tables are exchanged by device
copies exactly where the ranks would all-gather) against the unsharded one."""
import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu


def _run_swarms(hpe, swarms, torch):
    """drive shards that live on the same GPU; exchange = copy every shard's owned rows to the others"""
    def exchange():
        for src in swarms:
            for dst in swarms:
                if dst is src or src.hi == src.lo:
                    continue
                nbytes = (src.hi - src.lo) * src.ROW * 8
                off = src.lo * src.ROW * 8
                # device-to-device copy through torch's raw pointer view
                src_t = _as_tensor(torch, src.table_ptr() + off, nbytes // 8)
                dst_t = _as_tensor(torch, dst.table_ptr() + off, nbytes // 8)
                dst_t.copy_(src_t)
        torch.cuda.synchronize()

    for s in swarms:
        s.init()
    torch.cuda.synchronize()
    exchange()
    for s in swarms:
        s.reduce()
    for _ in range(swarms[0].maxiter - 1):
        for s in swarms:
            s.step()
        torch.cuda.synchronize()
        exchange()
        for s in swarms:
            s.reduce()
    torch.cuda.synchronize()
    return [s.best() for s in swarms]


def _as_tensor(torch, ptr, n):
    """float64 CUDA tensor aliasing n doubles at device pointer ptr (no ownership)"""
    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Arr(), device="cuda")


@pytest.fixture(scope="module")
def setup(hpe, orc, omodel, D):
    rng = np.random.default_rng(41)
    truth = synth_pose(D, rng)
    depth = orc.render_depth(omodel, truth)
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    return truth, depth, hand, obs, x0


def test_swarm_equals_one_cta_kernel_and_oracle(hpe, orc, omodel, D, setup):
    torch = pytest.importorskip("torch")
    truth, depth, hand, obs, x0 = setup
    num_p, maxiter, seed = 64, 12, 77
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    cf = hpe.costfunc(hand, obs)
    bestp, best, trace = pso.pso_evolve_batch(cf, x0[None], num_p, seeds=[seed])
    sw = hpe.Swarm(cf, pso.params, num_p, 0, num_p, x0, seed=seed)
    (bp, bc, tr), = _run_swarms(hpe, [sw], torch)
    assert np.array_equal(bp, bestp[0]) and bc == best[0]          # same arithmetic, bit for bit
    assert np.array_equal(tr, trace[0])
    bo, co, tro, _, _ = orc.pso_evolve(omodel, orc.Observation(depth), orc.pso_params(maxiter=maxiter), x0, num_p, seed)
    assert abs(bc - co) <= 1e-8 * co and np.allclose(tr, tro, rtol=1e-8, atol=0)
    assert np.abs(bp - bo).max() < 1e-7


def test_sharded_swarm_is_bit_identical(hpe, D, setup):
    torch = pytest.importorskip("torch")
    truth, depth, hand, obs, x0 = setup
    num_p, maxiter, seed = 1000, 8, 4
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    cf = hpe.costfunc(hand, obs)
    (ref,) = _run_swarms(hpe, [hpe.Swarm(cf, pso.params, num_p, 0, num_p, x0, seed=seed)], torch)
    for cuts in ([0, 500, 1000], [0, 130, 131, 700, 1000]):
        shards = [hpe.Swarm(cf, pso.params, num_p, a, b, x0, seed=seed) for a, b in zip(cuts[:-1], cuts[1:])]
        res = _run_swarms(hpe, shards, torch)
        for bp, bc, tr in res:
            assert np.array_equal(bp, ref[0]) and bc == ref[1] and np.array_equal(tr, ref[2])
    assert ref[2][-1] <= ref[2][0]


def test_cluster_kernel_is_bit_identical_to_one_cta_kernel(hpe, orc, omodel, D):
    """one swarm per 8-CTA cluster (latency path; swarm state in distributed shared memory = path 2, in global
    memory = path 3) == one swarm per CTA (throughput path = 1), for batches, odd particle counts, a swarm too
    large for the shared-memory variant, and chained sequences"""
    rng = np.random.default_rng(52)
    n_fr = 5
    truths = [synth_pose(D, rng) for _ in range(n_fr)]
    frames = np.stack([orc.render_depth(omodel, t) for t in truths])
    x0 = np.stack([t + rng.normal(size=26) * D.STD * 0.3 for t in truths])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=n_fr)
    obs.upload(frames)
    cf = hpe.costfunc(hand, obs)
    lib = hpe.lib()
    for num_p, maxiter in ((64, 10), (37, 6), (128, 5), (130, 4)):
        pso = hpe.PSO()
        pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
        res = {}
        for path in (1, 2, 3):
            assert lib.hpe_set_pso_path(path) == 0
            try:
                res[path] = (pso.pso_evolve_batch(cf, x0, num_p, seeds=[7, 8, 9, 10, 11]),
                             pso.track_sequence(cf, x0[0], num_p, n_fr, seed=1000))
            finally:
                lib.hpe_set_pso_path(0)
        (b1, c1, t1), (p1, k1) = res[1]
        (b2, c2, t2), (p2, k2) = res[2]
        assert np.array_equal(b1, b2) and np.array_equal(c1, c2) and np.array_equal(t1, t2)
        assert np.array_equal(p1, p2) and np.array_equal(k1, k2)
        (b3, c3, t3), (p3, k3) = res[3]
        assert np.array_equal(b1, b3) and np.array_equal(c1, c3) and np.array_equal(t1, t3)
        assert np.array_equal(p1, p3) and np.array_equal(k1, k3)


def test_one_launch_swarm_is_bit_identical_to_stepwise(hpe, D, setup):
    """hpe_swarm_run: the whole pso_evolve of a large swarm (topology, informants, move + evaluate,
    global argmin) in one cooperative kernel launch == the stepwise launches, bit for bit"""
    torch = pytest.importorskip("torch")
    truth, depth, hand, obs, x0 = setup
    pso = hpe.PSO()
    cf = hpe.costfunc(hand, obs)
    for num_p, maxiter, seed in ((1000, 8, 4), (64, 12, 77), (4099, 5, 9), (7, 3, 1)):
        pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
        (ref,) = _run_swarms(hpe, [hpe.Swarm(cf, pso.params, num_p, 0, num_p, x0, seed=seed)], torch)
        sw = hpe.Swarm(cf, pso.params, num_p, 0, num_p, x0, seed=seed)
        sw.run()
        bp, bc, tr = sw.best()
        assert np.array_equal(bp, ref[0]) and bc == ref[1] and np.array_equal(tr, ref[2]), (num_p, maxiter)
        sw.run()                                       # a second run on the same handle starts afresh
        bp2, bc2, tr2 = sw.best()
        assert np.array_equal(bp2, ref[0]) and bc2 == ref[1] and np.array_equal(tr2, ref[2])
