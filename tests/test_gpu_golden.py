"""GPU: the CUDA path against the committed golden file (tests/golden/golden_v1.npz, produced by
tests/golden/make_golden.py from the CPU oracle), and size-independent properties at the full
BASELINE sizes where the oracle would take too long."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


def golden_frame(g):
    fr = np.zeros((240, 320), dtype=np.float32)
    fr[g["frame_y"], g["frame_x"]] = g["frame_mm"]
    return fr


def test_cuda_path_reproduces_golden(hpe, golden, D):
    g = golden
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(golden_frame(g)[None])
    o = obs.get(0)
    assert o["n_hand"] == int(g["n_hand"]) and np.array_equal(o["cloud"], g["cloud"])
    assert abs(o["scale"] - float(g["scale"])) <= 1e-13 * float(g["scale"])
    assert o["dt_max"] == np.float32(g["dt_max"]) and np.array_equal(o["dt"][::16], g["dt_rows"])
    assert abs(o["dt"].astype(np.float64).sum() - float(g["dt_sum"])) < 1e-6
    S, J = hand.build_hand_models(g["poses"])
    assert np.abs(S - g["spheres"]).max() < 1e-11 and np.abs(J - g["joints"]).max() < 1e-11
    cf = hpe.costfunc(hand, obs)
    a, d, c = cf.cal_cost_batch(g["poses"], with_collision=True)
    assert np.allclose(a, g["align"], rtol=1e-9) and np.allclose(d, g["depth_term"], rtol=1e-9, atol=1e-12)
    assert np.allclose(c, g["colli"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(cf.compute_correspondences(g["poses"]), g["match"].astype(np.int32))
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, int(g["pso_maxiter"]), 1e-8, 1e-8)
    bp, best, tr = pso.pso_evolve_batch(cf, g["pso_x0"][None], int(g["pso_num_p"]), seeds=[int(g["pso_seed"])])
    assert np.allclose(tr[0], g["pso_trace"], rtol=1e-8) and np.abs(bp[0] - g["pso_bestp"]).max() < 1e-7


def test_full_size_cost_batch_properties(hpe, orc, omodel, golden, D):
    """configs[1] at its real size (65,536 poses): every pose against the oracle, and properties that
    do not need it: batch-split invariance, permutation equivariance, translation of the cloud."""
    g = golden
    truth = g["truth"]
    rng = np.random.default_rng(2)
    n = 65536
    th = truth + rng.normal(size=(n, 26)) * D.STD
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    fr = golden_frame(g)
    obs.upload(fr[None])
    cf = hpe.costfunc(hand, obs)
    a, d, _ = cf.cal_cost_batch(th)
    assert np.isfinite(a).all() and np.isfinite(d).all() and (a >= 0).all() and (d >= 0).all()
    ao, do, _ = orc.cost_batch(omodel, orc.Observation(fr), th, collision=False)      # all 65,536 poses
    rel = np.abs((a + d) - (ao + do)) / (ao + do)
    assert rel.max() < 1e-9                                    # north star asks for 1e-4
    # a pose's cost does not depend on the batch it is evaluated in, nor on its position in it
    a2, d2, _ = cf.cal_cost_batch(th[1000:1777])
    assert np.array_equal(a2, a[1000:1777]) and np.array_equal(d2, d[1000:1777])
    perm = rng.permutation(n)
    a3, d3, _ = cf.cal_cost_batch(th[perm])
    assert np.array_equal(a3, a[perm]) and np.array_equal(d3, d[perm])
    # the two host pipelines for large batches (kernel chasing the upload / one launch per uploaded chunk)
    # and odd upload chunk sizes return the same bits
    try:
        for mode, chunk in ((0, 0), (1, 0), (1, 5000), (1, 65536), (1, 1031)):
            hpe.set_cost_pipeline(mode, chunk)
            a4, d4, c4 = cf.cal_cost_batch(th[: n - 7], with_collision=True)
            assert np.array_equal(a4, a[: n - 7]) and np.array_equal(d4, d[: n - 7]), (mode, chunk)
    finally:
        hpe.set_cost_pipeline(0, 0)


def test_full_size_swarm_batch_properties(hpe, golden, D):
    """many independent 64 x 30 swarms: a swarm's result does not depend on the batch around it;
    the trace is non-increasing and ends at the reported best cost"""
    g = golden
    rng = np.random.default_rng(6)
    n_fr = 600
    truths = np.clip(D.X0 + rng.normal(size=(n_fr, 26)) * D.STD * 0.5, D.LB, D.UB)
    hand = hpe.handmodel()
    frames = hpe.render_frames(hand, truths)
    obs = hpe.observedmodel(capacity=n_fr)
    obs.upload(frames)
    cf = hpe.costfunc(hand, obs)
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
    x0 = truths + rng.normal(size=truths.shape) * D.STD * 0.3
    bp, best, tr = pso.pso_evolve_batch(cf, x0, 64)
    assert np.all(np.diff(tr, axis=1) <= 0) and np.array_equal(tr[:, -1], best)
    a, d, _ = cf.cal_cost_batch(bp[17:18], slot=17)
    assert abs(a[0] + d[0] - best[17]) <= 1e-12 * best[17]       # gbest_pos scores gbest_cost
    bp2, best2, _ = pso.pso_evolve_batch(cf, x0[100:110], 64, first=100)
    assert np.array_equal(bp2, bp[100:110]) and np.array_equal(best2, best[100:110])
    x0_cost = np.array([cf.cal_cost(x0[k], slot=k) for k in range(0, n_fr, 50)])
    assert np.all(best[::50] < x0_cost)


def test_cuda_path_reproduces_reference_fixtures(hpe, D):
    """tests/golden/golden_ref_v1.npz was written by the REFERENCE'S OWN CODE (oracle/_ref, see
    tests/golden/make_golden_ref.py): forward kinematics, observation, the three energy terms, the
    correspondences and two PSO::pso_evolve runs on the replayed random streams"""
    g = np.load(os.path.join(HERE, "golden", "golden_ref_v1.npz"))
    g1 = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(golden_frame(g1)[None])
    o = obs.get(0)
    assert np.array_equal(o["cloud"], g["cloud"])                                   # bit-exact
    assert abs(o["scale"] - float(g["scale"])) <= 1e-13 * float(g["scale"])
    assert o["dt_max"] == np.float32(g["dt_max"]) and np.array_equal(o["dt"][::16], g["dt_rows"])
    S, J = hand.build_hand_models(g["poses"])
    assert np.abs(S - g["spheres"]).max() < 1e-11 and np.abs(J - g["joints"]).max() < 1e-11
    cf = hpe.costfunc(hand, obs)
    a, d, c = cf.cal_cost_batch(g["poses"], with_collision=True)
    assert np.allclose(a + d, g["cal_cost"], rtol=1e-9) and np.allclose(a + d + c, g["cal_cost2"], rtol=1e-9)
    assert np.allclose(a, g["align"], rtol=1e-9) and np.allclose(d, g["depth_term"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(cf.compute_correspondences(g["poses"]), g["match"].astype(np.int32))
    pso = hpe.PSO()
    for tag in ("pso", "pso64"):
        pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, int(g[tag + "_maxiter"]), 1e-8, 1e-8)
        bp, best, _ = pso.pso_evolve_batch(cf, g["pso_x0"][None], int(g[tag + "_num_p"]), seeds=[int(g[tag + "_seed"])])
        assert np.abs(bp[0] - g[tag + "_bestp"]).max() < 1e-7 and abs(best[0] - float(g[tag + "_best"])) <= 1e-8 * best[0]
        # 1 mm joint criterion of the north star, on the reference's final pose
        _, Jg = hand.build_hand_models(np.stack([bp[0], g[tag + "_bestp"]]))
        assert np.abs(Jg[0] - Jg[1]).max() * 10.0 < 1e-3


def test_cost_kernel_self_scheduling(hpe, golden, D):
    """cost_kernel hands its poses out through counters in device memory (groups, then shrinking units) and the last
    warp re-arms the slot it used: every batch size must give the bits of the one big batch, also after more launches
    than there are slots, from two streams at once, and with NaN poses in the batch."""
    import torch
    g = golden
    rng = np.random.default_rng(11)
    n = 60000                                        # > 12 poses per warp (3,552 warps): groups + units of three + single poses
    th = g["truth"] + rng.normal(size=(n, 26)) * D.STD
    th[777] = np.nan
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(golden_frame(g)[None])
    cf = hpe.costfunc(hand, obs)
    a, d, _ = cf.cal_cost_batch(th)
    ok = np.ones(n, bool); ok[777] = False          # the NaN pose only has to come out the same way every time
    assert np.isfinite(a[ok]).all() and np.isfinite(d[ok]).all()
    sizes = [1, 2, 3, 4, 5, 6, 7, 31, 250, 3551, 3552, 3553, 21311, 21312, 21313, 42623, 42624, 42625, 50001]
    for k, m in enumerate(sizes):
        lo = (97 * k) % (n - m)
        a2, d2, _ = cf.cal_cost_batch(th[lo:lo + m])
        assert np.array_equal(a2, a[lo:lo + m], equal_nan=True) and np.array_equal(d2, d[lo:lo + m], equal_nan=True), m
    for rep in range(300):                           # slot reuse
        m = 1 + (rep * 37) % 200
        a2, d2, _ = cf.cal_cost_batch(th[rep:rep + m])
        assert np.array_equal(a2, a[rep:rep + m], equal_nan=True), rep
    # more launches than the ring has slots (4,096), queued without a host round trip in between
    th_small = torch.from_numpy(th[:64]).cuda()
    out_small = torch.empty(64, dtype=torch.float64, device="cuda")
    for rep in range(4500):
        cf.cal_cost_dev(th_small.data_ptr(), 64, out_small.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(out_small.cpu().numpy(), (a + d)[:64], equal_nan=True)
    # two launches in flight on different streams (different slots)
    th_dev = torch.from_numpy(th).cuda()
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(3):
        for o_, s_ in zip(outs, streams):
            cf.cal_cost_dev(th_dev.data_ptr(), n, o_.data_ptr(), stream=s_.cuda_stream)
    torch.cuda.synchronize()
    want = a + d
    for o_ in outs:
        assert np.array_equal(o_.cpu().numpy(), want, equal_nan=True)
