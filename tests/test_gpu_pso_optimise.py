"""GPU: PSO::pso_optimise (src/PSO.cpp:539-712) -- SURVEY 8(f) rank 3, the reference's hybrid of per-particle
coordinate descent (cal_gradient + goldstein on cal_cost2) and canonical PSO -- against the oracle (which is
bit-identical to the reference build, tests/test_ref_build.py) and against the fixtures the reference's own
code wrote (tests/golden/golden_ref_opt_v1.npz).

Like refine_init_pose this differentiates the energy numerically with eps = 1e-5: a 1e-13 relative
difference between two evaluations (summation order) moves a gradient by ~1e-7 and a position by up to
~1e-4 after ten descent steps (measured on the oracle by perturbing x0 by 1e-13), so the comparison is
"same decisions (equal evaluation counts), positions within 1e-2, personal-best costs within 1e-6 relative in the
median and 5e-3 at worst (first B200 run: 1e-12 ... 1e-8 typical, 1.3e-4 on one high-cost particle)" on nearly all
swarms, not bit for bit."""
import os

import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _setup(hpe, orc, omodel, D, rng, n):
    truths = np.stack([synth_pose(D, rng) for _ in range(n)])
    frames = np.stack([orc.render_depth(omodel, t) for t in truths])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=n)
    obs.upload(frames)
    return truths, frames, hpe.costfunc(hand, obs), hpe.PSO()


@pytest.mark.parametrize("num_p,maxiter", [(8, 2), (11, 4)])
def test_pso_optimise_matches_oracle(hpe, orc, omodel, D, num_p, maxiter):
    rng = np.random.default_rng(50 + num_p)
    n = 6
    truths, frames, cf, pso = _setup(hpe, orc, omodel, D, rng, n)
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    x0 = truths + rng.normal(size=(n, 26)) * D.STD * 0.3
    seeds = 1000 + np.arange(n)
    bg, cg, eg, fxg, fcg = pso.pso_optimise_batch(cf, x0, num_p, seeds=seeds, want_state=True)
    same = 0
    for k in range(n):
        oo = orc.Observation(frames[k])
        bo, co, eo, fxo, fco = orc.pso_optimise(omodel, oo, orc.pso_params(maxiter=maxiter), x0[k], num_p, int(seeds[k]))
        assert cg[k] == fcg[k].min()                       # gbest_cost is the smallest pcost
        if eg[k] != eo:
            # a different number of line-search evaluations is accepted only where the oracle met a near-tie (the smallest
            # relative gap of an Armijo / Goldstein comparison in this run; see tests/test_gpu_refine.py for the bound)
            assert orc.pso_optimise.last_margin < 1e-7, (k, eg[k], eo, orc.pso_optimise.last_margin)
        if eg[k] == eo:
            same += 1
            assert np.abs(fxg[k] - fxo).max() < 1e-2, (k, np.abs(fxg[k] - fxo).max())
            rel = np.abs(fcg[k] / fco - 1)
            assert rel.max() < 5e-3 and np.median(rel) < 1e-6, (k, rel.max(), np.median(rel))
            assert abs(cg[k] / co - 1) < 1e-3 and np.abs(bg[k] - bo).max() < 1e-2
        # whatever the path: pcost never exceeds the cost of the particle's start (pbest only improves)
        start = orc.generate_particles(orc.pso_params(), x0[k], num_p, int(seeds[k]))
        a0, d0, _ = orc.cost_batch(omodel, oo, start)
        assert np.all(fcg[k] <= (a0 + d0) * (1 + 1e-9))
    assert same >= n // 2, same
    # the single-swarm entry point is the batch entry
    assert np.array_equal(pso.pso_optimise(cf, x0[2], num_p, seed=int(seeds[2]), slot=2), bg[2])


def test_pso_optimise_uses_w_c1_c2(hpe, orc, omodel, D):
    """the only optimiser of the reference that reads omega, phip, phig (src/PSO.cpp:664-666)"""
    rng = np.random.default_rng(53)
    truths, frames, cf, pso = _setup(hpe, orc, omodel, D, rng, 1)
    x0 = truths + rng.normal(size=(1, 26)) * D.STD * 0.3
    oo = orc.Observation(frames[0])
    outs = []
    for w, c1, c2 in ((0.7298, 1.49618, 1.49618), (0.4, 0.9, 2.0)):
        pso.set_pso_params(D.UB, D.LB, D.STD, w, c1, c2, 3, 1e-8, 1e-8)
        bg, cg, eg = pso.pso_optimise_batch(cf, x0, 8, seeds=[77])
        bo, co, eo, _, _ = orc.pso_optimise(omodel, oo, orc.pso_params(maxiter=3), x0[0], 8, 77, w=w, c1=c1, c2=c2)
        if eg[0] == eo:
            assert abs(cg[0] / co - 1) < 1e-3 and np.abs(bg[0] - bo).max() < 1e-2
        outs.append(bg[0])
    assert not np.array_equal(outs[0], outs[1])


def test_pso_optimise_reproduces_reference_fixtures(hpe, orc, omodel, D):
    g = np.load(os.path.join(HERE, "golden", "golden_ref_opt_v1.npz"))
    close = 0
    for k, (num_p, maxiter, seed, w, c1, c2) in enumerate(g["cases"]):
        depth = orc.render_depth(omodel, g["truth%d" % k])
        hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
        obs.upload(depth[None])
        cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
        pso.set_pso_params(D.UB, D.LB, D.STD, w, c1, c2, int(maxiter), 1e-8, 1e-8)
        b = pso.pso_optimise(cf, g["x0_%d" % k], int(num_p), seed=int(seed))
        a, d, _ = cf.cal_cost_batch(b[None])
        if np.abs(b - g["bestp%d" % k]).max() < 1e-2 and abs((a[0] + d[0]) / float(g["bestp_cost%d" % k]) - 1) < 1e-3:
            close += 1
    assert close >= len(g["cases"]) - 1, close


def test_pso_optimise_edge_cases(hpe, orc, omodel, D):
    rng = np.random.default_rng(54)
    truths, frames, cf, pso = _setup(hpe, orc, omodel, D, rng, 2)
    x0 = truths + rng.normal(size=(2, 26)) * D.STD * 0.2
    # maxiter = 1: generate_particles + one evaluation, gbest = the first minimum (src/PSO.cpp:555-568)
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 1, 1e-8, 1e-8)
    bg, cg, eg = pso.pso_optimise_batch(cf, x0, 16, seeds=[3, 4])
    for k in range(2):
        oo = orc.Observation(frames[k])
        bo, co, eo, _, _ = orc.pso_optimise(omodel, oo, orc.pso_params(maxiter=1), x0[k], 16, 3 + k)
        assert eg[k] == eo == 16 and abs(cg[k] / co - 1) < 1e-9 and np.abs(bg[k] - bo).max() < 1e-9
    # no swarms: nothing to do; bad arguments are reported, not clamped
    b0, c0, e0 = pso.pso_optimise_batch(cf, np.empty((0, 26)), 8)
    assert b0.shape == (0, 26)
    with pytest.raises(hpe.HpeError):
        pso.pso_optimise_batch(cf, x0, 0)
    with pytest.raises(hpe.HpeError):
        pso.pso_optimise_batch(cf, np.tile(x0, (2, 1)), 8)          # more swarms than frames
    # a slot that holds no observation is reported (HPE_ERR_STATE), not fitted
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=2)
    obs.upload(frames[:1])
    with pytest.raises(hpe.HpeError) as e:
        pso.pso_optimise_batch(hpe.costfunc(hand, obs), x0, 4)
    assert e.value.code == 4


def test_pso_optimise_full_size_properties(hpe, orc, omodel, D):
    """64 particles x 30 iterations (the swarm size of BASELINE's configs), too long for the scalar oracle in a test:
    size-independent properties instead -- counter-based draws make a shorter run a prefix of a longer one, so the
    best cost can only fall with maxiter; a swarm's result does not depend on the batch it runs in; the returned
    cost is the smallest personal best."""
    rng = np.random.default_rng(55)
    truths, frames, cf, pso = _setup(hpe, orc, omodel, D, rng, 3)
    x0 = truths + rng.normal(size=(3, 26)) * D.STD * 0.3
    best = {}
    for maxiter in (5, 30):
        pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
        bg, cg, eg, fx, fc = pso.pso_optimise_batch(cf, x0, 64, seeds=[9, 9, 9], want_state=True)
        assert np.array_equal(cg, fc.min(axis=1)) and np.all(eg > 64 * maxiter)
        best[maxiter] = cg
        if maxiter == 30:
            alone = pso.pso_optimise_batch(cf, x0[1:2], 64, seeds=[9], first=1)
            assert np.array_equal(alone[0][0], bg[1]) and alone[1][0] == cg[1] and alone[2][0] == eg[1]
    assert np.all(best[30] <= best[5])
    # the fit is a real one: far below the cost of the start pose
    for k in range(3):
        assert best[30][k] < 0.2 * orc.cal_cost(omodel, orc.Observation(frames[k]), x0[k])
