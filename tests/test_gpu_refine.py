"""GPU: PSO::refine_init_pose (src/PSO.cpp:216-266) -- SURVEY 8(f) rank 1 -- against the oracle, against
the fixtures the reference's own code wrote (tests/golden/golden_ref_v1.npz), and the complete
per-frame loop of test_full (refine_init_pose -> pso_evolve -> x0 = bestp, src/testmodel.cpp:117-138).

The refinement differentiates the energy numerically with eps = 1e-5, so a relative difference of
1e-13 between two evaluations (summation order) already moves a gradient by ~1e-7 relative: the GPU
path agrees with the oracle to ~1e-6 in the refined pose, not bit for bit; the line-search decisions
(accept / grow / shrink) coincide, which the evaluation counts check."""
import os

import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
# Two evaluations of the same pose differ by ~1e-13 relative between the GPU and the oracle (summation order); the central
# differences (eps = 1e-5 on values of order 30) turn that into ~1e-8 relative in g.p, i.e. ~1e-8 relative in the Armijo /
# Goldstein thresholds.  A line-search comparison may therefore go the other way only if its two sides are closer than that.
MARGIN = 1e-7


def test_refine_matches_oracle(hpe, orc, omodel, D):
    rng = np.random.default_rng(21)
    n = 24
    truths = np.stack([synth_pose(D, rng) for _ in range(n)])
    frames = np.stack([orc.render_depth(omodel, t) for t in truths])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=n)
    obs.upload(frames)
    cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
    spread = np.array([0.05, 0.1, 0.3, 0.6])[np.arange(n) % 4]
    x0 = truths + rng.normal(size=(n, 26)) * D.STD * spread[:, None]
    xg, eg = pso.refine_init_pose_batch(x0, cf)
    same_path, worst_joint = 0, 0.0
    for k in range(n):
        oo = orc.Observation(frames[k])
        xo, eo, margin = orc.refine_init_pose(omodel, oo, x0[k], want_margin=True)
        assert np.array_equal(xg[k, 6:], x0[k, 6:])                   # only rotation and position move
        if eg[k] != eo:
            # a different search path is accepted ONLY where the oracle itself met a near-tie: some Armijo / Goldstein
            # comparison of this run had its two sides within MARGIN relative (the oracle records the smallest gap)
            assert margin < MARGIN, (k, eg[k], eo, margin)
        if eg[k] == eo:
            same_path += 1
            assert np.abs(xg[k] - xo).max() < 1e-5, (k, np.abs(xg[k] - xo).max())
            _, Jg = orc.fk(omodel, xg[k])
            _, Jo = orc.fk(omodel, xo)
            worst_joint = max(worst_joint, np.linalg.norm(Jg - Jo, axis=1).max() * 10.0)
        # whatever the path, the refined pose must not be worse than the start (Armijo condition)
        c0 = orc.cal_cost2(omodel, oo, x0[k])[0]
        c1 = orc.cal_cost2(omodel, oo, xg[k])[0]
        assert c1 <= c0 * (1 + 1e-9), (k, c0, c1)
    assert same_path >= n // 2, same_path           # flips need a near-tie (asserted above) and stay the exception
    assert worst_joint < 1e-3, worst_joint          # mm; north star: 1 mm
    # single-pose entry point = batch entry
    assert np.array_equal(pso.refine_init_pose(x0[3], cf, slot=3), xg[3])


def test_refine_reproduces_reference_fixtures(hpe, orc, omodel, D):
    g = np.load(os.path.join(HERE, "golden", "golden_ref_v1.npz"))
    depth = orc.render_depth(omodel, g["truth"])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
    for k in range(len(g["refine_x0"])):
        x = pso.refine_init_pose(g["refine_x0"][k], cf)
        assert np.abs(x - g["refine_out"][k]).max() < 1e-5, (k, np.abs(x - g["refine_out"][k]).max())
        a, d, c = cf.cal_cost_batch(x[None], with_collision=True)
        assert abs(a[0] + d[0] + c[0] - g["refine_cost"][k]) <= 1e-6 * g["refine_cost"][k]


def test_track_sequence_refined_matches_oracle(hpe, orc, omodel, D):
    rng = np.random.default_rng(22)
    n_fr, num_p, maxiter = 4, 32, 10
    pose = synth_pose(D, rng)
    truths = []
    for _ in range(n_fr):
        pose = np.clip(pose + rng.normal(size=26) * np.r_[[1.0] * 3, [0.1] * 3, [1.5] * 20], D.LB, D.UB)
        truths.append(pose.copy())
    frames = np.stack([orc.render_depth(omodel, t) for t in truths])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=n_fr)
    obs.upload(frames)
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    x0 = truths[0] + rng.normal(size=26) * D.STD * 0.2
    cf = hpe.costfunc(hand, obs)
    poses, costs, refined = pso.track_sequence_refined(cf, x0, num_p, n_fr, seed=1000)
    # the GPU loop equals its own pieces called one by one ...
    cur = x0
    for k in range(n_fr):
        r = pso.refine_init_pose(cur, cf, slot=k)
        assert np.array_equal(r, refined[k])
        b = pso.pso_evolve(cf, r, num_p, seed=1000, slot=k)
        assert np.array_equal(b, poses[k])
        cur = b
    # ... and the oracle's loop, to the 1 mm joint criterion of the north star
    prm = orc.pso_params(maxiter=maxiter)
    cur = x0
    for k in range(n_fr):
        oo = orc.Observation(frames[k])
        r, _ = orc.refine_init_pose(omodel, oo, cur)
        bo, co, _, _, _ = orc.pso_evolve(omodel, oo, prm, r, num_p, 1000)
        _, Jg = orc.fk(omodel, poses[k])
        _, Jo = orc.fk(omodel, bo)
        assert np.linalg.norm(Jg - Jo, axis=1).max() * 10.0 < 1.0, k
        cur = bo


def test_gnd_truth_err_matches_reference_fixture_and_oracle(hpe, orc, omodel, D):
    """costfunc::gnd_truth_err (src/costfunc.cpp:476-507), SURVEY 8(f) rank 2"""
    g = np.load(os.path.join(HERE, "golden", "golden_ref_v1.npz"))
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    cf = hpe.costfunc(hand, obs)
    hand.build_hand_model(g["truth"])
    for f in range(len(g["gt_rows"])):
        e = cf.gnd_truth_err(g["gt_rows"], f)
        assert abs(e - g["gt_err"][f]) <= 1e-12 * g["gt_err"][f]
    rng = np.random.default_rng(41)
    th = D.X0 + rng.normal(size=(500, 26)) * D.STD
    gt = rng.normal(size=(500, 63)) * 50
    e = cf.gnd_truth_err_batch(th, gt)
    _, J = orc.fk(omodel, th)
    eo = np.array([orc.gnd_truth_err(J[i], gt[i]) for i in range(500)])
    assert np.allclose(e, eo, rtol=1e-12)
    # a pose scored against its own joints (mm, y/z negated) has zero error
    own = (J[:4] * 10.0 * np.array([1.0, -1.0, -1.0])).reshape(4, 63)
    assert np.abs(cf.gnd_truth_err_batch(th[:4], own)).max() < 1e-9
