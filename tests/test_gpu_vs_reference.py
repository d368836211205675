"""GPU: the CUDA path against THE REFERENCE'S OWN CODE, live.  oracle/_ref/libhpe_ref.so (the reference's
src/*.cpp compiled unmodified against the stand-in Armadillo/OpenCV headers, built where /root/reference
exists) travels to the GPU box with the snapshot; these tests call it directly -- no oracle restatement in
between -- and skip when the library is absent (the committed fixtures it wrote are checked elsewhere)."""
import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R(orc):
    import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    r = ref.Reference(threads=1)
    yield r
    r.close()


def test_costs_match_the_reference_code(hpe, orc, omodel, R, D):
    """north star: per-particle cost values match the reference CostFunc within 1e-4 relative (here 1e-9),
    on poses near the truth, at PSO-initialisation spread and far out of the image"""
    rng = np.random.default_rng(71)
    truth = synth_pose(D, rng)
    depth = orc.render_depth(omodel, truth)
    R.load_depth(depth)
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    o = obs.get(0)
    cloud, dcm, dt = R.observation()
    assert np.array_equal(o["cloud"], cloud) and np.array_equal(o["dt"].astype(np.float64), dt)
    assert abs(o["scale"] - R.scale) <= 1e-13 * R.scale
    th = np.concatenate([truth[None], truth + rng.normal(size=(600, 26)) * D.STD,
                         truth + rng.normal(size=(200, 26)) * D.STD * 0.1,
                         truth + rng.normal(size=(200, 26)) * D.STD * 4.0])
    cf = hpe.costfunc(hand, obs)
    a, d, c = cf.cal_cost_batch(th, with_collision=True)
    m = cf.compute_correspondences(th)
    S, J = hand.build_hand_models(th)
    worst = 0.0
    for i in range(len(th)):
        t, mr = R.cost_terms(th[i])
        assert np.array_equal(m[i], mr), i                               # cv::BFMatcher's indices, bit-exact
        tot = t[0] + t[1]
        worst = max(worst, abs(a[i] + d[i] - tot) / tot, abs(c[i] - t[2]) / max(t[2], 1.0))
        if i % 50 == 0:
            Sr, Jr = R.fk(th[i])
            assert np.abs(S[i] - Sr).max() < 1e-11 and np.abs(J[i] - Jr).max() < 1e-11
    assert worst < 1e-9, worst


@pytest.mark.parametrize("num_p,maxiter", [(64, 30), (32, 12)])
def test_pso_evolve_matches_the_reference_code(hpe, orc, omodel, R, D, num_p, maxiter):
    """north star: joint positions from tracking driven by the replayed RNG match the reference within 1 mm.
    PSO::pso_evolve itself (src/PSO.cpp:717-886) runs on the counter-based streams the GPU draws."""
    import ref
    rng = np.random.default_rng(72 + num_p)
    hand = hpe.handmodel()
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    R.set_pso_params(maxiter=maxiter)
    pose = synth_pose(D, rng)
    x_gpu = x_ref = pose + rng.normal(size=26) * D.STD * 0.3
    for frame in range(3):                                   # a short tracked sequence, temporal init
        pose = np.clip(pose + rng.normal(size=26) * np.r_[[1.0] * 3, [0.1] * 3, [1.5] * 20], D.LB, D.UB)
        depth = orc.render_depth(omodel, pose)
        R.load_depth(depth)
        obs = hpe.observedmodel(capacity=1)
        obs.upload(depth[None])
        cf = hpe.costfunc(hand, obs)
        hook = ref.PhiloxReplay(1000)
        hook.num_p = num_p
        hook.install()
        try:
            b_ref = R.pso_evolve(x_ref, num_p)
        finally:
            ref.PhiloxReplay.uninstall()
        b_gpu = pso.pso_evolve(cf, x_gpu, num_p, seed=1000)
        _, Jg = hand.build_hand_models(np.stack([b_gpu, b_ref]))
        err_mm = np.linalg.norm(Jg[0] - Jg[1], axis=1).max() * 10.0
        assert err_mm < 1.0, (frame, err_mm)                 # the north star's criterion
        assert err_mm < 1e-5, (frame, err_mm)                # in fact the trajectories coincide
        assert abs(cf.cal_cost(b_gpu) - R.cal_cost(b_ref)) <= 1e-8 * R.cal_cost(b_ref)
        x_gpu, x_ref = b_gpu, b_ref


def test_refine_init_pose_matches_the_reference_code(hpe, orc, omodel, R, D):
    rng = np.random.default_rng(73)
    hand, pso = hpe.handmodel(), hpe.PSO()
    agree = 0
    for k in range(8):
        truth = synth_pose(D, rng)
        depth = orc.render_depth(omodel, truth)
        R.load_depth(depth)
        obs = hpe.observedmodel(capacity=1)
        obs.upload(depth[None])
        cf = hpe.costfunc(hand, obs)
        x0 = truth + rng.normal(size=26) * D.STD * [0.05, 0.1, 0.3, 0.6][k % 4]
        xr = R.refine_init_pose(x0)
        xg = pso.refine_init_pose(x0, cf)
        if np.abs(xg - xr).max() < 1e-5:
            agree += 1
        assert R.cal_cost2(xg)[0] <= R.cal_cost2(x0)[0] * (1 + 1e-9)
    assert agree >= 7, agree


def test_depthmatch_matches_the_reference_code(hpe, orc, omodel, R, D):
    """costfunc::depthMatch_penalty (src/costfunc.cpp:439-473), the reference's own function called live: same pixels
    drawn, background untouched, drawn values (the spheres' depths) within FK accuracy"""
    rng = np.random.default_rng(73)
    truth = synth_pose(D, rng)
    depth = orc.render_depth(omodel, truth)
    R.load_depth(depth)
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    th = np.concatenate([truth[None], truth + rng.normal(size=(20, 26)) * D.STD * 0.2, truth + rng.normal(size=(20, 26)) * D.STD * 2.0])
    img = cf.depthMatch_penalty_batch(th)
    _, dcm, _ = R.observation()
    for i in range(len(th)):
        want = R.depthmatch_penalty(th[i])
        hit = want != dcm
        assert np.array_equal(img[i] != dcm, hit), i
        assert np.array_equal(img[i][~hit], dcm[~hit]) and np.allclose(img[i][hit], want[hit], rtol=1e-12, atol=0), i
