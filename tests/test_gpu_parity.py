"""GPU parity: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs.  Tolerances: match indices / distance transform / point selection are
bit-exact; FP64 quantities agree to 1e-9 relative (the north-star bar is 1e-4)."""
import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_fk_matches_oracle(hpe, orc, omodel, D):
    rng = np.random.default_rng(11)
    th = D.X0 + rng.normal(size=(2048, 26)) * D.STD
    th[0] = D.X0
    hand = hpe.handmodel()
    S, J = hand.build_hand_models(th)
    So, Jo = orc.fk(omodel, th)
    assert np.abs(S - So).max() < 1e-11
    assert np.abs(J - Jo).max() < 1e-11
    # reference-style single call records hand_joints
    s0 = hand.build_hand_model(D.X0)
    assert np.abs(s0 - So[0]).max() < 1e-11
    assert np.abs(hand.hand_joints - Jo[0]).max() < 1e-11


def test_observation_matches_oracle(hpe, orc, omodel, D):
    rng = np.random.default_rng(12)
    frames = np.stack([orc.render_depth(omodel, synth_pose(D, rng)) for _ in range(6)])
    obs = hpe.observedmodel(capacity=6)
    obs.upload(frames)
    for k in range(6):
        o = orc.Observation(frames[k])
        g = obs.get(k)
        assert g["n_hand"] == o.n_hand
        assert np.array_equal(g["cloud"], o.cloud)             # same pixels, same IEEE ops
        assert np.array_equal(g["dt"].astype(np.float64), o.dt)  # chamfer DT bit-exact
        assert g["dt_max"] == np.float32(o.dt_max)
        assert abs(g["scale"] - o.scale) <= 1e-13 * o.scale


def test_observation_few_pixels_and_empty(hpe, orc, D):
    obs = hpe.observedmodel(capacity=2)
    fr = np.zeros((2, D.H, D.W), dtype=np.float32)
    fr[0, 100:103, 150:160] = 300.0          # 30 hand pixels < 250 samples: every sample is pixel 0
    fr[1, 5, 7] = 250.0
    obs.upload(fr)
    for k in range(2):
        o = orc.Observation(fr[k])
        g = obs.get(k)
        assert np.array_equal(g["cloud"], o.cloud)
        assert np.array_equal(g["dt"].astype(np.float64), o.dt)
    with pytest.raises(hpe.HpeError) as e:
        obs.upload(np.zeros((1, D.H, D.W), dtype=np.float32))
    assert e.value.code == 3


def test_cost_terms_match_oracle(hpe, orc, omodel, frame0, D):
    truth, depth, oobs = frame0
    rng = np.random.default_rng(2)
    n = 6000
    th = truth + rng.normal(size=(n, 26)) * D.STD
    th[:500] = truth + rng.normal(size=(500, 26)) * D.STD * 0.1   # near the optimum
    th[0] = truth
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    a, d, c = cf.cal_cost_batch(th, with_collision=True)
    ao, do, co = orc.cost_batch(omodel, oobs, th)
    assert _rel(a, ao).max() < RTOL
    assert np.all(np.abs(d - do) <= RTOL * np.maximum(np.abs(do), 1e-6))
    assert np.all(np.abs(c - co) <= RTOL * np.maximum(np.abs(co), 1e-6))
    assert _rel(a + d, ao + do).max() < RTOL
    # match indices are bit-exact
    m = cf.compute_correspondences(th[:300])
    for i in range(300):
        S, _ = orc.fk(omodel, th[i])
        assert np.array_equal(m[i], orc.correspondences(oobs.cloud, S))
    # reference-style scalar calls
    assert abs(cf.cal_cost(th[3]) - (ao[3] + do[3])) <= RTOL * (ao[3] + do[3])
    assert abs(cf.cal_cost2(th[3]) - (ao[3] + do[3] + co[3])) <= RTOL * (ao[3] + do[3] + co[3])
    # frozen correspondences (cal_cost2 with compute_corr=false)
    tot, terms, _ = orc.cal_cost2(omodel, oobs, th[5], match=m[7])
    af, df, cfz = cf.cal_cost_frozen_batch(th[5:6], m[7])
    assert abs(af[0] - terms[0]) <= RTOL * terms[0]
    assert abs(af[0] + df[0] + cfz[0] - tot) <= RTOL * tot


def test_cost_with_injected_observation(hpe, orc, omodel, frame0, D):
    """config 2 style: the oracle's preprocessed observation is injected, so only FK+energy differ"""
    truth, depth, oobs = frame0
    rng = np.random.default_rng(3)
    th = truth + rng.normal(size=(1000, 26)) * D.STD
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.set(0, depth, oobs.dt.astype(np.float32), oobs.cloud, oobs.scale)
    a, d, _ = hpe.costfunc(hand, obs).cal_cost_batch(th)
    ao, do, _ = orc.cost_batch(omodel, oobs, th)
    assert _rel(a + d, ao + do).max() < RTOL


def test_out_of_image_and_degenerate_poses(hpe, orc, omodel, frame0, D):
    truth, depth, oobs = frame0
    th = np.tile(truth, (4, 1))
    th[0, 3] = 90.0      # far right: most spheres project outside the image
    th[1, 5] = 3.0       # very close to the camera
    th[2, 4] = -95.0
    th[3, 0:3] = [170.0, -170.0, 90.0]
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    a, d, c = hpe.costfunc(hand, obs).cal_cost_batch(th, with_collision=True)
    ao, do, co = orc.cost_batch(omodel, oobs, th)
    assert _rel(a + d + c, ao + do + co).max() < RTOL


def test_pso_evolve_matches_oracle(hpe, orc, omodel, D):
    rng = np.random.default_rng(5)
    n_fr, num_p, maxiter = 5, 64, 30
    truths = [synth_pose(D, rng) for _ in range(n_fr)]
    frames = np.stack([orc.render_depth(omodel, t) for t in truths])
    x0 = np.stack([t + rng.normal(size=26) * D.STD * 0.3 for t in truths])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=n_fr)
    obs.upload(frames)
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    seeds = [1000 + k for k in range(n_fr)]
    bestp, best, trace = pso.pso_evolve_batch(hpe.costfunc(hand, obs), x0, num_p, seeds=seeds)
    prm = orc.pso_params(maxiter=maxiter)
    joints_err = []
    for k in range(n_fr):
        bo, co, tro, _, _ = orc.pso_evolve(omodel, orc.Observation(frames[k]), prm, x0[k], num_p, seeds[k])
        assert np.allclose(trace[k], tro, rtol=1e-8, atol=0), (k, trace[k], tro)
        assert abs(best[k] - co) <= 1e-8 * co
        _, Jg = orc.fk(omodel, bestp[k])
        _, Jo = orc.fk(omodel, bo)
        joints_err.append(np.linalg.norm(Jg - Jo, axis=1).max() * 10.0)   # mm
    assert max(joints_err) < 1.0, joints_err        # north star: joints within 1 mm
    assert max(joints_err) < 1e-6, joints_err       # in fact the trajectories coincide


def test_track_sequence_matches_oracle(hpe, orc, omodel, D):
    rng = np.random.default_rng(6)
    n_fr, num_p, maxiter = 4, 32, 12
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
    poses, costs = pso.track_sequence(hpe.costfunc(hand, obs), x0, num_p, n_fr, seed=1000)
    prm = orc.pso_params(maxiter=maxiter)
    cur = x0
    for k in range(n_fr):
        bo, co, _, _, _ = orc.pso_evolve(omodel, orc.Observation(frames[k]), prm, cur, num_p, 1000)
        assert abs(costs[k] - co) <= 1e-8 * co, k
        assert np.abs(poses[k] - bo).max() < 1e-7, k
        cur = bo


def test_render_matches_oracle_renderer(hpe, orc, omodel, D):
    rng = np.random.default_rng(7)
    th = np.stack([synth_pose(D, rng) for _ in range(3)])
    fr = hpe.render_frames(hpe.handmodel(), th)
    for k in range(3):
        ref = orc.render_depth(omodel, th[k])
        assert ((fr[k] > 0) == (ref > 0)).mean() > 0.9999
        both = (fr[k] > 0) & (ref > 0)
        assert np.abs(fr[k][both] - ref[both]).max() < 1e-3


def test_errors_are_reported(hpe, D):
    hand = hpe.handmodel()
    obs = hpe.observedmodel(capacity=1)
    with pytest.raises(hpe.HpeError) as e:
        hpe.costfunc(hand, obs).cal_cost(D.X0)        # slot never filled
    assert e.value.code == 4
    with pytest.raises(hpe.HpeError):
        hpe.handmodel(tb_spheres=[3, 2, 2, 2])
    with pytest.raises(hpe.HpeError):
        hpe.observedmodel(capacity=1, num_samples=300)


def test_certified_matcher_equals_direct_form(hpe, orc, omodel, frame0, D):
    """the two-sweep search (default) and the reference's arithmetic on all 48 spheres give the
    same indices, on random poses, near-optimal poses and poses far from the cloud"""
    truth, depth, oobs = frame0
    rng = np.random.default_rng(8)
    th = np.concatenate([truth + rng.normal(size=(3000, 26)) * D.STD,
                         truth + rng.normal(size=(3000, 26)) * D.STD * 0.05,
                         truth + rng.normal(size=(500, 26)) * D.STD * 4.0])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    lib = hpe.lib()
    try:
        assert lib.hpe_set_matcher(1) == 0
        m_direct = cf.compute_correspondences(th)
    finally:
        lib.hpe_set_matcher(0)
    m_fast = cf.compute_correspondences(th)
    assert np.array_equal(m_direct, m_fast)
    for i in (0, 1, 3000, 3001, 6000, 6499):
        S, _ = orc.fk(omodel, th[i])
        assert np.array_equal(m_fast[i], orc.correspondences(oobs.cloud, S))


def test_matcher_on_exact_ties(hpe, orc, omodel, D):
    """a cloud made of sphere centres and of midpoints between neighbouring spheres: exact and
    near ties, where only the first-minimum / sqrtf semantics decide"""
    rng = np.random.default_rng(9)
    truth = synth_pose(D, rng)
    depth = orc.render_depth(omodel, truth)
    oobs = orc.Observation(depth)
    S, _ = orc.fk(omodel, truth)
    pts = [S[j] for j in range(48)] + [(S[j] + S[j + 1]) / 2 for j in range(47)]
    pts += [S[j] + rng.normal(size=3) * 1e-7 for j in range(48)]
    pts += [(S[j] + S[(j + 10) % 48]) / 2 for j in range(48)]
    cloud = np.array(pts)[:250]
    cloud = np.concatenate([cloud, oobs.cloud[: 250 - len(cloud)]])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.set(0, depth, oobs.dt.astype(np.float32), cloud, oobs.scale)
    cf = hpe.costfunc(hand, obs)
    th = np.concatenate([truth[None], truth + rng.normal(size=(200, 26)) * 1e-6])
    m = cf.compute_correspondences(th)
    for i in range(len(th)):
        Si, _ = orc.fk(omodel, th[i])
        assert np.array_equal(m[i], orc.correspondences(cloud, Si)), i
