"""CPU: the optimiser half of the oracle -- RNG known answers, the reference's quirks
(src/PSO.cpp:358-377, 56-74, 717-886) and invariants of pso_evolve."""
import numpy as np
import pytest

from conftest import synth_pose


def test_philox_known_answers(orc):
    """Random123 kat_vectors for philox4x32-10 (counter, key -> output)"""
    def call(ctr, key):
        seed = key[0] | (key[1] << 32)
        return orc.philox(seed, *ctr)
    assert call((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert call((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert call((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_and_normal_streams(orc):
    u = np.array([orc.uniform2(7, e, p, 3, 1) for e in range(13) for p in range(200)]).ravel()
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1 / 12) < 0.01
    z = np.array([orc.normal2(7, e, p, 0) for e in range(13) for p in range(400)]).ravel()
    assert abs(z.mean()) < 0.04 and abs(z.std() - 1.0) < 0.04 and np.isfinite(z).all()
    # counter-based: any element is addressable, streams differ, seeds differ
    assert orc.uniform2(7, 3, 5, 2, 1) == orc.uniform2(7, 3, 5, 2, 1)
    assert orc.uniform2(7, 3, 5, 2, 1) != orc.uniform2(7, 3, 5, 2, 2)
    assert orc.uniform2(7, 3, 5, 2, 1) != orc.uniform2(8, 3, 5, 2, 1)
    w = orc.philox(7, 3, 5, 2, 1)
    assert orc.uniform2(7, 3, 5, 2, 1)[0] == (((w[0] >> 5) << 26) | (w[1] >> 6)) / 2.0 ** 53


def test_check_constraints_resets_upper_violation_to_lower_bound(orc, D):
    prm = orc.pso_params()
    th = (D.LB + D.UB) / 2
    v = np.ones(26)
    th[7] = D.UB[7] + 5.0         # above the upper bound -> LOWER bound (src/PSO.cpp:372)
    th[9] = D.LB[9] - 1.0         # below the lower bound -> lower bound
    th[11] = D.UB[11]             # exactly on the bound: untouched (strict comparisons)
    t2, v2 = orc.check_constraints(prm, th, v)
    assert t2[7] == D.LB[7] and v2[7] == 0 and t2[9] == D.LB[9] and v2[9] == 0
    assert t2[11] == D.UB[11] and v2[11] == 1
    keep = np.ones(26, bool); keep[[7, 9]] = False
    assert np.array_equal(t2[keep], th[keep]) and np.all(v2[keep] == 1)


def test_generate_particles_not_clamped_and_not_x0(orc, D):
    prm = orc.pso_params()
    x0 = D.UB.copy()                                   # at the bound: half the draws must land outside
    P = orc.generate_particles(prm, x0, 64, 1000)
    assert (P > D.UB).any() and not np.any(np.all(P == x0, axis=1))
    z = (P - x0) / D.STD
    assert abs(z.mean()) < 0.1 and abs(z.std() - 1) < 0.1
    assert np.array_equal(P, orc.generate_particles(prm, x0, 64, 1000))
    assert np.array_equal(P[:10], orc.generate_particles(prm, x0, 10, 1000))   # keyed by particle index


def test_pso_evolve_invariants(orc, omodel, frame0, D):
    truth, depth, obs = frame0
    rng = np.random.default_rng(3)
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    prm = orc.pso_params(maxiter=12)
    bp, best, tr, xs, pc = orc.pso_evolve(omodel, obs, prm, x0, 32, 1000)
    assert len(tr) == 11 and np.all(np.diff(tr) <= 0) and tr[-1] == best
    assert abs(orc.cal_cost(omodel, obs, bp) - best) <= 1e-12 * best       # gbest_pos scores gbest_cost
    assert pc.min() == best
    assert best < orc.cal_cost(omodel, obs, x0)
    assert np.all(xs >= D.LB - 1e-12) and np.all(xs <= D.UB + 1e-12)        # positions after clamping
    # same seed -> same run (the reference reseeds inside pso_evolve); other seed -> other run
    bp2, best2, tr2, _, _ = orc.pso_evolve(omodel, obs, prm, x0, 32, 1000)
    assert np.array_equal(bp, bp2) and np.array_equal(tr, tr2)
    _, _, tr3, _, _ = orc.pso_evolve(omodel, obs, prm, x0, 32, 1001)
    assert not np.array_equal(tr, tr3)


def test_maxiter_counts_the_initial_evaluation(orc, omodel, frame0, D):
    truth, depth, obs = frame0
    x0 = truth + 1.0
    P = orc.generate_particles(orc.pso_params(), x0, 16, 1000)
    a, d, _ = orc.cost_batch(omodel, obs, P, collision=False)
    bp, best, tr, xs, pc = orc.pso_evolve(omodel, obs, orc.pso_params(maxiter=1), x0, 16, 1000)
    assert len(tr) == 0 and np.array_equal(xs, P) and best == (a + d).min() and np.array_equal(bp, P[(a + d).argmin()])
    _, best2, tr2, _, _ = orc.pso_evolve(omodel, obs, orc.pso_params(maxiter=2), x0, 16, 1000)
    assert len(tr2) == 1 and best2 <= best


def test_pso_recovers_a_nearby_pose(orc, omodel, D):
    rng = np.random.default_rng(11)
    truth = synth_pose(D, rng)
    obs = orc.Observation(orc.render_depth(omodel, truth))
    x0 = truth + rng.normal(size=26) * D.STD * 0.2
    _, best, _, _, _ = orc.pso_evolve(omodel, obs, orc.pso_params(maxiter=30), x0, 64, 1000)
    assert best < 0.25 * orc.cal_cost(omodel, obs, x0)
