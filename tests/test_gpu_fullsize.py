"""GPU: the BASELINE.json configurations at their real sizes, through properties that do not need the (slow)
CPU oracle: a result must not depend on how the work is batched, chunked, chained or sharded, and spot checks
against the oracle on a few items."""
import numpy as np
import pytest

from conftest import synth_pose
from test_gpu_swarm import _run_swarms

pytestmark = pytest.mark.gpu


def _sequence(hpe, D, n, seed):
    """smooth random walk of the pose + its frames rendered on the device"""
    rng = np.random.default_rng(seed)
    pose = synth_pose(D, rng)
    walk = np.empty((n, 26))
    for k in range(n):
        pose = np.clip(pose + rng.uniform(-1, 1, size=26) * np.r_[[2.0] * 3, [0.2] * 3, [2.0] * 20], D.LB, D.UB)
        walk[k] = pose
    return walk, rng


def test_config2_sequential_tracking_1000_frames(hpe, orc, omodel, D):
    """configs[2]: 1,000-frame sequence, 64 x 30, temporal init.  The one-launch chained run equals frame-by-frame
    calls bit for bit; the first frames equal the oracle's loop; the tracker stays on the hand."""
    n, num_p, maxiter = 1000, 64, 30
    walk, rng = _sequence(hpe, D, n, 3)
    hand = hpe.handmodel()
    frames = hpe.render_frames(hand, walk)
    obs = hpe.observedmodel(capacity=n)
    obs.upload(frames)
    cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    x0 = walk[0] + rng.normal(size=26) * D.STD * 0.3
    poses, costs = pso.track_sequence(cf, x0, num_p, n, seed=1000)
    assert np.isfinite(poses).all() and np.isfinite(costs).all()
    cur = x0
    for k in range(40):                                   # frame-by-frame = chained, bit for bit
        b = pso.pso_evolve(cf, cur, num_p, seed=1000, slot=k)
        assert np.array_equal(b, poses[k]), k
        cur = b
    prm = orc.pso_params(maxiter=maxiter)
    cur = x0
    for k in range(3):                                    # the oracle's loop on the head of the sequence
        bo, co, _, _, _ = orc.pso_evolve(omodel, orc.Observation(frames[k]), prm, cur, num_p, 1000)
        assert np.abs(poses[k] - bo).max() < 1e-7 and abs(costs[k] - co) <= 1e-8 * co
        cur = bo
    # restarting in the middle from the tracked pose reproduces the tail
    tail, tail_c = pso.track_sequence(cf, poses[699], num_p, 300, seed=1000, first=700)
    assert np.array_equal(tail, poses[700:]) and np.array_equal(tail_c, costs[700:])
    # the tracker does not lose the hand: wrist + tips stay within a few cm of the truth over the whole sequence
    _, Jw = hand.build_hand_models(walk)
    gt = (Jw * 10.0 * np.array([1.0, -1.0, -1.0])).reshape(n, 63)
    err = cf.gnd_truth_err_batch(poses, gt) / 6.0
    assert np.median(err) < 25.0 and err.max() < 80.0, (np.median(err), err.max())


def test_config3_sharded_swarm_16384(hpe, D):
    """configs[3]: 16,384 particles x 30 generations on one frame: unsharded stepwise run == uneven shards ==
    the one-launch kernel, bit for bit"""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(4)
    truth = synth_pose(D, rng)
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(hpe.render_frames(hand, truth[None]))
    cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
    num_p, maxiter, seed = 16384, 30, 4
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    (ref,) = _run_swarms(hpe, [hpe.Swarm(cf, pso.params, num_p, 0, num_p, x0, seed=seed)], torch)
    cuts = [0, 2048, 2049, 9000, 16384]
    shards = [hpe.Swarm(cf, pso.params, num_p, a, b, x0, seed=seed) for a, b in zip(cuts[:-1], cuts[1:])]
    for bp, bc, tr in _run_swarms(hpe, shards, torch):
        assert np.array_equal(bp, ref[0]) and bc == ref[1] and np.array_equal(tr, ref[2])
    sw = hpe.Swarm(cf, pso.params, num_p, 0, num_p, x0, seed=seed)
    sw.run()
    bp, bc, tr = sw.best()
    assert np.array_equal(bp, ref[0]) and bc == ref[1] and np.array_equal(tr, ref[2])
    assert np.all(np.diff(tr) <= 0) and bc < cf.cal_cost(x0)
    a, d, _ = cf.cal_cost_batch(bp[None])
    assert abs(a[0] + d[0] - bc) <= 1e-12 * bc                 # gbest_pos scores gbest_cost


def test_config4_offline_batch_is_chunk_invariant(hpe, D):
    """configs[4] (262,144 frames in the bench): a 4,096-frame slice fitted in one batch == fitted in chunks of
    1,000 through a smaller observation buffer, bit for bit"""
    rng = np.random.default_rng(5)
    n = 4096
    truths = np.clip(D.X0 + rng.normal(size=(n, 26)) * D.STD * 0.5, D.LB, D.UB)
    x0 = truths + rng.normal(size=truths.shape) * D.STD * 0.3
    hand = hpe.handmodel()
    frames = hpe.render_frames(hand, truths)
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 30, 1e-8, 1e-8)
    obs = hpe.observedmodel(capacity=n)
    obs.upload(frames)
    bp, bc, _ = pso.pso_evolve_batch(hpe.costfunc(hand, obs), x0, 64, want_trace=False)
    small = hpe.observedmodel(capacity=1000)
    cfs = hpe.costfunc(hand, small)
    for a in range(0, n, 1000):
        b = min(a + 1000, n)
        small.upload(frames[a:b])
        bp2, bc2, _ = pso.pso_evolve_batch(cfs, x0[a:b], 64, want_trace=False)
        assert np.array_equal(bp2, bp[a:b]) and np.array_equal(bc2, bc[a:b]), a
    x0_cost = np.array([hpe.costfunc(hand, obs).cal_cost(x0[k], slot=k) for k in range(0, n, 256)])
    assert np.all(bc[::256] < x0_cost)


def test_streaming_fit_from_host_frames(hpe, D):
    """hpe_fit_frames: frames streamed from host memory through a small observation buffer (uploads overlapped
    with the fits) == upload + pso_evolve_batch on a buffer that holds them all, bit for bit; an empty frame
    comes back as NaN cost with its x0"""
    rng = np.random.default_rng(8)
    n = 1500
    truths = np.clip(D.X0 + rng.normal(size=(n, 26)) * D.STD * 0.5, D.LB, D.UB)
    x0 = truths + rng.normal(size=truths.shape) * D.STD * 0.3
    hand = hpe.handmodel()
    frames = hpe.render_frames(hand, truths)
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 12, 1e-8, 1e-8)
    seeds = np.arange(n, dtype=np.uint64) + 5
    big = hpe.observedmodel(capacity=n)
    big.upload(frames)
    bp, bc, _ = pso.pso_evolve_batch(hpe.costfunc(hand, big), x0, 32, seeds=seeds, want_trace=False)
    for cap in (256, 2 * n, 1300, 77, 1):        # two halves | six chunks on three streams | four on two | halves | one slot
        m = n if cap > 1 else 5
        small = hpe.observedmodel(capacity=cap)
        bp2, bc2 = pso.fit_frames(hpe.costfunc(hand, small), frames[:m], x0[:m], 32, seeds=seeds[:m])
        assert np.array_equal(bp2, bp[:m]) and np.array_equal(bc2, bc[:m]), cap
    # submit / wait: two calls in flight through one buffer (the second call's uploads start under the first call's
    # last fits), a third submit is refused, results are the blocking call's
    for cap in (2 * n, 300):
        cf2 = hpe.costfunc(hand, hpe.observedmodel(capacity=cap))
        j1 = pso.fit_frames_submit(cf2, frames[:900], x0[:900], 32, seeds=seeds[:900])
        j2 = pso.fit_frames_submit(cf2, frames[900:], x0[900:], 32, seeds=seeds[900:])
        with pytest.raises(hpe.HpeError) as e:
            pso.fit_frames_submit(cf2, frames[:5], x0[:5], 32, seeds=seeds[:5])
        assert e.value.code == 4                      # HPE_ERR_STATE
        p1, c1 = j1.wait()
        j3 = pso.fit_frames_submit(cf2, frames[:700], x0[:700], 32, seeds=seeds[:700])
        p2, c2 = j2.wait()
        p3, c3 = j3.wait()
        assert np.array_equal(np.concatenate([p1, p2]), bp) and np.array_equal(np.concatenate([c1, c2]), bc), cap
        assert np.array_equal(p3, bp[:700]) and np.array_equal(c3, bc[:700]), cap
    fr = frames[:10].copy()
    fr[3] = 0.0
    bp3, bc3 = pso.fit_frames(hpe.costfunc(hand, hpe.observedmodel(capacity=4)), fr, x0[:10], 32, seeds=seeds[:10])
    assert np.isnan(bc3[3]) and np.array_equal(bp3[3], x0[3])
    keep = np.arange(10) != 3
    assert np.array_equal(bp3[keep], bp[:10][keep]) and np.array_equal(bc3[keep], bc[:10][keep])
