"""GPU: the streaming form of the batched cost (hpe_cost_batch_submit / hpe_cost_batch_wait, include/hpe.h) --
several batches in flight, the upload of one under the kernels of the previous -- gives hpe_cost_batch's results
bit for bit, in any interleaving, and reports misuse instead of queueing without bound."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_streamed_batches_equal_the_one_call_form(hpe, orc, omodel, frame0, D):
    truth, depth, oobs = frame0
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    rng = np.random.default_rng(80)
    sizes = [40000, 5, 20000, 16384, 1]
    batches = [truth + rng.normal(size=(n, 26)) * D.STD for n in sizes]
    want = [cf.cal_cost_batch(b, with_collision=(k % 2 == 1)) for k, b in enumerate(batches)]
    # three in flight, waited out of order; then the remaining two
    jobs = [cf.cal_cost_submit(batches[k], with_collision=(k % 2 == 1)) for k in range(3)]
    assert sorted(j.id for j in jobs) == [0, 1, 2]
    with pytest.raises(hpe.HpeError) as e:                    # a fourth is refused, not queued
        cf.cal_cost_submit(batches[3])
    assert e.value.code == 4
    for k in (1, 0, 2):
        got = jobs[k].wait()
        for g, w in zip(got, want[k]):
            assert np.array_equal(g, w), k
    with pytest.raises(hpe.HpeError) as e:                    # a job can be waited for once
        jobs[0].wait()
    assert e.value.code in (1, 4)
    more = [cf.cal_cost_submit(batches[k], with_collision=(k % 2 == 1)) for k in (3, 4)]
    for k, j in zip((3, 4), more):
        for g, w in zip(j.wait(), want[k]):
            assert np.array_equal(g, w), k
    # spot check against the oracle, and the empty batch
    ao, do, _ = orc.cost_batch(omodel, oobs, batches[1])
    assert np.abs((want[1][0] + want[1][1]) / (ao + do) - 1).max() < 1e-9
    a, d, c = cf.cal_cost_submit(np.empty((0, 26))).wait()
    assert a.shape == d.shape == c.shape == (0,)


def test_dropped_jobs_release_their_slots(hpe, frame0, D):
    """a CostJob dropped without wait() waits in its finaliser (its pinned buffers are DMA targets) and frees its slot"""
    import gc
    truth, depth, _ = frame0
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    th = truth + np.random.default_rng(82).normal(size=(5000, 26)) * D.STD
    want = cf.cal_cost_batch(th)
    for _ in range(3):
        jobs = [cf.cal_cost_submit(th) for _ in range(3)]       # all three slots taken
        del jobs
        gc.collect()
    a, d, _ = cf.cal_cost_submit(th).wait()
    assert np.array_equal(a, want[0]) and np.array_equal(d, want[1])


def test_streamed_loop_two_in_flight(hpe, frame0, D):
    """the way bench.py drives it: submit step i, wait for step i-1"""
    truth, depth, _ = frame0
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    rng = np.random.default_rng(81)
    sets = [truth + rng.normal(size=(30000, 26)) * D.STD for _ in range(3)]
    want = [cf.cal_cost_batch(s) for s in sets]
    pending = []
    for i in range(12):
        pending.append((i % 3, cf.cal_cost_submit(sets[i % 3])))
        if len(pending) == 2:
            k, j = pending.pop(0)
            a, d, _ = j.wait()
            assert np.array_equal(a, want[k][0]) and np.array_equal(d, want[k][1])
    for k, j in pending:
        a, d, _ = j.wait()
        assert np.array_equal(a, want[k][0]) and np.array_equal(d, want[k][1])
