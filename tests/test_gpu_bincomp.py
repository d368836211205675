"""GPU: costfunc::bincomp_penalty (src/costfunc.cpp:384-436) -- SURVEY 8(f) rank 4, the silhouette-mismatch term
"used for testing" and the projected-sphere image behind the reference's debug views -- against the oracle
(bit-identical to the reference build, tests/test_ref_build.py; its filled circle pinned against cv2.circle,
tests/test_oracle_pins.py) and the reference-made fixture.  Integer pixel counts: equality is exact."""
import os

import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_bincomp_matches_oracle_bit_exact(hpe, orc, omodel, frame0, D):
    truth, depth, oobs = frame0
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    rng = np.random.default_rng(70)
    n = 1500
    th = truth + rng.normal(size=(n, 26)) * D.STD * np.array([0.0, 0.1, 1.0, 3.0])[np.arange(n) % 4][:, None]
    th[5::6, 3:6] += rng.normal(size=(len(th[5::6]), 3)) * 15.0          # discs clipped by the border / outside
    got, sil = cf.bincomp_penalty_batch(th, want_silhouette=True)
    for k in range(n):
        if k < 200:
            v, s = orc.bincomp_penalty(omodel, oobs, th[k], want_silhouette=True)
            assert np.array_equal(sil[k], s != 0), k
        else:
            v = orc.bincomp_penalty(omodel, oobs, th[k])
        assert got[k] == v, (k, got[k], v)
    assert got[0] == np.count_nonzero(sil[0] != (depth != 0))
    assert cf.bincomp_penalty(th[3]) == got[3]                            # single-pose entry point


def test_bincomp_reproduces_reference_fixtures(hpe, orc, omodel, D):
    g = np.load(os.path.join(HERE, "golden", "golden_ref_opt_v1.npz"))
    depth = orc.render_depth(omodel, g["bincomp_truth"])
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    assert np.array_equal(hpe.costfunc(hand, obs).bincomp_penalty_batch(g["bincomp_poses"]), g["bincomp"])


def test_bincomp_edge_cases(hpe, orc, omodel, frame0, D):
    truth, depth, oobs = frame0
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=3)
    # a frame with a handful of hand pixels in slot 1; slot 2 stays empty
    few = np.zeros((240, 320), dtype=np.float32)
    few[100:103, 150:160] = 300.0
    obs.upload(np.stack([depth, few]))
    cf = hpe.costfunc(hand, obs)
    v = cf.bincomp_penalty(truth, slot=1)
    assert v == orc.bincomp_penalty(omodel, orc.Observation(few), truth)
    assert cf.bincomp_penalty_batch(np.empty((0, 26))).shape == (0,)
    with pytest.raises(hpe.HpeError) as e:
        cf.bincomp_penalty(truth, slot=2)
    assert e.value.code == 4
    with pytest.raises(hpe.HpeError):
        cf.bincomp_penalty(truth, slot=3)


def test_bincomp_full_size_properties(hpe, frame0, D):
    """65,536 poses (BASELINE configs[1]'s batch): the result of a pose does not depend on its place in the batch
    or on how the batch is split"""
    truth, depth, _ = frame0
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    rng = np.random.default_rng(71)
    n = 65536
    th = truth + rng.normal(size=(n, 26)) * D.STD
    pen = cf.bincomp_penalty_batch(th)
    perm = rng.permutation(n)
    assert np.array_equal(cf.bincomp_penalty_batch(th[perm]), pen[perm])
    assert np.array_equal(np.concatenate([cf.bincomp_penalty_batch(th[:30001]), cf.bincomp_penalty_batch(th[30001:])]), pen)
    assert pen.min() >= 0 and pen.max() <= 240 * 320 and cf.bincomp_penalty(truth) < np.median(pen)
