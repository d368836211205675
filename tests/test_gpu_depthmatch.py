"""GPU: costfunc::depthMatch_penalty (src/costfunc.cpp:439-473) -- SURVEY 8(f) rank 4, second half -- through the C ABI
against the oracle: the depth map with the outlines of the projected spheres (cv::circle, thickness 2).  The RASTER is
bit-exact (the same pixels, overwritten in the same order -- integer work); the value drawn is the sphere's depth c_z,
a forward-kinematics output, which agrees with the oracle's to ~1e-15 relative like every FK quantity (GPU sincos vs
libm), so values are compared at 1e-12 and the untouched background exactly.
The oracle's restatement of OpenCV's outlined circle is pinned against cv2.circle and against the reference's own
depthMatch_penalty (oracle/_ref) in the CPU tests."""
import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu


def test_depthmatch_matches_oracle(hpe, orc, omodel, D):
    rng = np.random.default_rng(61)
    truth = synth_pose(D, rng)
    depth = orc.render_depth(omodel, truth)
    oobs = orc.Observation(depth)
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=1)
    obs.upload(depth[None])
    cf = hpe.costfunc(hand, obs)
    th = np.concatenate([truth[None],
                         truth + rng.normal(size=(60, 26)) * D.STD * 0.1,
                         truth + rng.normal(size=(60, 26)) * D.STD,
                         truth + rng.normal(size=(40, 26)) * D.STD * 3.0])          # outlines clipped by / outside the image
    th[5, 5] = 6.0                        # very close to the camera: large radii in pixels
    th[6, 3] = 60.0                       # far to the side: most outlines outside the image
    img = cf.depthMatch_penalty_batch(th)
    assert img.shape == (len(th), D.H, D.W)
    drawn = 0
    for i in range(len(th)):
        want = orc.depthmatch_penalty(omodel, oobs, th[i])
        hit = want != oobs.depth
        assert np.array_equal(img[i] != oobs.depth, hit), (i, int(((img[i] != oobs.depth) != hit).sum()))   # same pixels
        assert np.array_equal(img[i][~hit], oobs.depth[~hit])                                            # background untouched
        assert np.allclose(img[i][hit], want[hit], rtol=1e-12, atol=0), i                                # same sphere's depth
        drawn += int(hit.sum())
    assert drawn > 100 * len(th)          # the outlines are there
    # single-pose entry point = batch entry
    assert np.array_equal(cf.depthMatch_penalty(th[3]), img[3])


def test_depthmatch_other_sample_counts_and_errors(hpe, orc, omodel, D):
    rng = np.random.default_rng(62)
    truth = synth_pose(D, rng)
    depth = orc.render_depth(omodel, truth)
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=2, num_samples=256)
    obs.upload(depth[None], first=1)
    cf = hpe.costfunc(hand, obs)
    img = cf.depthMatch_penalty(truth, slot=1)
    want = orc.depthmatch_penalty(omodel, orc.Observation(depth, num_samples=256), truth)
    assert np.array_equal(img != want, np.zeros_like(img, dtype=bool)) or np.allclose(img, want, rtol=1e-12, atol=0)
    assert np.array_equal(img == 0, want == 0)
    with pytest.raises(hpe.HpeError) as e:
        cf.depthMatch_penalty(truth, slot=0)          # slot never filled
    assert e.value.code == 4
