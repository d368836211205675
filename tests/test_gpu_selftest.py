"""GPU: device self-tests of arithmetic shortcuts that must be exact."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_div10_shortcut_is_the_ieee_division_for_every_float(hpe):
    bad = C.c_ulonglong(123)
    assert hpe.lib().hpe_selftest_div10(C.byref(bad)) == 0
    assert bad.value == 0


def test_divide_free_back_projection_is_the_ieee_division(hpe, D):
    lib = hpe.lib()
    for f in (D.FOCAL, 241.42, 500.0, 3.0):
        bad = C.c_ulonglong(123)
        assert lib.hpe_selftest_div_const(C.c_double(f), C.byref(bad)) == 0
        assert bad.value == 0, f


def test_both_distance_transforms_equal_the_oracle(hpe, orc, omodel, D):
    """default = OpenCV's two raster passes with prefix-minimum row scans; 1 = closed-form row search;
    2 = the two passes as a skewed wavefront"""
    import numpy as np
    from conftest import synth_pose
    rng = np.random.default_rng(33)
    frames = np.stack([orc.render_depth(omodel, synth_pose(D, rng)) for _ in range(5)])
    frames[3, :, :] = 0; frames[3, 0, 0] = 400.0                     # single corner pixel: distances up to 398 px
    frames[4, 200:240, 0:30] = 350.0; frames[4, 0:3, 317:320] = 220.0  # blobs touching the borders
    lib = hpe.lib()
    res = []
    for mode in (0, 1, 2):
        assert lib.hpe_set_chamfer(mode) == 0
        try:
            obs = hpe.observedmodel(capacity=5)
            obs.upload(frames)
            res.append([obs.get(k) for k in range(5)])
        finally:
            lib.hpe_set_chamfer(0)
    for k in range(5):
        ref = orc.Observation(frames[k])
        for r in res:
            assert np.array_equal(r[k]["dt"].astype(np.float64), ref.dt), k
            assert r[k]["dt_max"] == np.float32(ref.dt_max)
