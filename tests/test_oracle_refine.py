"""CPU: the SURVEY 8(f) rows of the oracle -- refine_init_pose / cal_grad / goldstein, dim_restore,
gnd_truth_err -- against the reference's own code (oracle/_ref, when built) and against the fixtures
it wrote (always)."""
import os

import numpy as np
import pytest

from conftest import synth_pose

HERE = os.path.dirname(os.path.abspath(__file__))


def test_refine_fixtures(orc, omodel, D):
    g = np.load(os.path.join(HERE, "golden", "golden_ref_v1.npz"))
    obs = orc.Observation(orc.render_depth(omodel, g["truth"]))
    for k in range(len(g["refine_x0"])):
        x, n = orc.refine_init_pose(omodel, obs, g["refine_x0"][k])
        assert np.array_equal(x, g["refine_out"][k]) and n > 0                       # bit-exact
        assert orc.cal_cost2(omodel, obs, x)[0] == g["refine_cost"][k]
    _, J = orc.fk(omodel, g["truth"])
    for f in range(len(g["gt_rows"])):
        assert orc.gnd_truth_err(J, g["gt_rows"][f]) == g["gt_err"][f]
    assert np.array_equal(orc.dim_restore(g["dim22"]), g["dim26"])


def test_refine_against_reference_build(orc, omodel, D):
    import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    R = ref.Reference(threads=1)
    rng = np.random.default_rng(31)
    for k in range(10):
        truth = synth_pose(D, rng)
        depth = orc.render_depth(omodel, truth)
        oo = orc.Observation(depth)
        R.load_depth(depth)
        x0 = truth + rng.normal(size=26) * D.STD * [0.05, 0.1, 0.3, 0.6, 1.0][k % 5]
        xo, n = orc.refine_init_pose(omodel, oo, x0)
        assert np.array_equal(R.refine_init_pose(x0), xo), k
        assert 8 <= n <= 2 * 15 * (7 + 30)
    R.close()


def test_refine_properties(orc, omodel, frame0, D):
    truth, depth, obs = frame0
    rng = np.random.default_rng(32)
    x0 = truth + rng.normal(size=26) * D.STD * 0.2
    x, n = orc.refine_init_pose(omodel, obs, x0)
    assert np.array_equal(x[6:], x0[6:]) and not np.array_equal(x[:6], x0[:6])
    assert orc.cal_cost2(omodel, obs, x)[0] <= orc.cal_cost2(omodel, obs, x0)[0]
