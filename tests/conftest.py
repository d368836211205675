import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def omodel(orc):
    return orc.make_model()


@pytest.fixture(scope="session")
def D():
    from hpe_b200 import defaults
    return defaults


def synth_pose(D, rng, spread=0.5):
    """ground-truth pose of SURVEY 8(d): x0_ref + N(0, spread*std_ref), clipped to the bounds"""
    return np.clip(D.X0 + rng.normal(size=26) * D.STD * spread, D.LB, D.UB)


@pytest.fixture(scope="session")
def frame0(orc, omodel, D):
    """one synthetic Subject1-format frame + its oracle observation"""
    rng = np.random.default_rng(20140528)
    truth = synth_pose(D, rng)
    depth = orc.render_depth(omodel, truth)
    return truth, depth, orc.Observation(depth)


@pytest.fixture(scope="session")
def hpe():
    """the product library; GPU tests fail loudly when it is missing"""
    import hpe_b200
    hpe_b200.lib()
    assert hpe_b200.device_count() > 0, "no CUDA device visible"
    return hpe_b200
