"""GPU: a process may drive several devices -- the library's host-side caches (SM count, kernel attributes, streams,
scratch buffers, streaming jobs) are kept per device, and a handle is refused on a device it was not created on
(include/hpe.h, "Devices and threads").  Needs two GPUs; skipped on a one-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_devices_in_one_process(hpe, D):
    torch = pytest.importorskip("torch")
    if hpe.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(41)
    th = D.X0 + rng.normal(size=(3000, 26)) * D.STD
    x0 = D.X0 + rng.normal(size=(4, 26)) * D.STD * 0.3
    res = []
    objs = []
    for dev in (0, 1):
        torch.cuda.set_device(dev)
        hand = hpe.handmodel()
        obs = hpe.observedmodel(capacity=4)
        obs.upload(hpe.render_frames(hand, np.tile(D.X0, (4, 1))))
        cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
        pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 8, 1e-8, 1e-8)
        a, d, _ = cf.cal_cost_batch(th)
        bp, bc, _ = pso.pso_evolve_batch(cf, x0, 32)
        seq, _ = pso.track_sequence(cf, x0[0], 32, 4)
        res.append((a, d, bp, bc, seq))
        objs.append((hand, obs, cf, pso))
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y)                       # same bits on both devices
    # device 0's handles while device 1 is current: refused, not launched with foreign pointers
    with pytest.raises(hpe.HpeError) as e:
        objs[0][2].cal_cost_batch(th[:10])
    assert e.value.code == 4
    torch.cuda.set_device(0)
    a0, _, _ = objs[0][2].cal_cost_batch(th[:10])
    assert np.array_equal(a0, res[0][0][:10])
