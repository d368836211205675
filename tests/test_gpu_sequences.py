"""GPU: batched sequence tracking (hpe_track_sequences / hpe_track_sequences_refined).  The reference's loop is one
sequence (src/testmodel.cpp:117-138); the batched entry points run the same loop over several sequences at once and
must return, for every sequence, the bits hpe_track_sequence returns for it alone."""
import numpy as np
import pytest

from conftest import synth_pose

pytestmark = pytest.mark.gpu


def _walks(D, n_seq, n_frames, seed):
    rng = np.random.default_rng(seed)
    walks = np.empty((n_seq, n_frames, 26))
    for q in range(n_seq):
        pose = synth_pose(D, rng)
        for k in range(n_frames):
            pose = np.clip(pose + rng.uniform(-1, 1, size=26) * np.r_[[2.0] * 3, [0.2] * 3, [2.0] * 20], D.LB, D.UB)
            walks[q, k] = pose
    x0 = walks[:, 0] + rng.normal(size=(n_seq, 26)) * D.STD * 0.3
    return walks, x0


@pytest.mark.parametrize("n_seq,path", [(3, 0), (11, 0), (40, 0), (5, 1)])
def test_track_sequences_equals_one_at_a_time(hpe, orc, omodel, D, n_seq, path):
    n_frames, num_p, maxiter = 12, 32, 10
    walks, x0 = _walks(D, n_seq, n_frames, 21 + n_seq)
    hand = hpe.handmodel()
    frames = hpe.render_frames(hand, walks.reshape(-1, 26))                   # sequence after sequence
    obs = hpe.observedmodel(capacity=n_seq * n_frames)
    obs.upload(frames)
    cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    seeds = np.arange(n_seq, dtype=np.uint64) + 1000
    lib = hpe.lib()
    try:
        assert lib.hpe_set_pso_path(path) == 0
        poses, costs = pso.track_sequences(cf, x0, num_p, n_seq, n_frames, seeds=seeds)
        for q in range(n_seq):
            p1, c1 = pso.track_sequence(cf, x0[q], num_p, n_frames, seed=int(seeds[q]), first=q * n_frames)
            assert np.array_equal(p1, poses[q]) and np.array_equal(c1, costs[q]), q
    finally:
        lib.hpe_set_pso_path(0)
    # frame-interleaved slots give the same answer
    inter = hpe.observedmodel(capacity=n_seq * n_frames)
    inter.upload(frames.reshape(n_seq, n_frames, D.H, D.W).transpose(1, 0, 2, 3).reshape(-1, D.H, D.W))
    p2, c2 = pso.track_sequences(hpe.costfunc(hand, inter), x0, num_p, n_seq, n_frames, seeds=seeds, interleaved=True)
    assert np.array_equal(p2, poses) and np.array_equal(c2, costs)
    # the head of sequence 0 against the oracle's loop
    prm = orc.pso_params(maxiter=maxiter)
    cur = x0[0]
    for k in range(3):
        bo, co, _, _, _ = orc.pso_evolve(omodel, orc.Observation(frames[k]), prm, cur, num_p, int(seeds[0]))
        assert np.abs(poses[0, k] - bo).max() < 1e-7 and abs(costs[0, k] - co) <= 1e-8 * co
        cur = bo


def test_track_sequences_refined_equals_one_at_a_time(hpe, D):
    n_seq, n_frames, num_p, maxiter = 4, 6, 32, 8
    walks, x0 = _walks(D, n_seq, n_frames, 33)
    hand = hpe.handmodel()
    frames = hpe.render_frames(hand, walks.reshape(-1, 26))
    seqmaj = hpe.observedmodel(capacity=n_seq * n_frames)
    seqmaj.upload(frames)
    inter = hpe.observedmodel(capacity=n_seq * n_frames)
    inter.upload(frames.reshape(n_seq, n_frames, D.H, D.W).transpose(1, 0, 2, 3).reshape(-1, D.H, D.W))
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    seeds = np.arange(n_seq, dtype=np.uint64) + 7
    poses, costs, refined = pso.track_sequences_refined(hpe.costfunc(hand, inter), x0, num_p, n_seq, n_frames, seeds=seeds)
    cf = hpe.costfunc(hand, seqmaj)
    for q in range(n_seq):
        p1, c1, r1 = pso.track_sequence_refined(cf, x0[q], num_p, n_frames, seed=int(seeds[q]), first=q * n_frames)
        assert np.array_equal(p1, poses[q]) and np.array_equal(c1, costs[q]) and np.array_equal(r1, refined[q]), q


def test_track_sequences_rejects_overlapping_layouts(hpe, D):
    hand, obs = hpe.handmodel(), hpe.observedmodel(capacity=8)
    obs.upload(hpe.render_frames(hand, np.tile(D.X0, (8, 1))))
    pso = hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 4, 1e-8, 1e-8)
    import ctypes as C
    out = np.empty((8, 26))
    rc = hpe.lib().hpe_track_sequences(hand._h, obs._h, C.byref(pso.params), 16, np.tile(D.X0, (2, 1)).ctypes.data_as(C.c_void_p),
                                       0, 2, 4, 2, 1, None, out.ctypes.data_as(C.c_void_p), None)
    assert rc == 1          # HPE_ERR_ARG
