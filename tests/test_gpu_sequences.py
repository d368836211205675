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


def test_async_device_pointer_entry_points(hpe, D):
    """hpe_pso_evolve_dev / hpe_track_sequences_dev / hpe_refine_init_pose_dev: everything on a caller's stream, device
    pointers in and out, no synchronisation inside; same bits as the blocking host-pointer calls"""
    torch = pytest.importorskip("torch")
    import ctypes as C
    n_seq, n_frames, num_p, maxiter = 3, 5, 32, 8
    walks, x0 = _walks(D, n_seq, n_frames, 77)
    hand = hpe.handmodel()
    obs = hpe.observedmodel(capacity=n_seq * n_frames)
    obs.upload(hpe.render_frames(hand, walks.reshape(-1, 26)))
    cf, pso = hpe.costfunc(hand, obs), hpe.PSO()
    pso.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8)
    lib = hpe.lib()
    st = torch.cuda.Stream()
    n = n_seq * n_frames
    xall = walks.reshape(-1, 26) + np.random.default_rng(78).normal(size=(n, 26)) * D.STD * 0.2
    with torch.cuda.stream(st):
        x_dev = torch.from_numpy(xall).cuda()
        bp_dev = torch.empty(n, 26, dtype=torch.float64, device="cuda")
        bc_dev = torch.empty(n, dtype=torch.float64, device="cuda")
        # refine in place, then one swarm per frame from the refined poses, all queued on `st`
        assert lib.hpe_refine_init_pose_dev(hand._h, obs._h, C.c_void_p(x_dev.data_ptr()), 0, n, None, C.c_void_p(st.cuda_stream)) == 0
        assert lib.hpe_pso_evolve_dev(hand._h, obs._h, C.byref(pso.params), num_p, C.c_void_p(x_dev.data_ptr()), 0, n, None,
                                      C.c_void_p(bp_dev.data_ptr()), C.c_void_p(bc_dev.data_ptr()), None,
                                      C.c_void_p(st.cuda_stream)) == 0
        xs_dev = torch.from_numpy(x0).cuda()
        ps_dev = torch.empty(n, 26, dtype=torch.float64, device="cuda")
        cs_dev = torch.empty(n, dtype=torch.float64, device="cuda")
        assert lib.hpe_track_sequences_dev(hand._h, obs._h, C.byref(pso.params), num_p, C.c_void_p(xs_dev.data_ptr()), 0, n_seq,
                                           n_frames, n_frames, 1, None, C.c_void_p(ps_dev.data_ptr()),
                                           C.c_void_p(cs_dev.data_ptr()), C.c_void_p(st.cuda_stream)) == 0
    st.synchronize()
    xr, _ = pso.refine_init_pose_batch(xall, cf)
    assert np.array_equal(x_dev.cpu().numpy(), xr)
    bp, bc, _ = pso.pso_evolve_batch(cf, xr, num_p, want_trace=False)
    assert np.array_equal(bp_dev.cpu().numpy(), bp) and np.array_equal(bc_dev.cpu().numpy(), bc)
    poses, costs = pso.track_sequences(cf, x0, num_p, n_seq, n_frames)
    assert np.array_equal(ps_dev.cpu().numpy().reshape(n_seq, n_frames, 26), poses)
    assert np.array_equal(cs_dev.cpu().numpy().reshape(n_seq, n_frames), costs)
