"""CPU: the oracle restatement (oracle/hpe_oracle.c) against THE REFERENCE'S OWN CODE.

oracle/_ref/libhpe_ref.so is the reference's src/{fingermodel,thumbmodel,handmodel,costfunc,
observedmodel,PSO}.cpp compiled unmodified from /root/reference (oracle/Makefile) against the stand-in
Armadillo/OpenCV headers of oracle/shim/ -- the image has neither library.  The reference's logic runs
as written; what the stand-ins restate (BFMatcher, distanceTransform, Armadillo's summation order) is
pinned separately against cv2 in tests/test_oracle_pins.py.  The reference's RNG calls are fed the
counter-based streams of DESIGN.md "RNG" through a hook, so PSO::pso_evolve itself can be compared.

The .so is built where /root/reference exists and ships with the snapshot; tests that need it skip
when it is absent, the committed fixtures (tests/golden/golden_ref_v1.npz, written from it by
tests/golden/make_golden_ref.py) are always checked."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gref():
    return np.load(os.path.join(HERE, "golden", "golden_ref_v1.npz"))


@pytest.fixture(scope="module")
def R(orc):
    import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = ref.Reference(threads=1)
    yield r
    r.close()


# ---------------------------------------------------------------- fixtures written by the reference
def test_oracle_reproduces_reference_fixtures(orc, omodel, gref, D):
    """runs everywhere (also on the GPU box): the oracle against values the reference code produced"""
    g = gref
    depth = orc.render_depth(omodel, g["truth"])
    obs = orc.Observation(depth)
    assert np.array_equal(obs.cloud, g["cloud"]) and obs.scale == float(g["scale"])
    assert obs.dt_max == float(g["dt_max"]) and np.array_equal(obs.dt[::16].astype(np.float32), g["dt_rows"])
    assert obs.dt.sum() == float(g["dt_sum"]) and obs.depth.sum() == float(g["depth_cm_sum"])
    S, J = orc.fk(omodel, g["poses"])
    assert np.array_equal(S, g["spheres"]) and np.array_equal(J, g["joints"])          # bit-exact
    a, d, c = orc.cost_batch(omodel, obs, g["poses"])
    assert np.array_equal(a, g["align"]) and np.array_equal(d, g["depth_term"]) and np.array_equal(c, g["colli"])
    assert np.array_equal(a + d, g["cal_cost"]) and np.array_equal(a + d + c, g["cal_cost2"])
    for i in range(len(g["poses"])):
        assert np.array_equal(orc.correspondences(obs.cloud, S[i]), g["match"][i].astype(np.int32))
    tot, _, _ = orc.cal_cost2(omodel, obs, g["poses"][2], match=g["match"][3].astype(np.int32))
    assert tot == float(g["frozen_cost"])
    for tag in ("pso", "pso64"):
        prm = orc.pso_params(maxiter=int(g[tag + "_maxiter"]))
        bestp, best, _, _, _ = orc.pso_evolve(omodel, obs, prm, g["pso_x0"], int(g[tag + "_num_p"]), int(g[tag + "_seed"]))
        assert np.array_equal(bestp, g[tag + "_bestp"]) and best == float(g[tag + "_best"])


def test_oracle_golden_equals_reference_golden(gref):
    """the oracle-made fixture file and the reference-made one agree field by field"""
    a = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    shared = [k for k in gref.files if k in a.files]
    assert len(shared) >= 18
    for k in shared:
        assert np.array_equal(a[k], gref[k]), k


# ---------------------------------------------------------------- live comparison with the reference build
def test_fk_bit_exact(orc, omodel, R, D):
    rng = np.random.default_rng(11)
    S, J = R.fk(D.X0)
    # SURVEY B.4 sanity values
    assert np.allclose(S[0], [2.155593, -2.404259, -31.112313], atol=1e-6)
    assert np.allclose(S[47], [-7.823058, 4.986993, -25.010879], atol=1e-6)
    assert np.allclose(J[4], [2.754023, -5.852627, 21.259828], atol=1e-6)
    for k in range(300):
        th = D.X0 + rng.normal(size=26) * D.STD * (3.0 if k % 3 == 0 else 1.0)
        So, Jo = orc.fk(omodel, th)
        Sr, Jr = R.fk(th)
        assert np.array_equal(So, Sr) and np.array_equal(Jo, Jr)


def test_fk_and_costs_with_other_hand_geometries(orc, D):
    """not only the geometry of misc/hgeo.dat / rad.dat: random segment lengths, palm spacings, CMC angles and
    radii (the spacing goes through the reference's `float` member, src/fingermodel.h:43)"""
    import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(15)
    for trial in range(6):
        geo = D.HGEO_CM * rng.uniform(0.7, 1.3, size=20)
        spacing = D.SPACING + rng.normal(size=5) * 0.3
        cmc = D.CMC + rng.normal(size=5) * 8.0
        rad = D.RAD_CM * rng.uniform(0.8, 1.2, size=48)
        om = orc.make_model(geo_cm=geo, spacing=spacing, cmc=cmc, rad_cm=rad)
        Rg = ref.Reference(threads=1, geo_cm=geo, spacing=spacing, cmc=cmc, rad_cm=rad)
        truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
        depth = orc.render_depth(om, truth)
        oo = orc.Observation(depth)
        Rg.load_depth(depth)
        for k in range(40):
            th = truth + rng.normal(size=26) * D.STD * (0.2 if k % 2 else 1.5)
            So, Jo = orc.fk(om, th)
            Sr, Jr = Rg.fk(th)
            assert np.array_equal(So, Sr) and np.array_equal(Jo, Jr), (trial, k)
            tot, terms, _ = orc.cal_cost2(om, oo, th)
            t, _ = Rg.cost_terms(th)
            assert np.array_equal(terms, t), (trial, k)
        Rg.close()


def test_observation_bit_exact(orc, omodel, R, D):
    """load_data + depth_to_ptncloud + dist_transform through next_frame, three frames in a row
    (the frame is swapped under the live costfunc like src/testmodel.cpp:124)"""
    rng = np.random.default_rng(12)
    for k in range(3):
        truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.6, D.LB, D.UB)
        depth = orc.render_depth(omodel, truth)
        if k == 2:
            depth[:, :40] = 0; depth[200:, :] = 0          # cropped hand
        oo = orc.Observation(depth)
        assert R.load_depth(depth) == oo.n_pts == 250
        cloud, dcm, dt = R.observation()
        assert np.array_equal(cloud, oo.cloud) and np.array_equal(dcm, oo.depth) and np.array_equal(dt, oo.dt)
        assert R.scale == oo.scale
        th = truth + rng.normal(size=26) * D.STD * 0.5
        assert R.cal_cost(th) == orc.cal_cost(omodel, oo, th)


def test_cost_terms_bit_exact(orc, omodel, R, D):
    rng = np.random.default_rng(13)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(omodel, truth)
    oo = orc.Observation(depth)
    R.load_depth(depth)
    th = np.concatenate([truth[None], truth + rng.normal(size=(150, 26)) * D.STD,
                         truth + rng.normal(size=(50, 26)) * D.STD * 4.0,        # far away, out-of-image spheres
                         truth + rng.normal(size=(50, 26)) * D.STD * 0.05])
    a, d, c = orc.cost_batch(omodel, oo, th)
    for i in range(len(th)):
        t, m = R.cost_terms(th[i])
        _, _, mo = orc.cal_cost2(omodel, oo, th[i])
        assert np.array_equal(m, mo)
        assert t[0] == a[i] and t[1] == d[i] and t[2] == c[i]
        assert R.cal_cost(th[i]) == a[i] + d[i]
    # frozen correspondences
    _, m5 = R.cal_cost2(th[5])
    v, _ = R.cal_cost2(th[6], match=m5)
    vo, _, _ = orc.cal_cost2(omodel, oo, th[6], match=m5)
    assert v == vo


@pytest.mark.parametrize("num_p,maxiter,seed", [(16, 6, 1000), (64, 30, 1000), (33, 9, 77)])
def test_pso_evolve_on_replayed_streams(orc, omodel, R, D, num_p, maxiter, seed):
    """PSO::pso_evolve as written (informants, clamp to the LOWER bound, count/topology rule) ends at
    exactly the oracle's pose when both draw the same counter-based streams"""
    import ref
    rng = np.random.default_rng(14 + num_p)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(omodel, truth)
    oo = orc.Observation(depth)
    R.load_depth(depth)
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    R.set_pso_params(maxiter=maxiter)
    hook = ref.PhiloxReplay(seed)
    hook.num_p = num_p
    hook.install()
    try:
        b = R.pso_evolve(x0, num_p)
    finally:
        ref.PhiloxReplay.uninstall()
    assert hook.seeds_seen == [1000] and hook.gen == maxiter - 1     # arma_rng::set_seed(1000), src/PSO.cpp:722
    bo, co, _, _, _ = orc.pso_evolve(omodel, oo, orc.pso_params(maxiter=maxiter), x0, num_p, seed)
    assert np.array_equal(b, bo) and R.cal_cost(b) == co


def test_shim_primitives_against_numpy():
    """the stand-in Armadillo header itself, op by op, against numpy on random inputs: the dense product (left-to-right
    dot products: equal to numpy up to summation order), accu / mean / dot (Armadillo's two-accumulator order, restated
    here), min / max / floor / repmat / sum(A, dim) (exact).  The matcher and the chamfer transform of the OpenCV stand-in
    are pinned through test_observation_bit_exact / test_cost_terms_bit_exact and, bit for bit against cv2, in
    tests/test_oracle_pins.py."""
    import ctypes as C
    import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    L = ref.lib()
    for n in ("ref_create", "ref_load_frame", "ref_fk", "ref_cal_cost", "ref_cal_cost2", "ref_cost_terms",
              "ref_cost_batch", "ref_set_pso_params", "ref_pso_evolve", "ref_pso_optimise",
              "ref_refine_init_pose", "ref_gnd_truth_err", "ref_dim_restore", "ref_set_rng"):
        assert hasattr(L, n)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng = np.random.default_rng(71)
    # dense products of the shapes the hot files use (4x4 chains, 3x3 K times 3xN, Nx3 times 3x3)
    for m, k, n in ((4, 4, 4), (3, 3, 48), (250, 3, 3), (1, 26, 1), (7, 5, 9)):
        a, b = rng.normal(size=(m, k)) * 30, rng.normal(size=(k, n)) * 30
        out = np.empty((m, n))
        L.ref_shim_matmul(dp(a), m, k, dp(b), n, dp(out))
        want = a @ b
        assert np.allclose(out, want, rtol=1e-13, atol=1e-11 * np.abs(a).max() * np.abs(b).max())
        # the stated order: plain left-to-right dot products, no FMA (oracle/Makefile: -ffp-contract=off)
        ref_lr = np.array([[np.cumsum(a[i] * b[:, j])[-1] for j in range(n)] for i in range(m)])
        assert np.array_equal(out, ref_lr)

    def two_acc(x):        # arrayops::accumulate: even and odd elements in two accumulators, then their sum
        v1 = v2 = 0.0
        n2 = len(x) // 2 * 2
        for i in range(0, n2, 2):
            v1 += x[i]
            v2 += x[i + 1]
        if n2 < len(x):
            v1 += x[n2]
        return v1 + v2

    for n in (1, 2, 3, 48, 250, 1001):
        x = rng.normal(size=n) * 100
        out = np.empty(5)
        L.ref_shim_reduce(dp(x), n, dp(out))
        assert out[0] == two_acc(x) and abs(out[0] - np.sum(x)) <= 1e-12 * np.abs(x).sum()
        assert out[1] == two_acc(x) / n and abs(out[1] - x.mean()) <= 1e-12 * np.abs(x).mean() + 1e-300
        assert out[2] == x.min() and out[3] == x.max()
        assert out[4] == two_acc(x * x) and abs(out[4] - x @ x) <= 1e-12 * (x @ x)
    for m, k, nr, nc in ((1, 3, 5, 1), (6, 1, 1, 6), (3, 4, 2, 3)):
        a = rng.normal(size=(m, k)) * 50
        a.flat[0] = -3.0; a.flat[-1] = 7.0                          # exact integers: floor keeps them
        rep, fl = np.empty((m * nr, k * nc)), np.empty((m, k))
        L.ref_shim_repmat_floor(dp(a), m, k, nr, nc, dp(rep), dp(fl))
        assert np.array_equal(rep, np.tile(a, (nr, nc))) and np.array_equal(fl, np.floor(a))
        cs, rs = np.empty(k), np.empty(m)
        L.ref_shim_sums(dp(a), m, k, dp(cs), dp(rs))
        assert np.allclose(cs, a.sum(0), rtol=1e-13, atol=1e-11) and np.allclose(rs, a.sum(1), rtol=1e-13, atol=1e-11)


# ---------------------------------------------------------------- PSO::pso_optimise
def test_oracle_reproduces_reference_pso_optimise_fixtures(orc, omodel):
    """runs everywhere: orc_pso_optimise ends at exactly the pose the reference's PSO::pso_optimise returned
    (tests/golden/golden_ref_opt_v1.npz, written from oracle/_ref by make_golden_ref_opt.py)"""
    g = np.load(os.path.join(HERE, "golden", "golden_ref_opt_v1.npz"))
    for k, (num_p, maxiter, seed, w, c1, c2) in enumerate(g["cases"]):
        oo = orc.Observation(orc.render_depth(omodel, g["truth%d" % k]))
        bo, co, n, _, fc = orc.pso_optimise(omodel, oo, orc.pso_params(maxiter=int(maxiter)), g["x0_%d" % k], int(num_p),
                                            int(seed), w=w, c1=c1, c2=c2)
        assert np.array_equal(bo, g["bestp%d" % k])
        assert orc.cal_cost(omodel, oo, bo) == float(g["bestp_cost%d" % k])
        assert co == fc.min() and n > int(num_p) * int(maxiter)


@pytest.mark.parametrize("num_p,maxiter,seed,w", [(4, 3, 1000, 0.7298), (9, 4, 5, 0.5)])
def test_pso_optimise_on_replayed_streams(orc, omodel, R, D, num_p, maxiter, seed, w):
    """PSO::pso_optimise as written (src/PSO.cpp:539-712: ten coordinate-descent steps per particle with
    cal_gradient + goldstein on cal_cost2, correspondences recomputed twice in step 0 and frozen afterwards,
    gbest_pos = the CURRENT position of the best-pcost particle, canonical update with w, c1, c2) ends at
    exactly the oracle's pose on the same counter-based streams.  One OpenMP thread: the reference draws
    randi inside its parallel loop (:588) and races on gbest (:563-566)."""
    import ref
    rng = np.random.default_rng(40 + num_p)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(omodel, truth)
    oo = orc.Observation(depth)
    R.load_depth(depth)
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    R.set_pso_params(maxiter=maxiter, w=w, c1=1.3, c2=1.6)
    hook = ref.PhiloxReplayOptimise(seed)
    hook.num_p = num_p
    hook.install()
    try:
        b = R.pso_optimise(x0, num_p)
    finally:
        ref.PhiloxReplay.uninstall()
    assert hook.randi_calls == num_p * (maxiter - 1) and hook.gen == maxiter - 1
    bo, co, n, fx, fc = orc.pso_optimise(omodel, oo, orc.pso_params(maxiter=maxiter), x0, num_p, seed, w=w, c1=1.3, c2=1.6)
    assert np.array_equal(b, bo)
    assert co == fc.min() and n > num_p * maxiter


# ---------------------------------------------------------------- costfunc::bincomp_penalty
def test_oracle_reproduces_reference_bincomp_fixtures(orc, omodel):
    """runs everywhere: the silhouette-mismatch counts the reference's costfunc::bincomp_penalty returned"""
    g = np.load(os.path.join(HERE, "golden", "golden_ref_opt_v1.npz"))
    oo = orc.Observation(orc.render_depth(omodel, g["bincomp_truth"]))
    got = np.array([orc.bincomp_penalty(omodel, oo, p) for p in g["bincomp_poses"]])
    assert np.array_equal(got, g["bincomp"]) and got.min() > 0 and len(np.unique(got)) > 30


def test_bincomp_penalty_bit_exact(orc, omodel, R, D):
    """costfunc::bincomp_penalty as written (src/costfunc.cpp:384-436: centre and radius truncated, discs drawn by
    cv::circle, binarised, compared with depth != 0) against the oracle, incl. discs clipped by the image border"""
    rng = np.random.default_rng(60)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(omodel, truth)
    oo = orc.Observation(depth)
    R.load_depth(depth)
    for k in range(120):
        th = truth + rng.normal(size=26) * D.STD * [0.0, 0.1, 1.0, 3.0][k % 4]
        if k % 6 == 5:
            th[3:6] += rng.normal(size=3) * 15.0
        assert R.bincomp_penalty(th) == orc.bincomp_penalty(omodel, oo, th)
    v, sil = orc.bincomp_penalty(omodel, oo, truth, want_silhouette=True)
    assert v == np.count_nonzero((sil != 0) != (depth != 0))


def test_depthmatch_penalty_bit_exact(orc, omodel, D):
    """costfunc::depthMatch_penalty (src/costfunc.cpp:439-473) -- the reference's own function, compiled from its source
    against the stand-in headers -- returns exactly the image orc_depthmatch_penalty returns: near the truth, far from it,
    outlines clipped by the border and outside the image"""
    import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(91)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(omodel, truth)
    oo = orc.Observation(depth)
    R = ref.Reference(threads=1)
    R.load_depth(depth)
    for k in range(30):
        th = truth + rng.normal(size=26) * D.STD * (0.05, 1.0, 3.0)[k % 3]
        a, b = orc.depthmatch_penalty(omodel, oo, th), R.depthmatch_penalty(th)
        assert np.array_equal(a, b), (k, int((a != b).sum()))
        if k % 3 == 0:
            assert (a != oo.depth).sum() > 50         # near the truth the outlines are inside the image
