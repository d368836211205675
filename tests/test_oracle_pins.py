"""CPU: pins of the oracle (oracle/hpe_oracle.c).  The reference has no tests or golden vectors and
cannot be built here, so the oracle is pinned against (a) the third-party routines the reference
calls, through Python cv2, (b) an independent numpy restatement, (c) the sanity values the survey
recorded, (d) the committed golden file (drift guard)."""
import os

import numpy as np
import pytest

import np_restatement as R
from conftest import synth_pose

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


def golden_frame(g):
    fr = np.zeros((240, 320), dtype=np.float32)
    fr[g["frame_y"], g["frame_x"]] = g["frame_mm"]
    return fr


# ---- (a) third-party arithmetic -------------------------------------------------------------
def test_matcher_equals_cv2_bfmatcher(orc, omodel, D):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for trial in range(5):
        th = D.X0 + rng.normal(size=26) * D.STD
        S, _ = orc.fk(omodel, th)
        P = S[rng.integers(0, 48, size=250)] + rng.normal(size=(250, 3)) * (0.02 if trial % 2 else 1.5)
        m = cv2.BFMatcher(cv2.NORM_L2).match(P.astype(np.float32), S.astype(np.float32))
        assert np.array_equal(orc.correspondences(P, S), np.array([x.trainIdx for x in m]))


def test_matcher_tie_semantics_equal_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    # duplicated train rows: cv2 keeps the first of equal distances, so must the oracle
    S = np.tile(np.array([[1.0, 2.0, -30.0], [1.5, 2.0, -30.0], [1.0, 2.0, -30.0]]), (16, 1))
    P = np.array([[1.0, 2.1, -30.0], [1.5, 1.9, -29.0], [1.25, 2.0, -30.0]] * 10)
    m = cv2.BFMatcher(cv2.NORM_L2).match(P.astype(np.float32), S.astype(np.float32))
    assert np.array_equal(orc.correspondences(P, S), np.array([x.trainIdx for x in m]))


def test_chamfer_equals_cv2_distance_transform(orc, omodel, D):
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False)        # OpenCV's own fixed-point path (the IPP build differs by ~6e-4 px)
    rng = np.random.default_rng(3)
    masks = []
    for _ in range(4):
        depth = orc.render_depth(omodel, synth_pose(D, rng))
        masks.append(np.where(depth == 0, 255, 0).astype(np.uint8))
    for k in range(6):
        masks.append(((rng.random((240, 320)) > 0.002 * (k + 1)) * 255).astype(np.uint8))
    for corner in ((0, 0), (239, 319), (0, 319), (120, 160)):
        m = np.full((240, 320), 255, np.uint8)
        m[corner] = 0
        masks.append(m)
    for m in masks:
        assert np.array_equal(cv2.distanceTransform(m, cv2.DIST_L2, 5), orc.chamfer5(m))


def test_chamfer_closed_form(orc):
    """min over feature pixels of the Q16 chamfer metric == the two-pass result (what the GPU uses)"""
    rng = np.random.default_rng(5)
    for dens in (0.0005, 0.01):
        mask = ((rng.random((240, 320)) > dens) * 255).astype(np.uint8)
        ys, xs = np.nonzero(mask == 0)
        Y, X = np.mgrid[0:240, 0:320]
        best = np.full((240, 320), 2 ** 62, dtype=np.int64)
        for y, x in zip(ys, xs):
            dx, dy = abs(X - x), abs(Y - y)
            mx, mn = np.maximum(dx, dy), np.minimum(dx, dy)
            best = np.minimum(best, np.where(2 * mn <= mx, mn * 143976 + (mx - 2 * mn) * 65536,
                                             (mx - mn) * 143976 + (2 * mn - mx) * 91750))
        ref = best.astype(np.uint32).astype(np.float32) * np.float32(1 / 65536)
        assert np.array_equal(ref, orc.chamfer5(mask))


# ---- (b) independent restatement, (c) survey values -------------------------------------------
def test_fk_survey_values(orc, omodel, D):
    S, J = orc.fk(omodel, D.X0)
    expect = {0: (2.155593, -2.404259, -31.112313), 7: (7.792197, -0.585899, -26.038955),
              8: (1.831466, -2.997341, -31.674797), 17: (2.754023, 5.852627, -21.259828),
              47: (-7.823058, 4.986993, -25.010879)}
    for i, v in expect.items():
        assert np.allclose(S[i], v, atol=1e-6)
    assert np.allclose(J[0], (0, 3, 32)) and np.allclose(J[4], (2.754023, -5.852627, 21.259828), atol=1e-6)
    assert np.allclose(J[20], (7.792197, 0.585899, 26.038955), atol=1e-6)


def test_fk_equals_numpy_restatement(orc, omodel, D):
    rng = np.random.default_rng(1)
    th = np.concatenate([D.X0[None], D.X0 + rng.normal(size=(300, 26)) * D.STD, D.LB[None], D.UB[None]])
    S, J = orc.fk(omodel, th)
    for i in range(len(th)):
        S2, J2 = R.hand(th[i], D.HGEO_CM, D.SPACING, D.CMC)
        assert np.abs(S[i] - S2).max() < 1e-12 and np.abs(J[i] - J2).max() < 1e-12


def test_fk_structure(orc, omodel, D):
    """sphere layout of src/handmodel.cpp:282-286 and the lerp positions of buildSpheres"""
    rng = np.random.default_rng(2)
    th = D.X0 + rng.normal(size=26) * D.STD
    S, J = orc.fk(omodel, th)
    U = S * np.array([1, -1, -1])
    assert np.allclose(U[17], J[4]) and np.allclose(U[27], J[8]) and np.allclose(U[47], J[16]) and np.allclose(U[7], J[20])
    assert np.allclose(U[11], J[1])                      # 4th palm sphere of the index sits on joint 1
    assert np.allclose(U[9] - U[8], U[10] - U[9])        # equally spaced palm spheres
    assert np.allclose(U[12], (J[1] + J[2]) / 2)
    # segment lengths are the geometry file's
    for d in range(1, 5):
        jj = J[1 + 4 * (d - 1):5 + 4 * (d - 1)]
        assert np.allclose(np.linalg.norm(np.diff(jj, axis=0), axis=1), D.HGEO_CM[4 * d + 1:4 * d + 4])
    # global translation moves everything rigidly
    th2 = th.copy(); th2[3:6] += (1.0, -2.0, 3.0)
    S2, _ = orc.fk(omodel, th2)
    assert np.allclose(S2 - S, np.array([1.0, 2.0, -3.0]))


def test_cost_terms_equal_numpy_restatement(orc, omodel, frame0, D):
    truth, depth, obs = frame0
    rng = np.random.default_rng(4)
    th = np.concatenate([truth[None], truth + rng.normal(size=(40, 26)) * D.STD,
                         truth + rng.normal(size=(20, 26)) * D.STD * 0.1])
    a, d, c = orc.cost_batch(omodel, obs, th)
    cloud, dcm, dt = obs.cloud, obs.depth, obs.dt
    for i in range(len(th)):
        a2, d2, c2 = R.cost_terms(th[i], D.HGEO_CM, D.SPACING, D.CMC, D.RAD_CM, cloud, dcm, dt, obs.scale)
        assert abs(a[i] - a2) <= 1e-10 * max(a2, 1) and abs(d[i] - d2) <= 1e-10 * max(d2, 1)
        assert abs(c[i] - c2) <= 1e-10 * max(c2, 1)
    assert orc.cal_cost(omodel, obs, truth) == a[0] + d[0]
    tot, terms, m = orc.cal_cost2(omodel, obs, th[3])
    assert tot == terms.sum() and np.allclose(terms, (a[3], d[3], c[3]), rtol=1e-14)
    # frozen correspondences change only the alignment term
    tot2, terms2, _ = orc.cal_cost2(omodel, obs, th[4], match=m)
    assert terms2[1] == d[4] and terms2[2] == c[4]


def test_observation_against_numpy(orc, omodel, frame0, D):
    truth, depth, obs = frame0
    ys, xs = np.nonzero(depth)                       # raster order
    Z = depth[ys, xs].astype(np.float64) / 10.0
    X = (xs - 160.0) * Z / D.FOCAL
    Y = (ys - 120.0) * Z / D.FOCAL
    n = len(Z)
    assert obs.n_hand == n
    step = n // 250
    idx = np.arange(250) * step
    assert np.array_equal(obs.cloud, np.stack([X, -Y, -Z], 1)[idx])
    cu = np.floor((D.FOCAL * X + 160.0 * Z) / Z)
    eu = np.floor((D.FOCAL * (X + 2.0) + 160.0 * Z) / Z)
    dist = np.abs(eu - cu)
    assert abs(obs.scale - np.mean(2.0 / dist[dist != 0])) < 1e-14
    assert np.array_equal(obs.depth, depth.astype(np.float64) / 10.0)
    assert obs.dt[depth != 0].max() == 0 and obs.dt_max == obs.dt.max()
    assert 0.08 < obs.scale < 0.16                   # ~Z/f at 20-35 cm


def test_observation_edge_cases(orc, D):
    fr = np.zeros((240, 320), np.float32)
    with pytest.raises(ValueError):
        orc.Observation(fr)                          # the reference would index an empty matrix
    fr[10, 20] = 300.0
    o = orc.Observation(fr)                          # fewer pixels than samples: stride 0, all = pixel 0
    assert o.n_hand == 1 and np.all(o.cloud == o.cloud[0])
    assert o.dt[10, 20] == 0 and o.dt[10, 21] == 1.0 and abs(o.dt[11, 21] - 1.4) < 1e-4


# ---- (d) golden file -----------------------------------------------------------------------------
def test_oracle_reproduces_golden(orc, omodel, golden, D):
    g = golden
    fr = golden_frame(g)
    obs = orc.Observation(fr)
    assert obs.n_hand == int(g["n_hand"]) and np.array_equal(obs.cloud, g["cloud"])
    assert abs(obs.scale - float(g["scale"])) < 1e-15 and obs.dt_max == float(g["dt_max"])
    assert np.array_equal(obs.dt[::16].astype(np.float32), g["dt_rows"])
    S, J = orc.fk(omodel, g["poses"])
    assert np.abs(S - g["spheres"]).max() < 1e-12 and np.abs(J - g["joints"]).max() < 1e-12
    a, d, c = orc.cost_batch(omodel, obs, g["poses"])
    assert np.allclose(a, g["align"], rtol=1e-12) and np.allclose(d, g["depth_term"], rtol=1e-12, atol=1e-15)
    assert np.allclose(c, g["colli"], rtol=1e-12, atol=1e-15)
    for i in range(len(g["poses"])):
        assert np.array_equal(orc.correspondences(obs.cloud, S[i]), g["match"][i])
    prm = orc.pso_params(maxiter=int(g["pso_maxiter"]))
    bp, best, tr, _, _ = orc.pso_evolve(omodel, obs, prm, g["pso_x0"], int(g["pso_num_p"]), int(g["pso_seed"]))
    assert np.allclose(tr, g["pso_trace"], rtol=1e-10) and np.allclose(bp, g["pso_bestp"], atol=1e-8)
    assert list(g["philox_kat"]) == orc.philox(0x0123456789abcdef, 1, 2, 3, 4)


def test_fill_circle_equals_cv2_circle(orc):
    """cv::circle(..., thickness=-1), the third-party call of costfunc::bincomp_penalty (src/costfunc.cpp:423):
    the oracle's restatement of OpenCV's midpoint walk covers exactly cv2.circle's pixels, also when the disc
    is clipped by the border or lies outside, down to radius 0"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    for t in range(1500):
        w, h = (320, 240) if t % 3 else (int(rng.integers(1, 90)), int(rng.integers(1, 90)))
        r = int(rng.integers(0, 70 if t % 5 else 4))
        cx, cy = int(rng.integers(-90, w + 90)), int(rng.integers(-90, h + 90))
        img = np.zeros((h, w), dtype=np.float64)
        cv2.circle(img, (cx, cy), r, 7.5, -1)
        assert np.array_equal(img != 0, orc.fill_circle(w, h, cx, cy, r) != 0), (w, h, cx, cy, r)


def test_circle_outline_equals_cv2_circle(orc):
    """cv::circle(..., thickness=2), the third-party call of costfunc::depthMatch_penalty (src/costfunc.cpp:468): OpenCV
    draws an outline of thickness > 1 through EllipseEx -> ellipse2Poly -> PolyLine -> ThickLine -> FillConvexPoly (+ Line2
    after clipLine, + filled end caps); the oracle's restatement (oracle/orc_draw.h) covers exactly cv2.circle's pixels on
    a CV_64F image -- clipped by the border, outside the image, radius 0, small images -- for thickness 2 (the
    reference's) and 3"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(18)
    for t in range(2500):
        w, h = (320, 240) if t % 3 else (int(rng.integers(1, 90)), int(rng.integers(1, 90)))
        r = int(rng.integers(0, 90 if t % 5 else 4))
        cx, cy = int(rng.integers(-120, w + 120)), int(rng.integers(-120, h + 120))
        th = 2 if t % 4 else 3
        img = np.zeros((h, w), dtype=np.float64)
        cv2.circle(img, (cx, cy), r, 7.5, th)
        assert np.array_equal(img != 0, orc.circle_outline(w, h, cx, cy, r, th) != 0), (w, h, cx, cy, r, th)
    # every radius the sphere model can produce at the reference's depths, centred
    for r in range(0, 140):
        img = np.zeros((240, 320), dtype=np.float64)
        cv2.circle(img, (160, 120), r, 1.0, 2)
        assert np.array_equal(img != 0, orc.circle_outline(320, 240, 160, 120, r, 2) != 0), r
