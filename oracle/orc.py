"""ctypes wrapper over oracle/libhpe_oracle.so -- the CPU oracle: test infrastructure only.
Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by the product package."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
from hpe_b200 import defaults as D  # noqa: E402

_LIB = None
W, H = 320, 240


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)


class Model(C.Structure):
    _fields_ = [("geo", C.c_double * 20), ("spacing", C.c_float * 5), ("cmc", C.c_double * 5),
                ("tb_n", C.c_int * 4), ("fg_n", C.c_int * 4), ("rad", C.c_double * 48),
                ("base", C.c_double * 80), ("back", C.c_double * 80)]


class Obs(C.Structure):
    _fields_ = [("n_hand", C.c_int), ("n_pts", C.c_int), ("scale", C.c_double),
                ("focal", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("dt_max", C.c_double), ("depth", C.POINTER(C.c_double)),
                ("dt", C.POINTER(C.c_double)), ("cloud", C.POINTER(C.c_double))]


class PsoParams(C.Structure):
    _fields_ = [("lb", C.c_double * 26), ("ub", C.c_double * 26), ("std", C.c_double * 26),
                ("maxiter", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libhpe_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(
                os.path.join(ROOT, "oracle", "hpe_oracle.c")):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_cal_cost.restype = C.c_double
        _LIB.orc_cal_cost2.restype = C.c_double
        _LIB.orc_align.restype = C.c_double
        _LIB.orc_depth_penalty.restype = C.c_double
        _LIB.orc_collision.restype = C.c_double
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_model(geo_cm=None, spacing=None, cmc=None, tb_n=None, fg_n=None, rad_cm=None):
    m = Model()
    geo_cm = f64(D.HGEO_CM if geo_cm is None else geo_cm)
    spacing = np.ascontiguousarray(D.SPACING if spacing is None else spacing, dtype=np.float32)
    cmc = f64(D.CMC if cmc is None else cmc)
    tb_n = np.ascontiguousarray(D.TB_N if tb_n is None else tb_n, dtype=np.int32)
    fg_n = np.ascontiguousarray(D.FG_N if fg_n is None else fg_n, dtype=np.int32)
    rad_cm = f64(D.RAD_CM if rad_cm is None else rad_cm)
    lib().orc_model_init(C.byref(m), _p(geo_cm), _p(spacing, C.c_float), _p(cmc),
                         _p(tb_n, C.c_int), _p(fg_n, C.c_int), _p(rad_cm))
    return m


def fk(model, theta):
    theta = f64(theta)
    single = theta.ndim == 1
    th = theta.reshape(-1, 26)
    S = np.empty((th.shape[0], 48, 3))
    J = np.empty((th.shape[0], 21, 3))
    for i in range(th.shape[0]):
        lib().orc_build_hand_model(C.byref(model), _p(th[i]), _p(S[i]), _p(J[i]))
    return (S[0], J[0]) if single else (S, J)


class Observation:
    """Owns an orc_obs; numpy views are copies."""

    def __init__(self, depth_mm, focal=D.FOCAL, downsample=True, num_samples=D.NUM_SAMPLES):
        self.o = Obs()
        depth_mm = np.ascontiguousarray(depth_mm, dtype=np.float32).reshape(H, W)
        rc = lib().orc_obs_from_depth(_p(depth_mm, C.c_float), C.c_double(focal), int(downsample),
                                      int(num_samples), C.byref(self.o))
        if rc != 0:
            raise ValueError("frame has no hand pixel")
        self.depth_mm = depth_mm

    def __del__(self):
        try:
            lib().orc_obs_free(C.byref(self.o))
        except Exception:
            pass

    n_pts = property(lambda s: s.o.n_pts)
    n_hand = property(lambda s: s.o.n_hand)
    scale = property(lambda s: s.o.scale)
    dt_max = property(lambda s: s.o.dt_max)

    @property
    def cloud(self):
        return np.ctypeslib.as_array(self.o.cloud, shape=(self.o.n_pts, 3)).copy()

    @property
    def dt(self):
        return np.ctypeslib.as_array(self.o.dt, shape=(H, W)).copy()

    @property
    def depth(self):
        return np.ctypeslib.as_array(self.o.depth, shape=(H, W)).copy()


def chamfer5(mask):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    out = np.empty((h, w), dtype=np.float32)
    lib().orc_chamfer5(_p(mask, C.c_uint8), w, h, _p(out, C.c_float))
    return out


def correspondences(cloud, spheres):
    cloud, spheres = f64(cloud), f64(spheres)
    m = np.empty(cloud.shape[0], dtype=np.int32)
    lib().orc_correspondences(_p(cloud), cloud.shape[0], _p(spheres), _p(m, C.c_int32))
    return m


def cal_cost(model, obs, theta):
    theta = f64(theta)
    return lib().orc_cal_cost(C.byref(model), C.byref(obs.o), _p(theta))


def cal_cost2(model, obs, theta, match=None):
    theta = f64(theta)
    terms = np.empty(3)
    compute = match is None
    if compute:
        match = np.empty(obs.n_pts, dtype=np.int32)
    else:
        match = np.ascontiguousarray(match, dtype=np.int32)
    tot = lib().orc_cal_cost2(C.byref(model), C.byref(obs.o), _p(theta), _p(match, C.c_int32),
                              int(compute), _p(terms))
    return tot, terms, match


def cost_batch(model, obs, theta, diagnostics=False, collision=True):
    theta = f64(theta).reshape(-1, 26)
    n = theta.shape[0]
    a, d = np.empty(n), np.empty(n)
    c = np.empty(n) if collision else None
    fm = np.empty(n) if diagnostics else None
    mg = np.empty(n) if diagnostics else None
    lib().orc_cost_batch(C.byref(model), C.byref(obs.o), _p(theta), C.c_int64(n), _p(a), _p(d),
                         _p(c) if collision else None,
                         _p(fm) if diagnostics else None, _p(mg) if diagnostics else None)
    if c is None:
        c = np.zeros(n)
    return (a, d, c, fm, mg) if diagnostics else (a, d, c)


def pso_params(lb=None, ub=None, std=None, maxiter=30):
    p = PsoParams()
    p.lb[:] = list(D.LB if lb is None else lb)
    p.ub[:] = list(D.UB if ub is None else ub)
    p.std[:] = list(D.STD if std is None else std)
    p.maxiter = int(maxiter)
    return p


def generate_particles(params, x0, num_p, seed):
    x0 = f64(x0)
    out = np.empty((num_p, 26))
    lib().orc_generate_particles(C.byref(params), _p(x0), num_p, C.c_uint64(seed), _p(out))
    return out


def check_constraints(params, theta, vel):
    theta, vel = f64(theta).copy(), f64(vel).copy()
    lib().orc_check_constraints(C.byref(params), _p(theta), _p(vel))
    return theta, vel


def pso_evolve(model, obs, params, x0, num_p, seed):
    x0 = f64(x0)
    bestp = np.empty(26)
    best = C.c_double()
    trace = np.empty(max(params.maxiter - 1, 1))
    fin_x = np.empty((num_p, 26))
    fin_c = np.empty(num_p)
    lib().orc_pso_evolve(C.byref(model), C.byref(obs.o), C.byref(params), _p(x0), num_p,
                         C.c_uint64(seed), _p(bestp), C.byref(best), _p(trace), _p(fin_x), _p(fin_c))
    return bestp, best.value, trace[:params.maxiter - 1], fin_x, fin_c


def pso_optimise(model, obs, params, x0, num_p, seed, w=0.7298, c1=1.49618, c2=1.49618):
    """PSO::pso_optimise (src/PSO.cpp:539-712); returns (bestp, gbest_cost, evaluations, final particles, final pcost)"""
    x0 = f64(x0)
    bestp = np.empty(26)
    best = C.c_double()
    n = C.c_int64()
    fin_x = np.empty((num_p, 26))
    fin_c = np.empty(num_p)
    lib().orc_pso_optimise.restype = C.c_int
    lib().orc_margin_min.restype = C.c_double
    lib().orc_margin_reset()
    lib().orc_pso_optimise(C.byref(model), C.byref(obs.o), C.byref(params), C.c_double(w), C.c_double(c1),
                           C.c_double(c2), _p(x0), int(num_p), C.c_uint64(seed), _p(bestp), C.byref(best),
                           C.byref(n), _p(fin_x), _p(fin_c))
    pso_optimise.last_margin = lib().orc_margin_min()      # smallest relative gap of a line-search comparison in this run
    return bestp, best.value, n.value, fin_x, fin_c


def fill_circle(w, h, cx, cy, radius):
    """cv::circle(..., thickness=-1) restated (OpenCV drawing.cpp Circle); returns the uint8 mask (h, w)"""
    mask = np.zeros((h, w), dtype=np.uint8)
    lib().orc_fill_circle(_p(mask, C.c_uint8), int(w), int(h), int(cx), int(cy), int(radius))
    return mask


def bincomp_penalty(model, obs, theta, want_silhouette=False):
    """costfunc::bincomp_penalty (src/costfunc.cpp:384-436) of the spheres of pose theta"""
    S, _ = fk(model, theta)
    sil = np.zeros((H, W), dtype=np.uint8) if want_silhouette else None
    lib().orc_bincomp_penalty.restype = C.c_double
    v = lib().orc_bincomp_penalty(C.byref(model), C.byref(obs.o), _p(f64(S)), _p(sil, C.c_uint8) if want_silhouette else None)
    return (v, sil) if want_silhouette else v


def circle_outline(w, h, cx, cy, radius, thickness=2):
    """cv::circle(..., thickness >= 2) restated (OpenCV drawing.cpp EllipseEx / ThickLine / FillConvexPoly); uint8 mask (h, w)"""
    mask = np.zeros((h, w), dtype=np.uint8)
    lib().orc_circle_outline(_p(mask, C.c_uint8), int(w), int(h), int(cx), int(cy), int(radius), int(thickness))
    return mask


def depthmatch_penalty(model, obs, theta):
    """costfunc::depthMatch_penalty (src/costfunc.cpp:439-473): depth map (cm) with the outlines of the projected spheres"""
    S, _ = fk(model, theta)
    out = np.empty((H, W))
    lib().orc_depthmatch_penalty(C.byref(model), C.byref(obs.o), _p(f64(S)), _p(out))
    return out


def philox(seed, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(C.c_uint64(seed), c0, c1, c2, c3, out)
    return list(out)


def uniform2(seed, elem_pair, particle, gen, stream):
    u = (C.c_double * 2)()
    lib().orc_uniform2(C.c_uint64(seed), elem_pair, particle, gen, stream, u)
    return u[0], u[1]


def normal2(seed, elem_pair, particle, gen):
    z = (C.c_double * 2)()
    lib().orc_normal2(C.c_uint64(seed), elem_pair, particle, gen, z)
    return z[0], z[1]


def render_depth(model, theta, focal=D.FOCAL):
    """Synthetic Subject1-format frame (float32 mm, 240x320) of the sphere model at theta."""
    S, _ = fk(model, theta)
    S = S.copy()
    S[:, 1:] *= -1
    rad = f64(np.array(model.rad[:]))
    out = np.empty((H, W), dtype=np.float32)
    lib().orc_render_depth(_p(f64(S)), _p(rad), C.c_double(focal), _p(out, C.c_float))
    return out


def num_threads():
    return lib().orc_num_threads()


def refine_init_pose(model, obs, x0, want_margin=False):
    """PSO::refine_init_pose (src/PSO.cpp:216-266); returns (refined pose, number of cal_cost2 calls[, smallest relative gap
    of a Goldstein / Armijo comparison during the run])"""
    x = f64(x0).copy()
    n = C.c_int()
    lib().orc_margin_min.restype = C.c_double
    lib().orc_margin_reset()
    lib().orc_refine_init_pose(C.byref(model), C.byref(obs.o), _p(x), C.byref(n))
    return (x, n.value, lib().orc_margin_min()) if want_margin else (x, n.value)


def dim_restore(t22):
    o = np.empty(26)
    lib().orc_dim_restore(_p(f64(t22)), _p(o))
    return o


def gnd_truth_err(joints, gt_row):
    lib().orc_gnd_truth_err.restype = C.c_double
    return lib().orc_gnd_truth_err(_p(f64(joints)), _p(f64(gt_row)))
