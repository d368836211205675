"""ctypes wrapper over oracle/_ref/libhpe_ref.so -- the REFERENCE'S OWN hot-path sources, compiled
unmodified from /root/reference/src against the stand-in Armadillo/OpenCV headers of oracle/shim/
(see oracle/Makefile and oracle/ref_driver.cpp).  Test infrastructure only: imported by tests/,
tests/golden/make_golden_ref.py and bench.py's CPU-baseline / --impl reference legs; never by the
product package.  The .so is built in the development container (where /root/reference exists) and
travels to the GPU box with the snapshot; available() says whether it is there."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

import orc
from orc import D, f64, _p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libhpe_ref.so")
_LIB = None

FILL_T = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double))
SEED_T = C.CFUNCTYPE(None, C.c_uint)


def available():
    if os.path.exists(SO):
        return True
    if os.path.exists("/root/reference/src/PSO.cpp"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libhpe_ref.so is not built (needs /root/reference)")
        _LIB = C.CDLL(SO)
        _LIB.ref_create.restype = C.c_void_p
        for n in ("ref_cal_cost", "ref_cal_cost2", "ref_obs_scale", "ref_gnd_truth_err", "ref_bincomp_penalty"):
            getattr(_LIB, n).restype = C.c_double
    return _LIB


def write_frame(path, depth_mm):
    """Subject1 on-disk format: raw little-endian float32[240*320], mm, 0 = background
    (src/observedmodel.cpp:279-308)."""
    np.ascontiguousarray(depth_mm, dtype="<f4").reshape(240, 320).tofile(path)


class Reference:
    """handmodel + observedmodel + costfunc + PSO of the reference, constructed with the values of
    src/testmodel.cpp:34-52 unless overridden."""

    def __init__(self, threads=1, geo_cm=None, spacing=None, cmc=None, rad_cm=None):
        L = lib()
        self.h = C.c_void_p(L.ref_create(_p(f64(D.HGEO_CM if geo_cm is None else geo_cm)),
                                         _p(f64(D.SPACING if spacing is None else spacing)),
                                         _p(f64(D.CMC if cmc is None else cmc)),
                                         _p(f64(D.TB_N)), _p(f64(D.FG_N)),
                                         _p(f64(D.RAD_CM if rad_cm is None else rad_cm))))
        self.n_pts = 0
        self._tmp = tempfile.TemporaryDirectory(prefix="hpe_ref_")
        self._frame_no = 0
        L.ref_set_threads(int(threads))

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None
            self._tmp.cleanup()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def set_threads(n):
        lib().ref_set_threads(int(n))

    @staticmethod
    def max_threads():
        return lib().ref_max_threads()

    # ---- observation ----
    def load_depth(self, depth_mm, downsample=True):
        """writes the frame in Subject1 format and loads it through init_observation / next_frame"""
        name = "%06d_depth.bin" % self._frame_no
        self._frame_no += 1
        write_frame(os.path.join(self._tmp.name, name), depth_mm)
        self.n_pts = lib().ref_load_frame(self.h, (self._tmp.name + "/").encode(), name.encode(), int(downsample))
        return self.n_pts

    @property
    def scale(self):
        return lib().ref_obs_scale(self.h)

    def observation(self):
        cloud = np.empty((self.n_pts, 3))
        depth = np.empty((240, 320))
        dt = np.empty((240, 320))
        lib().ref_obs_get(self.h, _p(cloud), _p(depth), _p(dt))
        return cloud, depth, dt

    # ---- model / energy ----
    def fk(self, theta):
        S, J = np.empty((48, 3)), np.empty((21, 3))
        lib().ref_fk(self.h, _p(f64(theta)), _p(S), _p(J))
        return S, J

    def cal_cost(self, theta):
        return lib().ref_cal_cost(self.h, _p(f64(theta)))

    def cal_cost2(self, theta, match=None):
        compute = match is None
        m = np.zeros(self.n_pts, dtype=np.int32) if compute else np.ascontiguousarray(match, dtype=np.int32).copy()
        v = lib().ref_cal_cost2(self.h, _p(f64(theta)), _p(m, C.c_int), self.n_pts, int(compute))
        return v, m

    def cost_terms(self, theta):
        t = np.empty(3)
        m = np.zeros(self.n_pts, dtype=np.int32)
        lib().ref_cost_terms(self.h, _p(f64(theta)), _p(t), _p(m, C.c_int))
        return t, m

    def depthmatch_penalty(self, theta):
        """costfunc::depthMatch_penalty (src/costfunc.cpp:439-473): the depth map (cm) with the sphere outlines, (240, 320)"""
        out = np.empty((240, 320))
        lib().ref_depthmatch_penalty(self.h, _p(f64(theta)), _p(out))
        return out

    def bincomp_penalty(self, theta):
        """costfunc::bincomp_penalty (src/costfunc.cpp:384-436) of the spheres build_hand_model gives for theta"""
        return lib().ref_bincomp_penalty(self.h, _p(f64(theta)))

    def cost_batch(self, theta):
        theta = f64(theta).reshape(-1, 26)
        out = np.empty(theta.shape[0])
        lib().ref_cost_batch(self.h, _p(theta), C.c_long(theta.shape[0]), _p(out))
        return out

    # ---- optimiser ----
    def set_pso_params(self, ub=None, lb=None, std=None, w=0.7298, c1=1.49618, c2=1.49618, maxiter=30,
                       minstep=1e-8, minfunc=1e-8):
        lib().ref_set_pso_params(self.h, _p(f64(D.UB if ub is None else ub)), _p(f64(D.LB if lb is None else lb)),
                                 _p(f64(D.STD if std is None else std)), C.c_double(w), C.c_double(c1),
                                 C.c_double(c2), int(maxiter), C.c_double(minstep), C.c_double(minfunc))

    def pso_evolve(self, x0, num_p):
        b = np.empty(26)
        lib().ref_pso_evolve(self.h, _p(f64(x0)), int(num_p), _p(b))
        return b

    def pso_optimise(self, x0, num_p):
        b = np.empty(26)
        lib().ref_pso_optimise(self.h, _p(f64(x0)), int(num_p), _p(b))
        return b

    def refine_init_pose(self, x0):
        x = f64(x0).copy()
        lib().ref_refine_init_pose(self.h, _p(x))
        return x

    def dim_restore(self, t22):
        o = np.empty(26)
        lib().ref_dim_restore(self.h, _p(f64(t22)), _p(o))
        return o

    def gnd_truth_err(self, gt, frame):
        gt = f64(gt).reshape(-1, 63)
        return lib().ref_gnd_truth_err(self.h, _p(gt), gt.shape[0], int(frame))


class PhiloxReplay:
    """RNG hook that feeds the reference's randn/randu calls the counter-based streams of
    DESIGN.md "RNG" (the ones the oracle and the GPU draw), recognising each call of
    PSO::pso_evolve by its kind and shape, in the fixed order of src/PSO.cpp:67,781-793:
      randn(26, N)            -> stream 0 (initial normals)
      randu(26, N) twice      -> rp (stream 1) then rg (stream 2) of the next generation
      randu(N), randi(N), randu(N) -> rsu, id, sel: drawn by the reference, never used; zeros
      randu(3 N)              -> topology R (stream 3) of the current generation
    """

    def __init__(self, seed):
        self.seed = int(seed)
        self.gen = 0
        self.toggle = 0
        self.seeds_seen = []
        self._fill = FILL_T(self.fill)
        self._seed = SEED_T(self.reseed)

    def install(self):
        lib().ref_set_rng(self._fill, self._seed)

    @staticmethod
    def uninstall():
        lib().ref_set_rng(FILL_T(), SEED_T())

    def reseed(self, s):
        self.seeds_seen.append(int(s))
        self.gen = 0
        self.toggle = 0

    def fill(self, kind, rows, cols, a, b, out):
        n = rows * cols
        buf = np.ctypeslib.as_array(out, shape=(n,))
        if kind == 1 and rows == 26:
            for i in range(cols):
                for d in range(0, 26, 2):
                    buf[i * 26 + d], buf[i * 26 + d + 1] = orc.normal2(self.seed, d >> 1, i, 0)
        elif kind == 0 and rows == 26:
            if self.toggle == 0:
                self.gen += 1
            stream = 1 + self.toggle
            self.toggle ^= 1
            for i in range(cols):
                for d in range(0, 26, 2):
                    buf[i * 26 + d], buf[i * 26 + d + 1] = orc.uniform2(self.seed, d >> 1, i, self.gen, stream)
        elif kind == 0 and cols == 1 and rows % 3 == 0 and self._is_topology(rows):
            num_p = rows // 3
            for s in range(num_p):
                for k in range(3):
                    buf[s * 3 + k] = orc.uniform2(self.seed, k >> 1, s, self.gen, 3)[k & 1]
        elif kind == 2:
            buf[:] = a
        else:
            buf[:] = 0.0

    def _is_topology(self, rows):
        # rsu / sel are randu(N); the topology draw is randu(3 N) and comes right after them
        return rows == 3 * self.num_p

    num_p = 0


class PhiloxReplayOptimise(PhiloxReplay):
    """the same for PSO::pso_optimise (src/PSO.cpp:539-712), run with ONE OpenMP thread so that the
    per-particle randi(10, [0,25]) draws of :588 arrive in particle order: the k-th such call after
    the last rp/rg pair belongs to particle k of the coming generation and is fed stream 4 of
    oracle/hpe_oracle.c:orc_pso_optimise; randn(26,N) / randu(26,N) as in the parent."""

    def __init__(self, seed):
        super().__init__(seed)
        self.particle = 0
        self.randi_calls = 0

    def reseed(self, s):
        super().reseed(s)
        self.particle = 0

    def fill(self, kind, rows, cols, a, b, out):
        if kind == 2 and rows * cols == 10:
            buf = np.ctypeslib.as_array(out, shape=(10,))
            i, gen = self.particle, self.gen + 1
            for m in range(10):
                u = orc.uniform2(self.seed, m >> 1, i, gen, 4)[m & 1]
                buf[m] = min(int(np.floor(u * (b - a + 1))) + a, b)
            self.particle += 1
            self.randi_calls += 1
            return
        if kind == 0 and rows == 26:
            self.particle = 0
        super().fill(kind, rows, cols, a, b, out)
