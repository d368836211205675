"""ctypes binding of include/hpe.h and the Python mirror of the reference's host classes."""
import ctypes as C
import os

import numpy as np

from . import defaults as D

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class HpeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hpe status {code}: {msg}")
        self.code = code


def library_path():
    # HPE_B200_LIB selects another build of the same library (kernel tuning experiments)
    return os.environ.get("HPE_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libhpe_b200.so")


class PsoParams(C.Structure):
    """hpe_pso_params: the arguments of PSO::set_pso_params (src/PSO.cpp:38-54)."""
    _fields_ = [("ub", C.c_double * 26), ("lb", C.c_double * 26), ("std", C.c_double * 26),
                ("w", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("maxiter", C.c_int),
                ("minstep", C.c_double), ("minfunc", C.c_double)]


def lib():
    """Load libhpe_b200.so; fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise HpeError(-1, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C hand-pose-estimation_b200` (there is no CPU fallback)")
        L = C.CDLL(path)
        L.hpe_last_error_string.restype = C.c_char_p
        L.hpe_obs_depth_dev.restype = C.c_void_p
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise HpeError(rc, lib().hpe_last_error_string().decode())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def set_cost_pipeline(mode=0, chunk_poses=0):
    """host path of hpe_cost_batch: 0 = one launch per uploaded chunk (default), 1 = chase kernel for >= 32,768
    poses (needs concurrent copies and kernels: not under ncu / compute-sanitizer)"""
    _check(lib().hpe_set_cost_pipeline(int(mode), int(chunk_poses)))


def device_count():
    return int(lib().hpe_device_count())


class handmodel:
    """handmodel(h_geo, h_spacing, tb_spheres, fg_spheres, h_CMC, sphR), src/handmodel.cpp:10-27."""

    def __init__(self, h_geo=None, h_spacing=None, tb_spheres=None, fg_spheres=None, h_CMC=None, sphR=None):
        self._geo = _f64(D.HGEO_CM if h_geo is None else h_geo)
        self._spacing = _f64(D.SPACING if h_spacing is None else h_spacing)
        self._cmc = _f64(D.CMC if h_CMC is None else h_CMC)
        self._tb = np.ascontiguousarray(D.TB_N if tb_spheres is None else tb_spheres, dtype=np.int32)
        self._fg = np.ascontiguousarray(D.FG_N if fg_spheres is None else fg_spheres, dtype=np.int32)
        self._rad = _f64(D.RAD_CM if sphR is None else sphR)
        for name, arr, n in (("h_geo", self._geo, 20), ("h_spacing", self._spacing, 5), ("h_CMC", self._cmc, 5),
                             ("tb_spheres", self._tb, 4), ("fg_spheres", self._fg, 4), ("sphR", self._rad, 48)):
            if arr.shape != (n,):
                raise HpeError(1, f"{name} must have shape ({n},)")
        self._h = C.c_void_p()
        _check(lib().hpe_model_create(_p(self._geo), _p(self._spacing), _p(self._cmc), _p(self._tb, C.c_int),
                                      _p(self._fg, C.c_int), _p(self._rad), C.byref(self._h)))
        self.hand_joints = np.zeros((21, 3))
        self.last_theta = None

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.hpe_model_destroy(self._h)
            self._h = None

    def get_radii(self):
        return self._rad

    def get_spheres_radii(self):
        return self._rad.copy()

    def get_hand_geo(self):
        return self._geo.copy()

    def get_hand_CMC(self):
        return self._cmc.copy()

    def get_spacing(self):
        return self._spacing.copy()

    def build_hand_model(self, h_theta):
        """theta (26,) -> sphere centres (48,3), negated y,z; records hand_joints (21,3)
        (src/handmodel.cpp:259-298)."""
        S, J = self.build_hand_models(np.asarray(h_theta, dtype=np.float64).reshape(1, 26))
        self.hand_joints = J[0]
        self.last_theta = np.asarray(h_theta, dtype=np.float64).reshape(26).copy()
        return S[0]

    def build_hand_models(self, thetas):
        th = _f64(thetas).reshape(-1, 26)
        n = th.shape[0]
        S = np.empty((n, 48, 3))
        J = np.empty((n, 21, 3))
        _check(lib().hpe_fk_batch(self._h, _p(th), C.c_int64(n), _p(S), _p(J)))
        return S, J


class observedmodel:
    """A batch of observed frames on the device.  `capacity` frames; slot 0 is "the current
    frame" for the single-frame methods that mirror the reference."""

    def __init__(self, capacity=1, num_samples=D.NUM_SAMPLES, focal=D.FOCAL):
        self._h = C.c_void_p()
        self.capacity = int(capacity)
        self.num_samples = int(num_samples)
        self.focal = float(focal)
        self.path = ""
        _check(lib().hpe_obs_create(self.capacity, self.num_samples, C.c_double(self.focal), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.hpe_obs_destroy(self._h)
            self._h = None

    # -- reference-style single-frame API (src/observedmodel.cpp:42-64, 420-430) --
    def init_observation(self, dpath, dfname, mm_to_cm=True, imW=240, imH=320, foclen=D.FOCAL, downsample=True):
        if (imW, imH) != (240, 320) or not mm_to_cm or not downsample or float(foclen) != self.focal:
            raise HpeError(1, "only the reference configuration is supported: imW=240, imH=320, "
                              "mm_to_cm=true, downsample=true, focal as given to the constructor")
        self.path = dpath
        return self.next_frame(dfname)

    def next_frame(self, next_name):
        """load_data (raw float32[240*320] file) + get_observed, slot 0."""
        full = os.path.join(self.path, next_name) if self.path else next_name
        data = np.fromfile(full, dtype="<f4")
        if data.size != D.W * D.H:
            raise HpeError(1, f"{full}: expected {D.W * D.H} float32 values, found {data.size}")
        self.upload(data.reshape(1, D.H, D.W))
        return self

    # -- batched API --
    def upload(self, depth_mm, first=0):
        fr = np.ascontiguousarray(depth_mm, dtype=np.float32).reshape(-1, D.H, D.W)
        _check(lib().hpe_obs_upload(self._h, _p(fr, C.c_float), int(first), int(fr.shape[0])))

    def upload_dev(self, dev_ptr, first, count):
        _check(lib().hpe_obs_upload_dev(self._h, C.c_void_p(dev_ptr), int(first), int(count)))

    def set(self, slot, depth_mm, dt_px, cloud, scale):
        depth_mm = np.ascontiguousarray(depth_mm, dtype=np.float32).reshape(D.H, D.W)
        dt_px = np.ascontiguousarray(dt_px, dtype=np.float32).reshape(D.H, D.W)
        cloud = _f64(cloud).reshape(self.num_samples, 3)
        _check(lib().hpe_obs_set(self._h, int(slot), _p(depth_mm, C.c_float), _p(dt_px, C.c_float), _p(cloud),
                                 C.c_double(scale)))

    def get(self, slot=0):
        cloud = np.empty((self.num_samples, 3))
        dt = np.empty((D.H, D.W), dtype=np.float32)
        scale, n_hand, dt_max = C.c_double(), C.c_int(), C.c_float()
        _check(lib().hpe_obs_get(self._h, int(slot), _p(cloud), _p(dt, C.c_float), C.byref(scale), C.byref(n_hand),
                                 C.byref(dt_max)))
        return {"cloud": cloud, "dt": dt, "scale": scale.value, "n_hand": n_hand.value, "dt_max": dt_max.value}

    def depth_dev_ptr(self):
        return lib().hpe_obs_depth_dev(self._h)

    def get_ptncloud(self, slot=0):
        return self.get(slot)["cloud"]

    def get_disttran(self, slot=0):
        return self.get(slot)["dt"].astype(np.float64)

    def get_img_scale(self, slot=0):
        return self.get(slot)["scale"]

    def get_imgW(self):
        return 240

    def get_imgH(self):
        return 320

    def get_focal_len(self):
        return self.focal


class costfunc:
    """costfunc(handmodel*, observedmodel*): non-owning references (src/costfunc.cpp:16-29)."""

    def __init__(self, handM, observed):
        self.hand = handM
        self.observation = observed

    def cal_cost(self, theta, slot=0):
        a, d, _ = self.cal_cost_batch(np.asarray(theta, dtype=np.float64).reshape(1, 26), slot)
        return float(a[0] + d[0])

    def cal_cost2(self, theta, matchId=None, compute_corr=True, debug=False, slot=0):
        th = _f64(theta).reshape(1, 26)
        if compute_corr:
            if matchId is not None:
                matchId[:] = self.compute_correspondences(th, slot)[0]
            a, d, c = self.cal_cost_batch(th, slot, with_collision=True)
        else:
            a, d, c = self.cal_cost_frozen_batch(th, matchId, slot)
        if debug:
            print(a[0]); print(d[0]); print(c[0], "\n")
        return float(a[0] + d[0] + c[0])

    def cal_cost_batch(self, thetas, slot=0, with_collision=False):
        th = _f64(thetas).reshape(-1, 26)
        n = th.shape[0]
        a, d = np.empty(n), np.empty(n)
        c = np.empty(n) if with_collision else None
        _check(lib().hpe_cost_batch(self.hand._h, self.observation._h, int(slot), _p(th), C.c_int64(n), _p(a), _p(d),
                                    _p(c)))
        return a, d, (c if with_collision else np.zeros(n))

    def cal_cost_frozen_batch(self, thetas, matchId, slot=0):
        th = _f64(thetas).reshape(-1, 26)
        n = th.shape[0]
        m = np.ascontiguousarray(matchId, dtype=np.int32)
        a, d, c = np.empty(n), np.empty(n), np.empty(n)
        _check(lib().hpe_cost_batch_frozen(self.hand._h, self.observation._h, int(slot), _p(th), C.c_int64(n),
                                           _p(m, C.c_int32), _p(a), _p(d), _p(c)))
        return a, d, c

    def cal_cost_submit(self, thetas, slot=0, with_collision=False):
        """streaming form of cal_cost_batch (hpe_cost_batch_submit): queues the batch and returns a job whose
        wait() gives (align, depth, colli); up to 3 jobs in flight, the upload of one under the kernels of the
        previous.  Pass page-locked arrays for real overlap."""
        return CostJob(self, thetas, slot, with_collision)

    def depthMatch_penalty(self, theta, slot=0):
        """costfunc::depthMatch_penalty (src/costfunc.cpp:439-473) of the spheres of pose theta: the depth map (cm) with the
        outlines of the projected spheres, (240, 320)"""
        return self.depthMatch_penalty_batch(np.asarray(theta, dtype=np.float64).reshape(1, 26), slot=slot)[0]

    def depthMatch_penalty_batch(self, thetas, slot=0):
        th = _f64(thetas).reshape(-1, 26)
        n = th.shape[0]
        img = np.empty((n, D.H, D.W))
        _check(lib().hpe_depthmatch_batch(self.hand._h, self.observation._h, int(slot), _p(th), C.c_int64(n), _p(img)))
        return img

    def bincomp_penalty(self, theta, slot=0):
        """costfunc::bincomp_penalty (src/costfunc.cpp:384-436) of the spheres of pose theta: pixels where the
        silhouette of the projected spheres and the observed one differ"""
        return self.bincomp_penalty_batch(np.asarray(theta, dtype=np.float64).reshape(1, 26), slot=slot)[0]

    def bincomp_penalty_batch(self, thetas, slot=0, want_silhouette=False):
        """returns penalties (n,) [and the drawn silhouettes as a bool array (n, 240, 320)]"""
        th = _f64(thetas).reshape(-1, 26)
        n = th.shape[0]
        pen = np.empty(n)
        sil = np.zeros((n, D.H * D.W // 32), dtype=np.uint32) if want_silhouette else None
        _check(lib().hpe_bincomp_batch(self.hand._h, self.observation._h, int(slot), _p(th), C.c_int64(n), _p(pen),
                                       _p(sil, C.c_uint32)))
        if not want_silhouette:
            return pen
        bits = np.unpackbits(sil.view(np.uint8).reshape(n, -1), axis=1, bitorder="little")
        return pen, bits.reshape(n, D.H, D.W).astype(bool)

    def gnd_truth_err(self, gnd_truth, frame, theta=None):
        """costfunc::gnd_truth_err(gnd_truth, frame) (src/costfunc.cpp:476-507): gnd_truth = joint.txt
        as an (n_frames, 63) array; uses the pose of the last build_hand_model call unless theta is given"""
        th = self.hand.last_theta if theta is None else theta
        if th is None:
            raise HpeError(1, "gnd_truth_err: no pose (call build_hand_model first or pass theta)")
        return self.gnd_truth_err_batch(np.asarray(th, dtype=np.float64).reshape(1, 26),
                                        _f64(gnd_truth).reshape(-1, 63)[frame:frame + 1])[0]

    def gnd_truth_err_batch(self, thetas, gt_rows):
        th = _f64(thetas).reshape(-1, 26)
        gt = _f64(gt_rows).reshape(-1, 63)
        if gt.shape[0] != th.shape[0]:
            raise HpeError(1, "gnd_truth_err_batch: one joint.txt row per pose")
        err = np.empty(th.shape[0])
        _check(lib().hpe_gnd_truth_err(self.hand._h, _p(th), _p(gt), C.c_int64(th.shape[0]), _p(err)))
        return err

    def compute_correspondences(self, thetas, slot=0):
        th = _f64(thetas).reshape(-1, 26)
        n = th.shape[0]
        m = np.empty((n, self.observation.num_samples), dtype=np.int32)
        _check(lib().hpe_correspondences(self.hand._h, self.observation._h, int(slot), _p(th), C.c_int64(n),
                                         _p(m, C.c_int32)))
        return m

    def cal_cost_dev(self, theta_dev_ptr, n, cost_dev_ptr, terms_dev_ptr=None, slot=0, with_collision=False,
                     stream=None):
        """device pointers in, device pointers out, asynchronous on `stream`."""
        _check(lib().hpe_cost_batch_dev(self.hand._h, self.observation._h, int(slot), C.c_void_p(theta_dev_ptr),
                                        C.c_int64(n), int(with_collision), C.c_void_p(cost_dev_ptr),
                                        C.c_void_p(terms_dev_ptr or 0), C.c_void_p(stream or 0)))


_FIT_POOL = {}       # n -> [(bestp, best), ...] page-locked result buffers of finished fit jobs


class FitJob:
    """one hpe_fit_frames call in flight (hpe_fit_frames_submit / hpe_fit_frames_wait); keeps its host arrays alive.
    The page-locked result buffers are recycled: allocating page-locked memory synchronises the device, which would
    serialise the very calls this form is meant to overlap."""

    def __init__(self, pso, optfunc, fr, x0, num_particles, sd):
        n = fr.shape[0]
        self.keep = (fr, x0, sd, optfunc)
        self.n = n
        free = _FIT_POOL.get(n)
        self.bestp, self.best = free.pop() if free else (_Pinned((n, 26)), _Pinned((n,)))
        jid = C.c_int(-1)
        _check(lib().hpe_fit_frames_submit(optfunc.hand._h, optfunc.observation._h, C.byref(pso.params), int(num_particles),
                                           _p(fr, C.c_float), _p(x0), C.c_int64(n), _p(sd, C.c_uint64),
                                           _p(self.bestp.array), _p(self.best.array), C.byref(jid)))
        self.id = jid.value

    def wait(self):
        jid, self.id = self.id, -1
        if jid < 0:
            raise HpeError(4, "fit job already waited for")
        _check(lib().hpe_fit_frames_wait(int(jid)))
        out = self.bestp.array.copy(), self.best.array.copy()
        pool = _FIT_POOL.setdefault(self.n, [])
        if len(pool) < 2:
            pool.append((self.bestp, self.best))
        else:
            self.bestp.free(); self.best.free()
        self.bestp = self.best = self.keep = None
        return out

    def __del__(self):
        if getattr(self, "id", -1) >= 0 and _LIB is not None:
            try:
                _LIB.hpe_fit_frames_wait(int(self.id))       # in-flight DMA must finish before its buffers go away
            except Exception:
                pass


class _Pinned:
    """numpy view of a page-locked host buffer owned by the library (hpe_host_alloc / hpe_host_free)"""

    def __init__(self, shape, dtype=np.float64):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        _check(lib().hpe_host_alloc(C.c_size_t(n), C.byref(self._p)))
        buf = (C.c_ubyte * max(n, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if getattr(self, "_p", None) is not None and self._p.value and _LIB is not None:
            self.array = None
            _LIB.hpe_host_free(self._p)
            self._p = None

    __del__ = free


class CostJob:
    """one batch in flight (hpe_cost_batch_submit / hpe_cost_batch_wait).  The poses are copied into page-locked host
    buffers and the results land in page-locked buffers, so submit() returns while the upload is still running
    (a copy from pageable memory would block the caller); the buffers live until wait() -- or until the job is
    dropped, which waits first: a job must never outlive its buffers or keep its slot busy for ever."""

    def __init__(self, cf, thetas, slot, with_collision):
        th = _f64(thetas).reshape(-1, 26)
        n = th.shape[0]
        self._bufs = [_Pinned((max(n, 1), 26)), _Pinned((max(n, 1),)), _Pinned((max(n, 1),))]
        self.th, self.a, self.d = self._bufs[0].array[:n], self._bufs[1].array[:n], self._bufs[2].array[:n]
        self.th[:] = th
        self.c = None
        if with_collision:
            self._bufs.append(_Pinned((max(n, 1),)))
            self.c = self._bufs[3].array[:n]
            self.c[:] = 0.0
        self.id = None
        job = C.c_int(-1)
        _check(lib().hpe_cost_batch_submit(cf.hand._h, cf.observation._h, int(slot), _p(self.th), C.c_int64(n),
                                           _p(self.a), _p(self.d), _p(self.c), C.byref(job)))
        self.id = job.value
        self._keep = cf

    def wait(self):
        """blocks until the batch is done; returns copies (align, depth, collision) and releases the job"""
        if self.id is None:
            raise HpeError(4, "CostJob.wait: the job was already waited on")
        jid, self.id = self.id, None
        _check(lib().hpe_cost_batch_wait(int(jid)))
        out = (self.a.copy(), self.d.copy(), np.zeros(len(self.a)) if self.c is None else self.c.copy())
        self._release()
        return out

    def _release(self):
        self.th = self.a = self.d = self.c = None
        for b in self._bufs:
            b.free()
        self._bufs = []

    def __del__(self):
        try:
            if getattr(self, "id", None) is not None and _LIB is not None:
                _LIB.hpe_cost_batch_wait(int(self.id))       # in-flight DMA must finish before its buffers go away
                self.id = None
            self._release()
        except Exception:
            pass


class PSO:
    """PSO: set_pso_params + pso_evolve (src/PSO.cpp:38-54, 717-886)."""

    def __init__(self):
        self.params = PsoParams()
        self.set_pso_params(D.UB, D.LB, D.STD, 0.7298, 1.49618, 1.49618, 100, 1e-6, 1e-6)

    def set_pso_params(self, upperbound, lowerbound, std, omega, phip, phig, maxiter, minstep, minfunc):
        p = self.params
        p.ub[:] = list(_f64(upperbound))
        p.lb[:] = list(_f64(lowerbound))
        p.std[:] = list(_f64(std))
        p.w, p.c1, p.c2 = float(omega), float(phip), float(phig)
        p.maxiter = int(maxiter)
        p.minstep, p.minfunc = float(minstep), float(minfunc)

    def pso_evolve(self, optfunc, x0, num_particles, seed=1000, slot=0):
        """one swarm on frame `slot`; returns bestp (26,) like the reference's out-argument."""
        bestp, _, _ = self.pso_evolve_batch(optfunc, np.asarray(x0, dtype=np.float64).reshape(1, 26), num_particles,
                                            seeds=[seed], first=slot)
        return bestp[0]

    def pso_evolve_batch(self, optfunc, x0, num_particles, seeds=None, first=0, want_trace=True):
        x0 = _f64(x0).reshape(-1, 26)
        count = x0.shape[0]
        bestp = np.empty((count, 26))
        best = np.empty(count)
        tl = max(self.params.maxiter - 1, 1)
        trace = np.empty((count, tl)) if want_trace else None
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        _check(lib().hpe_pso_evolve(optfunc.hand._h, optfunc.observation._h, C.byref(self.params), int(num_particles),
                                    _p(x0), int(first), int(count), _p(sd, C.c_uint64), _p(bestp), _p(best),
                                    _p(trace)))
        if trace is not None:
            trace = trace[:, :self.params.maxiter - 1]
        return bestp, best, trace

    def pso_optimise(self, optfunc, x0, num_particles, seed=1000, slot=0):
        """PSO::pso_optimise(optfunc, x0, num_particles, bestp) (src/PSO.cpp:539-712): the hybrid of coordinate
        descent and canonical PSO with this object's w, c1, c2; returns bestp (26,)"""
        return self.pso_optimise_batch(optfunc, np.asarray(x0, dtype=np.float64).reshape(1, 26), num_particles,
                                       seeds=[seed], first=slot)[0][0]

    def pso_optimise_batch(self, optfunc, x0, num_particles, seeds=None, first=0, want_state=False):
        """one pso_optimise swarm per frame first+b; returns (bestp, gbest_cost, evaluations of the sequential
        loop[, final particles, final pcost])"""
        x0 = _f64(x0).reshape(-1, 26)
        count = x0.shape[0]
        bestp, best = np.empty((count, 26)), np.empty(count)
        evals = np.zeros(count, dtype=np.int64)
        fx = np.empty((count, int(num_particles), 26)) if want_state else None
        fc = np.empty((count, int(num_particles))) if want_state else None
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        _check(lib().hpe_pso_optimise(optfunc.hand._h, optfunc.observation._h, C.byref(self.params), int(num_particles),
                                      _p(x0), int(first), int(count), _p(sd, C.c_uint64), _p(bestp), _p(best),
                                      _p(evals, C.c_int64), _p(fx), _p(fc)))
        return (bestp, best, evals, fx, fc) if want_state else (bestp, best, evals)

    def track_sequence(self, optfunc, x0, num_particles, n_frames, seed=1000, first=0):
        x0 = _f64(x0).reshape(26)
        poses = np.empty((n_frames, 26))
        costs = np.empty(n_frames)
        _check(lib().hpe_track_sequence(optfunc.hand._h, optfunc.observation._h, C.byref(self.params),
                                        int(num_particles), _p(x0), int(first), int(n_frames), C.c_uint64(seed),
                                        _p(poses), _p(costs)))
        return poses, costs

    def fit_frames(self, optfunc, frames_mm, x0, num_particles, seeds=None):
        """offline batch from host frames: upload, observation and one swarm per frame, the upload of the next
        chunk overlapped with the fit of the current one (the observedmodel's capacity sets the chunk size)"""
        fr = np.ascontiguousarray(frames_mm, dtype=np.float32).reshape(-1, D.H, D.W)
        x0 = _f64(x0).reshape(-1, 26)
        n = fr.shape[0]
        if x0.shape[0] != n:
            raise HpeError(1, "fit_frames: one initial pose per frame")
        bestp, best = np.empty((n, 26)), np.empty(n)
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        _check(lib().hpe_fit_frames(optfunc.hand._h, optfunc.observation._h, C.byref(self.params), int(num_particles),
                                    _p(fr, C.c_float), _p(x0), C.c_int64(n), _p(sd, C.c_uint64), _p(bestp), _p(best)))
        return bestp, best

    def fit_frames_submit(self, optfunc, frames_mm, x0, num_particles, seeds=None):
        """submit / wait form of fit_frames (hpe_fit_frames_submit): returns a job whose wait() gives (bestp, best).
        Two jobs may be in flight; the frames of the second start to upload under the last fits of the first.  The
        arrays passed in are used as they are (page-locked ones make the overlap real) and must not change before wait()."""
        fr = np.ascontiguousarray(frames_mm, dtype=np.float32).reshape(-1, D.H, D.W)
        x0 = _f64(x0).reshape(-1, 26)
        if x0.shape[0] != fr.shape[0]:
            raise HpeError(1, "fit_frames: one initial pose per frame")
        return FitJob(self, optfunc, fr, x0, num_particles, None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64))

    def refine_init_pose(self, x0, optfunc, slot=0):
        """PSO::refine_init_pose(x0, optfunc) (src/PSO.cpp:216-266); returns the refined pose (the
        reference updates x0 in place)."""
        return self.refine_init_pose_batch(np.asarray(x0, dtype=np.float64).reshape(1, 26), optfunc, first=slot)[0][0]

    def refine_init_pose_batch(self, x0, optfunc, first=0):
        """x0[b] refined against frame first+b; returns (poses, number of cal_cost2 evaluations)"""
        x = _f64(x0).reshape(-1, 26).copy()
        ev = np.zeros(x.shape[0], dtype=np.int32)
        _check(lib().hpe_refine_init_pose(optfunc.hand._h, optfunc.observation._h, _p(x), int(first), int(x.shape[0]),
                                          _p(ev, C.c_int)))
        return x, ev

    def track_sequence_refined(self, optfunc, x0, num_particles, n_frames, seed=1000, first=0):
        """the loop of test_full as written (src/testmodel.cpp:117-138): refine_init_pose, pso_evolve,
        x0 = bestp; returns (bestp per frame, gbest cost per frame, refined start pose per frame)"""
        x0 = _f64(x0).reshape(26)
        poses, refined = np.empty((n_frames, 26)), np.empty((n_frames, 26))
        costs = np.empty(n_frames)
        _check(lib().hpe_track_sequence_refined(optfunc.hand._h, optfunc.observation._h, C.byref(self.params),
                                                int(num_particles), _p(x0), int(first), int(n_frames),
                                                C.c_uint64(seed), _p(poses), _p(costs), _p(refined)))
        return poses, costs, refined

    def track_sequences(self, optfunc, x0, num_particles, n_seq, n_frames, seeds=None, first=0, interleaved=False):
        """n_seq independent sequences tracked in one launch (hpe_track_sequences).  Frame k of sequence q lives in slot
        first + q*n_frames + k, or first + k*n_seq + q when interleaved.  Returns (poses [n_seq, n_frames, 26],
        costs [n_seq, n_frames])"""
        x0 = _f64(x0).reshape(n_seq, 26)
        poses, costs = np.empty((n_seq, n_frames, 26)), np.empty((n_seq, n_frames))
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        ss, fs = (1, n_seq) if interleaved else (n_frames, 1)
        _check(lib().hpe_track_sequences(optfunc.hand._h, optfunc.observation._h, C.byref(self.params), int(num_particles),
                                         _p(x0), int(first), int(n_seq), int(n_frames), int(ss), int(fs),
                                         _p(sd, C.c_uint64), _p(poses), _p(costs)))
        return poses, costs

    def track_sequences_refined(self, optfunc, x0, num_particles, n_seq, n_frames, seeds=None, first=0):
        """test_full's loop (refine_init_pose, pso_evolve, x0 = bestp) for n_seq sequences in lock step; slots are
        frame-interleaved (first + k*n_seq + q).  Returns (poses, costs, refined start poses), sequence-major"""
        x0 = _f64(x0).reshape(n_seq, 26)
        poses, refined = np.empty((n_seq, n_frames, 26)), np.empty((n_seq, n_frames, 26))
        costs = np.empty((n_seq, n_frames))
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        _check(lib().hpe_track_sequences_refined(optfunc.hand._h, optfunc.observation._h, C.byref(self.params),
                                                 int(num_particles), _p(x0), int(first), int(n_seq), int(n_frames),
                                                 _p(sd, C.c_uint64), _p(poses), _p(costs), _p(refined)))
        return poses, costs, refined


class Swarm:
    """One large swarm (pso_evolve semantics) whose particles [lo,hi) live on this GPU.  Drive it as
        init(); exchange(); reduce(); for _ in range(maxiter-1): step(); exchange(); reduce()
    where exchange() makes every rank's table complete (all-gather of the owned rows; a no-op
    when lo,hi cover the whole swarm)."""
    ROW = 27

    def __init__(self, optfunc, params, num_particles, lo, hi, x0, seed=1000, slot=0):
        self._h = C.c_void_p()
        self.num_particles, self.lo, self.hi = int(num_particles), int(lo), int(hi)
        self.maxiter = params.maxiter
        x0 = _f64(x0).reshape(26)
        self._keep = (optfunc, params)
        _check(lib().hpe_swarm_create(optfunc.hand._h, optfunc.observation._h, int(slot), C.byref(params),
                                      self.num_particles, self.lo, self.hi, _p(x0), C.c_uint64(seed),
                                      C.byref(self._h)))
        lib().hpe_swarm_table.restype = C.c_void_p

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.hpe_swarm_destroy(self._h)
            self._h = None

    def ipc_handle(self):
        """64-byte CUDA IPC handle of this rank's table (to be gathered over any host transport)"""
        buf = (C.c_ubyte * 64)()
        _check(lib().hpe_swarm_ipc_handle(self._h, buf))
        return bytes(buf)

    def attach_peers(self, handles, rank):
        """handles: list of `world` 64-byte handles (own entry ignored).  Afterwards init/step push the owned
        rows into the peers' tables over NVLink and reduce waits for theirs: no all-gather needed."""
        blob = b"".join(handles)
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _check(lib().hpe_swarm_attach_peers(self._h, arr, len(handles), int(rank)))

    def table_ptr(self):
        """device pointer of the [num_particles][27] float64 table (pcost, pbest)"""
        return lib().hpe_swarm_table(self._h)

    def init(self, stream=None):
        _check(lib().hpe_swarm_init(self._h, C.c_void_p(stream or 0)))

    def step(self, stream=None):
        _check(lib().hpe_swarm_step(self._h, C.c_void_p(stream or 0)))

    def run(self, stream=None):
        """init + all generations + exchanges in one cooperative launch (peers attached when world > 1)"""
        _check(lib().hpe_swarm_run(self._h, C.c_void_p(stream or 0)))

    def debug_times(self):
        """phase timestamps (ns) of the last run(), shape (maxiter, 8); the first call only enables recording"""
        t = np.zeros((self.maxiter, 8), dtype=np.uint64)
        _check(lib().hpe_swarm_debug_times(self._h, _p(t, C.c_uint64)))
        return t

    def reduce(self, stream=None):
        _check(lib().hpe_swarm_reduce(self._h, C.c_void_p(stream or 0)))

    def best(self):
        bestp = np.empty(26)
        cost = C.c_double()
        trace = np.empty(max(self.maxiter - 1, 1))
        _check(lib().hpe_swarm_best(self._h, _p(bestp), C.byref(cost), _p(trace)))
        return bestp, cost.value, trace[:self.maxiter - 1]


def measure_fp32_peak():
    """(TFLOP/s, implied SM MHz) of a register-only FFMA kernel on the current device."""
    tf, mhz = C.c_double(), C.c_double()
    _check(lib().hpe_measure_fp32_peak(C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def render_dev(hand, theta_dev_ptr, n, depth_dev_ptr, focal=D.FOCAL, stream=None):
    """device poses (n*26 f64) -> device frames (n*240*320 f32), asynchronous"""
    _check(lib().hpe_render_dev(hand._h, C.c_void_p(theta_dev_ptr), C.c_int64(n), C.c_double(focal),
                                C.c_void_p(depth_dev_ptr), C.c_void_p(stream or 0)))


def render_frames(hand, thetas, focal=D.FOCAL):
    """Synthetic Subject1-format frames (float32 mm, 0 = background) of the sphere model."""
    th = _f64(thetas).reshape(-1, 26)
    out = np.empty((th.shape[0], D.H, D.W), dtype=np.float32)
    _check(lib().hpe_render(hand._h, _p(th), C.c_int64(th.shape[0]), C.c_double(focal), _p(out, C.c_float)))
    return out
