"""CPU: the C-ABI library builds, loads without a GPU, exports every symbol include/hpe.h
declares, validates arguments, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hpe.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hpe_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import hpe_b200
    return hpe_b200.lib()


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("hpe_model_create", "hpe_fk_batch", "hpe_obs_upload", "hpe_obs_get", "hpe_cost_batch",
                 "hpe_cost_batch_dev", "hpe_cost_batch_frozen", "hpe_correspondences", "hpe_pso_evolve",
                 "hpe_track_sequence", "hpe_swarm_create", "hpe_swarm_step", "hpe_swarm_reduce",
                 "hpe_render_dev", "hpe_last_error_string", "hpe_refine_init_pose", "hpe_pso_optimise",
                 "hpe_bincomp_batch", "hpe_gnd_truth_err", "hpe_cost_batch_submit", "hpe_cost_batch_wait", "hpe_fit_frames",
                 "hpe_fit_frames_submit", "hpe_fit_frames_wait", "hpe_track_sequences", "hpe_track_sequences_refined"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    import hpe_b200
    out = subprocess.run(["nm", "-D", "--defined-only", hpe_b200.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hpe_[a-z0-9_]+)", out))
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, missing
    for n in declared_functions():
        assert getattr(lib, n) is not None


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: the header compiles as C and mentions no C++/torch types"""
    c = tmp_path / "t.c"
    c.write_text('#include "hpe.h"\nint main(void){hpe_pso_params p; (void)p; return HPE_OK;}\n')
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                    "-o", str(tmp_path / "t.o")], check=True)
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)     # declarations only
    assert "torch" not in text and "std::" not in text and "at::" not in text and "Tensor" not in text


def test_argument_validation_needs_no_gpu(lib):
    import hpe_b200
    D = hpe_b200.defaults
    lib.hpe_last_error_string.restype = C.c_char_p
    h = C.c_void_p()
    geo = np.ascontiguousarray(D.HGEO_CM); sp = np.ascontiguousarray(D.SPACING); cmc = np.ascontiguousarray(D.CMC)
    rad = np.ascontiguousarray(D.RAD_CM)
    tb = np.array([2, 2, 2, 2], np.int32); fg = np.array([4, 2, 2, 2], np.int32); bad = np.array([3, 2, 2, 2], np.int32)
    p = lambda a, t=C.c_double: a.ctypes.data_as(C.POINTER(t))
    assert lib.hpe_model_create(p(geo), p(sp), p(cmc), p(bad, C.c_int), p(fg, C.c_int), p(rad), C.byref(h)) == 1
    assert b"sphere layout" in lib.hpe_last_error_string()
    assert lib.hpe_model_create(None, p(sp), p(cmc), p(tb, C.c_int), p(fg, C.c_int), p(rad), C.byref(h)) == 1
    assert lib.hpe_model_create(p(geo), p(sp), p(cmc), p(tb, C.c_int), p(fg, C.c_int), p(rad), C.byref(h)) == 0
    assert lib.hpe_fk_batch(h, None, C.c_int64(0), None, None) == 0          # empty batch is fine
    assert lib.hpe_fk_batch(h, None, C.c_int64(-1), None, None) == 1
    assert lib.hpe_set_matcher(7) == 1 and lib.hpe_set_matcher(0) == 0
    # the later entry points validate before they touch the device: NULL handles, a job that is not in flight
    job = C.c_int(-1)
    assert lib.hpe_cost_batch_submit(None, None, 0, None, C.c_int64(4), None, None, None, C.byref(job)) == 1
    assert lib.hpe_cost_batch_submit(h, None, 0, None, C.c_int64(0), None, None, None, None) == 1      # no job output
    assert lib.hpe_cost_batch_wait(0) == 1 and lib.hpe_cost_batch_wait(99) == 1 and lib.hpe_cost_batch_wait(-1) == 1
    assert b"not in flight" in lib.hpe_last_error_string()
    assert lib.hpe_pso_optimise(h, None, None, 8, None, 0, 1, None, None, None, None, None, None) == 1
    assert lib.hpe_bincomp_batch(h, None, 0, None, C.c_int64(3), None, None) == 1
    assert lib.hpe_bincomp_batch(h, None, 0, None, C.c_int64(0), None, None) == 1                        # NULL observation
    lib.hpe_model_destroy(h)
    o = C.c_void_p()
    assert lib.hpe_obs_create(0, 250, C.c_double(241.42), C.byref(o)) == 1
    assert lib.hpe_obs_create(1, 300, C.c_double(241.42), C.byref(o)) == 1     # more points than 8 per lane


def test_no_cpu_fallback(lib):
    """without a CUDA device every compute entry point fails loudly with HPE_ERR_CUDA"""
    import hpe_b200
    if hpe_b200.device_count() > 0:
        pytest.skip("a GPU is visible: covered by the -m gpu tests")
    lib.hpe_last_error_string.restype = C.c_char_p
    o = C.c_void_p()
    assert lib.hpe_obs_create(1, 250, C.c_double(241.42), C.byref(o)) == 2
    assert b"no CPU fallback" in lib.hpe_last_error_string()
    hand = hpe_b200.handmodel()
    with pytest.raises(hpe_b200.HpeError) as e:
        hand.build_hand_model(hpe_b200.defaults.X0)
    assert e.value.code == 2
    with pytest.raises(hpe_b200.HpeError):
        hpe_b200.observedmodel(capacity=1)
    with pytest.raises(hpe_b200.HpeError):
        hpe_b200.render_frames(hand, hpe_b200.defaults.X0[None])


def test_product_never_touches_the_oracle():
    """only tests/, smoke() and bench.py's baseline legs may use oracle/: the package must not"""
    pkg = os.path.join(ROOT, "hand-pose-estimation_b200")
    for base, _, files in os.walk(pkg):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "hpe_oracle" not in text and "import orc" not in text and "oracle/" not in text, (base, f)
    # and the library links neither the oracle nor OpenMP
    import hpe_b200
    ldd = subprocess.run(["ldd", hpe_b200.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "gomp" not in ldd


def test_cost_kernel_plan_covers_every_pose_once(lib):
    """The self-scheduling plan of the scoring kernel (hpe_debug_cost_plan, host only) replayed with the kernel's fetch
    protocol and warps that arrive in random order: every pose must be handed out exactly once, whatever the batch size."""
    import ctypes as C
    import numpy as np
    rng = np.random.default_rng(0)
    sizes = [0, 1, 5, 6, 7, 35, 36, 3551, 3552, 21311, 21312, 42623, 42624, 42625, 65536, 100003, 1 << 20]
    sizes += [int(v) for v in rng.integers(1, 200000, size=40)]
    for sms in (148, 1, 7):
        for n in sizes:
            plan = (C.c_int * 6)()
            assert lib.hpe_debug_cost_plan(C.c_int64(n), sms, plan) == 0
            warps, group, unit, n1, n2, t2 = list(plan)
            assert 0 <= n1 <= n2 <= n and (warps > 0 or n == 0)
            seen = np.zeros(n, dtype=np.int32)
            nxt, tail = 0, 0
            exhausted = np.zeros(max(warps, 1), dtype=bool)
            active = list(range(warps))
            while active:
                w = active[int(rng.integers(len(active)))]          # any interleaving of the warps
                s = cnt = 0
                if not exhausted[w]:
                    s, nxt = nxt, nxt + group
                    cnt = min(group, n1 - s)
                    if cnt <= 0:
                        exhausted[w] = True
                if cnt <= 0:
                    t, tail = tail, tail + 1
                    if t < t2:
                        s = n1 + unit * t
                        cnt = min(unit, n2 - s)
                    else:
                        s = n2 + (t - t2)
                        cnt = 1 if s < n else 0
                if cnt <= 0:
                    active.remove(w)
                    continue
                seen[s:s + cnt] += 1
            assert (seen == 1).all(), (sms, n, list(plan))
