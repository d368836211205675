"""The C++ host classes (include/hpe_classes.hpp): a port of the reference's test_full driver
(tests/cpp/test_full.cpp) builds against libhpe_b200.so (CPU test) and, on the GPU, reproduces the
oracle's tracking loop on Subject1-format files."""
import os
import subprocess

import numpy as np
import pytest

from conftest import synth_pose

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hand-pose-estimation_b200")


def build_driver(tmp_path):
    exe = str(tmp_path / "test_full")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_full.cpp"), "-o", exe, "-L", PKG, "-lhpe_b200",
                    f"-Wl,-rpath,{PKG}"], check=True)
    return exe


def test_cpp_driver_builds_and_fails_loudly_without_gpu(tmp_path):
    import hpe_b200
    hpe_b200.lib()
    exe = build_driver(tmp_path)
    if hpe_b200.device_count() == 0:
        r = subprocess.run([exe, str(tmp_path) + "/", "1", "8", "3"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_tracks_like_the_oracle(tmp_path, hpe, orc, omodel, D):
    rng = np.random.default_rng(21)
    pose = synth_pose(D, rng)
    truths = []
    for _ in range(3):
        pose = np.clip(pose + rng.normal(size=26) * np.r_[[1.0] * 3, [0.1] * 3, [1.5] * 20], D.LB, D.UB)
        truths.append(pose.copy())
    for k, t in enumerate(truths):
        orc.render_depth(omodel, t).astype("<f4").tofile(str(tmp_path / f"{k:06d}_depth.bin"))   # Subject1 format
    x0 = truths[0] + rng.normal(size=26) * D.STD * 0.2
    np.savetxt(str(tmp_path / "x0.txt"), x0[None], fmt="%.17g")
    exe = build_driver(tmp_path)
    num_p, maxiter = 32, 10
    # without refine_init_pose the loop follows the oracle's trajectory exactly
    r = subprocess.run([exe, str(tmp_path) + "/", "3", str(num_p), str(maxiter), str(tmp_path / "x0.txt"), "norefine"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    prm = orc.pso_params(maxiter=maxiter)
    cur = x0
    for k in range(3):
        f = lines[k].split()
        cost, pose_out = float(f[1]), np.array([float(v) for v in f[3:]])
        obs = orc.Observation(np.fromfile(str(tmp_path / f"{k:06d}_depth.bin"), dtype="<f4").reshape(240, 320))
        bo, co, _, _, _ = orc.pso_evolve(omodel, obs, prm, cur, num_p, 1000)
        assert abs(cost - co) <= 1e-8 * co and np.abs(pose_out - bo).max() < 1e-7
        cur = bo
    # the driver as the reference wrote it (refine_init_pose before every pso_evolve, src/testmodel.cpp:126-130):
    # joints within the north star's 1 mm of the oracle's loop
    r2 = subprocess.run([exe, str(tmp_path) + "/", "3", str(num_p), str(maxiter), str(tmp_path / "x0.txt")],
                        capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    lines2 = r2.stdout.strip().splitlines()
    cur2 = x0
    for k in range(3):
        pose_out = np.array([float(v) for v in lines2[k].split()[3:]])
        obs2 = orc.Observation(np.fromfile(str(tmp_path / f"{k:06d}_depth.bin"), dtype="<f4").reshape(240, 320))
        ref_x, _ = orc.refine_init_pose(omodel, obs2, cur2)
        bo2, _, _, _, _ = orc.pso_evolve(omodel, obs2, prm, ref_x, num_p, 1000)
        _, Jg = orc.fk(omodel, pose_out)
        _, Jo = orc.fk(omodel, bo2)
        assert np.linalg.norm(Jg - Jo, axis=1).max() * 10.0 < 1.0, k
        cur2 = bo2
    tail = lines[3].split()
    S, J = orc.fk(omodel, cur)
    assert abs(float(tail[1]) - obs.scale) <= 1e-12 * obs.scale
    assert np.allclose([float(v) for v in tail[3:6]], J[4], atol=1e-9) and abs(float(tail[7]) - S[47, 2]) < 1e-9
