#!/usr/bin/env python
"""Writes tests/golden/golden_ref_opt_v1.npz (the SURVEY 8(f) rank 3 and 4 rows) from the REFERENCE ITSELF (oracle/_ref/libhpe_ref.so, the
reference's own sources compiled unmodified, see make_golden_ref.py): PSO::pso_optimise
(src/PSO.cpp:539-712) on the replayed counter-based streams (ref.PhiloxReplayOptimise, one OpenMP
thread so that the per-particle randi draws arrive in particle order).  Needs /root/reference: run in
the development container; also costfunc::bincomp_penalty (src/costfunc.cpp:384-436) of 64 poses through the
reference's code with the filled cv::circle of oracle/shim (pinned against cv2.circle).  The committed .npz is what the GPU box checks the oracle and the CUDA path
against.  Run: python tests/golden/make_golden_ref_opt.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import orc  # noqa: E402  (synthetic frame renderer + Philox streams only)
import ref  # noqa: E402
from hpe_b200 import defaults as D  # noqa: E402

CASES = ((8, 3, 1000, 0.7298, 1.49618, 1.49618), (12, 4, 7, 0.6, 1.2, 1.7), (16, 5, 1000, 0.7298, 1.49618, 1.49618))


def main():
    om = orc.make_model()
    rng = np.random.default_rng(20140529)
    R = ref.Reference(threads=1)
    out = dict(cases=np.array(CASES))
    for k, (num_p, maxiter, seed, w, c1, c2) in enumerate(CASES):
        truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
        depth = orc.render_depth(om, truth)          # input data, not a reference output
        R.load_depth(depth)
        x0 = truth + rng.normal(size=26) * D.STD * 0.3
        R.set_pso_params(maxiter=maxiter, w=w, c1=c1, c2=c2)
        hook = ref.PhiloxReplayOptimise(seed)
        hook.num_p = num_p
        hook.install()
        try:
            bestp = R.pso_optimise(x0, num_p)
        finally:
            ref.PhiloxReplay.uninstall()
        assert hook.randi_calls == num_p * (maxiter - 1) and hook.gen == maxiter - 1
        out["truth%d" % k] = truth
        out["x0_%d" % k] = x0
        out["bestp%d" % k] = bestp
        out["bestp_cost%d" % k] = np.array(R.cal_cost(bestp))     # cal_cost of the returned pose (not gbest_cost)
    # costfunc::bincomp_penalty on the last frame: near, far and partly-outside poses
    poses = truth + rng.normal(size=(64, 26)) * D.STD * np.array([0.0, 0.1, 1.0, 3.0])[np.arange(64) % 4][:, None]
    poses[7::8, 3:6] += rng.normal(size=(8, 3)) * 15.0
    out["bincomp_truth"] = truth
    out["bincomp_poses"] = poses
    out["bincomp"] = np.array([R.bincomp_penalty(p) for p in poses])
    np.savez_compressed(os.path.join(HERE, "golden_ref_opt_v1.npz"), **out)
    print("wrote golden_ref_opt_v1.npz:", {k: out["bestp_cost%d" % k].item() for k in range(len(CASES))})


if __name__ == "__main__":
    main()
