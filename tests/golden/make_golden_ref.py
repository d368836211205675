#!/usr/bin/env python
"""Regenerates tests/golden/golden_ref_v1.npz from the REFERENCE ITSELF: oracle/_ref/libhpe_ref.so is
the reference's own src/{fingermodel,thumbmodel,handmodel,costfunc,observedmodel,PSO}.cpp compiled
unmodified (oracle/Makefile) against the stand-in Armadillo/OpenCV headers in oracle/shim/.  Needs
/root/reference, so it runs in the development container only; the .npz it writes is committed and is
what the GPU box (which has no /root/reference) checks the CUDA path and the oracle against.

Inputs are the same as tests/golden/make_golden.py (the oracle-made file), so the two files can be
compared field by field.  Run: python tests/golden/make_golden_ref.py
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


def main():
    om = orc.make_model()
    rng = np.random.default_rng(20140528)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(om, truth)          # input data, not a reference output
    R = ref.Reference(threads=1)
    R.load_depth(depth)
    cloud, depth_cm, dt = R.observation()
    poses = np.concatenate([D.X0[None], truth[None], truth + rng.normal(size=(30, 26)) * D.STD,
                            truth + rng.normal(size=(32, 26)) * D.STD * 0.1])
    S = np.empty((len(poses), 48, 3))
    J = np.empty((len(poses), 21, 3))
    terms = np.empty((len(poses), 3))
    match = np.empty((len(poses), R.n_pts), dtype=np.int8)
    cal_cost = np.empty(len(poses))
    cal_cost2 = np.empty(len(poses))
    for i, th in enumerate(poses):
        S[i], J[i] = R.fk(th)
        terms[i], m = R.cost_terms(th)
        match[i] = m
        cal_cost[i] = R.cal_cost(th)
        cal_cost2[i], _ = R.cal_cost2(th)
    # frozen correspondences (cal_cost2 with compute_corr = false): poses[2] scored with poses[3]'s matches
    frozen_cost, _ = R.cal_cost2(poses[2], match=match[3].astype(np.int32))
    # pso_evolve on the replayed streams
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    out = dict(truth=truth, poses=poses, spheres=S, joints=J, align=terms[:, 0], depth_term=terms[:, 1],
               colli=terms[:, 2], match=match, cal_cost=cal_cost, cal_cost2=cal_cost2,
               frozen_cost=np.array(frozen_cost),
               cloud=cloud, scale=np.array(R.scale), dt_max=np.array(dt.max()), dt_sum=np.array(dt.sum()),
               dt_rows=dt[::16].astype(np.float32), depth_cm_sum=np.array(depth_cm.sum()), pso_x0=x0)
    for tag, num_p, maxiter, seed in (("pso", 16, 8, 1000), ("pso64", 64, 30, 1000)):
        R.set_pso_params(maxiter=maxiter)
        hook = ref.PhiloxReplay(seed)
        hook.num_p = num_p
        hook.install()
        bestp = R.pso_evolve(x0, num_p)
        ref.PhiloxReplay.uninstall()
        assert hook.seeds_seen == [1000] and hook.gen == maxiter - 1
        out.update({tag + "_bestp": bestp, tag + "_best": np.array(R.cal_cost(bestp)),
                    tag + "_num_p": np.array(num_p), tag + "_maxiter": np.array(maxiter), tag + "_seed": np.array(seed)})
    extra(R, om, rng, truth, depth, out)
    np.savez_compressed(os.path.join(HERE, "golden_ref_v1.npz"), **out)
    print("wrote golden_ref_v1.npz", os.path.getsize(os.path.join(HERE, "golden_ref_v1.npz")), "bytes")


def extra(R, om, rng, truth, depth, out):
    """rows added after the first version (kept separate so that the random stream above is unchanged)"""
    rng2 = np.random.default_rng(20140529)
    # PSO::refine_init_pose (src/PSO.cpp:216-266) from three starting poses
    starts = truth + rng2.normal(size=(3, 26)) * D.STD * np.array([0.1, 0.3, 0.6])[:, None]
    refined = np.stack([R.refine_init_pose(s) for s in starts])
    out.update(refine_x0=starts, refine_out=refined,
               refine_cost=np.array([R.cal_cost2(r)[0] for r in refined]))
    # costfunc::gnd_truth_err (src/costfunc.cpp:476-507) on a made-up joint.txt row block
    gt = rng2.normal(size=(4, 63)) * 30.0
    R.fk(truth)
    out.update(gt_rows=gt, gt_err=np.array([R.gnd_truth_err(gt, f) for f in range(4)]))
    # PSO::dim_restore (src/PSO.cpp:160-180)
    t22 = rng2.normal(size=22) * 10
    out.update(dim22=t22, dim26=R.dim_restore(t22))


if __name__ == "__main__":
    main()
