#!/usr/bin/env python
"""Regenerates tests/golden/golden_v1.npz from the CPU ORACLE (oracle/hpe_oracle.c).

These are oracle outputs: they pin the oracle against drift and give the GPU tests fixed expected
values.  The same inputs run through the REFERENCE'S OWN CODE (oracle/_ref) are in golden_ref_v1.npz
(make_golden_ref.py); tests/test_ref_build.py checks that the two files agree field by field.
Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import orc  # noqa: E402
from hpe_b200 import defaults as D  # noqa: E402


def main():
    om = orc.make_model()
    rng = np.random.default_rng(20140528)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(om, truth)
    obs = orc.Observation(depth)
    poses = np.concatenate([D.X0[None], truth[None], truth + rng.normal(size=(30, 26)) * D.STD,
                            truth + rng.normal(size=(32, 26)) * D.STD * 0.1])
    S, J = orc.fk(om, poses)
    a, d, c = orc.cost_batch(om, obs, poses)
    match = np.stack([orc.correspondences(obs.cloud, S[i]) for i in range(len(poses))]).astype(np.int8)
    prm = orc.pso_params(maxiter=8)
    x0 = truth + rng.normal(size=26) * D.STD * 0.3
    bestp, best, trace, _, _ = orc.pso_evolve(om, obs, prm, x0, 16, 1000)
    ys, xs = np.nonzero(depth)
    out = dict(truth=truth, poses=poses, spheres=S, joints=J, align=a, depth_term=d, colli=c, match=match,
               frame_y=ys.astype(np.int16), frame_x=xs.astype(np.int16), frame_mm=depth[ys, xs],
               cloud=obs.cloud, scale=np.array(obs.scale), n_hand=np.array(obs.n_hand),
               dt_max=np.array(obs.dt_max), dt_sum=np.array(obs.dt.sum()),
               dt_rows=obs.dt[::16].astype(np.float32),
               pso_x0=x0, pso_bestp=bestp, pso_best=np.array(best), pso_trace=trace,
               pso_num_p=np.array(16), pso_maxiter=np.array(8), pso_seed=np.array(1000),
               philox_kat=np.array(orc.philox(0x0123456789abcdef, 1, 2, 3, 4), dtype=np.uint32))
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote golden_v1.npz", os.path.getsize(os.path.join(HERE, "golden_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
