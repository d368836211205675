"""Pruning study for the nearest-centre search (numpy, uses the oracle as a development aid only):
how many sphere blocks survive a conservative cluster-level bound, for several cluster sizes."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import orc
from hpe_b200 import defaults as D

def kd_order(P, leaf):
    idx = np.arange(len(P))
    def rec(ids):
        if len(ids) <= leaf:
            return [ids]
        pts = P[ids]
        ax = np.argmax(pts.max(0) - pts.min(0))
        o = ids[np.argsort(pts[:, ax], kind="stable")]
        nl = ((len(ids) // leaf + 1) // 2) * leaf if len(ids) % leaf == 0 else ((-(-len(ids) // leaf) + 1) // 2) * leaf
        nl = min(nl, len(ids) - 1)
        return rec(o[:nl]) + rec(o[nl:])
    return rec(idx)

def main():
    om = orc.make_model()
    rng = np.random.default_rng(20140528)
    truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
    depth = orc.render_depth(om, truth)
    obs = orc.Observation(depth, num_samples=256)
    P = obs.cloud                      # (n,3) negated frame
    n = len(P)
    sig = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    poses = truth + np.random.default_rng(2).normal(size=(400, 26)) * D.STD * sig
    S, _ = orc.fk(om, poses)
    S = S.copy()                       # build_hand_model returns the negated frame the matcher sees
    # static sphere blocks of 4: (a) consecutive, (b) regrouped
    consec = [list(range(4*g, 4*g+4)) for g in range(12)]
    fb = [8, 18, 28, 38]
    regroup = [[0,1,2,3],[4,5,6,7]]
    regroup += [[fb[0]+0, fb[0]+1, fb[1]+0, fb[1]+1], [fb[0]+2, fb[0]+3, fb[1]+2, fb[1]+3],
                [fb[2]+0, fb[2]+1, fb[3]+0, fb[3]+1], [fb[2]+2, fb[2]+3, fb[3]+2, fb[3]+3]]
    regroup += [[b+4, b+5, b+6, b+7] for b in fb]
    regroup += [[fb[0]+8, fb[0]+9, fb[1]+8, fb[1]+9], [fb[2]+8, fb[2]+9, fb[3]+8, fb[3]+9]]
    assert sorted(sum(regroup, [])) == list(range(48))
    for name, blocks in (("consecutive", consec), ("regrouped", regroup)):
        B = np.array(blocks)
        for leaf, lanes_per_cluster in ((32, 8), (16, 8), (8, 8), (32, 32)):
            cl = kd_order(P, leaf)
            ctr = np.array([P[c].mean(0) for c in cl])
            # minimal enclosing-ish: centre = bbox centre, rho = max dist
            ctr = np.array([(P[c].max(0) + P[c].min(0)) / 2 for c in cl])
            rho = np.array([np.linalg.norm(P[c] - ctr[i], axis=1).max() for i, c in enumerate(cl)])
            d = np.linalg.norm(ctr[None, :, None, :] - S[:, None, :, :], axis=-1)      # pose, cluster, sphere
            dblk = d[:, :, B].min(-1)                                                     # pose, cluster, block
            dmin = d.min(-1)                                                             # pose, cluster
            keep = dblk <= dmin[..., None] + 2 * rho[None, :, None] + 1e-3
            K = keep.sum(-1)                                                             # pose, cluster
            # tighter: per-point dynamic bound using true nearest distance of the cluster's farthest point
            dn = np.linalg.norm(P[None, :, None, :] - S[:, None, :, :], axis=-1).min(-1)  # pose, point
            U = np.array([[dn[p, c].max() for c in cl] for p in range(len(S))])
            keep2 = dblk - rho[None, :, None] <= U[..., None] + 1e-3
            K2 = keep2.sum(-1)
            ncl = len(cl)
            # warp step = 4 clusters when lanes_per_cluster == 8, else 1
            per = 32 // lanes_per_cluster
            pad = (-ncl) % per
            def steps(Kx):
                Kp = np.pad(Kx, ((0, 0), (0, pad)))
                return Kp.reshape(len(S), -1, per).max(-1).sum(-1).mean()
            pts_per_lane = leaf // lanes_per_cluster
            print(f"{name:12s} leaf={leaf:3d} lanes/cl={lanes_per_cluster:2d} clusters={ncl:3d} rho={rho.mean():.2f} "
                  f"K_static={K.mean():.2f} K_dyn={K2.mean():.2f}  warp-iters static={steps(K):.1f} dyn={steps(K2):.1f} "
                  f"(pts/lane/iter={pts_per_lane}; full sweep = {12 * ncl / per:.0f} iters)")

if __name__ == "__main__":
    main()
