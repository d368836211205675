"""Sweep-step study for the pruned matcher (numpy; uses the oracle as a development aid only): block lists per cluster under the
box / bisector bound of prune_blocks (hpe_eval.cuh), and what the two matching passes cost for several ways of dealing the eight
clusters to the (pass, quarter) slots -- storage order, ranked by list length (shipped), concatenated pairs, half-warp clusters --
plus the ideal list (blocks that really hold a nearest sphere) and the bound with 2 / 3 reference spheres.
Usage: python tools/prune_balance_sim.py [sigma]   (1.0 = the benchmark poses, 0.2 = poses near the truth as in a swarm)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200")); sys.path.insert(0, os.path.join(ROOT,"tools"))
import orc
from hpe_b200 import defaults as D
from prune_sim import kd_order
om = orc.make_model()
rng = np.random.default_rng(20140528)
truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
depth = orc.render_depth(om, truth)
obs = orc.Observation(depth, num_samples=256)
P = obs.cloud.astype(np.float32).astype(np.float64)
sig = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
poses = truth + np.random.default_rng(2).normal(size=(2000, 26)) * D.STD * sig
S, _ = orc.fk(om, poses)
fb = [8, 18, 28, 38]
regroup = [[0,1,2,3],[4,5,6,7]]
regroup += [[fb[0]+0, fb[0]+1, fb[1]+0, fb[1]+1], [fb[0]+2, fb[0]+3, fb[1]+2, fb[1]+3],
            [fb[2]+0, fb[2]+1, fb[3]+0, fb[3]+1], [fb[2]+2, fb[2]+3, fb[3]+2, fb[3]+3]]
regroup += [[b+4, b+5, b+6, b+7] for b in fb]
regroup += [[fb[0]+8, fb[0]+9, fb[1]+8, fb[1]+9], [fb[2]+8, fb[2]+9, fb[3]+8, fb[3]+9]]
B = np.array(regroup)
q = P.mean(0)
Pp = P - q
def lists(leaf):
    cl = kd_order(P, leaf)
    lo = np.array([Pp[c].min(0) for c in cl]); hi = np.array([Pp[c].max(0) for c in cl])
    b = (lo+hi)/2; h = (hi-lo)/2
    pmax = np.array([ (Pp[c]**2).sum(1).max() for c in cl])
    C = S - q                                  # pose,48,3
    e = -2*C; w = (C**2).sum(-1)               # pose,48
    sb = w[:,None,:] + np.einsum('psi,ci->pcs', e, b)       # pose,cluster,sphere
    k = sb.argmin(-1)                                          # pose,cluster
    ek = np.take_along_axis(e[:,None,:,:].repeat(len(cl),1), k[...,None,None].repeat(3,-1), 2)  # pose,cl,1,3
    wk = np.take_along_axis(w[:,None,:].repeat(len(cl),1), k[...,None], 2)
    sk = np.take_along_axis(sb, k[...,None], 2)
    bound = (sb - sk) - (np.abs(e[:,None,:,:]-ek) * h[None,:,None,:]).sum(-1) - 4e-5*(w[:,None,:]+wk+2*pmax[None,:,None])
    keep_s = ~(bound > 0)
    keep_b = keep_s[:,:,B].any(-1)           # pose, cluster, block
    return keep_b.sum(-1), keep_s
K, ks = lists(32)
print("mean blocks/cluster", K.mean(), "std", K.std(), "spheres kept/cluster", ks.sum(-1).mean())
cur = K.reshape(len(K),2,4).max(-1).sum(-1)
srt = np.sort(K,1)[:,::-1]
bal = srt[:,0]+srt[:,4]
print("current steps", cur.mean(), " sorted", bal.mean(), " ideal sum/4", (K.sum(1)/4).mean())
# pair-concatenation: quarter handles 2 clusters sequentially, best pairing (largest+smallest)
pair = (srt[:,:4] + srt[:,7:3:-1]).max(1)
print("concat pairs (opt pairing)", pair.mean())
# half-warp: 16 lanes per cluster (2 pts/lane), 4 passes, sorted: l1+l3+l5+l7
print("halfwarp sorted", (srt[:,0]+srt[:,2]+srt[:,4]+srt[:,6]).mean())
K16,_ = lists(16)
print("leaf16 mean blocks", K16.mean(), "sum/8 ideal (2pt/lane eq to 4-step units):", (K16.sum(1)/8).mean())
s16 = np.sort(K16,1)[:,::-1]
print("leaf16, 4 passes x4 quarters (2pts/lane) sorted:", (s16[:,0]+s16[:,4]+s16[:,8]+s16[:,12]).mean(), " in 4pt-step equivalents /2")
# ideal: blocks containing the true nearest sphere of some point of the cluster
cl = kd_order(P, 32)
d = np.linalg.norm(P[None,:,None,:]-S[:,None,:,:],axis=-1)   # pose, pt, sph
nn = d.argmin(-1)
blk_of = np.zeros(48,int)
for bi,bb in enumerate(regroup):
    for s in bb: blk_of[s]=bi
nb = blk_of[nn]                                  # pose, pt
Kid = np.array([[len(set(nb[p,c])) for c in cl] for p in range(len(S))])
print("ideal blocks/cluster", Kid.mean(), " ideal current-scheme steps", Kid.reshape(len(S),2,4).max(-1).sum(-1).mean(), "sorted", (np.sort(Kid,1)[:,::-1][:,[0,4]]).sum(1).mean())
# two reference spheres
def lists2(nref):
    lo = np.array([Pp[c].min(0) for c in cl]); hi = np.array([Pp[c].max(0) for c in cl])
    b = (lo+hi)/2; h = (hi-lo)/2
    pmax = np.array([ (Pp[c]**2).sum(1).max() for c in cl])
    C = S - q; e = -2*C; w = (C**2).sum(-1)
    sb = w[:,None,:] + np.einsum('psi,ci->pcs', e, b)
    order = np.argsort(sb,-1)
    keep_s = np.ones_like(sb,bool)
    for r in range(nref):
        k = order[...,r]
        ek = np.take_along_axis(e[:,None,:,:].repeat(len(cl),1), k[...,None,None].repeat(3,-1), 2)
        wk = np.take_along_axis(w[:,None,:].repeat(len(cl),1), k[...,None], 2)
        sk = np.take_along_axis(sb, k[...,None], 2)
        bound = (sb - sk) - (np.abs(e[:,None,:,:]-ek) * h[None,:,None,:]).sum(-1) - 4e-5*(w[:,None,:]+wk+2*pmax[None,:,None])
        keep_s &= ~(bound > 0)
    keep_b = keep_s[:,:,B].any(-1)
    return keep_b.sum(-1)
for nref in (1,2,3):
    K2 = lists2(nref); s2=np.sort(K2,1)[:,::-1]
    print("nref",nref,"blocks/cluster",K2.mean()," cur",K2.reshape(len(S),2,4).max(-1).sum(-1).mean()," sorted",(s2[:,0]+s2[:,4]).mean())
