import sys, ctypes as C, numpy as np
sys.path.insert(0,'hand-pose-estimation_b200'); sys.path.insert(0,'oracle')
import hpe_b200, orc
D=hpe_b200.defaults; lib=hpe_b200.lib()
om=orc.make_model(); rng=np.random.default_rng(1)
truth=np.clip(D.X0+rng.normal(size=26)*D.STD*0.5,D.LB,D.UB)
depth=orc.render_depth(om,truth)
hand,obs=hpe_b200.handmodel(),hpe_b200.observedmodel(capacity=1); obs.upload(depth[None]); cf=hpe_b200.costfunc(hand,obs)
cnt=C.c_ulonglong()
for spread in (1.0,0.3,0.1,0.03,0.01,0.0):
    th=truth+rng.normal(size=(20000,26))*D.STD*spread
    lib.hpe_debug_fallback_points(C.byref(cnt),1)
    cf.cal_cost_batch(th)
    lib.hpe_debug_fallback_points(C.byref(cnt),1)
    print(spread, "fallback points per eval:", cnt.value/20000, "per point", cnt.value/20000/250)
