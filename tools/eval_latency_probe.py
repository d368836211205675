"""cycles of one cal_cost evaluation by one warp on an idle SM, per phase"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
D = hpe_b200.defaults
hand = hpe_b200.handmodel()
rng = np.random.default_rng(4)
truth = np.clip(D.X0 + rng.normal(size=26) * D.STD * 0.5, D.LB, D.UB)
obs = hpe_b200.observedmodel(capacity=1)
obs.upload(hpe_b200.render_frames(hand, truth[None]))
th = np.ascontiguousarray(truth + rng.normal(size=26) * D.STD * 0.3)
cy = (C.c_longlong * 3)()
rc = hpe_b200.lib().hpe_debug_eval_latency(hand._h, obs._h, 0, th.ctypes.data_as(C.POINTER(C.c_double)), cy)
assert rc == 0
tot = sum(cy)
print("cycles per evaluation (one warp, idle SM): FK %d | depth+staging %d | matching+alignment %d | total %d = %.2f us at 1.965 GHz"
      % (cy[0], cy[1], cy[2], tot, tot / 1965.0))
