import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hand-pose-estimation_b200"))
import hpe_b200
lib = hpe_b200.lib()
out = (C.c_double * 2)()
print("peak (TFLOP/s, implied MHz):", hpe_b200.binding.measure_fp32_peak())
assert lib.hpe_measure_fp32_patterns(out) == 0
print("3-register FFMA pattern: %.1f TFLOP/s; sweep-shaped pattern (FFMA only counted): %.1f TFLOP/s" % (out[0], out[1]))
out2 = (C.c_double * 2)()
assert lib.hpe_measure_fp32x2(out2) == 0
print("FFMA2 peak: %.1f TFLOP/s; FFMA2 sweep shape: %.3f T pairs/s (= %.1f TFLOP/s at 6 flop/pair; scalar sweep = %.3f T pairs/s)"
      % (out2[0], out2[1], out2[1] * 6, out[1] / 6))
