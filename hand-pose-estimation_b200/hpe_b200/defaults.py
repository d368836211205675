"""Constants of the reference's driver, kept as data (no code is derived from them).

Values and where the reference states them (paths relative to /root/reference):
  HGEO_MM   misc/hgeo.dat (20 segment lengths, mm; loaded at src/testmodel.cpp:46, /10 at :50)
  RAD_MM    misc/rad.dat  (48 sphere radii, mm; loaded at :47, /10 at :49)
  SPACING, CMC, TB_N, FG_N, X0       src/testmodel.cpp:34-40
  UB, LB, STD                        src/testmodel.cpp:74-98
  FOCAL, W, H                        src/testmodel.cpp:60-62 (imgH=320 is the width)
"""
import numpy as np

HGEO_MM = np.array([
    3.110740369831595231e+01, 2.728002169518929776e+01, 2.200017175069322306e+01, 2.199995952496276530e+01,
    6.896344533164219115e+01, 3.080009178775444667e+01, 2.310057694761108138e+01, 2.090010474582363287e+01,
    7.427700757419079025e+01, 3.629982271538526817e+01, 2.420042802617343725e+01, 2.199987377350150197e+01,
    7.744081578614793671e+01, 3.190023249683458673e+01, 2.090000473229609312e+01, 1.979954859510691279e+01,
    8.015924911455944368e+01, 2.310022144042995862e+01, 1.429931098235506148e+01, 1.539998521398316811e+01,
], dtype=np.float64)

RAD_MM = np.array([
    18.26, 15.18, 16.61, 14.96, 13.53, 12.10, 10.45, 8.80,
    12.10, 11.66, 11.8067, 11.9533, 11.11, 10.56, 10.12, 9.68, 9.46, 9.24,
    16.50, 15.0333, 13.5667, 12.10, 11.55, 11.00, 10.78, 10.56, 10.01, 9.46,
    16.50, 10.56, 12.54, 14.52, 10.12, 9.68, 9.46, 9.24, 8.80, 8.36,
    12.10, 11.2933, 10.4867, 9.68, 9.46, 9.24, 8.80, 8.36, 7.81, 7.26,
], dtype=np.float64)

# the driver converts both to cm by an Armadillo `vec / 10.` (src/testmodel.cpp:49-50)
HGEO_CM = HGEO_MM / 10.0
RAD_CM = RAD_MM / 10.0

SPACING = np.array([-1.86, -1.86, 0.0, 1.91, 3.84], dtype=np.float64)
CMC = np.array([150.0, 107.5, 89.8, 76.5, 59.6], dtype=np.float64)
TB_N = np.array([2, 2, 2, 2], dtype=np.int32)
FG_N = np.array([4, 2, 2, 2], dtype=np.int32)
X0 = np.array([0, -10, -40, 0, 3, 32, 6, 9, 8, 9, 3, 9, 9, 6, 1, 9, 8, 7, 4, 8, 7, 6, 2, 7, 7, 7],
              dtype=np.float64)

_digit_ub = [15.0, 90.0, 110.0, 90.0]
_digit_lb = [-15.0, 0.0, 0.0, 0.0]
UB = np.array([180.0] * 3 + [100.0] * 3 + _digit_ub * 5, dtype=np.float64)
LB = np.array([-180.0] * 3 + [-100.0] * 3 + _digit_lb * 5, dtype=np.float64)
STD = np.array([9.0] * 3 + [7.0] * 3 + [9.0] * 20, dtype=np.float64)

FOCAL = 241.42
W, H = 320, 240
NUM_SAMPLES = 250      # src/observedmodel.cpp:210
NSPH, NDOF, NJOINT = 48, 26, 21
