"""hpe_b200 -- Python binding of libhpe_b200.so (ctypes over the C ABI in include/hpe.h).

The classes mirror the reference's C++ host classes (same names, same method names and
argument meaning; numpy arrays instead of arma::vec):

  handmodel      src/handmodel.h:9-32       build_hand_model, hand_joints, get_radii
  observedmodel  src/observedmodel.h:34-68  next_frame, get_ptncloud, get_disttran, get_img_scale
  costfunc       src/costfunc.h:25-41       cal_cost, cal_cost2, compute_correspondences
  PSO            src/PSO.h:56-68            set_pso_params, pso_evolve, refine_init_pose

plus batched entry points (many poses / many frames per call), which is how the B200 is fed.
There is no CPU fallback: importing works anywhere, every compute call needs the CUDA library
and a GPU and raises HpeError otherwise.
"""
from . import defaults
from .binding import (HpeError, lib, library_path, device_count, handmodel, observedmodel, costfunc, PSO, Swarm,
                      PsoParams, render_frames, set_cost_pipeline)

__all__ = ["defaults", "HpeError", "lib", "library_path", "device_count", "handmodel", "observedmodel",
           "costfunc", "PSO", "Swarm", "PsoParams", "render_frames", "set_cost_pipeline"]
