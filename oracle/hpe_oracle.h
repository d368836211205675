/*
 * hpe_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, double-precision restatement of the hot path of
 * hjurong/hand-pose-estimation (reference mounted at /root/reference while
 * developing).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, link or call anything in
 * this directory; the product (libhpe_b200.so) never does.
 *
 * PARITY: the reference ships no tests, golden vectors or fixtures and needs Armadillo +
 * OpenCV C++, which the image lacks; oracle/_ref is therefore the reference's OWN hot-path
 * sources compiled unmodified against minimal stand-in headers (oracle/shim/, see oracle/Makefile
 * and oracle/ref_driver.cpp).  tests/test_ref_build.py and tests/test_oracle_refine.py hold this
 * restatement bit-identical to that build (FK, observation, energy terms, correspondences,
 * pso_evolve on replayed random streams, refine_init_pose, gnd_truth_err, dim_restore); what the
 * stand-ins restate of the two libraries is pinned separately:
 *   - cv2.BFMatcher(NORM_L2)  == orc_correspondences       (bit-exact)
 *   - cv2.distanceTransform(DIST_L2, 5) with IPP off == orc_chamfer5 (bit-exact)
 *   - Armadillo's sum/mean/dot order (two accumulators) and left-to-right small dense products;
 *     real Armadillo may hand products to BLAS, whose order is unspecified: that residue is
 *     unpinned and stated in DESIGN.md section 2.
 * Armadillo's RNG is replaced by a counter-based Philox4x32-10 stream that the
 * GPU replays ("RNG").
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef HPE_ORACLE_H
#define HPE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NSPH 48
#define ORC_NDOF 26
#define ORC_NJOINT 21
#define ORC_W 320
#define ORC_H 240

/* handmodel + 4 fingermodel + 1 thumbmodel state (src/handmodel.h:34-46,
 * src/fingermodel.h:35-45, src/thumbmodel.h:35-43).  base/back are the
 * pose-independent matrices the reference latches on the first call
 * (T01/T10 for a finger, Trf/T10 for the thumb), row-major 4x4. */
typedef struct {
    double geo[20];     /* cm: thumb L0..L3, index L4..L7, middle, ring, little */
    float  spacing[5];  /* stored as float in the reference (fingermodel.h:43) */
    double cmc[5];      /* degrees */
    int    tb_n[4];
    int    fg_n[4];
    double rad[ORC_NSPH];
    double base[5][16];
    double back[5][16];
} orc_model;

/* observedmodel state after next_frame (src/observedmodel.h:13-29). */
typedef struct {
    int     n_hand;       /* non-zero pixels */
    int     n_pts;        /* rows of the (subsampled) cloud */
    double  scale;        /* cm per pixel, observedmodel.cpp:195-202 */
    double  focal, cx, cy;
    double  dt_max;       /* disttrans.max(), costfunc.cpp:298 */
    double *depth;        /* H*W, cm, row-major depth[v*W+u] */
    double *dt;           /* H*W, px */
    double *cloud;        /* n_pts*3, (X,-Y,-Z) */
} orc_obs;

typedef struct {
    double lb[ORC_NDOF], ub[ORC_NDOF], std[ORC_NDOF];
    int    maxiter;       /* counts the initial evaluation, PSO.cpp:766,778 */
} orc_pso_params;

/* ---- model / forward kinematics ---- */
void orc_model_init(orc_model *m, const double geo_cm[20], const float spacing[5],
                    const double cmc_deg[5], const int tb_n[4], const int fg_n[4],
                    const double rad_cm[ORC_NSPH]);
void orc_build_hand_model(const orc_model *m, const double theta[ORC_NDOF],
                          double spheres[ORC_NSPH * 3], double joints[ORC_NJOINT * 3]);

/* ---- observation ---- */
int  orc_obs_from_depth(const float *depth_mm, double focal, int downsample,
                        int num_samples, orc_obs *o);
int  orc_obs_from_parts(const double *depth_cm, const double *dt_px, const double *cloud,
                        int n_pts, double scale, double focal, orc_obs *o);
void orc_obs_free(orc_obs *o);
void orc_chamfer5(const uint8_t *mask /*0 = feature*/, int w, int h, float *out);

/* ---- energy ---- */
void   orc_correspondences(const double *cloud, int n_pts, const double *spheres, int32_t *match);
double orc_align(const double *rad, const double *spheres, const double *cloud, int n_pts,
                 const int32_t *match);
double orc_depth_penalty(const orc_model *m, const orc_obs *o, const double *spheres_unneg);
double orc_collision(const double *spheres, const double *rad);
double orc_cal_cost(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF]);
double orc_cal_cost2(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF],
                     int32_t *match, int compute_corr, double terms[3]);
/* N poses; terms written separately; diagnostics optional (may be NULL):
 * floor_margin[i] = min distance of any projected centre coordinate to an
 * integer, match_gap[i] = min over points of (second best - best) float32
 * matcher distance.  OpenMP over poses. */
void orc_cost_batch(const orc_model *m, const orc_obs *o, const double *theta, int64_t n,
                    double *align, double *depth, double *colli,
                    double *floor_margin, double *match_gap);

/* ---- RNG (Philox4x32-10, DESIGN.md "RNG") ---- */
void   orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
void   orc_uniform2(uint64_t seed, uint32_t elem_pair, uint32_t particle, uint32_t gen,
                    uint32_t stream, double u[2]);
void   orc_normal2(uint64_t seed, uint32_t elem_pair, uint32_t particle, uint32_t gen,
                   double z[2]);

/* ---- optimiser ---- */
void orc_generate_particles(const orc_pso_params *p, const double x0[ORC_NDOF], int num_p,
                            uint64_t seed, double *particles /* num_p*26 */);
void orc_check_constraints(const orc_pso_params *p, double theta[ORC_NDOF], double vel[ORC_NDOF]);
/* returns 1 like the reference.  trace (maxiter-1) and final state are optional. */
int  orc_pso_evolve(const orc_model *m, const orc_obs *o, const orc_pso_params *p,
                    const double x0[ORC_NDOF], int num_p, uint64_t seed,
                    double bestp[ORC_NDOF], double *best_cost, double *trace,
                    double *final_particles, double *final_pcost);

/* ---- initial-pose refinement, ground-truth error ---- */
void orc_cal_grad(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF],
                  const int sel[ORC_NDOF], int32_t *match, double grad[ORC_NDOF]);
int  orc_goldstein(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF],
                   const double g_k[ORC_NDOF], int32_t *match, double f_k, double *tk, int maxiter);
/* cv::circle with thickness >= 2 (outline) as a mask, and costfunc::depthMatch_penalty (src/costfunc.cpp:439-473) */
void orc_circle_outline(uint8_t *mask, int w, int h, int cx, int cy, int radius, int thickness);
void orc_depthmatch_penalty(const orc_model *m, const orc_obs *o, const double *spheres, double *out);
/* smallest relative gap of a line-search comparison since the last reset (test instrumentation) */
void orc_margin_reset(void);
double orc_margin_min(void);
void orc_refine_init_pose(const orc_model *m, const orc_obs *o, double x0[ORC_NDOF], int *evals);
/* PSO::pso_optimise (src/PSO.cpp:539-712): coordinate descent + canonical PSO; streams in hpe_oracle.c */
int  orc_pso_optimise(const orc_model *m, const orc_obs *o, const orc_pso_params *p,
                      double w, double c1, double c2, const double x0[ORC_NDOF], int num_p,
                      uint64_t seed, double bestp[ORC_NDOF], double *best_cost, int64_t *evals,
                      double *final_particles, double *final_pcost);
/* costfunc::bincomp_penalty (src/costfunc.cpp:384-436) and the filled cv::circle it draws with */
void   orc_fill_circle(uint8_t *mask, int w, int h, int cx, int cy, int radius);
double orc_bincomp_penalty(const orc_model *m, const orc_obs *o, const double *spheres /* 48x3, negated y,z */,
                           uint8_t *silhouette /* H*W or NULL */);
void orc_dim_restore(const double in[22], double out[ORC_NDOF]);
double orc_gnd_truth_err(const double joints[ORC_NJOINT * 3], const double gt_row[63]);

/* ---- synthetic data (test helper, not a reference function) ---- */
void orc_render_depth(const double *spheres_unneg, const double *rad, double focal,
                      float *depth_mm /* H*W */);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
