/*
 * hpe.h -- C ABI of libhpe_b200.so, the B200 (sm_100a) implementation of the
 * particle-swarm hot path of hjurong/hand-pose-estimation.
 *
 * The reference has no FFI layer: its boundary is the C++ class API
 * (handmodel / observedmodel / costfunc / PSO).  Each entry point below names
 * the reference interface it replaces (file:line relative to the reference
 * root).  hpe_classes.hpp wraps these in C++ classes with the reference's
 * names; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function
 * returns an hpe_status (0 = ok); the caller owns every buffer; opaque handles
 * own device memory on the CUDA device that was current when they were
 * created.  There is no CPU fallback: without a CUDA device every compute call
 * returns HPE_ERR_CUDA.
 *
 * Devices and threads: a process may drive several GPUs (cudaSetDevice between
 * calls).  A handle is only valid while ITS device is current -- a call with a
 * handle of another device returns HPE_ERR_STATE instead of launching with
 * foreign pointers.  The library's own caches (SM count, kernel attributes,
 * streams, events, scratch buffers, the streaming-job slots of
 * hpe_cost_batch_submit) are kept per device, so calls on different devices
 * are independent.  Calls on ONE device must come from one host thread at a
 * time: the launches of a device share its scratch buffers and streams.
 *
 * Units follow the reference after its mm->cm conversion: lengths in cm,
 * angles in degrees, depth frames in mm (the on-disk format), distance
 * transform in pixels.
 */
#ifndef HPE_B200_ABI_H
#define HPE_B200_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPE_NDOF 26
#define HPE_NSPH 48
#define HPE_NJOINT 21
#define HPE_IMG_W 320
#define HPE_IMG_H 240

typedef enum {
    HPE_OK = 0,
    HPE_ERR_ARG = 1,          /* bad pointer / size / unsupported sphere layout */
    HPE_ERR_CUDA = 2,         /* CUDA runtime error or no device */
    HPE_ERR_EMPTY_FRAME = 3,  /* a frame without hand pixels (reference: Armadillo bounds error) */
    HPE_ERR_STATE = 4         /* handle used before it was filled */
} hpe_status;

typedef struct hpe_model hpe_model; /* handmodel + 4 fingermodel + thumbmodel geometry */
typedef struct hpe_obs hpe_obs;     /* a batch of observedmodel states (frames) on the device */
typedef struct hpe_swarm hpe_swarm; /* one large swarm stepped generation by generation */

/* PSO::set_pso_params (src/PSO.cpp:38-54).  w,c1,c2,minstep,minfunc are accepted
 * for signature fidelity and ignored, exactly as pso_evolve ignores them
 * (src/PSO.cpp:769-774 hard-codes W1, C1, C2, nK=3). */
typedef struct {
    double ub[HPE_NDOF];
    double lb[HPE_NDOF];
    double std[HPE_NDOF];
    double w, c1, c2;
    int    maxiter;          /* counts the initial evaluation (src/PSO.cpp:766,778) */
    double minstep, minfunc;
} hpe_pso_params;

const char *hpe_last_error_string(void);
int hpe_version(void);
/* number of CUDA devices visible; 0 when there is none */
int hpe_device_count(void);

/* ---- handmodel (src/handmodel.h:9-32) --------------------------------- */
/* handmodel::handmodel(h_geo, h_spacing, tb_spheres, fg_spheres, h_CMC, sphR),
 * src/handmodel.cpp:10-27.  spacing is narrowed to float as fingermodel.h:43 does.
 * Only the reference's sphere layout {2,2,2,2}/{4,2,2,2} is accepted (its
 * fixed-size matrices allow no other, src/handmodel.cpp:272-274). */
int hpe_model_create(const double geo_cm[20], const double spacing_cm[5], const double cmc_deg[5],
                     const int tb_spheres[4], const int fg_spheres[4], const double rad_cm[HPE_NSPH],
                     hpe_model **out);
int hpe_model_destroy(hpe_model *m);

/* handmodel::build_hand_model + hand_joints (src/handmodel.cpp:259-298) for n poses.
 * theta: host n*26.  spheres: host n*48*3 in the negated-y,z frame (may be NULL).
 * joints: host n*21*3, not negated (may be NULL). */
int hpe_fk_batch(const hpe_model *m, const double *theta, int64_t n, double *spheres, double *joints);

/* ---- observedmodel (src/observedmodel.h:34-68) ------------------------ */
/* Capacity for `capacity` frames; num_samples = 250 in the reference
 * (src/observedmodel.cpp:210); focal = 241.42 (src/testmodel.cpp:60). */
int hpe_obs_create(int capacity, int num_samples, double focal, hpe_obs **out);
int hpe_obs_destroy(hpe_obs *o);
/* observedmodel::next_frame for `count` frames (src/observedmodel.cpp:420-430 =
 * load_data :272-310 + depth_to_ptncloud :110-219 + dist_transform :340-369), all on
 * the device.  depth_mm: HOST float32 [count][240][320] in the file layout
 * (value at y*320+x, mm >= 0 as in the unsigned file format, 0 = background).  Frames are stored in slots
 * first..first+count-1.  On the device the distance transform is kept FOLDED with the depth map (hand pixel -> minus
 * its depth, background -> its distance: one gather per projected sphere in the depth term); hpe_obs_get returns the
 * plain transform. */
int hpe_obs_upload(hpe_obs *o, const float *depth_mm, int first, int count);
/* same with the frames already in device memory */
int hpe_obs_upload_dev(hpe_obs *o, const float *depth_mm_dev, int first, int count);
/* inject an externally preprocessed observation into one slot (host pointers):
 * depth_mm [240*320], dt_px [240*320] (the plain transform; folded on the device), cloud [n_pts*3] = (X,-Y,-Z), scale. */
int hpe_obs_set(hpe_obs *o, int slot, const float *depth_mm, const float *dt_px, const double *cloud,
                double scale);
/* getters: get_ptncloud / get_disttran / get_img_scale (src/observedmodel.cpp:375-417).
 * Any output may be NULL.  cloud: n_pts*3, dt_px: 240*320. */
int hpe_obs_get(const hpe_obs *o, int slot, double *cloud, float *dt_px, double *scale, int *n_hand,
                float *dt_max);
int hpe_obs_num_samples(const hpe_obs *o);
/* device pointer to the stored depth frames [capacity][240][320] (for on-device generators) */
float *hpe_obs_depth_dev(hpe_obs *o);

/* ---- costfunc (src/costfunc.h:25-41) ---------------------------------- */
/* costfunc::cal_cost / cal_cost2 for n poses against frame `slot`
 * (src/costfunc.cpp:89-127, 31-86): align = align_models (:346-377) after
 * compute_correspondences (:306-343), depth = depth_penalty (:227-304),
 * colli = self_collision_penalty (:130-197, only computed when colli != NULL).
 * cal_cost = align + depth; cal_cost2 = align + depth + colli.  HOST pointers; any
 * output may be NULL. */
int hpe_cost_batch(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n,
                   double *align, double *depth, double *colli);
/* Streaming form for callers that score many independent batches (the CostFunc batch check, offline
 * scoring): submit queues the upload, the kernels and the download of one batch and returns a job id,
 * wait blocks until that batch's results are in its align/depth/colli buffers.  Up to HPE_COST_JOBS
 * batches may be in flight (a further submit returns HPE_ERR_STATE); they run in submission order, the
 * upload of the next under the kernels of the current.  theta and the output arrays must stay valid
 * until the wait returns and should be page-locked (pageable memory makes the copies synchronous).
 * Results are those of hpe_cost_batch, bit for bit.  One submitting host thread; the frame in `slot`
 * must not be replaced (hpe_obs_upload / hpe_obs_set) while a job that reads it is in flight. */
#define HPE_COST_JOBS 3
int hpe_cost_batch_submit(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n,
                          double *align, double *depth, double *colli, int *job);
int hpe_cost_batch_wait(int job);
/* same on DEVICE pointers, asynchronous on `stream` (a cudaStream_t, may be NULL).
 * cost receives align+depth(+colli when with_collision); terms (may be NULL) receives three
 * arrays of n: align | depth | colli. */
int hpe_cost_batch_dev(const hpe_model *m, const hpe_obs *o, int slot, const double *theta_dev,
                       int64_t n, int with_collision, double *cost_dev, double *terms_dev,
                       void *stream);
/* cal_cost2(theta, matchId, compute_corr=false): frozen correspondences (src/costfunc.cpp:53-58).
 * match: host n_pts int32 shared by all n poses. */
int hpe_cost_batch_frozen(const hpe_model *m, const hpe_obs *o, int slot, const double *theta,
                          int64_t n, const int32_t *match, double *align, double *depth, double *colli);
/* compute_correspondences for n poses: match host n*n_pts int32 */
int hpe_correspondences(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n,
                        int32_t *match);

/* costfunc::gnd_truth_err (src/costfunc.cpp:476-507) for n poses: the joints of theta[i] (x10 -> mm,
 * y and z negated) against gt_mm[i] = one row of the data set's joint.txt (63 values, x y z of the 21
 * joints, mm); err[i] = summed distance over the wrist and the five finger tips (joints 0,4,8,12,16,20).
 * HOST pointers. */
int hpe_gnd_truth_err(const hpe_model *m, const double *theta, const double *gt_mm, int64_t n, double *err);
/* costfunc::bincomp_penalty (src/costfunc.cpp:384-436; "a new cost term used for testing") for n poses
 * against frame `slot`: the 48 spheres of theta[i] drawn as filled discs (centre = trunc(K c / c_z),
 * radius = trunc(r / img_scale), cv::circle with thickness -1) into an empty 240x320 image;
 * penalty[i] = number of pixels where that silhouette and the observed one (depth != 0) differ.
 * silhouette (may be NULL): n*2400 words, bit (x & 31) of word y*10 + x/32 = pixel (x, y) drawn --
 * the projected-sphere image the reference's debug views show. */
int hpe_bincomp_batch(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n,
                      double *penalty, uint32_t *silhouette);

/* costfunc::depthMatch_penalty (src/costfunc.cpp:439-473) for n poses against frame `slot`: image[i] (240*320 doubles,
 * row-major) = the observed depth map in cm with the OUTLINE of every projected sphere of theta[i] drawn into it in
 * sphere order -- centre = trunc(K c / c_z), radius = trunc(r / img_scale), cv::circle with thickness 2, value = the
 * sphere's depth c_z.  The reference returns this cv::Mat for its debug views.  The raster (which pixels, overwritten in
 * which order) is bit-identical; the drawn values are forward-kinematics outputs (1e-15 relative). */
int hpe_depthmatch_batch(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n,
                         double *image);

/* ---- PSO (src/PSO.h:56-68) -------------------------------------------- */
/* PSO::pso_evolve (src/PSO.cpp:717-886) for `count` independent swarms, swarm b fitted to
 * frame slot first+b with initial pose x0[b].  seeds: count values or NULL (every swarm
 * uses 1000, the constant at src/PSO.cpp:722).  Outputs (host): bestp count*26,
 * best_cost count (gbest_cost), trace count*(maxiter-1) = bcost_evo (may be NULL). */
int hpe_pso_evolve(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                   const double *x0, int first, int count, const uint64_t *seeds, double *bestp,
                   double *best_cost, double *trace);
/* the tracking loop of test_full without refine_init_pose (src/testmodel.cpp:117-138):
 * frame k starts from the best pose of frame k-1.  poses: n_frames*26, costs: n_frames. */
int hpe_track_sequence(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p,
                       int num_particles, const double x0[HPE_NDOF], int first, int n_frames,
                       uint64_t seed, double *poses, double *costs);

/* n_seq independent sequences tracked AT ONCE (one launch; one thread-block cluster -- or CTA -- per sequence, so the
 * whole GPU works where hpe_track_sequence occupies 16 SMs).  The reference's loop (src/testmodel.cpp:117-138,
 * src/visualiser.cpp:262-296) is one sequence; this is the same loop over n_seq of them.  Frame k of sequence q lives in
 * slot first + q*seq_stride + k*frame_stride with (seq_stride, frame_stride) = (n_frames, 1) (sequence after sequence) or
 * (1, n_seq) (frame-interleaved); x0: n_seq*26; seeds: n_seq values (every frame of sequence q uses seeds[q]) or NULL
 * (1000); poses: n_seq*n_frames*26 and costs: n_seq*n_frames, sequence-major.  Each sequence's result is bit-identical
 * to hpe_track_sequence on that sequence alone. */
int hpe_track_sequences(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                        const double *x0, int first, int n_seq, int n_frames, int seq_stride, int frame_stride,
                        const uint64_t *seeds, double *poses, double *costs);
/* Asynchronous device-pointer forms of the three calls above (and of hpe_refine_init_pose below): every pointer is a
 * device pointer, nothing is copied or synchronised, the work is queued on `stream` (a cudaStream_t; NULL = the default
 * stream).  seeds_dev of hpe_track_sequences_dev holds one value per FRAME (n_seq*n_frames, sequence-major) or is NULL.
 * The slots must already hold usable observations. */
int hpe_pso_evolve_dev(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                       const double *x0_dev, int first, int count, const uint64_t *seeds_dev,
                       double *bestp_dev, double *best_cost_dev, double *trace_dev, void *stream);
int hpe_track_sequences_dev(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                            const double *x0_dev, int first, int n_seq, int n_frames, int seq_stride,
                            int frame_stride, const uint64_t *seeds_dev, double *poses_dev, double *costs_dev,
                            void *stream);
int hpe_refine_init_pose_dev(const hpe_model *m, const hpe_obs *o, double *x0_dev, int first, int count,
                             int *evals_dev, void *stream);
/* PSO::refine_init_pose (src/PSO.cpp:216-266) with cal_grad (:183-214) and goldstein (:438-480) for
 * `count` frames: x0 (host, count*26) is refined in place against frame slot first+b -- gradient
 * descent with a Goldstein line search on theta[0:3], then on theta[3:6], every evaluation a
 * cal_cost2 (frozen correspondences for the probes and the line search).  evals (may be NULL)
 * receives the number of cal_cost2 evaluations the reference's sequential loop would have made. */
int hpe_refine_init_pose(const hpe_model *m, const hpe_obs *o, double *x0, int first, int count, int *evals);
/* PSO::pso_optimise (src/PSO.cpp:539-712; the reference's "gradient descent + pso" test method, call
 * site src/visualiser.cpp:521) for `count` independent swarms, swarm b fitted to frame slot first+b
 * from x0[b]: per iteration 10 coordinate-descent steps per particle (cal_gradient :380-405 +
 * goldstein :438-480 on cal_cost2), then the canonical update with p->w, p->c1, p->c2 towards
 * pbest and gbest, check_constraints and cal_cost.  Outputs (host): bestp count*26 (gbest_pos),
 * best_cost count (gbest_cost; may be NULL), evals count (cal_cost + cal_cost2 calls of the
 * sequential loop; may be NULL), final_particles count*num_particles*26 and final_pcost
 * count*num_particles (may be NULL). */
int hpe_pso_optimise(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                     const double *x0, int first, int count, const uint64_t *seeds, double *bestp,
                     double *best_cost, int64_t *evals, double *final_particles, double *final_pcost);
/* the tracking loop of test_full as written (src/testmodel.cpp:117-138): per frame
 * refine_init_pose(x0), pso_evolve(x0) -> bestp, x0 = bestp.  poses: n_frames*26 (bestp), costs:
 * n_frames (gbest_cost), refined (may be NULL): n_frames*26, the pose each swarm started from. */
int hpe_track_sequence_refined(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p,
                               int num_particles, const double x0[HPE_NDOF], int first, int n_frames,
                               uint64_t seed, double *poses, double *costs, double *refined);
/* the same loop for n_seq sequences in lock step; slots are frame-interleaved: frame k of sequence q in slot
 * first + k*n_seq + q.  x0: n_seq*26, seeds: n_seq or NULL (1000); poses / refined: n_seq*n_frames*26 and costs:
 * n_seq*n_frames, sequence-major.  Per sequence bit-identical to hpe_track_sequence_refined. */
int hpe_track_sequences_refined(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p,
                                int num_particles, const double *x0, int first, int n_seq, int n_frames,
                                const uint64_t *seeds, double *poses, double *costs, double *refined);

/* Offline batch from HOST frames: count Subject1-format frames (float32 mm, count*240*320) are uploaded,
 * observed (observedmodel::next_frame) and fitted with one pso_evolve swarm each (x0: count*26), streaming
 * through the observation buffer `o` in chunks so that the upload of a chunk runs under the fits of the chunks
 * before it: two halves of the buffer, or -- from a capacity of 8 (12) swarm CTAs per SM, 1,184 (1,776) frames on a
 * B200 -- four (six) chunks of two CTAs per SM whose swarm kernels run concurrently on two (three) streams.  bestp: count*26, best_cost: count (NaN for a frame without hand pixels, whose
 * bestp is its x0).  Afterwards `o` holds the last chunks.  HOST pointers (pinned memory makes the overlap real). */
int hpe_fit_frames(const hpe_model *m, hpe_obs *o, const hpe_pso_params *p, int num_particles,
                   const float *depth_mm, const double *x0, int64_t count, const uint64_t *seeds,
                   double *bestp, double *best_cost);
/* The same in two steps, up to two calls in flight per device: submit queues everything (uploads, observations, swarms,
 * the copy of the results) and returns a job id, wait blocks until the job's results are in bestp / best_cost.  With a
 * second call submitted before the first is waited for -- same `o`, whose frames neither call may touch meanwhile -- the
 * first uploads of the second run under the last swarms of the first (src/testmodel.cpp:117-138 has no counterpart: the
 * reference fits one frame at a time).  All host buffers must stay valid (and should be page-locked) until the wait. */
int hpe_fit_frames_submit(const hpe_model *m, hpe_obs *o, const hpe_pso_params *p, int num_particles,
                          const float *depth_mm, const double *x0, int64_t count, const uint64_t *seeds,
                          double *bestp, double *best_cost, int *job);
int hpe_fit_frames_wait(int job);

/* ---- one large swarm, particles [lo,hi) of num_particles owned by this GPU ---- */
/* State lives on the device; the caller (one process per GPU) exchanges the
 * pbest table between generations (all-gather of hpe_swarm_table rows). */
int hpe_swarm_create(const hpe_model *m, const hpe_obs *o, int slot, const hpe_pso_params *p,
                     int num_particles, int lo, int hi, const double x0[HPE_NDOF], uint64_t seed,
                     hpe_swarm **out);
int hpe_swarm_destroy(hpe_swarm *s);
/* device pointer to the global table [num_particles][27] doubles = (pcost, pbest[26]);
 * rows [lo,hi) are written by this GPU, the others must be filled by the exchange.  The rows are
 * double buffered: call this again after every hpe_swarm_init / hpe_swarm_step, the pointer alternates. */
double *hpe_swarm_table(hpe_swarm *s);
/* generation 0: generate_particles + evaluation of the owned rows (src/PSO.cpp:745-763) */
int hpe_swarm_init(hpe_swarm *s, void *stream);
/* after the exchange: global argmin / count bookkeeping (src/PSO.cpp:864-876) */
int hpe_swarm_reduce(hpe_swarm *s, void *stream);
/* one generation for the owned rows: topology, informants, velocity, clamp, evaluation,
 * pbest (src/PSO.cpp:781-861).  Requires a complete table. */
int hpe_swarm_step(hpe_swarm *s, void *stream);
/* Device-initiated exchange instead of the caller's all-gather (one process per GPU on one NVLink
 * node): every rank publishes the 64-byte CUDA IPC handle of its table, gathers all `world` handles (any
 * host transport) and attaches.  From then on hpe_swarm_init / hpe_swarm_step store the owned rows
 * directly into every peer's table over NVLink and raise a flag there, hpe_swarm_reduce waits on the
 * flags on the device; the caller issues init, reduce, (step, reduce) x (maxiter-1) with no exchange and
 * no host synchronisation in between.  Results are bit-identical to the all-gather path. */
int hpe_swarm_ipc_handle(hpe_swarm *s, unsigned char handle[64]);
int hpe_swarm_attach_peers(hpe_swarm *s, const unsigned char *handles /* world*64 */, int world, int rank);
/* The whole run -- hpe_swarm_init, then (hpe_swarm_step, exchange, hpe_swarm_reduce) x (maxiter-1) -- as ONE
 * cooperative kernel launch per GPU: rows go to the peers by NVLink stores from the warp that produced
 * them, ranks meet on device-side flags, 2 grid barriers per generation.  Peers must be attached when more
 * than one rank takes part (hpe_swarm_attach_peers); every rank calls it once.  Bit-identical to the
 * stepwise path.  Asynchronous on `stream`. */
int hpe_swarm_run(hpe_swarm *s, void *stream);
/* bound of the device-side waits on peer flags, in polls of ~50-100 ns (default 2^28): a peer that never delivers makes
 * hpe_swarm_best return HPE_ERR_STATE instead of hanging the GPU; tests lower the bound */
int hpe_swarm_set_wait_limit(long long polls);
/* diagnostics: first call switches phase-timestamp recording on; later calls copy the stamps of the last
 * hpe_swarm_run to out (maxiter*8 values, ns) */
int hpe_swarm_debug_times(hpe_swarm *s, unsigned long long *out);
/* host copies of gbest (and of bcost_evo, maxiter-1 values, may be NULL) after the last hpe_swarm_reduce */
int hpe_swarm_best(hpe_swarm *s, double bestp[HPE_NDOF], double *best_cost, double *trace);

/* ---- synthetic frames (also the overlay of src/costfunc.cpp:439-473) ---- */
/* ray-cast the sphere model at n poses into depth frames (mm, 0 = background).
 * theta_dev: device n*26; depth_mm_dev: device n*240*320. */
int hpe_render_dev(const hpe_model *m, const double *theta_dev, int64_t n, double focal,
                   float *depth_mm_dev, void *stream);
/* host convenience: theta host n*26 -> depth_mm host n*240*320 */
int hpe_render(const hpe_model *m, const double *theta, int64_t n, double focal, float *depth_mm);

/* Page-locked host buffers for the streaming entry points (hpe_cost_batch_submit, hpe_fit_frames): copies from / to
 * pageable memory are staged by the driver and block the calling thread, so uploads only run under kernels when the
 * caller's buffers are pinned.  Any pinned memory works (cudaHostAlloc, torch pin_memory); these two are for callers
 * without a CUDA runtime of their own.  A buffer must outlive the jobs that use it. */
int hpe_host_alloc(size_t bytes, void **out);
int hpe_host_free(void *p);

/* ---- diagnostics ---- */
/* Nearest-centre search used by the cost kernels launched through hpe_cost_batch*,
 * hpe_correspondences: 0 (default) = certified two-sweep search, 1 = the reference's arithmetic
 * on all 48 spheres.  Both return cv::BFMatcher's indices; 1 exists to cross-check 0. */
int hpe_set_matcher(int mode);
/* host path of hpe_cost_batch: mode 0 (default) = upload / compute / download streams, one kernel launch per
 * uploaded chunk of 16,384 poses; mode 1 (batches of >= 32,768 poses) = ONE kernel launched before the upload
 * that scores the chunks as they land ("chase"; chunk_poses = poses per upload chunk, 0 = a decreasing
 * schedule).  Mode 1 needs copies and kernels to run concurrently: under tools that serialise them (ncu,
 * compute-sanitizer, CUDA_LAUNCH_BLOCKING=1) the kernel gives up after ~2 s and the call returns
 * HPE_ERR_CUDA.  Results are identical. */
int hpe_set_cost_pipeline(int mode, int chunk_poses);
/* kernel behind hpe_pso_evolve / hpe_track_sequence: 0 (default) = automatic (one 8-CTA cluster per swarm
 * for sequences and batches of <= 222 swarms or sequences, one CTA per swarm otherwise), 1 = always one CTA per swarm,
 * 2 = always one cluster per swarm, 3 = one cluster per swarm with the swarm state in global memory instead of
 * distributed shared memory (cross-check; the default for swarms of more than 128 particles).  Results are
 * bit-identical. */
int hpe_set_pso_path(int mode);
/* distance transform used by hpe_obs_upload*: 0 (default) = OpenCV's two raster passes, a row at a time with
 * the in-row recurrence as a prefix-minimum scan; 1 = closed-form row search; 2 = the two passes as a skewed
 * wavefront.  Bit-identical results; 1 and 2 exist to cross-check 0. */
int hpe_set_chamfer(int mode);
/* cycles of ONE evaluation by ONE warp on an otherwise idle SM, per phase: forward kinematics | depth term +
 * staging | matching + alignment (the latency that bounds the one-particle-per-warp paths) */
int hpe_debug_eval_latency(const hpe_model *m, const hpe_obs *o, int slot, const double theta[HPE_NDOF],
                           long long cycles[3]);
/* host only (no device needed): how hpe_cost_batch* hands n poses out to the warps of its kernel on a device with `sms`
 * SMs -- plan = { warps, poses per group, poses per tail unit, n1, n2, t2 }: a warp takes groups from a counter that runs
 * over [0, n1) and, once that is used up, tickets t from a second counter: ticket t < t2 is the unit of `poses per tail
 * unit` poses starting at n1 + unit * t (cut at n2), later tickets are the single poses n2 + (t - t2) < n. */
int hpe_debug_cost_plan(int64_t n, int sms, int plan[6]);
/* number of points the certified search handed to its exact fallback since the last reset */
int hpe_debug_fallback_points(unsigned long long *count, int reset);

/* device self-test: the division-free mm/10 of the depth term against the IEEE division over every
 * finite float; *mismatches must come back 0 */
int hpe_selftest_div10(unsigned long long *mismatches);
/* the same for the divide-free x / focal of the back-projection (src/observedmodel.cpp:139-143): x = (column -
 * 160) * (mm / 10) for every float mm in [1, 4096) and a spread of columns; *mismatches must come back 0 */
int hpe_selftest_div_const(double divisor, unsigned long long *mismatches);

/* ---- measurement ---- */
/* FP32 (non-tensor) FFMA peak of the current device in TFLOP/s, measured with a
 * register-only FFMA kernel; the denominator of the cost kernel's roofline (the
 * driver-written MEASURED_PEAKS.json carries only HBM and bf16-tensor figures). */
int hpe_measure_fp32_peak(double *tflops, double *sm_mhz_est);
/* issue-behaviour probes quoted in DESIGN.md: out[0] 3-register FFMA pattern, out[1] sweep-shaped scalar loop (TFLOP/s) */
int hpe_measure_fp32_patterns(double out[2]);
/* out[0] FFMA2 peak (TFLOP/s), out[1] sweep-shaped FFMA2 loop (1e12 point-sphere pairs/s) */
int hpe_measure_fp32x2(double out[2]);

#ifdef __cplusplus
}
#endif
#endif
