// hpe_internal.h -- host-side handle layouts and launch wrappers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hpe.h"
#include "hpe_device.cuh"

// per-frame scalars produced by the observation kernels (device resident)
struct FrameMeta {
    double scale;     // cm per pixel (src/observedmodel.cpp:195-202)
    double dt_max;    // disttrans.max() (src/costfunc.cpp:298)
    int    n_hand;    // non-zero pixels
    int    ok;        // 1 once the slot holds a usable observation
};

// device view of an hpe_obs, passed to kernels by value
struct ObsView {
    float     *depth_mm;  // [capacity][240][320]
    float     *dt_px;     // [capacity][240][320] folded: distance transform on the background, -depth_mm on hand pixels (fold_depth_kernel)
    double    *cloud;     // [capacity][3][n_pts]
    FrameMeta *meta;      // [capacity]
    unsigned  *wf_halo;   // [capacity][2][320] scratch of the wavefront chamfer (band halo rows)
    unsigned  *mask;      // [capacity][2400] hand-pixel bits, 32 consecutive pixels per word (cloud_kernel -> chamfer_scan_kernel)
    unsigned char *perm;  // [capacity][HPE_MAXPTS] cluster order of the cloud: sorted position -> point index (cluster_kernel)
    unsigned char *image; // [capacity][HPE_IMAGE_BYTES] what every consumer CTA stages of the frame's cloud (cluster_kernel)
    double focal, cx, cy;
    int n_pts;
    int capacity;
};

__device__ __forceinline__ FrameView frame_view(const ObsView &o, int slot)
{
    FrameView f;
    f.depth_mm = o.depth_mm + (size_t)slot * HPE_W * HPE_H;
    f.dt_px = o.dt_px + (size_t)slot * HPE_W * HPE_H;
    f.cloud = o.cloud + (size_t)slot * 3 * o.n_pts;
    f.perm = o.perm + (size_t)slot * HPE_MAXPTS;
    f.image = reinterpret_cast<const uint4 *>(o.image + (size_t)slot * HPE_IMAGE_BYTES);
    f.scale = o.meta[slot].scale;
    f.dt_max = o.meta[slot].dt_max;
    f.focal = o.focal; f.cx = o.cx; f.cy = o.cy;
    f.n_pts = o.n_pts;
    return f;
}

struct hpe_model {
    DevModel dm;
    int device;
};

struct hpe_obs {
    ObsView v;
    int device;
    void *scratch;        // chamfer / compaction workspace
    size_t scratch_bytes;
};

struct PsoDev {           // constants of one pso_evolve call, passed by value
    double lb[HPE_NDOF], ub[HPE_NDOF], std[HPE_NDOF];
    int maxiter;
    int num_p;
};

// device-resident control block of a large swarm (hpe_swarm.cu)
struct SwarmCtl {
    double gbest_cost;
    double gbest[HPE_NDOF];
    int count;
    int iter;                         // evaluations done so far (1 after init)
    int error;                        // sticky: a bounded wait of the peer exchange gave up (nothing else writes it)
    int pad;
};

// arguments of the fused update + evaluate + pbest kernel (swarm_step_kernel, hpe_cost.cu)
struct SwarmStepArgs {
    PsoDev prm;
    uint64_t seed;
    const double *table_in;           // [num_p][27] rows of the previous generation
    double *table_out;                // rows of this generation (owned rows written here)
    const int *inf;                   // informant of every particle
    const SwarmCtl *ctl;
    double *x, *v;                    // owned positions / velocities [own][26]
    double W1, C1;
    int lo, own, update;
};

// arguments of the one-launch sharded swarm (swarm_run_kernel, hpe_cost.cu)
#define HPE_MAX_PEERS 16
struct SwarmMin {                              // a rank's first minimum of pcost over its own rows
    unsigned long long key;                    // IEEE bits of the cost (costs are >= 0)
    int idx, pad;
};
struct SwarmResult {                           // global first minimum of a generation, published by CTA 0 to its GPU's CTAs
    int idx, lost;
    unsigned gen, pad;
};
struct SwarmRunArgs {
    PsoDev prm;
    uint64_t seed;
    double W1, C1;
    int num_p, lo, hi, maxiter;
    double *table[2];                          // local row buffers [num_p][27]
    double *x, *v;                             // owned positions / velocities
    const double *x0;
    int *R, *inf;
    unsigned long long *key;
    unsigned long long *key2;                  // [2][own][2] informant keys of hpe_swarm_run: (index, cost bits) per owned particle
    long long spin_limit;                      // bound of the flag waits (a lost peer must not hang the GPU)
    volatile SwarmResult *result;              // [2] by generation parity, local
    SwarmCtl *ctl;
    double *trace;
    double *peer_table[2][HPE_MAX_PEERS];      // the peers' row buffers (NULL for self / absent)
    volatile unsigned *peer_flags[HPE_MAX_PEERS];   // the peers' flag arrays: [rank] is written by `rank`
    volatile unsigned *my_flags;               // this rank's flag array (written by the peers)
    SwarmMin *peer_min[HPE_MAX_PEERS];         // the peers' minimum boards [2][HPE_MAX_PEERS], entry [parity][rank] is ours
    volatile SwarmMin *my_min;                 // this rank's board (written by the peers)
    unsigned long long *cta_key;               // per-CTA minima of the step phase
    int *cta_idx;
    int world, rank;
    unsigned gen_base;                         // flags are monotonic across runs
    unsigned long long *times;                 // optional debug timestamps, 8 per generation (NULL: off)
};

int hpe_set_error(int code, const char *fmt, ...);
int hpe_check_device_of(int handle_device, const char *what);

// Host-side caches (SM count, kernel attributes, streams, events, scratch buffers) are kept PER DEVICE: a process may
// drive several GPUs (cudaSetDevice between calls); a handle is only valid on the device it was created on, which the
// entry points check (hpe_check_device_of).  Calls on ONE device are not thread-safe (the scratch buffers are shared
// by the launches of that device), calls on different devices are independent.
#define HPE_MAX_DEVICES 64
inline int hpe_device_index()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 ? 0 : (dev >= HPE_MAX_DEVICES ? HPE_MAX_DEVICES - 1 : dev);
}
template <class T> struct PerDevice {
    T v[HPE_MAX_DEVICES];                 // zero-initialised (static storage)
    T &get() { return v[hpe_device_index()]; }
};
inline int hpe_sm_count()
{
    static PerDevice<int> sms;
    int &n = sms.get();
    if (!n) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, hpe_device_index());
    return n;
}
extern int g_hpe_chamfer;   // hpe_set_chamfer(): 0 = wavefront two-pass, 1 = row-search closed form
extern int g_hpe_matcher;   // hpe_set_matcher(): 0 = certified two-sweep search, 1 = direct form
#define HPE_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return hpe_set_error(HPE_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                   \
                                 cudaGetErrorString(e_), __FILE__, __LINE__);                    \
    } while (0)

// launch wrappers (defined in the kernel .cu files); all asynchronous on `st`
cudaError_t hpe_launch_fk(const DevModel &m, const double *theta_dev, int64_t n, double *spheres_dev,
                          double *joints_dev, cudaStream_t st);
cudaError_t hpe_launch_cost(const DevModel &m, const ObsView &o, int slot, const double *theta_dev,
                            int64_t n, bool with_colli, double *cost_dev, double *terms_dev,
                            const int32_t *frozen_match_dev, int32_t *match_out_dev, cudaStream_t st,
                            int64_t terms_stride = 0 /* distance between the three term arrays; 0 = n */);
#define HPE_MAX_CHUNKS 64
struct ChaseChunks {            // upload schedule of the chase kernel: chunk c holds poses [end[c-1], end[c])
    int64_t end[HPE_MAX_CHUNKS];
    int count;
};
// the same for poses that are still being uploaded chunk by chunk (arrived_dev[c] != 0 once chunk c has landed)
cudaError_t hpe_launch_cost_chase(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int64_t n,
                                  bool with_colli, double *terms_dev, int64_t terms_stride, const int32_t *frozen_dev,
                                  int32_t *match_out_dev, const unsigned *arrived_dev, const ChaseChunks &chunks, int *timeout_dev,
                                  cudaStream_t st);
cudaError_t hpe_launch_eval_latency(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int reps,
                                    long long *cycles_dev, double *cost_dev, cudaStream_t st);
cudaError_t hpe_launch_gt_err(const double *joints_dev, const double *gt_dev, int64_t n, double *err_dev, cudaStream_t st);
cudaError_t hpe_launch_observe(const ObsView &o, int first, int count, void *scratch, cudaStream_t st);
cudaError_t hpe_launch_dtmax(const ObsView &o, int slot, cudaStream_t st);
// cluster order of the clouds of slots [first, first+count) (run by hpe_launch_observe; separately for injected clouds)
cudaError_t hpe_launch_cluster(const ObsView &o, int first, int count, cudaStream_t st);
cudaError_t hpe_launch_fold_depth(const ObsView &o, int slot, cudaStream_t st);     // injected maps: dt_px becomes the folded map
cudaError_t hpe_launch_pso(const DevModel &m, const ObsView &o, const PsoDev &p, const double *x0_dev,
                           int first, int count, const uint64_t *seeds_dev, double *bestp_dev,
                           double *best_cost_dev, double *trace_dev, int chain, cudaStream_t st, int seq_stride = 0,
                           int frame_stride = 1, int scratch_id = 0);
// scratch_id (0..2): which of the device's three swarm-state scratch buffers the launch uses -- launches that may run
// concurrently (hpe_fit_frames' compute streams) must not share one
// chain: 0 = `count` independent swarms on slots first + b; L > 0 = count / L sequences of L frames (count a multiple of
// L): frame k of sequence q is swarm q * L + k on slot first + q * seq_stride + k * frame_stride (seq_stride 0 = L), starts
// from the best pose of frame k - 1, and x0_dev holds one pose per SEQUENCE
// the same for one swarm per 8-CTA cluster (latency path: sequences, small batches)
cudaError_t hpe_launch_pso_cluster(const DevModel &m, const ObsView &o, const PsoDev &p, const double *x0_dev,
                                   int first, int count, const uint64_t *seeds_dev, double *bestp_dev,
                                   double *best_cost_dev, double *trace_dev, int chain, cudaStream_t st, int seq_stride = 0,
                                   int frame_stride = 1);
// PSO::refine_init_pose for frames [first, first+count): x_dev count*26 in place, evals_dev count or NULL
cudaError_t hpe_launch_refine(const DevModel &m, const ObsView &o, int first, int count, double *x_dev, int *evals_dev,
                              cudaStream_t st);
unsigned long long hpe_fallbacks_refine(int reset);
// PSO::pso_optimise for `count` independent swarms (hpe_pso_opt.cu); evals / final_* may be NULL
cudaError_t hpe_launch_pso_optimise(const DevModel &m, const ObsView &o, const PsoDev &p, double w, double c1, double c2,
                                    const double *x0_dev, int first, int count, const uint64_t *seeds_dev,
                                    double *bestp_dev, double *best_cost_dev, long long *evals_dev, double *final_x_dev,
                                    double *final_pcost_dev, cudaStream_t st);
unsigned long long hpe_fallbacks_pso_opt(int reset);
// costfunc::bincomp_penalty of n poses against frame `slot` (hpe_bincomp.cu); silhouette_dev: n*2400 words or NULL
cudaError_t hpe_launch_bincomp(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int64_t n,
                               double *out_dev, unsigned *silhouette_dev, cudaStream_t st);
// costfunc::depthMatch_penalty of n poses against frame `slot` (hpe_depthmatch.cu); out_dev: n*240*320 doubles
cudaError_t hpe_launch_depthmatch(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int64_t n,
                                  double *out_dev, cudaStream_t st);
cudaError_t hpe_launch_render(const DevModel &m, const double *theta_dev, int64_t n, double focal,
                              float *depth_dev, cudaStream_t st);
size_t hpe_observe_scratch_bytes(int count);
cudaError_t hpe_launch_swarm_step(const DevModel &m, const ObsView &o, int slot, const SwarmStepArgs &a, cudaStream_t st);
cudaError_t hpe_launch_swarm_run(const DevModel &m, const ObsView &o, int slot, const SwarmRunArgs &a, cudaStream_t st);
unsigned long long hpe_fallbacks_cost(int reset);
unsigned long long hpe_fallbacks_pso(int reset);
unsigned long long hpe_fallbacks_pso_cluster(int reset);
extern int g_hpe_cluster_dsm; // 1 = the cluster kernel keeps the swarm state in distributed shared memory (num_p <= 128)
extern int g_hpe_pso_path;   // hpe_set_pso_path(): 0 = automatic, 1 = one CTA per swarm, 2 = one cluster per swarm
