// hpe_api.cu -- the C ABI of include/hpe.h: handles, host<->device staging, launches.
// No compute happens on the host and nothing here falls back to a CPU path.
#include <ctime>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "hpe_internal.h"

static thread_local char g_err[512] = "ok";

int hpe_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

namespace {

// grow-only device workspace per host thread (staging for the host-pointer entry points)
struct Workspace {
    void *p = nullptr;
    size_t bytes = 0;
    int device = -1;
    cudaError_t reserve(size_t want)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev == device && want <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0; device = dev;
        e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
};
thread_local Workspace g_ws;

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

}  // namespace

// a handle is only valid on the device it was created on (hpe.h): refuse instead of launching with foreign pointers
int hpe_check_device_of(int handle_device, const char *what)
{
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) {      // no device at all: the compute call that follows reports it
        cudaGetLastError();
        return HPE_OK;
    }
    if (cur != handle_device)
        return hpe_set_error(HPE_ERR_STATE, "%s handle belongs to device %d but the current device is %d", what, handle_device, cur);
    return HPE_OK;
}

namespace {

int check_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return hpe_set_error(HPE_ERR_CUDA, "no CUDA device: libhpe_b200 has no CPU fallback (%s)",
                             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return HPE_OK;
}

int check_slot(const hpe_obs *o, int first, int count)
{
    if (!o) return hpe_set_error(HPE_ERR_ARG, "observation handle is NULL");
    if (int rc = hpe_check_device_of(o->device, "observation")) return rc;
    if (first < 0 || count < 0 || first + count > o->v.capacity)
        return hpe_set_error(HPE_ERR_ARG, "slots [%d,%d) outside capacity %d", first, first + count, o->v.capacity);
    return HPE_OK;
}

// slots must hold a usable observation before any cost / PSO call
int check_filled(const hpe_obs *o, int first, int count)
{
    std::vector<FrameMeta> meta(count);
    HPE_CUDA(cudaMemcpy(meta.data(), o->v.meta + first, sizeof(FrameMeta) * count, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; ++i)
        if (!meta[i].ok)
            return hpe_set_error(HPE_ERR_STATE, "observation slot %d is empty or has no hand pixel", first + i);
    return HPE_OK;
}

PsoDev make_pso(const hpe_pso_params *p, int num_p)
{
    PsoDev d;
    memcpy(d.lb, p->lb, sizeof d.lb);
    memcpy(d.ub, p->ub, sizeof d.ub);
    memcpy(d.std, p->std, sizeof d.std);
    d.maxiter = p->maxiter;
    d.num_p = num_p;
    return d;
}

}  // namespace

// process-wide debug / cross-check switches (hpe_set_pso_path, hpe_set_cost_pipeline): C++ linkage like their declarations
int g_hpe_pso_path = 0;
int g_hpe_cost_pipeline = 0;   // 0 = one launch per uploaded chunk (default), 1 = chase kernel for large host batches
int g_hpe_cost_chunk = 0;      // poses per upload chunk of the chase path (0 = default)

extern "C" {

const char *hpe_last_error_string(void) { return g_err; }
int hpe_version(void) { return 100; }

int hpe_set_matcher(int mode)
{
    if (mode != 0 && mode != 1) return hpe_set_error(HPE_ERR_ARG, "matcher mode must be 0 or 1");
    g_hpe_matcher = mode;
    return HPE_OK;
}

int hpe_set_cost_pipeline(int mode, int chunk_poses)
{
    if (mode < 0 || mode > 1 || chunk_poses < 0) return hpe_set_error(HPE_ERR_ARG, "hpe_set_cost_pipeline: bad argument");
    g_hpe_cost_pipeline = mode;
    g_hpe_cost_chunk = chunk_poses;
    return HPE_OK;
}

int hpe_set_pso_path(int mode)
{
    if (mode < 0 || mode > 3) return hpe_set_error(HPE_ERR_ARG, "pso path must be 0, 1, 2 or 3");
    g_hpe_pso_path = mode == 3 ? 2 : mode;
    g_hpe_cluster_dsm = mode == 3 ? 0 : 1;
    return HPE_OK;
}

int hpe_set_chamfer(int mode)
{
    if (mode < 0 || mode > 2) return hpe_set_error(HPE_ERR_ARG, "chamfer mode must be 0, 1 or 2");
    g_hpe_chamfer = mode;
    return HPE_OK;
}

int hpe_debug_fallback_points(unsigned long long *count, int reset)
{
    if (!count) return hpe_set_error(HPE_ERR_ARG, "hpe_debug_fallback_points: NULL output");
    *count = hpe_fallbacks_cost(reset) + hpe_fallbacks_pso(reset) + hpe_fallbacks_pso_cluster(reset) +
             hpe_fallbacks_refine(reset) + hpe_fallbacks_pso_opt(reset);
    return HPE_OK;
}

// page-locked host memory for the streaming entry points: cudaMemcpyAsync from / to pageable memory is staged by the
// driver and blocks the caller, so the uploads of hpe_cost_batch_submit only overlap with kernels when the caller's
// buffers are pinned
int hpe_host_alloc(size_t bytes, void **out)
{
    if (!out) return hpe_set_error(HPE_ERR_ARG, "hpe_host_alloc: NULL output");
    *out = nullptr;
    if (int rc = check_device()) return rc;
    HPE_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return HPE_OK;
}

int hpe_host_free(void *p)
{
    if (p) HPE_CUDA(cudaFreeHost(p));
    return HPE_OK;
}

int hpe_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ------------------------------------------------------------------------------------------
// handmodel
// ------------------------------------------------------------------------------------------
int hpe_model_create(const double geo_cm[20], const double spacing_cm[5], const double cmc_deg[5],
                     const int tb_spheres[4], const int fg_spheres[4], const double rad_cm[HPE_NSPH],
                     hpe_model **out)
{
    if (!geo_cm || !spacing_cm || !cmc_deg || !tb_spheres || !fg_spheres || !rad_cm || !out)
        return hpe_set_error(HPE_ERR_ARG, "hpe_model_create: NULL argument");
    static const int tb_ref[4] = {2, 2, 2, 2}, fg_ref[4] = {4, 2, 2, 2};
    if (memcmp(tb_spheres, tb_ref, sizeof tb_ref) || memcmp(fg_spheres, fg_ref, sizeof fg_ref))
        return hpe_set_error(HPE_ERR_ARG, "sphere layout must be thumb {2,2,2,2}, finger {4,2,2,2} "
                                          "(the reference's fixed 8x3/10x3 matrices, src/handmodel.cpp:272-274)");
    hpe_model *m = new (std::nothrow) hpe_model;
    if (!m) return hpe_set_error(HPE_ERR_ARG, "out of host memory");
    memset(m, 0, sizeof *m);
    cudaGetDevice(&m->device);
    DevModel &d = m->dm;
    const double PI = std::acos(-1.0);
    memcpy(d.geo, geo_cm, sizeof d.geo);
    for (int i = 0; i < HPE_NSPH; ++i) { d.rad[i] = rad_cm[i]; d.radf[i] = (float)rad_cm[i]; }
    // the matrices the digits latch behind setCMCTrans: src/fingermodel.cpp:106-132,
    // src/thumbmodel.cpp:112-138 (spacing is a float member, so spacing*spacing is a float product)
    for (int k = 0; k < 5; ++k) {
        double nCMC = cmc_deg[k] / 180.0 * PI;
        double L = geo_cm[4 * k];
        float sp = (float)spacing_cm[k];
        double a = std::sqrt(L * L + sp * sp - 2 * L * sp * std::cos(nCMC));
        double beta = std::asin(std::sin(nCMC) * sp / a);
        double *b = d.base[k];
        b[0] = std::cos(nCMC); b[1] = -std::sin(nCMC); b[2] = 0; b[3] = L * std::cos(nCMC);
        b[4] = std::sin(nCMC); b[5] = std::cos(nCMC);  b[6] = 0; b[7] = L * std::sin(nCMC);
        b[8] = 0; b[9] = 0; b[10] = 1; b[11] = 0;
        if (k == 0) {
            d.backt[k][0] = -a * std::cos(beta);
            d.backt[k][1] = -a * std::sin(beta);
        } else {
            d.backt[k][0] = -L * std::sin(nCMC) * std::cos(beta);
            d.backt[k][1] = -L * std::sin(nCMC) * std::sin(beta);
        }
    }
    double pCMC = cmc_deg[0] / 180.0 * PI + PI;     // src/thumbmodel.cpp:149
    d.tc = std::cos(pCMC);
    d.ts = std::sin(pCMC);
    *out = m;
    return HPE_OK;
}

int hpe_model_destroy(hpe_model *m)
{
    delete m;
    return HPE_OK;
}

int hpe_fk_batch(const hpe_model *m, const double *theta, int64_t n, double *spheres, double *joints)
{
    if (!m || (!theta && n > 0) || n < 0) return hpe_set_error(HPE_ERR_ARG, "hpe_fk_batch: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (n == 0) return HPE_OK;
    if (int rc = check_device()) return rc;
    size_t b_th = align_up(sizeof(double) * n * HPE_NDOF), b_sp = align_up(sizeof(double) * n * HPE_NSPH * 3),
           b_jt = align_up(sizeof(double) * n * HPE_NJOINT * 3);
    HPE_CUDA(g_ws.reserve(b_th + b_sp + b_jt));
    char *base = (char *)g_ws.p;
    double *d_th = (double *)base, *d_sp = (double *)(base + b_th), *d_jt = (double *)(base + b_th + b_sp);
    HPE_CUDA(cudaMemcpyAsync(d_th, theta, sizeof(double) * n * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(hpe_launch_fk(m->dm, d_th, n, spheres ? d_sp : nullptr, joints ? d_jt : nullptr, 0));
    if (spheres) HPE_CUDA(cudaMemcpyAsync(spheres, d_sp, sizeof(double) * n * HPE_NSPH * 3, cudaMemcpyDeviceToHost, 0));
    if (joints) HPE_CUDA(cudaMemcpyAsync(joints, d_jt, sizeof(double) * n * HPE_NJOINT * 3, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

// diagnostics: cycles of one evaluation by one warp on an idle SM: FK | depth term + staging | matching + alignment
int hpe_debug_eval_latency(const hpe_model *m, const hpe_obs *o, int slot, const double theta[HPE_NDOF], long long cycles[3])
{
    if (!m || !theta || !cycles) return hpe_set_error(HPE_ERR_ARG, "hpe_debug_eval_latency: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, slot, 1)) return rc;
    if (int rc = check_filled(o, slot, 1)) return rc;
    HPE_CUDA(g_ws.reserve(1024));
    double *d_th = (double *)g_ws.p;
    long long *d_cy = (long long *)((char *)g_ws.p + 512);
    double *d_c = (double *)((char *)g_ws.p + 768);
    HPE_CUDA(cudaMemcpy(d_th, theta, sizeof(double) * HPE_NDOF, cudaMemcpyHostToDevice));
    HPE_CUDA(hpe_launch_eval_latency(m->dm, o->v, slot, d_th, 50, d_cy, d_c, 0));
    HPE_CUDA(cudaMemcpy(cycles, d_cy, sizeof(long long) * 3, cudaMemcpyDeviceToHost));
    return HPE_OK;
}

// costfunc::gnd_truth_err for n (pose, joint.txt row) pairs
int hpe_gnd_truth_err(const hpe_model *m, const double *theta, const double *gt_mm, int64_t n, double *err)
{
    if (!m || n < 0 || ((!theta || !gt_mm || !err) && n > 0)) return hpe_set_error(HPE_ERR_ARG, "hpe_gnd_truth_err: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (n == 0) return HPE_OK;
    if (int rc = check_device()) return rc;
    size_t b_th = align_up(sizeof(double) * n * HPE_NDOF), b_jt = align_up(sizeof(double) * n * HPE_NJOINT * 3),
           b_gt = align_up(sizeof(double) * n * 63), b_er = align_up(sizeof(double) * n);
    HPE_CUDA(g_ws.reserve(b_th + b_jt + b_gt + b_er));
    char *base = (char *)g_ws.p;
    double *d_th = (double *)base, *d_jt = (double *)(base + b_th), *d_gt = (double *)(base + b_th + b_jt),
           *d_er = (double *)(base + b_th + b_jt + b_gt);
    HPE_CUDA(cudaMemcpyAsync(d_th, theta, sizeof(double) * n * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(cudaMemcpyAsync(d_gt, gt_mm, sizeof(double) * n * 63, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(hpe_launch_fk(m->dm, d_th, n, nullptr, d_jt, 0));
    HPE_CUDA(hpe_launch_gt_err(d_jt, d_gt, n, d_er, 0));
    HPE_CUDA(cudaMemcpyAsync(err, d_er, sizeof(double) * n, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

// costfunc::bincomp_penalty for n poses against one frame (hpe_bincomp.cu)
int hpe_depthmatch_batch(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n, double *image)
{
    if (!m || n < 0 || ((!theta || !image) && n > 0)) return hpe_set_error(HPE_ERR_ARG, "hpe_depthmatch_batch: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, slot, 1)) return rc;
    if (n == 0) return HPE_OK;
    if (int rc = check_filled(o, slot, 1)) return rc;
    const size_t npix = (size_t)HPE_W * HPE_H;
    size_t b_th = align_up(sizeof(double) * n * HPE_NDOF), b_out = align_up(sizeof(double) * n * npix);
    HPE_CUDA(g_ws.reserve(b_th + b_out));
    char *base = (char *)g_ws.p;
    double *d_th = (double *)base, *d_out = (double *)(base + b_th);
    HPE_CUDA(cudaMemcpyAsync(d_th, theta, sizeof(double) * n * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(hpe_launch_depthmatch(m->dm, o->v, slot, d_th, n, d_out, 0));
    HPE_CUDA(cudaMemcpyAsync(image, d_out, sizeof(double) * n * npix, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

int hpe_bincomp_batch(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n, double *penalty,
                      uint32_t *silhouette)
{
    if (!m || n < 0 || ((!theta || !penalty) && n > 0)) return hpe_set_error(HPE_ERR_ARG, "hpe_bincomp_batch: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, slot, 1)) return rc;
    if (n == 0) return HPE_OK;
    if (int rc = check_filled(o, slot, 1)) return rc;
    const size_t words = (size_t)HPE_W * HPE_H / 32;
    size_t b_th = align_up(sizeof(double) * n * HPE_NDOF), b_out = align_up(sizeof(double) * n),
           b_sil = silhouette ? align_up(sizeof(uint32_t) * n * words) : 0;
    HPE_CUDA(g_ws.reserve(b_th + b_out + b_sil));
    char *base = (char *)g_ws.p;
    double *d_th = (double *)base, *d_out = (double *)(base + b_th);
    unsigned *d_sil = silhouette ? (unsigned *)(base + b_th + b_out) : nullptr;
    HPE_CUDA(cudaMemcpyAsync(d_th, theta, sizeof(double) * n * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(hpe_launch_bincomp(m->dm, o->v, slot, d_th, n, d_out, d_sil, 0));
    HPE_CUDA(cudaMemcpyAsync(penalty, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, 0));
    if (silhouette) HPE_CUDA(cudaMemcpyAsync(silhouette, d_sil, sizeof(uint32_t) * n * words, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

// ------------------------------------------------------------------------------------------
// observedmodel
// ------------------------------------------------------------------------------------------
int hpe_obs_create(int capacity, int num_samples, double focal, hpe_obs **out)
{
    if (!out || capacity <= 0) return hpe_set_error(HPE_ERR_ARG, "hpe_obs_create: bad argument");
    if (num_samples <= 0 || num_samples > HPE_MAXPTS)
        return hpe_set_error(HPE_ERR_ARG, "num_samples must be in [1,%d]", HPE_MAXPTS);
    if (int rc = check_device()) return rc;
    hpe_obs *o = new (std::nothrow) hpe_obs;
    if (!o) return hpe_set_error(HPE_ERR_ARG, "out of host memory");
    memset(o, 0, sizeof *o);
    cudaGetDevice(&o->device);
    ObsView &v = o->v;
    v.capacity = capacity;
    v.n_pts = num_samples;
    v.focal = focal;
    v.cx = HPE_W / 2.;     // img_center << imH/2. << imW/2., src/observedmodel.cpp:56
    v.cy = HPE_H / 2.;
    size_t npix = (size_t)HPE_W * HPE_H;
    cudaError_t e = cudaMalloc(&v.depth_mm, sizeof(float) * npix * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&v.dt_px, sizeof(float) * npix * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&v.cloud, sizeof(double) * 3 * num_samples * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&v.meta, sizeof(FrameMeta) * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&v.wf_halo, sizeof(unsigned) * 2 * HPE_W * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&v.mask, sizeof(unsigned) * (npix / 32) * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&v.perm, (size_t)HPE_MAXPTS * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&v.image, (size_t)HPE_IMAGE_BYTES * capacity);
    if (e == cudaSuccess) e = cudaMemset(v.image, 0, (size_t)HPE_IMAGE_BYTES * capacity);     // slots never observed: zeros, not garbage
    if (e == cudaSuccess) e = cudaMemset(v.meta, 0, sizeof(FrameMeta) * capacity);
    if (e != cudaSuccess) {
        hpe_obs_destroy(o);
        return hpe_set_error(HPE_ERR_CUDA, "hpe_obs_create(%d frames): %s", capacity, cudaGetErrorString(e));
    }
    *out = o;
    return HPE_OK;
}

int hpe_obs_destroy(hpe_obs *o)
{
    if (!o) return HPE_OK;
    cudaFree(o->v.depth_mm); cudaFree(o->v.dt_px); cudaFree(o->v.cloud); cudaFree(o->v.meta); cudaFree(o->v.wf_halo); cudaFree(o->v.mask); cudaFree(o->v.perm); cudaFree(o->v.image);
    delete o;
    return HPE_OK;
}

int hpe_obs_num_samples(const hpe_obs *o) { return o ? o->v.n_pts : 0; }
float *hpe_obs_depth_dev(hpe_obs *o) { return o ? o->v.depth_mm : nullptr; }

static int observe_and_check(hpe_obs *o, int first, int count)
{
    HPE_CUDA(hpe_launch_observe(o->v, first, count, nullptr, 0));
    std::vector<FrameMeta> meta(count);
    HPE_CUDA(cudaMemcpy(meta.data(), o->v.meta + first, sizeof(FrameMeta) * count, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; ++i)
        if (!meta[i].ok)
            return hpe_set_error(HPE_ERR_EMPTY_FRAME, "frame in slot %d has no hand pixel", first + i);
    return HPE_OK;
}

int hpe_obs_upload(hpe_obs *o, const float *depth_mm, int first, int count)
{
    if (int rc = check_slot(o, first, count)) return rc;
    if (!depth_mm && count > 0) return hpe_set_error(HPE_ERR_ARG, "hpe_obs_upload: NULL frames");
    if (count == 0) return HPE_OK;
    size_t npix = (size_t)HPE_W * HPE_H;
    HPE_CUDA(cudaMemcpyAsync(o->v.depth_mm + npix * first, depth_mm, sizeof(float) * npix * count,
                             cudaMemcpyHostToDevice, 0));
    return observe_and_check(o, first, count);
}

int hpe_obs_upload_dev(hpe_obs *o, const float *depth_mm_dev, int first, int count)
{
    if (int rc = check_slot(o, first, count)) return rc;
    if (count == 0) return HPE_OK;
    size_t npix = (size_t)HPE_W * HPE_H;
    if (depth_mm_dev && depth_mm_dev != o->v.depth_mm + npix * first)
        HPE_CUDA(cudaMemcpyAsync(o->v.depth_mm + npix * first, depth_mm_dev, sizeof(float) * npix * count,
                                 cudaMemcpyDeviceToDevice, 0));
    return observe_and_check(o, first, count);
}

int hpe_obs_set(hpe_obs *o, int slot, const float *depth_mm, const float *dt_px, const double *cloud, double scale)
{
    if (int rc = check_slot(o, slot, 1)) return rc;
    if (!depth_mm || !dt_px || !cloud) return hpe_set_error(HPE_ERR_ARG, "hpe_obs_set: NULL argument");
    size_t npix = (size_t)HPE_W * HPE_H;
    const int np = o->v.n_pts;
    std::vector<double> soa(3 * (size_t)np);
    for (int k = 0; k < np; ++k)
        for (int c = 0; c < 3; ++c) soa[(size_t)c * np + k] = cloud[k * 3 + c];
    int n_hand = 0;
    for (size_t k = 0; k < npix; ++k) n_hand += depth_mm[k] != 0.0f;
    FrameMeta meta;
    meta.scale = scale; meta.dt_max = 0.0; meta.n_hand = n_hand; meta.ok = 1;
    HPE_CUDA(cudaMemcpy(o->v.depth_mm + npix * slot, depth_mm, sizeof(float) * npix, cudaMemcpyHostToDevice));
    HPE_CUDA(cudaMemcpy(o->v.dt_px + npix * slot, dt_px, sizeof(float) * npix, cudaMemcpyHostToDevice));
    HPE_CUDA(cudaMemcpy(o->v.cloud + (size_t)3 * np * slot, soa.data(), sizeof(double) * 3 * np, cudaMemcpyHostToDevice));
    HPE_CUDA(cudaMemcpy(o->v.meta + slot, &meta, sizeof meta, cudaMemcpyHostToDevice));
    HPE_CUDA(hpe_launch_dtmax(o->v, slot, 0));
    HPE_CUDA(hpe_launch_fold_depth(o->v, slot, 0));      // after the maximum: hand pixels become -depth
    HPE_CUDA(hpe_launch_cluster(o->v, slot, 1, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

int hpe_obs_get(const hpe_obs *o, int slot, double *cloud, float *dt_px, double *scale, int *n_hand, float *dt_max)
{
    if (int rc = check_slot(o, slot, 1)) return rc;
    size_t npix = (size_t)HPE_W * HPE_H;
    const int np = o->v.n_pts;
    FrameMeta meta;
    HPE_CUDA(cudaMemcpy(&meta, o->v.meta + slot, sizeof meta, cudaMemcpyDeviceToHost));
    if (scale) *scale = meta.scale;
    if (n_hand) *n_hand = meta.n_hand;
    if (dt_max) *dt_max = (float)meta.dt_max;
    if (!meta.ok) return hpe_set_error(HPE_ERR_STATE, "observation slot %d is empty or has no hand pixel", slot);
    if (cloud) {
        std::vector<double> soa(3 * (size_t)np);
        HPE_CUDA(cudaMemcpy(soa.data(), o->v.cloud + (size_t)3 * np * slot, sizeof(double) * 3 * np, cudaMemcpyDeviceToHost));
        for (int k = 0; k < np; ++k)
            for (int c = 0; c < 3; ++c) cloud[k * 3 + c] = soa[(size_t)c * np + k];
    }
    if (dt_px) {
        HPE_CUDA(cudaMemcpy(dt_px, o->v.dt_px + npix * slot, sizeof(float) * npix, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < npix; ++k)
            if (dt_px[k] < 0.0f) dt_px[k] = 0.0f;      // the device map is folded: hand pixels hold minus their depth
    }
    return HPE_OK;
}

// ------------------------------------------------------------------------------------------
// costfunc
// ------------------------------------------------------------------------------------------
int hpe_cost_batch_dev(const hpe_model *m, const hpe_obs *o, int slot, const double *theta_dev, int64_t n,
                       int with_collision, double *cost_dev, double *terms_dev, void *stream)
{
    if (!m || n < 0) return hpe_set_error(HPE_ERR_ARG, "hpe_cost_batch_dev: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, slot, 1)) return rc;
    if (n == 0) return HPE_OK;
    if (!theta_dev) return hpe_set_error(HPE_ERR_ARG, "hpe_cost_batch_dev: NULL theta");
    HPE_CUDA(hpe_launch_cost(m->dm, o->v, slot, theta_dev, n, with_collision != 0, cost_dev, terms_dev, nullptr,
                             nullptr, (cudaStream_t)stream));
    return HPE_OK;
}

static int cost_host(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n,
                     const int32_t *frozen, int32_t *match_out, double *align, double *depth, double *colli)
{
    if (!m || n < 0 || (!theta && n > 0)) return hpe_set_error(HPE_ERR_ARG, "cost: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, slot, 1)) return rc;
    if (n == 0) return HPE_OK;
    if (int rc = check_filled(o, slot, 1)) return rc;
    const int np = o->v.n_pts;
    size_t b_th = align_up(sizeof(double) * n * HPE_NDOF), b_tm = align_up(sizeof(double) * n * 3),
           b_fr = align_up(sizeof(int32_t) * np), b_mo = match_out ? align_up(sizeof(int32_t) * n * np) : 0;
    HPE_CUDA(g_ws.reserve(b_th + b_tm + b_fr + b_mo + 512));      // + arrival flags of the chase kernel
    char *base = (char *)g_ws.p;
    double *d_th = (double *)base, *d_tm = (double *)(base + b_th);
    int32_t *d_fr = (int32_t *)(base + b_th + b_tm), *d_mo = (int32_t *)(base + b_th + b_tm + b_fr);
    if (frozen) HPE_CUDA(cudaMemcpyAsync(d_fr, frozen, sizeof(int32_t) * np, cudaMemcpyHostToDevice, 0));
    // Three-stage pipeline over chunks of poses: an upload stream copies the chunks back to back (the poses
    // are 208 B each; PCIe is the slowest stage of a large batch and 3.4 MB copies reach its full rate), a
    // compute stream scores chunk i as soon as it has landed, a download stream returns the three term arrays
    // -- the kernel writes them as separate arrays that go straight into the caller's buffers.
    enum { MAX_CHUNKS = HPE_MAX_CHUNKS };
    struct Pipe { cudaStream_t pipe[3]; cudaEvent_t ready, up_done[MAX_CHUNKS], k_done[MAX_CHUNKS]; };
    static PerDevice<Pipe> pipe_pd;                                    // per device (hpe_internal.h)
    Pipe &P = pipe_pd.get();
    cudaStream_t *pipe = P.pipe;                                       // upload, compute, download
    cudaEvent_t &ready = P.ready, *up_done = P.up_done, *k_done = P.k_done;
    if (!pipe[0]) {
        for (int i = 0; i < 3; ++i) HPE_CUDA(cudaStreamCreateWithFlags(&pipe[i], cudaStreamNonBlocking));
        HPE_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        for (int i = 0; i < MAX_CHUNKS; ++i) {
            HPE_CUDA(cudaEventCreateWithFlags(&up_done[i], cudaEventDisableTiming));
            HPE_CUDA(cudaEventCreateWithFlags(&k_done[i], cudaEventDisableTiming));
        }
    }
    HPE_CUDA(cudaEventRecord(ready, 0));          // prior default-stream work (frozen copy, observation kernels)
    HPE_CUDA(cudaStreamWaitEvent(pipe[0], ready, 0));
    HPE_CUDA(cudaStreamWaitEvent(pipe[1], ready, 0));
    // upload schedule: chunks of ~16,384 poses (3.4 MB copies run at the full PCIe rate; every extra copy costs
    // ~20 us of fixed overhead on this platform and other sizes were measured slower -- 4 x 18,432 + 10,240 poses:
    // -8 % -- so the default path keeps them uniform).  The chase path ends
    // with halving chunks so that little is left to score when the last one lands.  hpe_set_cost_pipeline can fix
    // the chunk size.
    ChaseChunks cks;
    cks.count = 0;
    if (n <= 24576 && g_hpe_cost_chunk == 0) {
        cks.end[cks.count++] = n;
    } else {
        const bool halving = g_hpe_cost_chunk == 0 && g_hpe_cost_pipeline == 1;
        int64_t big = g_hpe_cost_chunk > 0 ? g_hpe_cost_chunk : 16384;
        if ((n + big - 1) / big > MAX_CHUNKS - 8) big = (n + MAX_CHUNKS - 9) / (MAX_CHUNKS - 8);
        int64_t at = 0;
        while (at < n && cks.count < MAX_CHUNKS - 1) {
            int64_t left = n - at, take = big;
            if (halving && left <= big) take = left > 4096 ? (left / 2 + 1023) / 1024 * 1024 : left;
            if (take > left) take = left;
            at += take;
            cks.end[cks.count++] = at;
        }
        if (at < n) cks.end[cks.count - 1] = n;
    }
    if (n >= 32768 && g_hpe_cost_pipeline == 1) {
        timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
        const double t_entry = ts0.tv_sec * 1e6 + ts0.tv_nsec * 1e-3;
        // Large batch: ONE kernel launched before the upload chases the chunks as they land (cost_chase_kernel):
        // no per-chunk launch, no lock-step of small launches; the flags are raised by 4-byte DMA copies queued
        // behind each chunk on the upload stream.
        static PerDevice<unsigned *> ones_pd;         // pinned source of the flag copies
        unsigned *&ones = ones_pd.get();
        if (!ones) {
            HPE_CUDA(cudaHostAlloc((void **)&ones, sizeof(unsigned) * MAX_CHUNKS, cudaHostAllocDefault));
            for (int i = 0; i < MAX_CHUNKS; ++i) ones[i] = 1u;
        }
        const int nck = cks.count;
        unsigned *d_flags = (unsigned *)(base + b_th + b_tm + b_fr + b_mo);
        int *d_timeout = (int *)(d_flags + MAX_CHUNKS);
        HPE_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(unsigned) * (MAX_CHUNKS + 1), pipe[1]));
        HPE_CUDA(cudaEventRecord(k_done[0], pipe[1]));
        HPE_CUDA(cudaStreamWaitEvent(pipe[0], k_done[0], 0));           // flags are cleared before any is raised
        HPE_CUDA(hpe_launch_cost_chase(m->dm, o->v, slot, d_th, n, colli != nullptr, d_tm, n, frozen ? d_fr : nullptr,
                                       match_out ? d_mo : nullptr, d_flags, cks, d_timeout, pipe[1]));
        for (int c = 0; c < nck; ++c) {
            const int64_t c0 = c ? cks.end[c - 1] : 0, cn = cks.end[c] - c0;
            HPE_CUDA(cudaMemcpyAsync(d_th + c0 * HPE_NDOF, theta + c0 * HPE_NDOF, sizeof(double) * cn * HPE_NDOF,
                                     cudaMemcpyHostToDevice, pipe[0]));
            HPE_CUDA(cudaMemcpyAsync(d_flags + c, ones + c, sizeof(unsigned), cudaMemcpyHostToDevice, pipe[0]));
        }
        cudaStream_t st = pipe[1];
        static PerDevice<int *> timed_out_pd;
        int *&timed_out_pin = timed_out_pd.get();
        if (!timed_out_pin) HPE_CUDA(cudaHostAlloc((void **)&timed_out_pin, sizeof(int), cudaHostAllocDefault));
        if (align) HPE_CUDA(cudaMemcpyAsync(align, d_tm, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        if (depth) HPE_CUDA(cudaMemcpyAsync(depth, d_tm + n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        if (colli) HPE_CUDA(cudaMemcpyAsync(colli, d_tm + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        if (match_out) HPE_CUDA(cudaMemcpyAsync(match_out, d_mo, sizeof(int32_t) * n * np, cudaMemcpyDeviceToHost, st));
        HPE_CUDA(cudaMemcpyAsync(timed_out_pin, d_timeout, sizeof(int), cudaMemcpyDeviceToHost, st));
        const bool dbg = getenv("HPE_DEBUG_PIPE") != nullptr;
        auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
        const double t_q = dbg ? now() : 0;
        HPE_CUDA(cudaStreamSynchronize(pipe[0]));
        const double t_u = dbg ? now() : 0;
        HPE_CUDA(cudaStreamSynchronize(pipe[1]));
        if (dbg) fprintf(stderr, "chase: queued at +%.0f us, uploads done +%.0f us, all done +%.0f us\n", t_q - t_entry, t_u - t_entry, now() - t_entry);
        if (*timed_out_pin) return hpe_set_error(HPE_ERR_CUDA, "hpe_cost_batch: the pose upload did not arrive (kernel gave up waiting)");
        return HPE_OK;
    }
    for (int k = 0; k < cks.count; ++k) {
        const int64_t c0 = k ? cks.end[k - 1] : 0, cn = cks.end[k] - c0;
        HPE_CUDA(cudaMemcpyAsync(d_th + c0 * HPE_NDOF, theta + c0 * HPE_NDOF, sizeof(double) * cn * HPE_NDOF,
                                 cudaMemcpyHostToDevice, pipe[0]));
        HPE_CUDA(cudaEventRecord(up_done[k], pipe[0]));
    }
    for (int k = 0; k < cks.count; ++k) {
        const int64_t c0 = k ? cks.end[k - 1] : 0, cn = cks.end[k] - c0;
        HPE_CUDA(cudaStreamWaitEvent(pipe[1], up_done[k], 0));
        HPE_CUDA(hpe_launch_cost(m->dm, o->v, slot, d_th + c0 * HPE_NDOF, cn, colli != nullptr, nullptr, d_tm + c0,
                                 frozen ? d_fr : nullptr, match_out ? d_mo + c0 * np : nullptr, pipe[1], n));
        HPE_CUDA(cudaEventRecord(k_done[k], pipe[1]));
        cudaStream_t st = pipe[2];
        HPE_CUDA(cudaStreamWaitEvent(st, k_done[k], 0));
        if (align) HPE_CUDA(cudaMemcpyAsync(align + c0, d_tm + c0, sizeof(double) * cn, cudaMemcpyDeviceToHost, st));
        if (depth) HPE_CUDA(cudaMemcpyAsync(depth + c0, d_tm + n + c0, sizeof(double) * cn, cudaMemcpyDeviceToHost, st));
        if (colli) HPE_CUDA(cudaMemcpyAsync(colli + c0, d_tm + 2 * n + c0, sizeof(double) * cn, cudaMemcpyDeviceToHost, st));
        if (match_out)
            HPE_CUDA(cudaMemcpyAsync(match_out + c0 * np, d_mo + c0 * np, sizeof(int32_t) * cn * np, cudaMemcpyDeviceToHost, st));
    }
    HPE_CUDA(cudaStreamSynchronize(pipe[2]));
    HPE_CUDA(cudaStreamSynchronize(pipe[1]));
    return HPE_OK;
}

int hpe_cost_batch(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n, double *align,
                   double *depth, double *colli)
{
    return cost_host(m, o, slot, theta, n, nullptr, nullptr, align, depth, colli);
}

// ------------------------------------------------------------------------------------------
// Streaming form of hpe_cost_batch: submit queues the upload / kernels / download of one batch and returns,
// wait blocks until its results are in the caller's buffers.  With two or three batches in flight the upload of
// batch k+1 runs under the kernels of batch k (the copy engine never idles between calls), which the one-call
// form cannot do: it has to drain the last chunk's kernel and the download before it returns.
// ------------------------------------------------------------------------------------------
namespace {
struct CostJob {
    Workspace ws;
    cudaEvent_t up_done[HPE_MAX_CHUNKS], k_done[HPE_MAX_CHUNKS], done;
    bool events = false, busy = false;
};
// per device: the job slots, the three pipeline streams (FIFO across the device's jobs) and the "prior work" event.
// A job id encodes its device: id = device * HPE_COST_JOBS + slot.
struct JobSet {
    CostJob jobs[HPE_COST_JOBS];
    cudaStream_t pipe[3];          // upload, compute, download
    cudaEvent_t ready;
};
JobSet *g_jobsets[HPE_MAX_DEVICES];
std::mutex g_jobs_mutex;
JobSet *jobset(int dev)
{
    std::lock_guard<std::mutex> lk(g_jobs_mutex);
    if (!g_jobsets[dev]) g_jobsets[dev] = new (std::nothrow) JobSet();
    return g_jobsets[dev];
}
}  // namespace

static int cost_batch_submit_impl(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n, double *align,
                                  double *depth, double *colli, int *job)
{
    if (!m || !job || n < 0 || (!theta && n > 0)) return hpe_set_error(HPE_ERR_ARG, "hpe_cost_batch_submit: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, slot, 1)) return rc;
    if (int rc = check_filled(o, slot, 1)) return rc;
    const int dev = hpe_device_index();
    JobSet *S = jobset(dev);
    if (!S) return hpe_set_error(HPE_ERR_ARG, "out of host memory");
    int id = -1;
    for (int i = 0; i < HPE_COST_JOBS && id < 0; ++i)
        if (!S->jobs[i].busy) id = i;
    if (id < 0) return hpe_set_error(HPE_ERR_STATE, "hpe_cost_batch_submit: %d batches already in flight, wait for one", HPE_COST_JOBS);
    CostJob &J = S->jobs[id];
    cudaStream_t *g_job_pipe = S->pipe;
    cudaEvent_t &g_job_ready = S->ready;
    if (!g_job_pipe[0]) {
        for (int i = 0; i < 3; ++i) HPE_CUDA(cudaStreamCreateWithFlags(&g_job_pipe[i], cudaStreamNonBlocking));
        HPE_CUDA(cudaEventCreateWithFlags(&g_job_ready, cudaEventDisableTiming));
    }
    if (!J.events) {
        for (int i = 0; i < HPE_MAX_CHUNKS; ++i) {
            HPE_CUDA(cudaEventCreateWithFlags(&J.up_done[i], cudaEventDisableTiming));
            HPE_CUDA(cudaEventCreateWithFlags(&J.k_done[i], cudaEventDisableTiming));
        }
        HPE_CUDA(cudaEventCreateWithFlags(&J.done, cudaEventDisableTiming));
        J.events = true;
    }
    cudaStream_t up = g_job_pipe[0], run = g_job_pipe[1], down = g_job_pipe[2];
    size_t b_th = align_up(sizeof(double) * (n > 0 ? n : 1) * HPE_NDOF), b_tm = align_up(sizeof(double) * (n > 0 ? n : 1) * 3);
    HPE_CUDA(J.ws.reserve(b_th + b_tm));
    double *d_th = (double *)J.ws.p, *d_tm = (double *)((char *)J.ws.p + b_th);
    HPE_CUDA(cudaEventRecord(g_job_ready, 0));       // prior default-stream work (observation kernels)
    HPE_CUDA(cudaStreamWaitEvent(up, g_job_ready, 0));
    HPE_CUDA(cudaStreamWaitEvent(run, g_job_ready, 0));
    // chunks of 65,536 poses (13.6 MB): with other batches in flight nothing is gained by splitting a batch further --
    // measured end to end on 65,536-pose batches, 3 in flight: one chunk 2.53e8 evaluations/s (52.6 GB/s of uploads, the
    // resident kernel rate), 32,768: 2.37e8, 16,384: 2.14e8, 8,192: 1.49e8 (tools/stream_probe.py)
    int64_t big = g_hpe_cost_chunk > 0 ? g_hpe_cost_chunk : 65536;
    if ((n + big - 1) / big > HPE_MAX_CHUNKS) big = (n + HPE_MAX_CHUNKS - 1) / HPE_MAX_CHUNKS;
    int k = 0;
    for (int64_t c0 = 0; c0 < n; c0 += big, ++k) {
        const int64_t cn = n - c0 < big ? n - c0 : big;
        HPE_CUDA(cudaMemcpyAsync(d_th + c0 * HPE_NDOF, theta + c0 * HPE_NDOF, sizeof(double) * cn * HPE_NDOF,
                                 cudaMemcpyHostToDevice, up));
        HPE_CUDA(cudaEventRecord(J.up_done[k], up));
        HPE_CUDA(cudaStreamWaitEvent(run, J.up_done[k], 0));
        HPE_CUDA(hpe_launch_cost(m->dm, o->v, slot, d_th + c0 * HPE_NDOF, cn, colli != nullptr, nullptr, d_tm + c0, nullptr,
                                 nullptr, run, n));
        HPE_CUDA(cudaEventRecord(J.k_done[k], run));
        HPE_CUDA(cudaStreamWaitEvent(down, J.k_done[k], 0));
        if (align) HPE_CUDA(cudaMemcpyAsync(align + c0, d_tm + c0, sizeof(double) * cn, cudaMemcpyDeviceToHost, down));
        if (depth) HPE_CUDA(cudaMemcpyAsync(depth + c0, d_tm + n + c0, sizeof(double) * cn, cudaMemcpyDeviceToHost, down));
        if (colli) HPE_CUDA(cudaMemcpyAsync(colli + c0, d_tm + 2 * n + c0, sizeof(double) * cn, cudaMemcpyDeviceToHost, down));
    }
    HPE_CUDA(cudaEventRecord(J.done, down));
    J.busy = true;
    *job = dev * HPE_COST_JOBS + id;
    return HPE_OK;
}

int hpe_cost_batch_submit(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n, double *align,
                          double *depth, double *colli, int *job)
{
    const int rc = cost_batch_submit_impl(m, o, slot, theta, n, align, depth, colli, job);
    if (rc != HPE_OK) {
        // part of the batch may already be queued: let it drain so that nothing untracked keeps writing into the
        // caller's buffers (the job slot was never marked busy)
        if (JobSet *S = g_jobsets[hpe_device_index()])
            for (int i = 0; i < 3; ++i)
                if (S->pipe[i]) cudaStreamSynchronize(S->pipe[i]);
    }
    return rc;
}

int hpe_cost_batch_wait(int job)
{
    const int dev = job / HPE_COST_JOBS, id = job % HPE_COST_JOBS;
    JobSet *S = (job >= 0 && dev < HPE_MAX_DEVICES) ? g_jobsets[dev] : nullptr;
    if (!S || !S->jobs[id].busy)
        return hpe_set_error(HPE_ERR_ARG, "hpe_cost_batch_wait: job %d is not in flight", job);
    S->jobs[id].busy = false;
    HPE_CUDA(cudaEventSynchronize(S->jobs[id].done));        // events may be waited on from any current device
    return HPE_OK;
}

int hpe_cost_batch_frozen(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n,
                          const int32_t *match, double *align, double *depth, double *colli)
{
    if (!match) return hpe_set_error(HPE_ERR_ARG, "hpe_cost_batch_frozen: NULL match");
    // matchId indexes the 48 sphere centres (src/costfunc.cpp:346-377); Armadillo would raise on anything else
    if (o)
        for (int k = 0; k < o->v.n_pts; ++k)
            if (match[k] < 0 || match[k] >= HPE_NSPH)
                return hpe_set_error(HPE_ERR_ARG, "hpe_cost_batch_frozen: match[%d] = %d is not a sphere index", k, match[k]);
    return cost_host(m, o, slot, theta, n, match, nullptr, align, depth, colli);
}

int hpe_correspondences(const hpe_model *m, const hpe_obs *o, int slot, const double *theta, int64_t n, int32_t *match)
{
    if (!match) return hpe_set_error(HPE_ERR_ARG, "hpe_correspondences: NULL match");
    return cost_host(m, o, slot, theta, n, nullptr, match, nullptr, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------
// PSO
// ------------------------------------------------------------------------------------------
// chain = 0: `count` independent swarms; chain = L > 0: count / L sequences of L frames each (see hpe_launch_pso)
static int pso_host(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                    const double *x0, int n_x0, int first, int count, const uint64_t *seeds, double *bestp,
                    double *best_cost, double *trace, int chain, int seq_stride = 0, int frame_stride = 1)
{
    if (!m || !p || !x0 || !bestp || num_particles < 1 || p->maxiter < 1)
        return hpe_set_error(HPE_ERR_ARG, "pso: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, first, count)) return rc;     // the sequences' slots tile [first, first + count)
    if (count == 0) return HPE_OK;
    if (int rc = check_filled(o, first, count)) return rc;
    const int tl = p->maxiter > 1 ? p->maxiter - 1 : 1;
    size_t b_x0 = align_up(sizeof(double) * n_x0 * HPE_NDOF), b_sd = align_up(sizeof(uint64_t) * count),
           b_bp = align_up(sizeof(double) * count * HPE_NDOF), b_bc = align_up(sizeof(double) * count),
           b_tr = align_up(sizeof(double) * count * tl);
    HPE_CUDA(g_ws.reserve(b_x0 + b_sd + b_bp + b_bc + b_tr));
    char *base = (char *)g_ws.p;
    double *d_x0 = (double *)base;
    uint64_t *d_sd = (uint64_t *)(base + b_x0);
    double *d_bp = (double *)(base + b_x0 + b_sd), *d_bc = (double *)(base + b_x0 + b_sd + b_bp),
           *d_tr = (double *)(base + b_x0 + b_sd + b_bp + b_bc);
    HPE_CUDA(cudaMemcpyAsync(d_x0, x0, sizeof(double) * n_x0 * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    if (seeds) HPE_CUDA(cudaMemcpyAsync(d_sd, seeds, sizeof(uint64_t) * count, cudaMemcpyHostToDevice, 0));
    // few swarms (or a sequence): one swarm per 8-CTA cluster; many swarms: one per CTA
    const int units = chain ? count / chain : count;
    // clusters walk their units in rounds of 37 (8-CTA clusters, two CTAs per SM): measured on B200 (tools/track_multi_probe.py,
    // 64 x 30) a cluster finishes a frame in 0.73 ms, a CTA of the one-CTA kernel in 2-6 ms depending on how full the SMs
    // are, so the cluster kernel wins up to ~250 units (37: 50.6k tracked frames/s, 148: 50.6k against 28k; 296: 45k against 49k; 444 CTAs: 117k)
    const bool use_cluster = g_hpe_pso_path == 2 || (g_hpe_pso_path == 0 && units <= 222);
    cudaError_t e = use_cluster
        ? hpe_launch_pso_cluster(m->dm, o->v, make_pso(p, num_particles), d_x0, first, count, seeds ? d_sd : nullptr,
                                 d_bp, d_bc, trace ? d_tr : nullptr, chain, 0, seq_stride, frame_stride)
        : hpe_launch_pso(m->dm, o->v, make_pso(p, num_particles), d_x0, first, count, seeds ? d_sd : nullptr,
                         d_bp, d_bc, trace ? d_tr : nullptr, chain, 0, seq_stride, frame_stride);
    if (e != cudaSuccess)
        return hpe_set_error(e == cudaErrorInvalidValue ? HPE_ERR_ARG : HPE_ERR_CUDA,
                             "pso launch (num_particles=%d): %s", num_particles, cudaGetErrorString(e));
    HPE_CUDA(cudaMemcpyAsync(bestp, d_bp, sizeof(double) * count * HPE_NDOF, cudaMemcpyDeviceToHost, 0));
    if (best_cost) HPE_CUDA(cudaMemcpyAsync(best_cost, d_bc, sizeof(double) * count, cudaMemcpyDeviceToHost, 0));
    if (trace && p->maxiter > 1)
        HPE_CUDA(cudaMemcpyAsync(trace, d_tr, sizeof(double) * count * (p->maxiter - 1), cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

int hpe_pso_evolve(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                   const double *x0, int first, int count, const uint64_t *seeds, double *bestp, double *best_cost,
                   double *trace)
{
    return pso_host(m, o, p, num_particles, x0, count, first, count, seeds, bestp, best_cost, trace, 0);
}

int hpe_track_sequence(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                       const double x0[HPE_NDOF], int first, int n_frames, uint64_t seed, double *poses, double *costs)
{
    if (n_frames < 0) return hpe_set_error(HPE_ERR_ARG, "hpe_track_sequence: bad frame count");
    std::vector<uint64_t> seeds((size_t)(n_frames > 0 ? n_frames : 1), seed);
    return pso_host(m, o, p, num_particles, x0, 1, first, n_frames, seeds.data(), poses, costs, nullptr, n_frames);
}

// n_seq independent sequences tracked at once, one thread-block cluster (or CTA) per sequence, all of them in one
// launch; each sequence's result is bit-identical to hpe_track_sequence on it alone
int hpe_track_sequences(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                        const double *x0, int first, int n_seq, int n_frames, int seq_stride, int frame_stride,
                        const uint64_t *seeds, double *poses, double *costs)
{
    if (n_seq < 0 || n_frames < 0 || (int64_t)n_seq * n_frames > 0x7fffffff)
        return hpe_set_error(HPE_ERR_ARG, "hpe_track_sequences: bad sequence / frame count");
    if (n_seq == 0 || n_frames == 0) return HPE_OK;
    // the two layouts whose slots tile [first, first + n_seq * n_frames) without overlap
    const bool seq_major = seq_stride == n_frames && frame_stride == 1, frame_major = seq_stride == 1 && frame_stride == n_seq;
    if (!seq_major && !frame_major)
        return hpe_set_error(HPE_ERR_ARG, "hpe_track_sequences: slot(q, k) = first + q * seq_stride + k * frame_stride needs "
                             "(seq_stride, frame_stride) = (n_frames, 1) or (1, n_seq)");
    std::vector<uint64_t> sd((size_t)n_seq * n_frames);
    for (int q = 0; q < n_seq; ++q)
        for (int k = 0; k < n_frames; ++k) sd[(size_t)q * n_frames + k] = seeds ? seeds[q] : 1000ull;
    return pso_host(m, o, p, num_particles, x0, n_seq, first, n_seq * n_frames, sd.data(), poses, costs, nullptr, n_frames,
                    seq_stride, frame_stride);
}

// Asynchronous device-pointer forms ("every entry takes a stream"): nothing is copied, nothing is
// synchronised; inputs must be valid on `stream` when the call is made, results are valid once `stream` reaches the
// end of the queued work.  The slots must hold usable observations (not checked here: that would need a read-back).
int hpe_pso_evolve_dev(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                       const double *x0_dev, int first, int count, const uint64_t *seeds_dev, double *bestp_dev,
                       double *best_cost_dev, double *trace_dev, void *stream)
{
    if (!m || !p || !x0_dev || !bestp_dev || !best_cost_dev || num_particles < 1 || p->maxiter < 1)
        return hpe_set_error(HPE_ERR_ARG, "hpe_pso_evolve_dev: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, first, count)) return rc;
    if (count == 0) return HPE_OK;
    const bool use_cluster = g_hpe_pso_path == 2 || (g_hpe_pso_path == 0 && count <= 222);
    cudaError_t e = use_cluster
        ? hpe_launch_pso_cluster(m->dm, o->v, make_pso(p, num_particles), x0_dev, first, count, seeds_dev, bestp_dev,
                                 best_cost_dev, trace_dev, 0, (cudaStream_t)stream)
        : hpe_launch_pso(m->dm, o->v, make_pso(p, num_particles), x0_dev, first, count, seeds_dev, bestp_dev,
                         best_cost_dev, trace_dev, 0, (cudaStream_t)stream);
    if (e != cudaSuccess)
        return hpe_set_error(e == cudaErrorInvalidValue ? HPE_ERR_ARG : HPE_ERR_CUDA,
                             "pso launch (num_particles=%d): %s", num_particles, cudaGetErrorString(e));
    return HPE_OK;
}

int hpe_track_sequences_dev(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                            const double *x0_dev, int first, int n_seq, int n_frames, int seq_stride, int frame_stride,
                            const uint64_t *seeds_dev, double *poses_dev, double *costs_dev, void *stream)
{
    if (!m || !p || !x0_dev || !poses_dev || !costs_dev || num_particles < 1 || p->maxiter < 1 || n_seq < 0 || n_frames < 0 ||
        (int64_t)n_seq * n_frames > 0x7fffffff)
        return hpe_set_error(HPE_ERR_ARG, "hpe_track_sequences_dev: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (n_seq == 0 || n_frames == 0) return HPE_OK;
    const bool seq_major = seq_stride == n_frames && frame_stride == 1, frame_major = seq_stride == 1 && frame_stride == n_seq;
    if (!seq_major && !frame_major)
        return hpe_set_error(HPE_ERR_ARG, "hpe_track_sequences_dev: (seq_stride, frame_stride) must be (n_frames, 1) or (1, n_seq)");
    if (int rc = check_slot(o, first, n_seq * n_frames)) return rc;
    const bool use_cluster = g_hpe_pso_path == 2 || (g_hpe_pso_path == 0 && n_seq <= 222);
    const int count = n_seq * n_frames;
    cudaError_t e = use_cluster
        ? hpe_launch_pso_cluster(m->dm, o->v, make_pso(p, num_particles), x0_dev, first, count, seeds_dev, poses_dev, costs_dev,
                                 nullptr, n_frames, (cudaStream_t)stream, seq_stride, frame_stride)
        : hpe_launch_pso(m->dm, o->v, make_pso(p, num_particles), x0_dev, first, count, seeds_dev, poses_dev, costs_dev, nullptr,
                         n_frames, (cudaStream_t)stream, seq_stride, frame_stride);
    if (e != cudaSuccess)
        return hpe_set_error(e == cudaErrorInvalidValue ? HPE_ERR_ARG : HPE_ERR_CUDA,
                             "pso launch (num_particles=%d): %s", num_particles, cudaGetErrorString(e));
    return HPE_OK;
}

int hpe_refine_init_pose_dev(const hpe_model *m, const hpe_obs *o, double *x0_dev, int first, int count, int *evals_dev,
                             void *stream)
{
    if (!m || !x0_dev) return hpe_set_error(HPE_ERR_ARG, "hpe_refine_init_pose_dev: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, first, count)) return rc;
    if (count == 0) return HPE_OK;
    HPE_CUDA(hpe_launch_refine(m->dm, o->v, first, count, x0_dev, evals_dev, (cudaStream_t)stream));
    return HPE_OK;
}

// PSO::refine_init_pose for `count` frames (hpe_refine.cu)
int hpe_refine_init_pose(const hpe_model *m, const hpe_obs *o, double *x0, int first, int count, int *evals)
{
    if (!m || !x0) return hpe_set_error(HPE_ERR_ARG, "hpe_refine_init_pose: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, first, count)) return rc;
    if (count == 0) return HPE_OK;
    if (int rc = check_filled(o, first, count)) return rc;
    size_t b_x = align_up(sizeof(double) * count * HPE_NDOF), b_e = align_up(sizeof(int) * count);
    HPE_CUDA(g_ws.reserve(b_x + b_e));
    double *d_x = (double *)g_ws.p;
    int *d_e = (int *)((char *)g_ws.p + b_x);
    HPE_CUDA(cudaMemcpyAsync(d_x, x0, sizeof(double) * count * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(hpe_launch_refine(m->dm, o->v, first, count, d_x, d_e, 0));
    HPE_CUDA(cudaMemcpyAsync(x0, d_x, sizeof(double) * count * HPE_NDOF, cudaMemcpyDeviceToHost, 0));
    if (evals) HPE_CUDA(cudaMemcpyAsync(evals, d_e, sizeof(int) * count, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

// PSO::pso_optimise for `count` independent swarms (hpe_pso_opt.cu)
int hpe_pso_optimise(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                     const double *x0, int first, int count, const uint64_t *seeds, double *bestp, double *best_cost,
                     int64_t *evals, double *final_particles, double *final_pcost)
{
    if (!m || !p || !x0 || !bestp || num_particles < 1 || p->maxiter < 1)
        return hpe_set_error(HPE_ERR_ARG, "hpe_pso_optimise: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_slot(o, first, count)) return rc;
    if (count == 0) return HPE_OK;
    if (int rc = check_filled(o, first, count)) return rc;
    const size_t np = (size_t)num_particles;
    size_t b_x0 = align_up(sizeof(double) * count * HPE_NDOF), b_sd = align_up(sizeof(uint64_t) * count), b_bp = b_x0,
           b_bc = align_up(sizeof(double) * count), b_ev = align_up(sizeof(long long) * count),
           b_fx = final_particles ? align_up(sizeof(double) * count * np * HPE_NDOF) : 0,
           b_fc = final_pcost ? align_up(sizeof(double) * count * np) : 0;
    HPE_CUDA(g_ws.reserve(b_x0 + b_sd + b_bp + b_bc + b_ev + b_fx + b_fc));
    char *base = (char *)g_ws.p;
    double *d_x0 = (double *)base;
    uint64_t *d_sd = (uint64_t *)(base + b_x0);
    double *d_bp = (double *)(base + b_x0 + b_sd), *d_bc = (double *)(base + b_x0 + b_sd + b_bp);
    long long *d_ev = (long long *)(base + b_x0 + b_sd + b_bp + b_bc);
    double *d_fx = final_particles ? (double *)(base + b_x0 + b_sd + b_bp + b_bc + b_ev) : nullptr;
    double *d_fc = final_pcost ? (double *)(base + b_x0 + b_sd + b_bp + b_bc + b_ev + b_fx) : nullptr;
    HPE_CUDA(cudaMemcpyAsync(d_x0, x0, sizeof(double) * count * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    if (seeds) HPE_CUDA(cudaMemcpyAsync(d_sd, seeds, sizeof(uint64_t) * count, cudaMemcpyHostToDevice, 0));
    cudaError_t e = hpe_launch_pso_optimise(m->dm, o->v, make_pso(p, num_particles), p->w, p->c1, p->c2, d_x0, first, count,
                                            seeds ? d_sd : nullptr, d_bp, d_bc, d_ev, d_fx, d_fc, 0);
    if (e != cudaSuccess)
        return hpe_set_error(e == cudaErrorInvalidValue ? HPE_ERR_ARG : HPE_ERR_CUDA,
                             "pso_optimise launch (num_particles=%d): %s", num_particles, cudaGetErrorString(e));
    HPE_CUDA(cudaMemcpyAsync(bestp, d_bp, sizeof(double) * count * HPE_NDOF, cudaMemcpyDeviceToHost, 0));
    if (best_cost) HPE_CUDA(cudaMemcpyAsync(best_cost, d_bc, sizeof(double) * count, cudaMemcpyDeviceToHost, 0));
    if (evals) HPE_CUDA(cudaMemcpyAsync(evals, d_ev, sizeof(long long) * count, cudaMemcpyDeviceToHost, 0));
    if (final_particles)
        HPE_CUDA(cudaMemcpyAsync(final_particles, d_fx, sizeof(double) * count * np * HPE_NDOF, cudaMemcpyDeviceToHost, 0));
    if (final_pcost) HPE_CUDA(cudaMemcpyAsync(final_pcost, d_fc, sizeof(double) * count * np, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

// the complete per-frame loop of test_full (src/testmodel.cpp:117-138): refine_init_pose, pso_evolve, x0 = bestp, for
// n_seq sequences in lock step.  Slots are frame-interleaved (frame k of sequence q in slot first + k * n_seq + q), so that
// time step k is ONE refine launch and ONE swarm launch over n_seq contiguous slots; everything is queued on one stream
// (two launches and one copy per time step), the host synchronises once at the end.
static int track_refined(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles, const double *x0,
                         int first, int n_seq, int n_frames, const uint64_t *seeds, double *poses, double *costs,
                         double *refined)
{
    if (!m || !p || !x0 || !poses || num_particles < 1 || p->maxiter < 1 || n_frames < 0 || n_seq < 1 ||
        (int64_t)n_seq * n_frames > 0x7fffffff)
        return hpe_set_error(HPE_ERR_ARG, "hpe_track_sequence(s)_refined: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    const int total = n_seq * n_frames;
    if (int rc = check_slot(o, first, total)) return rc;
    if (n_frames == 0) return HPE_OK;
    if (int rc = check_filled(o, first, total)) return rc;
    const size_t row = sizeof(double) * HPE_NDOF, step = row * n_seq;
    size_t b_x = align_up(row * total), b_bp = b_x, b_bc = align_up(sizeof(double) * total),
           b_sd = align_up(sizeof(uint64_t) * n_seq);
    HPE_CUDA(g_ws.reserve(b_x + b_bp + b_bc + b_sd));
    char *base = (char *)g_ws.p;
    double *d_x = (double *)base, *d_bp = (double *)(base + b_x), *d_bc = (double *)(base + b_x + b_bp);   // [k][q]
    uint64_t *d_sd = (uint64_t *)(base + b_x + b_bp + b_bc);
    std::vector<uint64_t> sd((size_t)n_seq);
    for (int q = 0; q < n_seq; ++q) sd[q] = seeds ? seeds[q] : 1000ull;
    HPE_CUDA(cudaMemcpyAsync(d_sd, sd.data(), sizeof(uint64_t) * n_seq, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(cudaMemcpyAsync(d_x, x0, step, cudaMemcpyHostToDevice, 0));
    const PsoDev prm = make_pso(p, num_particles);
    const bool one_cta = g_hpe_pso_path == 1 || (g_hpe_pso_path == 0 && n_seq > 222);
    for (int k = 0; k < n_frames; ++k) {
        double *xk = d_x + (size_t)k * n_seq * HPE_NDOF, *bk = d_bp + (size_t)k * n_seq * HPE_NDOF;
        if (k > 0) HPE_CUDA(cudaMemcpyAsync(xk, bk - (size_t)n_seq * HPE_NDOF, step, cudaMemcpyDeviceToDevice, 0));
        HPE_CUDA(hpe_launch_refine(m->dm, o->v, first + k * n_seq, n_seq, xk, nullptr, 0));
        cudaError_t e = one_cta
            ? hpe_launch_pso(m->dm, o->v, prm, xk, first + k * n_seq, n_seq, d_sd, bk, d_bc + (size_t)k * n_seq, nullptr, 0, 0)
            : hpe_launch_pso_cluster(m->dm, o->v, prm, xk, first + k * n_seq, n_seq, d_sd, bk, d_bc + (size_t)k * n_seq, nullptr, 0, 0);
        if (e != cudaSuccess)
            return hpe_set_error(e == cudaErrorInvalidValue ? HPE_ERR_ARG : HPE_ERR_CUDA,
                                 "pso launch (num_particles=%d): %s", num_particles, cudaGetErrorString(e));
    }
    // results come back frame-major and are handed out sequence-major
    std::vector<double> h_bp((size_t)total * HPE_NDOF), h_x(refined ? (size_t)total * HPE_NDOF : 0), h_bc(costs ? (size_t)total : 0);
    HPE_CUDA(cudaMemcpyAsync(h_bp.data(), d_bp, row * total, cudaMemcpyDeviceToHost, 0));
    if (costs) HPE_CUDA(cudaMemcpyAsync(h_bc.data(), d_bc, sizeof(double) * total, cudaMemcpyDeviceToHost, 0));
    if (refined) HPE_CUDA(cudaMemcpyAsync(h_x.data(), d_x, row * total, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    for (int q = 0; q < n_seq; ++q)
        for (int k = 0; k < n_frames; ++k) {
            const size_t src = (size_t)k * n_seq + q, dst = (size_t)q * n_frames + k;
            memcpy(poses + dst * HPE_NDOF, h_bp.data() + src * HPE_NDOF, row);
            if (costs) costs[dst] = h_bc[src];
            if (refined) memcpy(refined + dst * HPE_NDOF, h_x.data() + src * HPE_NDOF, row);
        }
    return HPE_OK;
}

int hpe_track_sequence_refined(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                               const double x0[HPE_NDOF], int first, int n_frames, uint64_t seed, double *poses,
                               double *costs, double *refined)
{
    return track_refined(m, o, p, num_particles, x0, first, 1, n_frames, &seed, poses, costs, refined);
}

int hpe_track_sequences_refined(const hpe_model *m, const hpe_obs *o, const hpe_pso_params *p, int num_particles,
                                const double *x0, int first, int n_seq, int n_frames, const uint64_t *seeds, double *poses,
                                double *costs, double *refined)
{
    return track_refined(m, o, p, num_particles, x0, first, n_seq, n_frames, seeds, poses, costs, refined);
}

// Offline batch from HOST frames (BASELINE "offline batch"): count frames stream through the observation buffer in
// chunks -- the upload of a chunk (307 KB per frame over PCIe) runs under the observation + swarms of the chunks before
// it; initial poses go up and results come back once.  A buffer of at least 8 (12) swarm CTAs per SM is cut into FOUR (SIX)
// chunks whose swarms rotate over two (three) compute streams: only a quarter (sixth) of the buffer's upload is exposed at the
// start of a call, and the swarm kernels of consecutive chunks (such a chunk does not fill the device on its own) run
// concurrently, each with its own state scratch.  Smaller buffers keep two halves on one compute stream.
// Submit / wait form: up to two calls in flight per device.  The chunk events outlive a call, so the first uploads of call
// j + 1 run under the last fits of call j (same observation buffer, same chunking) instead of after its results are back.
namespace {
constexpr int FIT_MAXB = 6, FIT_MAXS = 3, FIT_JOBS = 2;
struct FitJob {
    bool busy = false;
    Workspace ws;                       // poses in, poses / costs out, seeds
    void *stage = nullptr;              // page-locked copy of the caller's initial poses and seeds: an asynchronous copy from
    size_t stage_bytes = 0;             // pageable memory would block the submitting thread behind the uploads already queued
    cudaEvent_t done = nullptr;
};
struct Fit {
    cudaStream_t up = nullptr, comp[FIT_MAXS] = {};
    cudaEvent_t landed[FIT_MAXB] = {}, fitted[FIT_MAXB] = {}, ready = nullptr, poses_up = nullptr, other_done[FIT_MAXS] = {};
    bool used[FIT_MAXB] = {};           // the chunk buffer has been fitted on before: an upload into it waits for fitted[b]
    const hpe_obs *last_obs = nullptr;  // what `used` / `fitted` refer to
    int last_nbuf = 0, last_csize = 0;
    int next_chunk = 0;                 // chunk counter across calls (buffer and stream rotation continue from call to call)
    FitJob jobs[FIT_JOBS];
};
PerDevice<Fit> g_fit;

int fit_drain(Fit &F)
{
    HPE_CUDA(cudaStreamSynchronize(F.up));
    for (int i = 0; i < FIT_MAXS; ++i) HPE_CUDA(cudaStreamSynchronize(F.comp[i]));
    return HPE_OK;
}
}  // namespace

static int fit_submit(const hpe_model *m, hpe_obs *o, const hpe_pso_params *p, int num_particles, const float *depth_mm,
                      const double *x0, int64_t count, const uint64_t *seeds, double *bestp, double *best_cost, int *job);

int hpe_fit_frames_submit(const hpe_model *m, hpe_obs *o, const hpe_pso_params *p, int num_particles, const float *depth_mm,
                          const double *x0, int64_t count, const uint64_t *seeds, double *bestp, double *best_cost, int *job)
{
    if (!m || !o || !p || !depth_mm || !x0 || !bestp || !job || num_particles < 1 || p->maxiter < 1 || count < 0)
        return hpe_set_error(HPE_ERR_ARG, "hpe_fit_frames_submit: bad argument");
    *job = -1;
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (int rc = check_device()) return rc;
    const int rc = fit_submit(m, o, p, num_particles, depth_mm, x0, count, seeds, bestp, best_cost, job);
    if (rc != HPE_OK && rc != HPE_ERR_STATE) {
        // a failure half way leaves copies and kernels queued that nothing tracks: let them finish before the caller
        // gets its buffers back, and forget the chunk history (the message of the failure stays the last error)
        Fit &F = g_fit.get();
        if (F.up) {
            cudaStreamSynchronize(F.up);
            for (int i = 0; i < FIT_MAXS; ++i) cudaStreamSynchronize(F.comp[i]);
        }
        for (int i = 0; i < FIT_MAXB; ++i) F.used[i] = false;
        F.last_obs = nullptr;
    }
    return rc;
}

static int fit_submit(const hpe_model *m, hpe_obs *o, const hpe_pso_params *p, int num_particles, const float *depth_mm,
                      const double *x0, int64_t count, const uint64_t *seeds, double *bestp, double *best_cost, int *job)
{
    Fit &F = g_fit.get();
    int j = 0;
    while (j < FIT_JOBS && F.jobs[j].busy) ++j;
    if (j == FIT_JOBS) return hpe_set_error(HPE_ERR_STATE, "hpe_fit_frames_submit: %d calls already in flight", FIT_JOBS);
    FitJob &J = F.jobs[j];
    if (!F.up) {
        HPE_CUDA(cudaStreamCreateWithFlags(&F.up, cudaStreamNonBlocking));
        for (int i = 0; i < FIT_MAXS; ++i) {
            HPE_CUDA(cudaStreamCreateWithFlags(&F.comp[i], cudaStreamNonBlocking));
            HPE_CUDA(cudaEventCreateWithFlags(&F.other_done[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < FIT_MAXB; ++i) {
            HPE_CUDA(cudaEventCreateWithFlags(&F.landed[i], cudaEventDisableTiming));
            HPE_CUDA(cudaEventCreateWithFlags(&F.fitted[i], cudaEventDisableTiming));
        }
        HPE_CUDA(cudaEventCreateWithFlags(&F.ready, cudaEventDisableTiming));
        HPE_CUDA(cudaEventCreateWithFlags(&F.poses_up, cudaEventDisableTiming));
        for (int i = 0; i < FIT_JOBS; ++i) HPE_CUDA(cudaEventCreateWithFlags(&F.jobs[i].done, cudaEventDisableTiming));
    }
    if (count == 0) {                                   // nothing to do: a job that is already complete
        HPE_CUDA(cudaEventRecord(J.done, F.comp[0]));
        J.busy = true;
        *job = j;
        return HPE_OK;
    }
    const int cap = o->v.capacity, wave = 2 * hpe_sm_count();           // chunks of two swarm CTAs per SM
    const int nstreams = cap >= 6 * wave ? 3 : (cap >= 4 * wave ? 2 : 1);
    const int nbuf = nstreams > 1 ? 2 * nstreams : (cap >= 2 ? 2 : 1);
    // chunks of two swarm CTAs per SM measured best with concurrent kernels (1,184 / 4,736 host frames, frames/s: 6 x 296 on
    // three streams 101k / 123k, 6 x 444: 97k / 109k, 4 x 296 on two streams 101k / 109k, two halves of 444 on one stream
    // 95k / 108-113k, two halves of 592: 81k / 97k -- tools/fit_probe.py); a larger buffer is not used in full
    const int csize = nstreams > 1 && cap / nbuf > wave ? wave : cap / nbuf;
    if (F.last_obs != o || F.last_nbuf != nbuf || F.last_csize != csize) {
        // another buffer or another chunking: the chunk events of the calls before mean nothing here
        if (int rc = fit_drain(F)) return rc;
        for (int i = 0; i < FIT_MAXB; ++i) F.used[i] = false;
        F.last_obs = o; F.last_nbuf = nbuf; F.last_csize = csize; F.next_chunk = 0;
    }
    const size_t row = sizeof(double) * HPE_NDOF, npix = (size_t)HPE_W * HPE_H;
    size_t b_x = align_up(row * count), b_bc = align_up(sizeof(double) * count), b_sd = align_up(sizeof(uint64_t) * count);
    HPE_CUDA(J.ws.reserve(2 * b_x + b_bc + b_sd));      // the job's own buffers (the previous user of this job slot has been waited for)
    char *base = (char *)J.ws.p;
    double *d_x0 = (double *)base, *d_bp = (double *)(base + b_x), *d_bc = (double *)(base + 2 * b_x);
    uint64_t *d_sd = (uint64_t *)(base + 2 * b_x + b_bc);
    HPE_CUDA(cudaEventRecord(F.ready, 0));
    HPE_CUDA(cudaStreamWaitEvent(F.up, F.ready, 0));
    for (int i = 0; i < FIT_MAXS; ++i) HPE_CUDA(cudaStreamWaitEvent(F.comp[i], F.ready, 0));
    // the initial poses travel on the upload stream, ahead of the frames (not behind the results of the call before),
    // from the job's page-locked staging copy
    const size_t need = row * count + (seeds ? sizeof(uint64_t) * count : 0);
    if (need > J.stage_bytes) {
        if (J.stage) cudaFreeHost(J.stage);
        J.stage = nullptr; J.stage_bytes = 0;
        HPE_CUDA(cudaHostAlloc(&J.stage, need, cudaHostAllocDefault));
        J.stage_bytes = need;
    }
    memcpy(J.stage, x0, row * count);
    HPE_CUDA(cudaMemcpyAsync(d_x0, J.stage, row * count, cudaMemcpyHostToDevice, F.up));
    if (seeds) {
        memcpy((char *)J.stage + row * count, seeds, sizeof(uint64_t) * count);
        HPE_CUDA(cudaMemcpyAsync(d_sd, (char *)J.stage + row * count, sizeof(uint64_t) * count, cudaMemcpyHostToDevice, F.up));
    }
    HPE_CUDA(cudaEventRecord(F.poses_up, F.up));
    for (int i = 0; i < FIT_MAXS; ++i) HPE_CUDA(cudaStreamWaitEvent(F.comp[i], F.poses_up, 0));
    const PsoDev prm = make_pso(p, num_particles);
    for (int64_t c0 = 0; c0 < count; c0 += csize) {
        const int k = F.next_chunk++;
        const int n = (int)(count - c0 < csize ? count - c0 : csize);
        const int b = k % nbuf, first = b * csize, sid = k % nstreams;
        cudaStream_t comp = F.comp[sid];
        if (F.used[b]) HPE_CUDA(cudaStreamWaitEvent(F.up, F.fitted[b], 0));      // the chunk's slots are free again
        HPE_CUDA(cudaMemcpyAsync(o->v.depth_mm + npix * first, depth_mm + npix * c0, sizeof(float) * npix * n,
                                 cudaMemcpyHostToDevice, F.up));
        HPE_CUDA(cudaEventRecord(F.landed[b], F.up));
        HPE_CUDA(cudaStreamWaitEvent(comp, F.landed[b], 0));
        HPE_CUDA(hpe_launch_observe(o->v, first, n, nullptr, comp));
        cudaError_t e = hpe_launch_pso(m->dm, o->v, prm, d_x0 + c0 * HPE_NDOF, first, n, seeds ? d_sd + c0 : nullptr,
                                       d_bp + c0 * HPE_NDOF, d_bc + c0, nullptr, 0, comp, 0, 1, sid);
        if (e != cudaSuccess) {
            return hpe_set_error(e == cudaErrorInvalidValue ? HPE_ERR_ARG : HPE_ERR_CUDA,
                                 "pso launch (num_particles=%d): %s", num_particles, cudaGetErrorString(e));
        }
        HPE_CUDA(cudaEventRecord(F.fitted[b], comp));
        F.used[b] = true;
    }
    for (int i = 1; i < FIT_MAXS; ++i) {
        HPE_CUDA(cudaEventRecord(F.other_done[i], F.comp[i]));
        HPE_CUDA(cudaStreamWaitEvent(F.comp[0], F.other_done[i], 0));
    }
    HPE_CUDA(cudaMemcpyAsync(bestp, d_bp, row * count, cudaMemcpyDeviceToHost, F.comp[0]));
    if (best_cost) HPE_CUDA(cudaMemcpyAsync(best_cost, d_bc, sizeof(double) * count, cudaMemcpyDeviceToHost, F.comp[0]));
    HPE_CUDA(cudaEventRecord(J.done, F.comp[0]));
    J.busy = true;
    *job = j;
    return HPE_OK;
}

int hpe_fit_frames_wait(int job)
{
    if (int rc = check_device()) return rc;
    Fit &F = g_fit.get();
    if (job < 0 || job >= FIT_JOBS || !F.jobs[job].busy) return hpe_set_error(HPE_ERR_ARG, "hpe_fit_frames_wait: no such job");
    cudaError_t e = cudaEventSynchronize(F.jobs[job].done);
    F.jobs[job].busy = false;
    if (e != cudaSuccess) return hpe_set_error(HPE_ERR_CUDA, "hpe_fit_frames_wait: %s", cudaGetErrorString(e));
    return HPE_OK;
}

int hpe_fit_frames(const hpe_model *m, hpe_obs *o, const hpe_pso_params *p, int num_particles, const float *depth_mm,
                   const double *x0, int64_t count, const uint64_t *seeds, double *bestp, double *best_cost)
{
    int job = -1;
    if (int rc = hpe_fit_frames_submit(m, o, p, num_particles, depth_mm, x0, count, seeds, bestp, best_cost, &job)) return rc;
    return hpe_fit_frames_wait(job);
}

// ------------------------------------------------------------------------------------------
// synthetic frames
// ------------------------------------------------------------------------------------------
int hpe_render_dev(const hpe_model *m, const double *theta_dev, int64_t n, double focal, float *depth_mm_dev,
                   void *stream)
{
    if (!m || n < 0 || ((!theta_dev || !depth_mm_dev) && n > 0))
        return hpe_set_error(HPE_ERR_ARG, "hpe_render_dev: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    HPE_CUDA(hpe_launch_render(m->dm, theta_dev, n, focal, depth_mm_dev, (cudaStream_t)stream));
    return HPE_OK;
}

int hpe_render(const hpe_model *m, const double *theta, int64_t n, double focal, float *depth_mm)
{
    if (!m || n < 0 || ((!theta || !depth_mm) && n > 0)) return hpe_set_error(HPE_ERR_ARG, "hpe_render: bad argument");
    if (int rc = hpe_check_device_of(m->device, "model")) return rc;
    if (n == 0) return HPE_OK;
    if (int rc = check_device()) return rc;
    size_t npix = (size_t)HPE_W * HPE_H;
    size_t b_th = align_up(sizeof(double) * n * HPE_NDOF);
    HPE_CUDA(g_ws.reserve(b_th + sizeof(float) * npix * n));
    double *d_th = (double *)g_ws.p;
    float *d_fr = (float *)((char *)g_ws.p + b_th);
    HPE_CUDA(cudaMemcpyAsync(d_th, theta, sizeof(double) * n * HPE_NDOF, cudaMemcpyHostToDevice, 0));
    HPE_CUDA(hpe_launch_render(m->dm, d_th, n, focal, d_fr, 0));
    HPE_CUDA(cudaMemcpyAsync(depth_mm, d_fr, sizeof(float) * npix * n, cudaMemcpyDeviceToHost, 0));
    HPE_CUDA(cudaStreamSynchronize(0));
    return HPE_OK;
}

}  // extern "C"
