// hpe_classes.hpp -- the reference's C++ host classes (handmodel, observedmodel, costfunc, PSO)
// re-expressed over the C ABI of hpe.h, header-only.  Same class names, method names and argument
// order as src/handmodel.h:9-32, src/observedmodel.h:34-68, src/costfunc.h:25-41, src/PSO.h:56-68
// of hjurong/hand-pose-estimation; Armadillo types become std::vector<double> (vec) and row-major
// std::vector<double> with explicit shapes (mat).  A reference driver such as test_full
// (src/testmodel.cpp:27-146) ports by replacing `arma::vec` with `hpe::vec`.
//
// Errors: the reference aborts or prints; these classes throw hpe::error carrying the C status
// code.  There is no CPU path: every compute call needs libhpe_b200.so and a CUDA device.
#ifndef HPE_B200_CLASSES_HPP
#define HPE_B200_CLASSES_HPP

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "hpe.h"

namespace hpe {

typedef std::vector<double> vec;
typedef std::vector<int32_t> uvec;

struct error : std::runtime_error {
    int code;
    error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc)
{
    if (rc != HPE_OK) throw error(rc, hpe_last_error_string());
}

// row-major matrix with the two Armadillo accessors the drivers use
struct mat {
    int n_rows, n_cols;
    std::vector<double> data;
    mat(int r = 0, int c = 0) : n_rows(r), n_cols(c), data((size_t)r * c, 0.0) {}
    double &operator()(int r, int c) { return data[(size_t)r * n_cols + c]; }
    double operator()(int r, int c) const { return data[(size_t)r * n_cols + c]; }
};

// ---- src/handmodel.h ---------------------------------------------------------------------
class handmodel {
public:
    // handmodel(h_geo, h_spacing, tb_spheres, fg_spheres, h_CMC, sphR), src/handmodel.cpp:10-27
    handmodel(const vec &h_geo, const vec &h_spacing, const vec &tb_spheres, const vec &fg_spheres, const vec &h_CMC,
              const vec &sphR)
        : hand_joints(HPE_NJOINT, 3), handle_(nullptr), geo_(h_geo), spacing_(h_spacing), cmc_(h_CMC), radii_(sphR)
    {
        if (h_geo.size() != 20 || h_spacing.size() != 5 || h_CMC.size() != 5 || tb_spheres.size() != 4 ||
            fg_spheres.size() != 4 || sphR.size() != HPE_NSPH)
            throw error(HPE_ERR_ARG, "handmodel: h_geo(20) h_spacing(5) tb_spheres(4) fg_spheres(4) h_CMC(5) sphR(48)");
        int tb[4], fg[4];
        for (int i = 0; i < 4; ++i) { tb[i] = (int)tb_spheres[i]; fg[i] = (int)fg_spheres[i]; }
        check(hpe_model_create(h_geo.data(), h_spacing.data(), h_CMC.data(), tb, fg, sphR.data(), &handle_));
    }
    ~handmodel() { hpe_model_destroy(handle_); }
    handmodel(const handmodel &) = delete;
    handmodel &operator=(const handmodel &) = delete;

    // build_hand_model(h_theta, sphere_centres): 48x3 centres, y,z negated; records hand_joints
    // (src/handmodel.cpp:259-298)
    void build_hand_model(const vec &h_theta, mat &sphere_centres)
    {
        if (h_theta.size() != HPE_NDOF) throw error(HPE_ERR_ARG, "build_hand_model: theta must have 26 entries");
        sphere_centres = mat(HPE_NSPH, 3);
        check(hpe_fk_batch(handle_, h_theta.data(), 1, sphere_centres.data.data(), hand_joints.data.data()));
    }
    vec get_hand_geo() const { return geo_; }
    vec get_hand_CMC() const { return cmc_; }
    vec get_spacing() const { return spacing_; }
    vec get_spheres_radii() const { return radii_; }
    vec *get_radii() { return &radii_; }
    const hpe_model *c_handle() const { return handle_; }

    mat hand_joints;   // 21x3, not negated (src/handmodel.cpp:291-296)

private:
    hpe_model *handle_;
    vec geo_, spacing_, cmc_, radii_;
};

// ---- src/observedmodel.h -----------------------------------------------------------------
class observedmodel {
public:
    // capacity: frames this object can hold on the device; the reference API works on slot 0
    explicit observedmodel(int capacity = 1, int num_samples = 250, double focal = 241.42)
        : handle_(nullptr), capacity_(capacity), num_samples_(num_samples), focal_(focal)
    {
        check(hpe_obs_create(capacity, num_samples, focal, &handle_));
    }
    ~observedmodel() { hpe_obs_destroy(handle_); }
    observedmodel(const observedmodel &) = delete;
    observedmodel &operator=(const observedmodel &) = delete;

    // init_observation(dpath, dfname, mm_to_cm, imW, imH, foclen, downsample), src/observedmodel.cpp:42-64
    void init_observation(const std::string &dpath, const std::string &dfname, bool mm_to_cm, int imW, int imH,
                          double foclen, bool downsampling)
    {
        if (!mm_to_cm || imW != 240 || imH != 320 || !downsampling || foclen != focal_)
            throw error(HPE_ERR_ARG, "init_observation: only the reference configuration (240, 320, cm, downsample) is supported");
        path_ = dpath;
        next_frame(dfname);
    }
    // next_frame(next): load_data + get_observed (src/observedmodel.cpp:420-430); the file is the
    // raw float32[240*320] Subject1 format (src/observedmodel.cpp:272-310)
    observedmodel &next_frame(const std::string &next)
    {
        std::string full = path_ + next;
        std::vector<float> buf((size_t)HPE_IMG_W * HPE_IMG_H);
        FILE *f = std::fopen(full.c_str(), "rb");
        if (!f) throw error(HPE_ERR_ARG, "error: open file for input failed! " + full);
        size_t got = std::fread(buf.data(), sizeof(float), buf.size(), f);
        std::fclose(f);
        if (got != buf.size()) throw error(HPE_ERR_ARG, full + ": not a 320x240 float32 frame");
        check(hpe_obs_upload(handle_, buf.data(), 0, 1));
        return *this;
    }
    // batched: `count` frames (host, mm) into slots first..first+count-1, preprocessing on the device
    void upload(const float *depth_mm, int first, int count) { check(hpe_obs_upload(handle_, depth_mm, first, count)); }

    mat get_ptncloud(int slot = 0) const
    {
        mat c(num_samples_, 3);
        check(hpe_obs_get(handle_, slot, c.data.data(), nullptr, nullptr, nullptr, nullptr));
        return c;
    }
    std::vector<float> get_disttran(int slot = 0) const
    {
        std::vector<float> dt((size_t)HPE_IMG_W * HPE_IMG_H);
        check(hpe_obs_get(handle_, slot, nullptr, dt.data(), nullptr, nullptr, nullptr));
        return dt;
    }
    double get_img_scale(int slot = 0) const
    {
        double s = 0;
        check(hpe_obs_get(handle_, slot, nullptr, nullptr, &s, nullptr, nullptr));
        return s;
    }
    double get_focal_len() const { return focal_; }
    int get_imgH() const { return 320; }
    int get_imgW() const { return 240; }
    int capacity() const { return capacity_; }
    const hpe_obs *c_handle() const { return handle_; }

private:
    hpe_obs *handle_;
    int capacity_, num_samples_;
    double focal_;
    std::string path_;
};

// ---- src/costfunc.h ----------------------------------------------------------------------
class costfunc {
public:
    // non-owning pointers, like the reference (src/costfunc.cpp:16-29)
    costfunc(handmodel *handM, observedmodel *observed) : hand(handM), observation(observed) {}

    double cal_cost(const vec &theta, int slot = 0)
    {
        double a = 0, d = 0;
        check(hpe_cost_batch(hand->c_handle(), observation->c_handle(), slot, theta.data(), 1, &a, &d, nullptr));
        return a + d;                       // no collision term, src/costfunc.cpp:122
    }
    double cal_cost2(const vec &theta, uvec &matchId, bool compute_corr, bool debug = false, int slot = 0)
    {
        double a = 0, d = 0, c = 0;
        if (compute_corr) {
            matchId.assign(hpe_obs_num_samples(observation->c_handle()), 0);
            check(hpe_correspondences(hand->c_handle(), observation->c_handle(), slot, theta.data(), 1, matchId.data()));
            check(hpe_cost_batch(hand->c_handle(), observation->c_handle(), slot, theta.data(), 1, &a, &d, &c));
        } else {
            check(hpe_cost_batch_frozen(hand->c_handle(), observation->c_handle(), slot, theta.data(), 1, matchId.data(),
                                        &a, &d, &c));
        }
        if (debug) std::printf("%g\n%g\n%g\n\n", a, d, c);
        return a + d + c;
    }
    // many poses per call: thetas is n*26, outputs n each (colli may be NULL)
    void cal_cost_batch(const double *thetas, long n, double *align, double *depth, double *colli, int slot = 0)
    {
        check(hpe_cost_batch(hand->c_handle(), observation->c_handle(), slot, thetas, n, align, depth, colli));
    }
    // streaming form of cal_cost_batch (hpe_cost_batch_submit / hpe_cost_batch_wait): returns a job id; up to
    // HPE_COST_JOBS batches in flight, buffers (page-locked for real overlap) must outlive cal_cost_wait(job)
    int cal_cost_submit(const double *thetas, long n, double *align, double *depth, double *colli, int slot = 0)
    {
        int job = -1;
        check(hpe_cost_batch_submit(hand->c_handle(), observation->c_handle(), slot, thetas, n, align, depth, colli, &job));
        return job;
    }
    static void cal_cost_wait(int job) { check(hpe_cost_batch_wait(job)); }
    void compute_correspondences(const vec &theta, uvec &matchId, int slot = 0)
    {
        matchId.assign(hpe_obs_num_samples(observation->c_handle()), 0);
        check(hpe_correspondences(hand->c_handle(), observation->c_handle(), slot, theta.data(), 1, matchId.data()));
    }

    // cv::Mat depthMatch_penalty(spheres), src/costfunc.cpp:439-473: the depth map (cm) with the outlines of the projected
    // spheres (cv::circle, thickness 2), 240 x 320 row-major; takes the pose like bincomp_penalty
    std::vector<double> depthMatch_penalty(const vec &theta, int slot = 0)
    {
        if (theta.size() != HPE_NDOF) throw error(HPE_ERR_ARG, "depthMatch_penalty: theta must have 26 entries");
        std::vector<double> img((size_t)HPE_IMG_W * HPE_IMG_H);
        check(hpe_depthmatch_batch(hand->c_handle(), observation->c_handle(), slot, theta.data(), 1, img.data()));
        return img;
    }
    // double bincomp_penalty(spheres), src/costfunc.cpp:384-436: silhouette mismatch in pixels; takes the pose
    // (the reference takes the spheres build_hand_model made of it)
    double bincomp_penalty(const vec &theta, int slot = 0)
    {
        double c = 0;
        check(hpe_bincomp_batch(hand->c_handle(), observation->c_handle(), slot, theta.data(), 1, &c, nullptr));
        return c;
    }
    // double gnd_truth_err(gnd_truth, frame), src/costfunc.cpp:476-507: gnd_truth = joint.txt, n_frames x 63
    // row-major (mm); theta = the pose whose joints are compared (the reference reads hand_joints left by the
    // last build_hand_model call)
    double gnd_truth_err(const double *gnd_truth, int frame, const vec &theta)
    {
        double e = 0;
        check(hpe_gnd_truth_err(hand->c_handle(), theta.data(), gnd_truth + (size_t)frame * 63, 1, &e));
        return e;
    }

    handmodel *hand;
    observedmodel *observation;
};

// ---- src/PSO.h ---------------------------------------------------------------------------
class PSO {
public:
    PSO()
    {
        std::memset(&p_, 0, sizeof p_);
        p_.w = 0.7298; p_.c1 = 1.49618; p_.c2 = 1.49618; p_.maxiter = 100; p_.minstep = 1e-6; p_.minfunc = 1e-6;
    }
    // set_pso_params(upperbound, lowerbound, std, omega, phip, phig, maxiter, minstep, minfunc), src/PSO.cpp:38-54
    void set_pso_params(const vec &upperbound, const vec &lowerbound, const vec &std_, double omega, double phip,
                        double phig, int maxiter, double minstep, double minfunc)
    {
        if (upperbound.size() != HPE_NDOF || lowerbound.size() != HPE_NDOF || std_.size() != HPE_NDOF)
            throw error(HPE_ERR_ARG, "set_pso_params: bounds and std must have 26 entries");
        std::memcpy(p_.ub, upperbound.data(), sizeof p_.ub);
        std::memcpy(p_.lb, lowerbound.data(), sizeof p_.lb);
        std::memcpy(p_.std, std_.data(), sizeof p_.std);
        p_.w = omega; p_.c1 = phip; p_.c2 = phig; p_.maxiter = maxiter; p_.minstep = minstep; p_.minfunc = minfunc;
    }
    // int pso_evolve(optfunc, x0, num_particles, bestp), src/PSO.cpp:717-886; returns 1 like the
    // reference.  seed replaces arma_rng::set_seed(1000) at :722.
    int pso_evolve(costfunc &optfunc, const vec &x0, int num_particles, vec &bestp, unsigned long long seed = 1000,
                   int slot = 0, double *best_cost = nullptr)
    {
        bestp.assign(HPE_NDOF, 0.0);
        uint64_t s = seed;
        check(hpe_pso_evolve(optfunc.hand->c_handle(), optfunc.observation->c_handle(), &p_, num_particles, x0.data(),
                             slot, 1, &s, bestp.data(), best_cost, nullptr));
        return 1;
    }
    // `count` independent swarms on frames first..first+count-1 (x0 and bestp are count*26)
    void pso_evolve_batch(costfunc &optfunc, const double *x0, int num_particles, int first, int count, double *bestp,
                          double *best_cost, const uint64_t *seeds = nullptr)
    {
        check(hpe_pso_evolve(optfunc.hand->c_handle(), optfunc.observation->c_handle(), &p_, num_particles, x0, first,
                             count, seeds, bestp, best_cost, nullptr));
    }
    // int pso_optimise(optfunc, x0, num_particles, bestp), src/PSO.cpp:539-712 (coordinate descent + canonical
    // PSO with omega, phip, phig); returns 1 like the reference
    int pso_optimise(costfunc &optfunc, const vec &x0, int num_particles, vec &bestp, unsigned long long seed = 1000,
                     int slot = 0, double *best_cost = nullptr)
    {
        bestp.assign(HPE_NDOF, 0.0);
        uint64_t s = seed;
        check(hpe_pso_optimise(optfunc.hand->c_handle(), optfunc.observation->c_handle(), &p_, num_particles, x0.data(),
                               slot, 1, &s, bestp.data(), best_cost, nullptr, nullptr, nullptr));
        return 1;
    }
    // the loop of test_full without refine_init_pose (src/testmodel.cpp:117-138)
    void track_sequence(costfunc &optfunc, const vec &x0, int num_particles, int first, int n_frames, double *poses,
                        double *costs, unsigned long long seed = 1000)
    {
        check(hpe_track_sequence(optfunc.hand->c_handle(), optfunc.observation->c_handle(), &p_, num_particles, x0.data(),
                                 first, n_frames, seed, poses, costs));
    }
    // the same loop over n_seq independent sequences at once (one launch, one thread-block cluster per sequence): frame k
    // of sequence q in slot first + q * n_frames + k, or first + k * n_seq + q when interleaved; x0: n_seq * 26, seeds:
    // n_seq values or nullptr; poses: n_seq * n_frames * 26, costs: n_seq * n_frames
    void track_sequences(costfunc &optfunc, const double *x0, int num_particles, int first, int n_seq, int n_frames,
                         double *poses, double *costs, const uint64_t *seeds = nullptr, bool interleaved = false)
    {
        check(hpe_track_sequences(optfunc.hand->c_handle(), optfunc.observation->c_handle(), &p_, num_particles, x0, first,
                                  n_seq, n_frames, interleaved ? 1 : n_frames, interleaved ? n_seq : 1, seeds, poses, costs));
    }
    // offline batch from host frames (count * 240 * 320 floats, mm): upload, observation, one swarm per frame,
    // uploads overlapped with the fits; the observedmodel's capacity sets the chunk size
    void fit_frames(costfunc &optfunc, const float *depth_mm, const double *x0, long count, int num_particles,
                    double *bestp, double *best_cost, const uint64_t *seeds = nullptr)
    {
        check(hpe_fit_frames(optfunc.hand->c_handle(), const_cast<hpe_obs *>(optfunc.observation->c_handle()), &p_,
                             num_particles, depth_mm, x0,
                             count, seeds, bestp, best_cost));
    }
    // the same in two steps (hpe_fit_frames_submit / hpe_fit_frames_wait): returns a job id; two calls may be in flight
    // per device, the uploads of the second then run under the fits of the first.  All buffers (page-locked for real
    // overlap) must outlive fit_frames_wait(job); the observedmodel must not be touched in between
    int fit_frames_submit(costfunc &optfunc, const float *depth_mm, const double *x0, long count, int num_particles,
                          double *bestp, double *best_cost, const uint64_t *seeds = nullptr)
    {
        int job = -1;
        check(hpe_fit_frames_submit(optfunc.hand->c_handle(), const_cast<hpe_obs *>(optfunc.observation->c_handle()), &p_,
                                    num_particles, depth_mm, x0, count, seeds, bestp, best_cost, &job));
        return job;
    }
    static void fit_frames_wait(int job) { check(hpe_fit_frames_wait(job)); }
    // void refine_init_pose(x0, optfunc), src/PSO.cpp:216-266: x0 is updated in place
    void refine_init_pose(vec &x0, costfunc &optfunc, int slot = 0)
    {
        if (x0.size() != HPE_NDOF) throw error(HPE_ERR_ARG, "refine_init_pose: x0 must have 26 entries");
        check(hpe_refine_init_pose(optfunc.hand->c_handle(), optfunc.observation->c_handle(), x0.data(), slot, 1, nullptr));
    }
    // the loop of test_full as written: refine_init_pose, pso_evolve, x0 = bestp (src/testmodel.cpp:117-138)
    void track_sequence_refined(costfunc &optfunc, const vec &x0, int num_particles, int first, int n_frames,
                                double *poses, double *costs, double *refined = nullptr, unsigned long long seed = 1000)
    {
        check(hpe_track_sequence_refined(optfunc.hand->c_handle(), optfunc.observation->c_handle(), &p_, num_particles,
                                         x0.data(), first, n_frames, seed, poses, costs, refined));
    }
    // ... for n_seq sequences in lock step; slots frame-interleaved (first + k * n_seq + q)
    void track_sequences_refined(costfunc &optfunc, const double *x0, int num_particles, int first, int n_seq, int n_frames,
                                 double *poses, double *costs, double *refined = nullptr, const uint64_t *seeds = nullptr)
    {
        check(hpe_track_sequences_refined(optfunc.hand->c_handle(), optfunc.observation->c_handle(), &p_, num_particles, x0,
                                          first, n_seq, n_frames, seeds, poses, costs, refined));
    }
    const hpe_pso_params &params() const { return p_; }

private:
    hpe_pso_params p_;
};

}  // namespace hpe
#endif
