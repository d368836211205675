// oracle/ref_driver.cpp -- C entry points over the REFERENCE'S OWN classes (TEST INFRASTRUCTURE).
//
// oracle/Makefile (target _ref) compiles the reference's hot-path sources unmodified from
// $(REF)/src -- fingermodel.cpp thumbmodel.cpp handmodel.cpp costfunc.cpp observedmodel.cpp PSO.cpp --
// against the stand-in Armadillo/OpenCV headers in oracle/shim/ and links them with this file into
// oracle/_ref/libhpe_ref.so.  Nothing here restates the algorithm: every function below only
// marshals plain arrays into the reference's arma types and calls the reference's public methods
// (handmodel::build_hand_model, observedmodel::init_observation/next_frame, costfunc::cal_cost/
// cal_cost2/..., PSO::set_pso_params/pso_evolve/refine_init_pose/pso_optimise).
// Users: tests/test_ref_build.py (pins oracle/hpe_oracle.c against the reference code),
// tests/golden/make_golden_ref.py (fixtures for the GPU box) and bench.py's CPU baseline.
#include "PSO.h"
#include "costfunc.h"
#include "handmodel.h"
#include "observedmodel.h"

#include <omp.h>
#include <string>

namespace {
struct ref_ctx {
    handmodel *hand;
    observedmodel *obs;
    costfunc *cost;
    PSO *pso;
    bool frame_loaded;
};

vec to_vec(const double *p, int n)
{
    vec v(n);
    for (int i = 0; i < n; ++i) v(i) = p[i];
    return v;
}
}  // namespace

extern "C" {

// handmodel::handmodel(h_geo, h_spacing, tb_spheres, fg_spheres, h_CMC, sphR), src/handmodel.cpp:10;
// same argument values as src/testmodel.cpp:42-52 when called with the defaults
void *ref_create(const double geo_cm[20], const double spacing[5], const double cmc[5], const double tb_n[4],
                 const double fg_n[4], const double rad_cm[48])
{
    ref_ctx *c = new ref_ctx;
    c->hand = new handmodel(to_vec(geo_cm, 20), to_vec(spacing, 5), to_vec(tb_n, 4), to_vec(fg_n, 4), to_vec(cmc, 5),
                            to_vec(rad_cm, 48));
    c->obs = new observedmodel();
    c->cost = new costfunc(c->hand, c->obs);
    c->pso = new PSO();
    c->frame_loaded = false;
    return c;
}

void ref_destroy(void *h)
{
    ref_ctx *c = (ref_ctx *)h;
    delete c->pso;
    delete c->cost;
    delete c->obs;
    delete c->hand;
    delete c;
}

void ref_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
int ref_max_threads(void) { return omp_get_max_threads(); }

// the hook every randu/randn/randi/set_seed of the reference goes through (oracle/shim/armadillo)
void ref_set_rng(void (*fill)(int, int, int, int, int, double *), void (*seed)(unsigned))
{
    arma::shim_rng().fill = fill ? fill : arma::shim_default_fill;
    arma::shim_rng().seed = seed ? seed : arma::shim_default_seed;
}

// first frame: observedmodel::init_observation(path, name, true, 240, 320, 241.42, downsample)
// (src/testmodel.cpp:55-68); later frames: next_frame(name) (src/testmodel.cpp:124)
int ref_load_frame(void *h, const char *dir, const char *name, int downsample)
{
    ref_ctx *c = (ref_ctx *)h;
    if (!c->frame_loaded) {
        c->obs->init_observation(std::string(dir), std::string(name), true, 240, 320, 241.42, downsample != 0);
        c->frame_loaded = true;
    } else {
        c->obs->next_frame(std::string(name));
    }
    return (int)c->obs->get_ptncloud()->n_rows;
}

double ref_obs_scale(void *h) { return ((ref_ctx *)h)->obs->get_img_scale(); }

// row-major copies of the observation: cloud n x 3, depth and distance transform 240 x 320
void ref_obs_get(void *h, double *cloud, double *depth, double *dt)
{
    ref_ctx *c = (ref_ctx *)h;
    mat &P = *c->obs->get_ptncloud();
    mat &D = *c->obs->get_depthmap();
    mat &T = *c->obs->get_disttran();
    if (cloud)
        for (uword i = 0; i < P.n_rows; ++i)
            for (uword k = 0; k < 3; ++k) cloud[i * 3 + k] = P(i, k);
    if (depth)
        for (uword r = 0; r < D.n_rows; ++r)
            for (uword k = 0; k < D.n_cols; ++k) depth[r * D.n_cols + k] = D(r, k);
    if (dt)
        for (uword r = 0; r < T.n_rows; ++r)
            for (uword k = 0; k < T.n_cols; ++k) dt[r * T.n_cols + k] = T(r, k);
}

// handmodel::build_hand_model + the public hand_joints member, row-major 48 x 3 and 21 x 3
void ref_fk(void *h, const double theta[26], double spheres[48 * 3], double joints[21 * 3])
{
    ref_ctx *c = (ref_ctx *)h;
    vec th = to_vec(theta, 26);
    mat::fixed<48, 3> S;
    c->hand->build_hand_model(th, S);
    for (int i = 0; i < 48; ++i)
        for (int k = 0; k < 3; ++k) spheres[i * 3 + k] = S(i, k);
    if (joints)
        for (int i = 0; i < 21; ++i)
            for (int k = 0; k < 3; ++k) joints[i * 3 + k] = c->hand->hand_joints(i, k);
}

double ref_cal_cost(void *h, const double theta[26])
{
    vec th = to_vec(theta, 26);
    return ((ref_ctx *)h)->cost->cal_cost(th);
}

// costfunc::cal_cost2(theta, matchId, compute_corr); match holds n_pts indices (in when
// compute_corr == 0, out otherwise)
double ref_cal_cost2(void *h, const double theta[26], int *match, int n_pts, int compute_corr)
{
    ref_ctx *c = (ref_ctx *)h;
    vec th = to_vec(theta, 26);
    uvec m = zeros<uvec>(n_pts);
    if (!compute_corr)
        for (int i = 0; i < n_pts; ++i) m(i) = (uword)match[i];
    double v = c->cost->cal_cost2(th, m, compute_corr != 0, false);
    if (compute_corr)
        for (int i = 0; i < n_pts && i < (int)m.n_rows; ++i) match[i] = (int)m(i);
    return v;
}

// the three terms through the reference's public term methods, in cal_cost2's call order
// (src/costfunc.cpp:50-72); terms = (align, depth, collision)
void ref_cost_terms(void *h, const double theta[26], double terms[3], int *match)
{
    ref_ctx *c = (ref_ctx *)h;
    vec th = to_vec(theta, 26);
    mat::fixed<48, 3> S;
    c->hand->build_hand_model(th, S);
    uvec m;
    c->cost->compute_correspondences(*c->obs->get_ptncloud(), S, m);
    terms[0] = c->cost->align_models(*c->hand->get_radii(), S, *c->obs->get_ptncloud(), m);
    terms[1] = c->cost->depth_penalty(*c->obs->get_camera_mat(), *c->obs->get_depthmap(), S, *c->obs->get_disttran(),
                                      c->obs->get_img_scale());
    terms[2] = c->cost->self_collision_penalty(S, *c->hand->get_radii());
    if (match)
        for (uword i = 0; i < m.n_rows; ++i) match[i] = (int)m(i);
}

// costfunc::bincomp_penalty (src/costfunc.cpp:384-436) of the spheres of pose theta
double ref_bincomp_penalty(void *h, const double theta[26])
{
    ref_ctx *c = (ref_ctx *)h;
    vec th = to_vec(theta, 26);
    mat::fixed<48, 3> S;
    c->hand->build_hand_model(th, S);
    mat Sm = S;
    return c->cost->bincomp_penalty(Sm);
}

// costfunc::depthMatch_penalty (src/costfunc.cpp:439-473) of the spheres of pose theta: out = 240 x 320 doubles, row-major
void ref_depthmatch_penalty(void *h, const double theta[26], double *out)
{
    ref_ctx *c = (ref_ctx *)h;
    vec th = to_vec(theta, 26);
    mat::fixed<48, 3> S;
    c->hand->build_hand_model(th, S);
    mat Sm = S;
    cv::Mat img = c->cost->depthMatch_penalty(Sm);
    for (int y = 0; y < img.rows; ++y)
        for (int x = 0; x < img.cols; ++x) out[(long)y * img.cols + x] = img.at<double>(y, x);
}

// n poses through costfunc::cal_cost, parallelised over poses the way PSO::pso_evolve does
// (#pragma omp parallel for around cal_cost, src/PSO.cpp:848-852); used as the CPU baseline
void ref_cost_batch(void *h, const double *theta, long n, double *total)
{
    ref_ctx *c = (ref_ctx *)h;
#pragma omp parallel for
    for (long i = 0; i < n; ++i) {
        vec th = to_vec(theta + i * 26, 26);
        total[i] = c->cost->cal_cost(th);
    }
}

// PSO::set_pso_params, src/PSO.cpp:38-54 (argument order: upper, lower, std, ...)
void ref_set_pso_params(void *h, const double ub[26], const double lb[26], const double sd[26], double w, double c1,
                        double c2, int maxiter, double minstep, double minfunc)
{
    ref_ctx *c = (ref_ctx *)h;
    vec u = to_vec(ub, 26), l = to_vec(lb, 26), s = to_vec(sd, 26);
    c->pso->set_pso_params(u, l, s, w, c1, c2, maxiter, minstep, minfunc);
}

int ref_pso_evolve(void *h, const double x0[26], int num_p, double bestp[26])
{
    ref_ctx *c = (ref_ctx *)h;
    vec x = to_vec(x0, 26), b;
    int rc = c->pso->pso_evolve(*c->cost, x, num_p, b);
    for (int i = 0; i < 26; ++i) bestp[i] = b(i);
    return rc;
}

int ref_pso_optimise(void *h, const double x0[26], int num_p, double bestp[26])
{
    ref_ctx *c = (ref_ctx *)h;
    vec x = to_vec(x0, 26), b;
    int rc = c->pso->pso_optimise(*c->cost, x, num_p, b);
    for (int i = 0; i < 26; ++i) bestp[i] = b(i);
    return rc;
}

void ref_refine_init_pose(void *h, double x0[26])
{
    ref_ctx *c = (ref_ctx *)h;
    vec x = to_vec(x0, 26);
    c->pso->refine_init_pose(x, *c->cost);
    for (int i = 0; i < 26; ++i) x0[i] = x(i);
}

void ref_dim_restore(void *h, const double in22[22], double out26[26])
{
    ref_ctx *c = (ref_ctx *)h;
    vec a = to_vec(in22, 22), b = zeros<vec>(26);
    c->pso->dim_restore(a, b);
    for (int i = 0; i < 26; ++i) out26[i] = b(i);
}

// costfunc::gnd_truth_err(gnd_truth, frame), src/costfunc.cpp:476-507: gt is n_frames x 63 row-major
// (joint.txt rows); uses the joints of the last build_hand_model call
double ref_gnd_truth_err(void *h, const double *gt, int n_frames, int frame)
{
    ref_ctx *c = (ref_ctx *)h;
    mat G(n_frames, 63);
    for (int r = 0; r < n_frames; ++r)
        for (int k = 0; k < 63; ++k) G(r, k) = gt[r * 63 + k];
    return c->cost->gnd_truth_err(G, frame);
}

// ---- the stand-in Armadillo header itself, op by op (pinned against numpy in tests/test_ref_build.py) ----
// row-major plain arrays in and out; the ops are exactly the ones the reference's hot files call
void ref_shim_matmul(const double *a, int m, int k, const double *b, int n, double *out)
{
    mat A(m, k), B(k, n);
    for (int i = 0; i < m; ++i) for (int j = 0; j < k; ++j) A(i, j) = a[i * k + j];
    for (int i = 0; i < k; ++i) for (int j = 0; j < n; ++j) B(i, j) = b[i * n + j];
    mat Cm = A * B;
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) out[i * n + j] = Cm(i, j);
}
// out = {accu, mean, min, max, dot(x, x)} of a vector
void ref_shim_reduce(const double *x, int n, double out[5])
{
    vec v(n);
    for (int i = 0; i < n; ++i) v(i) = x[i];
    out[0] = accu(v); out[1] = mean(v); out[2] = v.min(); out[3] = v.max(); out[4] = dot(v, v);
}
// repmat(A, nr, nc) and floor(A) of an m x k matrix; out_rep is (m nr) x (k nc), out_floor m x k
void ref_shim_repmat_floor(const double *a, int m, int k, int nr, int nc, double *out_rep, double *out_floor)
{
    mat A(m, k);
    for (int i = 0; i < m; ++i) for (int j = 0; j < k; ++j) A(i, j) = a[i * k + j];
    mat R = repmat(A, nr, nc), F = floor(A);
    for (int i = 0; i < m * nr; ++i) for (int j = 0; j < k * nc; ++j) out_rep[i * (k * nc) + j] = R(i, j);
    for (int i = 0; i < m; ++i) for (int j = 0; j < k; ++j) out_floor[i * k + j] = F(i, j);
}
// column sums, sum(A, 0), and row sums, sum(A, 1)
void ref_shim_sums(const double *a, int m, int k, double *col_sums, double *row_sums)
{
    mat A(m, k);
    for (int i = 0; i < m; ++i) for (int j = 0; j < k; ++j) A(i, j) = a[i * k + j];
    mat c = sum(A, 0), r = sum(A, 1);
    for (int j = 0; j < k; ++j) col_sums[j] = c(0, j);
    for (int i = 0; i < m; ++i) row_sums[i] = r(i, 0);
}

}  // extern "C"
