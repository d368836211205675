/*
 * hpe_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See hpe_oracle.h for the rules on who may call this and for how it is
 * pinned against the reference build (oracle/_ref).  Build: make -C oracle  (gcc -O2 -fopenmp
 * -ffp-contract=off: IEEE double/float, no FMA contraction, so that the
 * evaluation order written here is the arithmetic that runs).
 *
 * All file:line citations are relative to /root/reference/.
 */
#include "hpe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#include "orc_draw.h"
#endif

/* ------------------------------------------------------------------ */
/* 4x4 helpers (row-major).  Armadillo multiplies 4x4 matrices with its */
/* "tiny square" kernel: each output element is the left-to-right sum   */
/* a0*b0 + a1*b1 + a2*b2 + a3*b3.                                       */
/* ------------------------------------------------------------------ */
static void m4_mul(double out[16], const double A[16], const double B[16])
{
    double t[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            t[r * 4 + c] = ((A[r * 4 + 0] * B[0 * 4 + c] + A[r * 4 + 1] * B[1 * 4 + c]) +
                            A[r * 4 + 2] * B[2 * 4 + c]) +
                           A[r * 4 + 3] * B[3 * 4 + c];
    memcpy(out, t, sizeof t);
}

static void m4_set(double T[16], double a00, double a01, double a02, double a03, double a10,
                   double a11, double a12, double a13, double a20, double a21, double a22,
                   double a23)
{
    T[0] = a00; T[1] = a01; T[2] = a02; T[3] = a03;
    T[4] = a10; T[5] = a11; T[6] = a12; T[7] = a13;
    T[8] = a20; T[9] = a21; T[10] = a22; T[11] = a23;
    T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;
}

/* src/fingermodel.cpp:8,203-205 and src/thumbmodel.cpp:8,219-221 */
static double deg2rad(double deg)
{
    const double PI = acos(-1);
    return deg / 180.0 * PI;
}

/* planar DH link [[c,-s,0,Lc],[s,c,0,Ls],..], e.g. src/fingermodel.cpp:142-145 */
static void m4_planar(double T[16], double ang, double L)
{
    m4_set(T, cos(ang), -sin(ang), 0, L * cos(ang),
              sin(ang),  cos(ang), 0, L * sin(ang),
              0, 0, 1, 0);
}

/* twist [[c,0,-s,0],[s,0,c,0],[0,-1,0,0]], src/fingermodel.cpp:137-140, src/thumbmodel.cpp:144-147 */
static void m4_twist(double T[16], double ang)
{
    m4_set(T, cos(ang), 0, -sin(ang), 0,
              sin(ang), 0,  cos(ang), 0,
              0, -1, 0, 0);
}

/* T00 * (Rz*Ry*Rx), src/fingermodel.cpp:157-180 + :289, src/thumbmodel.cpp:166-193 + :290 */
static void global_pose(double cur[16], const double gb_trans[3], const double g_pos[3])
{
    double TWS = deg2rad(gb_trans[0] + 180);
    double ANG = deg2rad(gb_trans[1]);
    double ROT = deg2rad(gb_trans[2]);
    double Rz[16], Ry[16], Rx[16], T00[16], Tgb[16];
    m4_set(Rz, cos(TWS), -sin(TWS), 0, 0, sin(TWS), cos(TWS), 0, 0, 0, 0, 1, 0);
    m4_set(Ry, cos(ANG), 0, sin(ANG), 0, 0, 1, 0, 0, -sin(ANG), 0, cos(ANG), 0);
    m4_set(Rx, 1, 0, 0, 0, 0, cos(ROT), -sin(ROT), 0, 0, sin(ROT), cos(ROT), 0);
    m4_set(T00, 1, 0, 0, g_pos[0], 0, 1, 0, g_pos[1], 0, 0, 1, g_pos[2]);
    m4_mul(Tgb, Rz, Ry);   /* Armadillo evaluates A*B*C of equal sizes as (A*B)*C */
    m4_mul(Tgb, Tgb, Rx);
    m4_mul(cur, T00, Tgb);
}

/*
 * One-time geometry staging: handmodel ctor / init_hand / process_geo
 * (src/handmodel.cpp:10-27, 69-85, 107-121) and the matrices the digits latch
 * behind setCMCTrans (src/fingermodel.cpp:106-132, src/thumbmodel.cpp:112-138).
 */
void orc_model_init(orc_model *m, const double geo_cm[20], const float spacing[5],
                    const double cmc_deg[5], const int tb_n[4], const int fg_n[4],
                    const double rad_cm[ORC_NSPH])
{
    memcpy(m->geo, geo_cm, sizeof m->geo);
    memcpy(m->spacing, spacing, sizeof m->spacing);
    memcpy(m->cmc, cmc_deg, sizeof m->cmc);
    memcpy(m->tb_n, tb_n, sizeof m->tb_n);
    memcpy(m->fg_n, fg_n, sizeof m->fg_n);
    memcpy(m->rad, rad_cm, sizeof m->rad);
    for (int d = 0; d < 5; ++d) {
        double nCMC = deg2rad(m->cmc[d]);
        double L = m->geo[4 * d];            /* L0 (thumb) or L4 (finger) */
        float sp = m->spacing[d];
        /* `spacing*spacing` is a float*float product in the reference */
        double a = sqrt(L * L + sp * sp - 2 * L * sp * cos(nCMC));
        double beta = asin(sin(nCMC) * sp / a);
        m4_planar(m->base[d], nCMC, L);
        if (d == 0) {
            /* src/thumbmodel.cpp:132-135 */
            m4_set(m->back[d], cos(beta), -sin(beta), 0, -a * cos(beta),
                               sin(beta),  cos(beta), 0, -a * sin(beta),
                               0, 0, 1, 0);
        } else {
            /* src/fingermodel.cpp:126-129 */
            m4_set(m->back[d], cos(beta), -sin(beta), 0, -L * sin(nCMC) * cos(beta),
                               sin(beta),  cos(beta), 0, -L * sin(nCMC) * sin(beta),
                               0, 0, 1, 0);
        }
    }
}

static void col3(const double T[16], double out[3])
{
    out[0] = T[3]; out[1] = T[7]; out[2] = T[11];
}

/* src/fingermodel.cpp:208-267 (finger) and src/thumbmodel.cpp:227-274 (thumb) */
static int build_spheres(const double joints[5][3], const int nsph[4], int is_thumb,
                         double *out /* rows of 3 */)
{
    int cnt = 0;
    for (int i = 0; i < 4; ++i) {
        const double *j1 = joints[i], *j2 = joints[i + 1];
        int j0, jend;
        double t;
        if (!is_thumb && i == 0) {
            t = 1. / (nsph[i] - 1);
            j0 = 0; jend = nsph[i];
        } else {
            t = 1. / nsph[i];
            j0 = 1; jend = nsph[i] + 1;
        }
        for (int j = j0; j < jend; ++j) {
            double a = 1. - t * j, b = t * j;
            for (int k = 0; k < 3; ++k) out[cnt * 3 + k] = a * j1[k] + b * j2[k];
            cnt++;
        }
    }
    return cnt;
}

/* src/fingermodel.cpp:270-317 */
static void finger_fk(const orc_model *m, int d, const double th[4], const double gb[3],
                      const double gp[3], double joints[5][3])
{
    const double *geo = &m->geo[4 * d];
    double T12[16], T23[16], T34[16], T45[16], T123[16], cur[16], basep[16];
    m4_twist(T12, deg2rad(th[0]));
    m4_planar(T23, deg2rad(th[1]), geo[1]);
    m4_planar(T34, deg2rad(th[2]), geo[2]);
    m4_planar(T45, deg2rad(th[3]), geo[3]);
    m4_mul(T123, T12, T23);
    global_pose(cur, gb, gp);
    m4_mul(cur, cur, m->base[d]);  col3(cur, joints[1]);
    m4_mul(basep, cur, m->back[d]); col3(basep, joints[0]);
    m4_mul(cur, cur, T123);        col3(cur, joints[2]);
    m4_mul(cur, cur, T34);         col3(cur, joints[3]);
    m4_mul(cur, cur, T45);         col3(cur, joints[4]);
}

/* src/thumbmodel.cpp:276-318 with the matrices of :140-164 */
static void thumb_fk(const orc_model *m, const double th[4], const double gb[3],
                     const double gp[3], double joints[5][3])
{
    const double PI = acos(-1);
    const double *geo = &m->geo[0];
    double nCMC = deg2rad(m->cmc[0]);
    double TMC2 = deg2rad(th[1]);
    double pCMC = nCMC + PI;
    double T01[16], T12[16], T23[16], T34[16], T012[16], cur[16], basep[16];
    m4_twist(T01, deg2rad(th[0]));
    m4_set(T12, cos(TMC2), -sin(TMC2) * cos(pCMC),  sin(TMC2) * sin(pCMC), geo[1] * cos(TMC2),
                sin(TMC2),  cos(TMC2) * cos(pCMC), -cos(TMC2) * sin(pCMC), geo[1] * sin(TMC2),
                0, sin(pCMC), cos(pCMC), 0);
    m4_planar(T23, deg2rad(th[2]), geo[2]);
    m4_planar(T34, deg2rad(th[3]), geo[3]);
    m4_mul(T012, T01, T12);
    global_pose(cur, gb, gp);
    m4_mul(cur, cur, m->base[0]);  col3(cur, joints[1]);
    m4_mul(basep, cur, m->back[0]); col3(basep, joints[0]);
    m4_mul(cur, cur, T012);        col3(cur, joints[2]);
    m4_mul(cur, cur, T23);         col3(cur, joints[3]);
    m4_mul(cur, cur, T34);         col3(cur, joints[4]);
}

/*
 * handmodel::build_hand_model, src/handmodel.cpp:259-298 (+ process_theta
 * :123-149).  spheres come back in the negated-y,z frame (:288); joints are
 * NOT negated (:291-296).
 */
void orc_build_hand_model(const orc_model *m, const double theta[ORC_NDOF],
                          double spheres[ORC_NSPH * 3], double joints[ORC_NJOINT * 3])
{
    const double *gb = &theta[0], *gp = &theta[3];
    double dj[5][5][3];
    int cnt = 0;
    thumb_fk(m, &theta[6], gb, gp, dj[0]);
    cnt += build_spheres((const double(*)[3])dj[0], m->tb_n, 1, &spheres[cnt * 3]);
    for (int d = 1; d < 5; ++d) {
        finger_fk(m, d, &theta[6 + 4 * d], gb, gp, dj[d]);
        cnt += build_spheres((const double(*)[3])dj[d], m->fg_n, 0, &spheres[cnt * 3]);
    }
    for (int i = 0; i < ORC_NSPH; ++i) {
        spheres[i * 3 + 1] *= -1;
        spheres[i * 3 + 2] *= -1;
    }
    if (joints) {
        for (int k = 0; k < 3; ++k) joints[k] = gp[k];
        for (int d = 1; d < 5; ++d)           /* index, middle, ring, little -> rows 1..16 */
            for (int j = 1; j < 5; ++j)
                for (int k = 0; k < 3; ++k)
                    joints[(1 + (d - 1) * 4 + (j - 1)) * 3 + k] = dj[d][j][k];
        for (int j = 1; j < 5; ++j)           /* thumb -> rows 17..20 */
            for (int k = 0; k < 3; ++k) joints[(17 + j - 1) * 3 + k] = dj[0][j][k];
    }
}

/* ------------------------------------------------------------------ */
/* Observation                                                          */
/* ------------------------------------------------------------------ */

/* arma::mean -> accumulate()/n with Armadillo's two-accumulator loop */
static double arma_accumulate(const double *x, long n)
{
    double acc1 = 0, acc2 = 0;
    long i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
        acc1 += x[i];
        acc2 += x[j];
    }
    if (i < n) acc1 += x[i];
    return acc1 + acc2;
}

/*
 * cv::distanceTransform(src, dst, CV_DIST_L2, 5) as called at
 * src/observedmodel.cpp:351 -- OpenCV is not under /root/reference (the
 * README asks for "OpenCV 3.0", unpinned); this restates its published
 * two-pass 5x5 chamfer algorithm: 16-bit fixed-point weights a=1, b=1.4,
 * c=2.1969 (-> 65536, 91750, 143976), border of 2 initialised to INT_MAX>>2,
 * forward then backward raster pass, result converted uint -> float (RN) and
 * scaled by 2^-16.  tests/test_oracle_pins.py checks it bit-for-bit against
 * cv2.distanceTransform with IPP disabled.
 */
void orc_chamfer5(const uint8_t *mask, int w, int h, float *out)
{
    const int B = 2;
    const unsigned HV = 65536u, DG = 91750u, LG = 143976u;
    const unsigned DIST_MAX = (unsigned)(0x7fffffff >> 2);
    const int step = w + 2 * B;
    unsigned *tmp = (unsigned *)malloc(sizeof(unsigned) * (size_t)step * (h + 2 * B));
    for (long k = 0; k < (long)step * (h + 2 * B); ++k) tmp[k] = DIST_MAX;
    for (int i = 0; i < h; ++i) {
        unsigned *t = tmp + (i + B) * step + B;
        for (int j = 0; j < w; ++j) {
            if (!mask[i * w + j]) {
                t[j] = 0;
            } else {
                unsigned t0 = t[j - step * 2 - 1] + LG, v;
                v = t[j - step * 2 + 1] + LG; if (t0 > v) t0 = v;
                v = t[j - step - 2] + LG;     if (t0 > v) t0 = v;
                v = t[j - step - 1] + DG;     if (t0 > v) t0 = v;
                v = t[j - step] + HV;         if (t0 > v) t0 = v;
                v = t[j - step + 1] + DG;     if (t0 > v) t0 = v;
                v = t[j - step + 2] + LG;     if (t0 > v) t0 = v;
                v = t[j - 1] + HV;            if (t0 > v) t0 = v;
                t[j] = t0;
            }
        }
    }
    const float scale = 1.f / 65536.f;
    for (int i = h - 1; i >= 0; --i) {
        unsigned *t = tmp + (i + B) * step + B;
        for (int j = w - 1; j >= 0; --j) {
            unsigned t0 = t[j], v;
            if (t0 > HV) {
                v = t[j + step * 2 + 1] + LG; if (t0 > v) t0 = v;
                v = t[j + step * 2 - 1] + LG; if (t0 > v) t0 = v;
                v = t[j + step + 2] + LG;     if (t0 > v) t0 = v;
                v = t[j + step + 1] + DG;     if (t0 > v) t0 = v;
                v = t[j + step] + HV;         if (t0 > v) t0 = v;
                v = t[j + step - 1] + DG;     if (t0 > v) t0 = v;
                v = t[j + step - 2] + LG;     if (t0 > v) t0 = v;
                v = t[j + 1] + HV;            if (t0 > v) t0 = v;
                t[j] = t0;
            }
            t0 = (t0 > DIST_MAX) ? DIST_MAX : t0;
            out[i * w + j] = (float)t0 * scale;
        }
    }
    free(tmp);
}

static void obs_alloc(orc_obs *o, int n_pts)
{
    o->depth = (double *)malloc(sizeof(double) * ORC_W * ORC_H);
    o->dt = (double *)malloc(sizeof(double) * ORC_W * ORC_H);
    o->cloud = (double *)malloc(sizeof(double) * 3 * (n_pts > 0 ? n_pts : 1));
}

void orc_obs_free(orc_obs *o)
{
    free(o->depth); free(o->dt); free(o->cloud);
    o->depth = o->dt = o->cloud = 0;
}

/*
 * observedmodel::next_frame = load_data + depth_to_ptncloud + dist_transform
 * (src/observedmodel.cpp:420-430, 272-310, 110-219, 313-369) for the
 * reference's configuration imgH=320 (width!), imgW=240, to_cm=true
 * (src/testmodel.cpp:61-64).  depth_mm is the file payload: float32[240*320],
 * value at y*320+x = depth of pixel (x,y) in mm, 0 = background.
 * Returns 0, or -1 when the frame has no hand pixel (the reference would
 * index an empty matrix at :215).
 */
int orc_obs_from_depth(const float *depth_mm, double focal, int downsample, int num_samples,
                       orc_obs *o)
{
    const int W = ORC_W, H = ORC_H;
    memset(o, 0, sizeof *o);
    o->focal = focal;
    o->cx = W / 2.;   /* img_center << imH/2. << imW/2.  (:56) */
    o->cy = H / 2.;
    int n = 0;
    for (int k = 0; k < W * H; ++k) n += depth_mm[k] != 0.0f;
    if (n == 0) return -1;
    int n_pts = downsample ? num_samples : n;
    obs_alloc(o, n_pts);
    o->n_hand = n;
    o->n_pts = n_pts;
    /* load_data :295-308: widen to double, /10, depthmap(y,x) = file[y*320+x] */
    for (int k = 0; k < W * H; ++k) o->depth[k] = (double)depth_mm[k] / 10.;

    /* depth_to_ptncloud :131-161: raster-order list of non-zero pixels */
    double *full = (double *)malloc(sizeof(double) * 3 * n);
    double *cmpp = (double *)malloc(sizeof(double) * n);
    long nnz = 0;
    int r = 0;
    const double rad = 2.0;
    for (int v = 0; v < H; ++v) {
        for (int u = 0; u < W; ++u) {
            double Z = o->depth[v * W + u];
            if (Z == 0.0) continue;
            double X = (u - o->cx) * Z / focal;
            double Y = (v - o->cy) * Z / focal;
            full[r * 3 + 0] = X;
            full[r * 3 + 1] = Y * -1;
            full[r * 3 + 2] = Z * -1;
            /* scale estimate :171-202.  K*(X,Y,Z): row0 = f*X + 0*Y + cx*Z */
            double cu = floor((focal * X + o->cx * Z) / Z);
            double cv = floor((focal * Y + o->cy * Z) / Z);
            double eu = floor((focal * (X + rad) + o->cx * Z) / Z);
            double ev = floor((focal * Y + o->cy * Z) / Z);
            double dist = sqrt((eu - cu) * (eu - cu) + (ev - cv) * (ev - cv));
            if (dist != 0.0) cmpp[nnz++] = rad / dist;
            r++;
        }
    }
    o->scale = arma_accumulate(cmpp, nnz) / (double)nnz;
    /* :204-217 stride subsample, integer division */
    if (downsample) {
        int sample_f = n / num_samples;
        for (int k = 0; k < num_samples; ++k)
            memcpy(&o->cloud[k * 3], &full[(long)k * sample_f * 3], 3 * sizeof(double));
    } else {
        memcpy(o->cloud, full, sizeof(double) * 3 * n);
    }
    free(full); free(cmpp);

    /* invert_depthmap + dist_transform :313-369 */
    uint8_t *mask = (uint8_t *)malloc(W * H);
    float *dtf = (float *)malloc(sizeof(float) * W * H);
    for (int k = 0; k < W * H; ++k) mask[k] = o->depth[k] == 0.0 ? 255 : 0;
    orc_chamfer5(mask, W, H, dtf);
    double mx = dtf[0];
    for (int k = 0; k < W * H; ++k) {
        o->dt[k] = (double)dtf[k];
        if (o->dt[k] > mx) mx = o->dt[k];
    }
    o->dt_max = mx;
    free(mask); free(dtf);
    return 0;
}

int orc_obs_from_parts(const double *depth_cm, const double *dt_px, const double *cloud,
                       int n_pts, double scale, double focal, orc_obs *o)
{
    memset(o, 0, sizeof *o);
    obs_alloc(o, n_pts);
    o->n_pts = n_pts; o->scale = scale; o->focal = focal;
    o->cx = ORC_W / 2.; o->cy = ORC_H / 2.;
    memcpy(o->depth, depth_cm, sizeof(double) * ORC_W * ORC_H);
    memcpy(o->dt, dt_px, sizeof(double) * ORC_W * ORC_H);
    memcpy(o->cloud, cloud, sizeof(double) * 3 * n_pts);
    double mx = dt_px[0];
    for (int k = 0; k < ORC_W * ORC_H; ++k) if (dt_px[k] > mx) mx = dt_px[k];
    o->dt_max = mx;
    return 0;
}

/* ------------------------------------------------------------------ */
/* Energy                                                               */
/* ------------------------------------------------------------------ */

/*
 * costfunc::compute_correspondences, src/costfunc.cpp:306-343.  The cloud and
 * the centres are converted to float32 (:321-322) and handed to
 * cv::BFMatcher(cv::NORM_L2).match (:331-334).  OpenCV is not vendored; its
 * brute-force matcher computes, per pair, d = ((t0*t0 + t1*t1) + t2*t2) in
 * float32 with t = a - b, takes sqrtf, and keeps the first strictly smaller
 * value (ties -> lowest train index).  Checked bit-for-bit against cv2 in
 * tests/test_oracle_pins.py.  second (optional) receives the runner-up
 * distance for the margin diagnostics.
 */
static void correspondences_ex(const double *cloud, int n_pts, const double *spheres,
                               int32_t *match, float *best, float *second)
{
    float c[ORC_NSPH][3];
    for (int j = 0; j < ORC_NSPH; ++j)
        for (int k = 0; k < 3; ++k) c[j][k] = (float)spheres[j * 3 + k];
    for (int i = 0; i < n_pts; ++i) {
        float p0 = (float)cloud[i * 3], p1 = (float)cloud[i * 3 + 1], p2 = (float)cloud[i * 3 + 2];
        float b = INFINITY, s = INFINITY;
        int bi = 0;
        for (int j = 0; j < ORC_NSPH; ++j) {
            float t0 = p0 - c[j][0], t1 = p1 - c[j][1], t2 = p2 - c[j][2];
            float d = 0.f;
            d += t0 * t0;
            d += t1 * t1;
            d += t2 * t2;
            float w = sqrtf(d);
            if (w < b) { s = b; b = w; bi = j; }
            else if (w < s) s = w;
        }
        match[i] = bi;
        if (best) best[i] = b;
        if (second) second[i] = s;
    }
}

void orc_correspondences(const double *cloud, int n_pts, const double *spheres, int32_t *match)
{
    correspondences_ex(cloud, n_pts, spheres, match, 0, 0);
}

/* costfunc::align_models, src/costfunc.cpp:346-377 */
double orc_align(const double *rad, const double *spheres, const double *cloud, int n_pts,
                 const int32_t *match)
{
    double *dif = (double *)malloc(sizeof(double) * (n_pts > 0 ? n_pts : 1));
    for (int i = 0; i < n_pts; ++i) {
        const double *c = &spheres[match[i] * 3], *p = &cloud[i * 3];
        double a = p[0] - c[0], b = p[1] - c[1], e = p[2] - c[2];
        double nor = sqrt(a * a + b * b + e * e);
        double df = fabs(nor - rad[match[i]]);
        dif[i] = df * df;
    }
    double lambda = (double)ORC_NSPH / n_pts;
    double dis = arma_accumulate(dif, n_pts) * lambda;
    free(dif);
    return dis;
}

/*
 * costfunc::depth_penalty, src/costfunc.cpp:227-304.  The caller passes the
 * centres with y,z already un-negated (the reference flips them in place at
 * :249).  K*c: row0 = f*x + 0*y + cx*z, row1 = 0*x + f*y + cy*z, row2 = z.
 */
static double depth_penalty_ex(const orc_model *m, const orc_obs *o, const double *sp,
                               double *floor_margin)
{
    double pen = 0.0, margin = 1.0;
    for (int i = 0; i < ORC_NSPH; ++i) {
        double x = sp[i * 3], y = sp[i * 3 + 1], z = sp[i * 3 + 2];
        double pu = (o->focal * x + o->cx * z) / z;
        double pv = (o->focal * y + o->cy * z) / z;
        double dx = floor(pu), dy = floor(pv);
        if (floor_margin) {
            double a = pu - dx, b = pv - dy;
            if (a > 0.5) a = 1 - a;
            if (b > 0.5) b = 1 - b;
            if (a < margin) margin = a;
            if (b < margin) margin = b;
        }
        int xb = dx >= 0 && dx < ORC_W;
        int yb = dy >= 0 && dy < ORC_H;
        if (xb && yb) {
            long at = (long)dy * ORC_W + (long)dx;
            double d_jc = o->depth[at];
            if (d_jc != 0.0) {
                double diff = fmax(0.0, d_jc - z);
                pen += diff * diff;
            } else {
                double ddis = o->dt[at] * o->scale + m->rad[i];
                pen += ddis * ddis;
            }
        } else {
            double maxdis = o->dt_max * o->scale + m->rad[i];
            pen += maxdis * maxdis;
        }
    }
    if (floor_margin) *floor_margin = margin;
    return pen;
}

double orc_depth_penalty(const orc_model *m, const orc_obs *o, const double *spheres_unneg)
{
    return depth_penalty_ex(m, o, spheres_unneg, 0);
}

/*
 * costfunc::self_collision_penalty, src/costfunc.cpp:130-197: last 6 spheres
 * of each digit, the 4 adjacent digit pairs, 36 sphere pairs each in
 * (a of digit i) major, (b of digit i+1) minor order.
 */
double orc_collision(const double *sp, const double *rad)
{
    static const int first[5] = {2, 12, 22, 32, 42};
    double penalty = 0.0;
    for (int i = 0; i < 4; ++i) {
        double sq[36];
        int n = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) {
                int ia = first[i] + a, ib = first[i + 1] + b;
                double d0 = sp[ib * 3] - sp[ia * 3], d1 = sp[ib * 3 + 1] - sp[ia * 3 + 1],
                       d2 = sp[ib * 3 + 2] - sp[ia * 3 + 2];
                double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                dist = rad[ib] + rad[ia] - dist;
                if (dist > 0) sq[n++] = dist * dist;
            }
        penalty += arma_accumulate(sq, n);
    }
    return penalty;
}

/* costfunc::cal_cost2, src/costfunc.cpp:31-86 */
double orc_cal_cost2(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF],
                     int32_t *match, int compute_corr, double terms[3])
{
    double S[ORC_NSPH * 3];
    orc_build_hand_model(m, theta, S, 0);
    if (compute_corr) orc_correspondences(o->cloud, o->n_pts, S, match);
    double align_cost = orc_align(m->rad, S, o->cloud, o->n_pts, match);
    for (int i = 0; i < ORC_NSPH; ++i) { S[i * 3 + 1] *= -1; S[i * 3 + 2] *= -1; }
    double depth_cost = orc_depth_penalty(m, o, S);
    double colli_cost = orc_collision(S, m->rad);
    if (terms) { terms[0] = align_cost; terms[1] = depth_cost; terms[2] = colli_cost; }
    return align_cost + depth_cost + colli_cost;
}

/* costfunc::cal_cost, src/costfunc.cpp:89-127 (no collision term, :122) */
double orc_cal_cost(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF])
{
    double S[ORC_NSPH * 3];
    int32_t *match = (int32_t *)malloc(sizeof(int32_t) * (o->n_pts > 0 ? o->n_pts : 1));
    orc_build_hand_model(m, theta, S, 0);
    orc_correspondences(o->cloud, o->n_pts, S, match);
    double align_cost = orc_align(m->rad, S, o->cloud, o->n_pts, match);
    for (int i = 0; i < ORC_NSPH; ++i) { S[i * 3 + 1] *= -1; S[i * 3 + 2] *= -1; }
    double depth_cost = orc_depth_penalty(m, o, S);
    free(match);
    return align_cost + depth_cost;
}

void orc_cost_batch(const orc_model *m, const orc_obs *o, const double *theta, int64_t n,
                    double *align, double *depth, double *colli, double *floor_margin,
                    double *match_gap)
{
#pragma omp parallel
    {
        int np = o->n_pts > 0 ? o->n_pts : 1;
        int32_t *match = (int32_t *)malloc(sizeof(int32_t) * np);
        float *b = (float *)malloc(sizeof(float) * np), *s = (float *)malloc(sizeof(float) * np);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            double S[ORC_NSPH * 3];
            orc_build_hand_model(m, &theta[i * ORC_NDOF], S, 0);
            correspondences_ex(o->cloud, o->n_pts, S, match, b, s);
            double a = orc_align(m->rad, S, o->cloud, o->n_pts, match);
            for (int k = 0; k < ORC_NSPH; ++k) { S[k * 3 + 1] *= -1; S[k * 3 + 2] *= -1; }
            double fm;
            double d = depth_penalty_ex(m, o, S, &fm);
            if (align) align[i] = a;
            if (depth) depth[i] = d;
            if (colli) colli[i] = orc_collision(S, m->rad);
            if (floor_margin) floor_margin[i] = fm;
            if (match_gap) {
                double g = INFINITY;
                for (int k = 0; k < o->n_pts; ++k)
                    if ((double)s[k] - (double)b[k] < g) g = (double)s[k] - (double)b[k];
                match_gap[i] = g;
            }
        }
        free(match); free(b); free(s);
    }
}

/* ------------------------------------------------------------------ */
/* RNG: Philox4x32-10 (Salmon et al., SC'11).  Replaces Armadillo's      */
/* process-global generator (arma_rng::set_seed, src/PSO.cpp:722).       */
/* counter = (elem_pair, particle, generation, stream); key = seed.      */
/* ------------------------------------------------------------------ */
void orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u53(uint32_t hi, uint32_t lo)
{
    uint64_t k = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return (double)k * (1.0 / 9007199254740992.0);
}

void orc_uniform2(uint64_t seed, uint32_t elem_pair, uint32_t particle, uint32_t gen,
                  uint32_t stream, double u[2])
{
    uint32_t w[4];
    orc_philox(seed, elem_pair, particle, gen, stream, w);
    u[0] = u53(w[0], w[1]);
    u[1] = u53(w[2], w[3]);
}

/* Box-Muller on stream 0: z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln u1),
 * u1 shifted to (0,1]. */
void orc_normal2(uint64_t seed, uint32_t elem_pair, uint32_t particle, uint32_t gen, double z[2])
{
    double u[2];
    orc_uniform2(seed, elem_pair, particle, gen, 0u, u);
    double u1 = u[0] + 1.0 / 9007199254740992.0;
    double r = sqrt(-2.0 * log(u1));
    double a = 6.283185307179586476925286766559 * u[1];
    z[0] = r * cos(a);
    z[1] = r * sin(a);
}

/* ------------------------------------------------------------------ */
/* Optimiser                                                            */
/* ------------------------------------------------------------------ */

/* PSO::generate_particles, src/PSO.cpp:56-74: x0 + randn % std, no clamp */
void orc_generate_particles(const orc_pso_params *p, const double x0[ORC_NDOF], int num_p,
                            uint64_t seed, double *particles)
{
    for (int i = 0; i < num_p; ++i)
        for (int d = 0; d < ORC_NDOF; d += 2) {
            double z[2];
            orc_normal2(seed, (uint32_t)(d >> 1), (uint32_t)i, 0u, z);
            particles[i * ORC_NDOF + d] = x0[d] + z[0] * p->std[d];
            particles[i * ORC_NDOF + d + 1] = x0[d + 1] + z[1] * p->std[d + 1];
        }
}

/* PSO::check_constraints, src/PSO.cpp:358-377.  Note :372 resets an
 * upper-bound violation to the LOWER bound. */
void orc_check_constraints(const orc_pso_params *p, double theta[ORC_NDOF], double vel[ORC_NDOF])
{
    for (int d = 0; d < ORC_NDOF; ++d) {
        int lo = theta[d] < p->lb[d], hi = theta[d] > p->ub[d];
        if (lo) { theta[d] = p->lb[d]; vel[d] = 0.; }
        if (hi) { theta[d] = p->lb[d]; vel[d] = 0.; }
    }
}

/* PSO::pso_evolve, src/PSO.cpp:717-886 */
int orc_pso_evolve(const orc_model *m, const orc_obs *o, const orc_pso_params *p,
                   const double x0[ORC_NDOF], int num_p, uint64_t seed,
                   double bestp[ORC_NDOF], double *best_cost, double *trace,
                   double *final_particles, double *final_pcost)
{
    const int D = ORC_NDOF;
    double *x = (double *)calloc((size_t)num_p * D, sizeof(double));
    double *v = (double *)calloc((size_t)num_p * D, sizeof(double));
    double *pb = (double *)calloc((size_t)num_p * D, sizeof(double));
    double *pcost = (double *)calloc(num_p, sizeof(double));
    double gbest[ORC_NDOF] = {0};
    double gbest_cost = 1e100;
    const int nK = 3;
    int32_t *R = (int32_t *)calloc((size_t)num_p * nK, sizeof(int32_t));
    int32_t *informant = (int32_t *)calloc(num_p, sizeof(int32_t));

    orc_generate_particles(p, x0, num_p, seed, x);

    /* :748-763 (the OpenMP race on gbest is resolved as the serial order) */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < num_p; ++i) {
        pcost[i] = orc_cal_cost(m, o, &x[i * D]);
        memcpy(&pb[i * D], &x[i * D], D * sizeof(double));
    }
    for (int i = 0; i < num_p; ++i)
        if (pcost[i] < gbest_cost) {
            gbest_cost = pcost[i];
            memcpy(gbest, &x[i * D], D * sizeof(double));
        }

    int iter = 1, count = 100;
    const double W1 = 1. / (2 * log(2));
    const double C1 = 0.5 + log(2);
    const double C2 = C1;

    while (iter < p->maxiter) {
        iter += 1;
        uint32_t gen = (uint32_t)(iter - 1);

        /* :790-803 topology; R(s*nK+k) = floor(U*(num_p-1)+0.5) */
        if (count > 0) {
            for (int s = 0; s < num_p; ++s)
                for (int k = 0; k < nK; ++k) {
                    double u[2];
                    orc_uniform2(seed, (uint32_t)(k >> 1), (uint32_t)s, gen, 3u, u);
                    R[s * nK + k] = (int32_t)floor(u[k & 1] * (num_p - 1) + 0.5);
                }
        }
        /* :810-812 informant = first minimum of pcost over {s : L(s,i)=1}, L = eye + links */
        for (int i = 0; i < num_p; ++i) informant[i] = i;
        for (int i = 0; i < num_p; ++i) {
            int best = -1;
            for (int s = 0; s < num_p; ++s) {
                int linked = (s == i);
                for (int k = 0; k < nK && !linked; ++k) linked = (R[s * nK + k] == i);
                if (linked && (best < 0 || pcost[s] < pcost[best])) best = s;
            }
            informant[i] = best;
        }
        /* :807-845 */
        for (int i = 0; i < num_p; ++i) {
            int inf = informant[i];
            for (int d = 0; d < D; d += 2) {
                double rp[2], rg[2];
                orc_uniform2(seed, (uint32_t)(d >> 1), (uint32_t)i, gen, 1u, rp);
                orc_uniform2(seed, (uint32_t)(d >> 1), (uint32_t)i, gen, 2u, rg);
                for (int e = 0; e < 2; ++e) {
                    int dd = d + e;
                    double xi = x[i * D + dd];
                    double nv = v[i * D + dd] * W1 + (rp[e] * C1) * (pb[i * D + dd] - xi);
                    if (inf != i) nv = nv + (rg[e] * C2) * (pb[inf * D + dd] - xi);
                    v[i * D + dd] = nv;
                    x[i * D + dd] = xi + nv;
                }
            }
            orc_check_constraints(p, &x[i * D], &v[i * D]);
        }
        /* :848-861 */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < num_p; ++i) {
            double fx = orc_cal_cost(m, o, &x[i * D]);
            if (fx < pcost[i]) {
                pcost[i] = fx;
                memcpy(&pb[i * D], &x[i * D], D * sizeof(double));
            }
        }
        /* :864-878 */
        int fmin_id = 0;
        for (int i = 1; i < num_p; ++i) if (pcost[i] < pcost[fmin_id]) fmin_id = i;
        double fmin = pcost[fmin_id];
        if (fmin < gbest_cost) {
            memcpy(gbest, &x[fmin_id * D], D * sizeof(double));
            gbest_cost = fmin;
            count = 0;
        } else {
            count += 1;
        }
        if (trace) trace[iter - 2] = gbest_cost;
    }
    memcpy(bestp, gbest, D * sizeof(double));
    if (best_cost) *best_cost = gbest_cost;
    if (final_particles) memcpy(final_particles, x, sizeof(double) * num_p * D);
    if (final_pcost) memcpy(final_pcost, pcost, sizeof(double) * num_p);
    free(x); free(v); free(pb); free(pcost); free(R); free(informant);
    return 1;
}

/* ------------------------------------------------------------------ */
/* PSO::refine_init_pose and its helpers            */
/* ------------------------------------------------------------------ */

/* op_dot::direct_dot_arma: two partial dot products over even / odd elements */
static double arma_dot(const double *a, const double *b, int n)
{
    double v1 = 0, v2 = 0;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) { v1 += a[i] * b[i]; v2 += a[j] * b[j]; }
    if (i < n) v1 += a[i] * b[i];
    return v1 + v2;
}

/* PSO::cal_grad, src/PSO.cpp:183-214: central differences (eps = 1e-5) of cal_cost2 with FROZEN
 * correspondences over the selected parameters, zero elsewhere */
void orc_cal_grad(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF],
                  const int sel[ORC_NDOF], int32_t *match, double grad[ORC_NDOF])
{
    for (int i = 0; i < ORC_NDOF; ++i) {
        grad[i] = 0.0;
        if (!sel[i]) continue;
        const double eps = 1e-5;
        double xph[ORC_NDOF], xmh[ORC_NDOF];
        memcpy(xph, theta, sizeof xph);
        memcpy(xmh, theta, sizeof xmh);
        xph[i] += eps;
        xmh[i] -= eps;
        double f_xph = orc_cal_cost2(m, o, xph, match, 0, 0);
        double f_xmh = orc_cal_cost2(m, o, xmh, match, 0, 0);
        grad[i] = (f_xph - f_xmh) / (2 * eps);
    }
}

/* Decision margins (test instrumentation, not part of the reference): the line searches decide by comparing two
 * doubles; a comparison whose two sides differ by less than the GPU path's evaluation noise (~1e-13 relative in the
 * costs, amplified to ~1e-8 by the eps = 1e-5 difference quotients) may legitimately go the other way there.  The
 * smallest relative gap |lhs - rhs| / max(|lhs|, |rhs|) met since the last reset is kept so that a test can accept a
 * diverging search path ONLY where the oracle itself saw a near-tie. */
static double g_min_margin = 1e300;
static void note_margin(double lhs, double rhs)
{
    double sc = fabs(lhs) > fabs(rhs) ? fabs(lhs) : fabs(rhs);
    double m = sc > 0 ? fabs(lhs - rhs) / sc : 0.0;
#pragma omp critical(orc_margin)
    {
        if (m < g_min_margin) g_min_margin = m;
    }
}
void orc_margin_reset(void) { g_min_margin = 1e300; }
double orc_margin_min(void) { return g_min_margin; }

/* PSO::goldstein, src/PSO.cpp:438-480: bracketing line search on the Armijo / Goldstein pair;
 * returns the iteration count, *tk = accepted step or 0 */
int orc_goldstein(const orc_model *m, const orc_obs *o, const double theta[ORC_NDOF],
                  const double g_k[ORC_NDOF], int32_t *match, double f_k, double *tk, int maxiter)
{
    double a = 0, b = 1e100, alpha = 0.5;
    const double t = 2, c = 0.25;
    double p_k[ORC_NDOF];
    for (int i = 0; i < ORC_NDOF; ++i) p_k[i] = -1 * g_k[i];
    int iter = 0;
    while (iter < maxiter) {
        iter += 1;
        double theta1[ORC_NDOF];
        for (int i = 0; i < ORC_NDOF; ++i) theta1[i] = theta[i] + alpha * p_k[i];
        double f_k1 = orc_cal_cost2(m, o, theta1, match, 0, 0);
        double armijo = f_k + c * alpha * arma_dot(g_k, p_k, ORC_NDOF);
        double goldstein = f_k + (1 - c) * alpha * arma_dot(g_k, p_k, ORC_NDOF);
        note_margin(f_k1, armijo);
        if (f_k1 <= armijo) {
            note_margin(f_k1, goldstein);
            if (f_k1 >= goldstein) {
                *tk = alpha;
                return iter;
            }
            a = alpha;
            {
                /* std::min(t*alpha, 0.5*(alpha+b)): the first argument unless the second is smaller */
                double x = t * alpha, y = 0.5 * (alpha + b);
                alpha = y < x ? y : x;
            }
        } else {
            b = alpha;
            alpha = 0.5 * (a + alpha);
        }
    }
    *tk = 0;
    return iter;
}

/* PSO::refine_init_pose, src/PSO.cpp:216-266: gradient descent on theta[0:3] (rotation), then on
 * theta[3:6] (position); each group: up to 15 iterations of (cal_cost2 with fresh correspondences,
 * numerical gradient on the frozen ones, Goldstein step), stopping after the first failed line
 * search or when |grad| <= 1e-6.  evals (optional) counts cal_cost2 calls. */
void orc_refine_init_pose(const orc_model *m, const orc_obs *o, double x0[ORC_NDOF], int *evals)
{
    const int len = 2;
    const int start_idx[7] = {0, 3, 6, 10, 14, 18, 22};
    const int end_idx[7] = {2, 5, 9, 13, 17, 21, 25};
    int32_t *match = (int32_t *)malloc(sizeof(int32_t) * (o->n_pts > 0 ? o->n_pts : 1));
    int n_evals = 0;
    for (int i = 0; i < len; ++i) {
        const double eps = 1e-6;
        double tol = 1;
        int cnt = 0, iter = 0;
        const int maxiter = 15;
        while (tol > eps && iter < maxiter && cnt < 1) {
            int sel[ORC_NDOF] = {0};
            for (int d = start_idx[i]; d <= end_idx[i]; ++d) sel[d] = 1;
            double grad[ORC_NDOF];
            double f_k = orc_cal_cost2(m, o, x0, match, 1, 0);
            orc_cal_grad(m, o, x0, sel, match, grad);
            double tk = 0;
            int ls = orc_goldstein(m, o, x0, grad, match, f_k, &tk, 30);
            n_evals += 1 + 2 * (end_idx[i] - start_idx[i] + 1) + ls;
            if (tk == 0) cnt += 1;
            for (int d = 0; d < ORC_NDOF; ++d) x0[d] = x0[d] - tk * grad[d];
            double sq[ORC_NDOF];
            for (int d = 0; d < ORC_NDOF; ++d) sq[d] = grad[d] * grad[d];
            tol = sqrt(arma_accumulate(sq, ORC_NDOF));
            iter += 1;
        }
    }
    if (evals) *evals = n_evals;
    free(match);
}

/* ------------------------------------------------------------------ */
/* PSO::pso_optimise (gradient descent + canonical PSO) */
/* ------------------------------------------------------------------ */

/* PSO::pso_optimise, src/PSO.cpp:539-712.  Per iteration every particle first runs `graditer` = 10
 * coordinate-descent steps (:592-626): parameter permu(m) = randi[0,25], f_k = cal_cost2 (fresh
 * correspondences only for m == 0, :600-604), one-sided-parameter central difference cal_gradient
 * (:380-405), Goldstein step (:438-480), f_k again -- for m == 0 this second call recomputes the
 * correspondences at the moved pose (:616, cal_corr is still true), the later steps reuse them --
 * and a pbest update against pcost (:619-625; the velocity fill is overwritten by :630).  Then the
 * canonical update with w, c1, c2 towards pbest and gbest (:664-666), check_constraints, cal_cost
 * and the second gbest test.  gbest_pos is the CURRENT position of the particle holding the
 * smallest pcost (:637, :697), not its pbest, exactly as written.  The OpenMP races of :556-568
 * (gbest) and :588 (randi from the shared generator) are resolved as the serial order; the random
 * draws are counter-based: stream 0 initial normals, 1 rp, 2 rg (as pso_evolve), 4 permu
 * (element m of particle i in generation gen = floor(26 u), u = uniform2(m>>1, i, gen, 4)[m&1]).
 * evals (optional) counts cal_cost + cal_cost2 calls. */
int orc_pso_optimise(const orc_model *m, const orc_obs *o, const orc_pso_params *p,
                     double w, double c1, double c2, const double x0[ORC_NDOF], int num_p,
                     uint64_t seed, double bestp[ORC_NDOF], double *best_cost, int64_t *evals,
                     double *final_particles, double *final_pcost)
{
    const int D = ORC_NDOF, graditer = 10;
    double *x = (double *)calloc((size_t)num_p * D, sizeof(double));
    double *v = (double *)calloc((size_t)num_p * D, sizeof(double));
    double *pb = (double *)calloc((size_t)num_p * D, sizeof(double));
    double *pcost = (double *)calloc(num_p, sizeof(double));
    double gbest[ORC_NDOF] = {0};
    double gbest_cost = 1e100;
    int64_t n_evals = 0;
    const int npts = o->n_pts > 0 ? o->n_pts : 1;

    orc_generate_particles(p, x0, num_p, seed, x);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < num_p; ++i) {
        pcost[i] = orc_cal_cost(m, o, &x[i * D]);
        memcpy(&pb[i * D], &x[i * D], D * sizeof(double));
    }
    n_evals += num_p;
    for (int i = 0; i < num_p; ++i)
        if (pcost[i] < gbest_cost) {
            gbest_cost = pcost[i];
            memcpy(gbest, &x[i * D], D * sizeof(double));
        }

    int iter = 1;
    while (iter < p->maxiter) {
        iter += 1;
        const uint32_t gen = (uint32_t)(iter - 1);

        /* :586-633 coordinate descent of every particle (particles are independent here) */
#pragma omp parallel for schedule(dynamic) reduction(+ : n_evals)
        for (int i = 0; i < num_p; ++i) {
            int32_t *match = (int32_t *)malloc(sizeof(int32_t) * npts);
            double ctheta[ORC_NDOF], cveloc[ORC_NDOF];
            memcpy(ctheta, &x[i * D], sizeof ctheta);
            memcpy(cveloc, &v[i * D], sizeof cveloc);
            for (int mm = 0; mm < graditer; ++mm) {
                double u[2];
                orc_uniform2(seed, (uint32_t)(mm >> 1), (uint32_t)i, gen, 4u, u);
                int param = (int)floor(u[mm & 1] * D);
                if (param > D - 1) param = D - 1;
                const int cal_corr = mm == 0;
                double f_k = orc_cal_cost2(m, o, ctheta, match, cal_corr, 0);
                int sel[ORC_NDOF] = {0};
                double g_k[ORC_NDOF];
                sel[param] = 1;
                orc_cal_grad(m, o, ctheta, sel, match, g_k);      /* cal_gradient :380-405: same arithmetic */
                double tk = 0;
                int ls = orc_goldstein(m, o, ctheta, g_k, match, f_k, &tk, 30);
                for (int d = 0; d < D; ++d) ctheta[d] = ctheta[d] - tk * g_k[d];
                f_k = orc_cal_cost2(m, o, ctheta, match, cal_corr, 0);
                n_evals += 4 + ls;
                if (f_k < pcost[i]) {
                    pcost[i] = f_k;
                    memcpy(&pb[i * D], ctheta, sizeof ctheta);
                }
            }
            orc_check_constraints(p, ctheta, cveloc);
            memcpy(&x[i * D], ctheta, sizeof ctheta);
            memcpy(&v[i * D], cveloc, sizeof cveloc);
            free(match);
        }
        /* :635-642 */
        int fmin_id = 0;
        for (int i = 1; i < num_p; ++i) if (pcost[i] < pcost[fmin_id]) fmin_id = i;
        if (pcost[fmin_id] < gbest_cost) {
            memcpy(gbest, &x[fmin_id * D], D * sizeof(double));
            gbest_cost = pcost[fmin_id];
        }
        /* :644-677 canonical velocity / position update */
        for (int i = 0; i < num_p; ++i) {
            for (int d = 0; d < D; d += 2) {
                double rp[2], rg[2];
                orc_uniform2(seed, (uint32_t)(d >> 1), (uint32_t)i, gen, 1u, rp);
                orc_uniform2(seed, (uint32_t)(d >> 1), (uint32_t)i, gen, 2u, rg);
                for (int e = 0; e < 2; ++e) {
                    int dd = d + e;
                    double xi = x[i * D + dd];
                    double nv = (w * v[i * D + dd] + (c1 * rp[e]) * (pb[i * D + dd] - xi)) +
                                (c2 * rg[e]) * (gbest[dd] - xi);
                    v[i * D + dd] = nv;
                    x[i * D + dd] = xi + nv;
                }
            }
            orc_check_constraints(p, &x[i * D], &v[i * D]);
        }
        /* :679-690 */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < num_p; ++i) {
            double fx = orc_cal_cost(m, o, &x[i * D]);
            if (fx < pcost[i]) {
                pcost[i] = fx;
                memcpy(&pb[i * D], &x[i * D], D * sizeof(double));
            }
        }
        n_evals += num_p;
        /* :693-704 */
        fmin_id = 0;
        for (int i = 1; i < num_p; ++i) if (pcost[i] < pcost[fmin_id]) fmin_id = i;
        if (pcost[fmin_id] < gbest_cost) {
            memcpy(gbest, &x[fmin_id * D], D * sizeof(double));
            gbest_cost = pcost[fmin_id];
        }
    }
    memcpy(bestp, gbest, D * sizeof(double));
    if (best_cost) *best_cost = gbest_cost;
    if (evals) *evals = n_evals;
    if (final_particles) memcpy(final_particles, x, sizeof(double) * num_p * D);
    if (final_pcost) memcpy(final_pcost, pcost, sizeof(double) * num_p);
    free(x); free(v); free(pb); free(pcost);
    return 1;
}

/* ------------------------------------------------------------------ */
/* costfunc::bincomp_penalty (silhouette mismatch)   */
/* ------------------------------------------------------------------ */

/* cv::circle(img, centre, radius, colour, thickness = -1) with line_type 8 and shift 0, OpenCV
 * imgproc drawing.cpp `Circle(..., fill = 1)` (third-party, not under /root/reference; call site
 * src/costfunc.cpp:423): the midpoint circle walk -- dy grows from 0, dx shrinks from the radius
 * while dx >= dy -- and for every step the four horizontal runs [cx-dx, cx+dx] on rows cy-+dy and
 * [cx-dy, cx+dy] on rows cy-+dx, clipped to the image.  mask[y*w + x] = 1 for every covered pixel.
 * Pinned bit for bit against cv2.circle (4.13) in tests/test_oracle_pins.py. */
void orc_fill_circle(uint8_t *mask, int w, int h, int cx, int cy, int radius)
{
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        const int ys[4] = {cy - dy, cy + dy, cy - dx, cy + dx};
        const int hw[4] = {dx, dx, dy, dy};
        for (int k = 0; k < 4; ++k) {
            if (ys[k] < 0 || ys[k] >= h) continue;
            int xa = cx - hw[k], xb = cx + hw[k];
            if (xa < 0) xa = 0;
            if (xb > w - 1) xb = w - 1;
            for (int x = xa; x <= xb; ++x) mask[(long)ys[k] * w + x] = 1;
        }
        dy++;
        err += plus;
        plus += 2;
        if (err > 0) {            /* mask = (err <= 0) - 1 */
            err -= minus;
            dx -= 1;
            minus -= 2;
        }
    }
}

/* double -> int as the reference's implicit conversions do it (truncation); values no image can
 * hold are clamped so that the circle falls outside instead of invoking the undefined conversion */
static int trunc_to_int(double v)
{
    if (!(v > -1048576.0)) return -1048576;      /* also NaN */
    if (v > 1048576.0) return 1048576;
    return (int)v;
}

/* costfunc::bincomp_penalty, src/costfunc.cpp:384-436: the 48 spheres are drawn as filled discs
 * (centre = trunc(K c / c_z), radius = trunc(r / scale)) into an empty 240x320 image; the penalty
 * is the number of pixels where that silhouette and the observed one (depth != 0) differ.
 * spheres: 48x3 as build_hand_model returns them (negated y, z; un-negated at :401).
 * silhouette (optional, H*W) receives the drawn mask. */
double orc_bincomp_penalty(const orc_model *m, const orc_obs *o, const double *spheres, uint8_t *silhouette)
{
    uint8_t *mask = (uint8_t *)calloc((size_t)ORC_W * ORC_H, 1);
    for (int i = 0; i < ORC_NSPH; ++i) {
        double x = spheres[i * 3], y = -spheres[i * 3 + 1], z = -spheres[i * 3 + 2];
        double pu = (o->focal * x + o->cx * z) / z;
        double pv = (o->focal * y + o->cy * z) / z;
        int rad = trunc_to_int(m->rad[i] / o->scale);
        if (rad < 0) continue;                    /* cv::circle asserts radius >= 0 */
        orc_fill_circle(mask, ORC_W, ORC_H, trunc_to_int(pu), trunc_to_int(pv), rad);
    }
    long c = 0;
    for (long p = 0; p < (long)ORC_W * ORC_H; ++p) c += (mask[p] != 0) != (o->depth[p] != 0.0);
    if (silhouette) memcpy(silhouette, mask, (size_t)ORC_W * ORC_H);
    free(mask);
    return (double)c;
}

/* cv::circle(img, (cx, cy), radius, colour, thickness >= 2): OpenCV's outlined circle (oracle/orc_draw.h), as a mask;
 * pinned bit for bit against cv2.circle (4.13) in tests/test_oracle_pins.py */
typedef struct { uint8_t *mask; int w; } mask_ctx;
static void mask_run2(void *ctx, int y, int x1, int x2)
{
    mask_ctx *c = (mask_ctx *)ctx;
    for (int x = x1; x <= x2; ++x) c->mask[(long)y * c->w + x] = 1;
}
void orc_circle_outline(uint8_t *mask, int w, int h, int cx, int cy, int radius, int thickness)
{
    mask_ctx c = {mask, w};
    orcd_target t = {w, h, &c, mask_run2};
    orcd_circle_outline(&t, cx, cy, radius, thickness);
}

/* costfunc::depthMatch_penalty, src/costfunc.cpp:439-473: the observed depth map (cm) with the outline of every
 * projected sphere drawn into it -- centre = trunc(K c / c_z), radius = trunc(r / scale), cv::circle(..., thickness 2),
 * value = the sphere's depth c_z -- in sphere order (a later outline overwrites an earlier one).
 * spheres: 48x3 as build_hand_model returns them (negated y, z; un-negated at :449).  out: H*W doubles. */
typedef struct { double *img; double value; } dm_ctx;
static void dm_run(void *ctx, int y, int x1, int x2)
{
    dm_ctx *c = (dm_ctx *)ctx;
    for (int x = x1; x <= x2; ++x) c->img[(long)y * ORC_W + x] = c->value;
}
void orc_depthmatch_penalty(const orc_model *m, const orc_obs *o, const double *spheres, double *out)
{
    memcpy(out, o->depth, sizeof(double) * ORC_W * ORC_H);
    for (int i = 0; i < ORC_NSPH; ++i) {
        const double x = spheres[i * 3], y = -spheres[i * 3 + 1], z = -spheres[i * 3 + 2];
        const double pu = (o->focal * x + o->cx * z) / z;
        const double pv = (o->focal * y + o->cy * z) / z;
        const int rad = trunc_to_int(m->rad[i] / o->scale);
        if (rad < 0) continue;                    /* cv::circle asserts radius >= 0 */
        dm_ctx c = {out, z};
        orcd_target t = {ORC_W, ORC_H, &c, dm_run};
        orcd_circle_outline(&t, trunc_to_int(pu), trunc_to_int(pv), rad, 2);
    }
}

/* PSO::dim_restore, src/PSO.cpp:160-180: 22 -> 26 parameters with DIP = 2/3 PIP */
void orc_dim_restore(const double in[22], double out[ORC_NDOF])
{
    for (int i = 0; i <= 9; ++i) out[i] = in[i];
    for (int f = 0; f < 4; ++f) {
        for (int k = 0; k < 3; ++k) out[10 + 4 * f + k] = in[10 + 3 * f + k];
        out[13 + 4 * f] = 2. / 3 * in[12 + 3 * f];
    }
}

/* costfunc::gnd_truth_err, src/costfunc.cpp:476-507: joints (21 x 3, cm, as build_hand_model left
 * them) x 10 -> mm, y and z negated; gt_row = one joint.txt row (63 values: x,y,z per joint, mm);
 * sum of the distances of the wrist and the five finger tips (rows 0,4,8,12,16,20) */
double orc_gnd_truth_err(const double joints[ORC_NJOINT * 3], const double gt_row[63])
{
    static const int sj[6] = {0, 4, 8, 12, 16, 20};
    double dist[6];
    for (int q = 0; q < 6; ++q) {
        int j = sj[q];
        double hx = joints[j * 3] * 10.0, hy = joints[j * 3 + 1] * 10.0 * -1, hz = joints[j * 3 + 2] * 10.0 * -1;
        double dx = gt_row[j * 3] - hx, dy = gt_row[j * 3 + 1] - hy, dz = gt_row[j * 3 + 2] - hz;
        dist[q] = sqrt(dx * dx + dy * dy + dz * dz);
    }
    return arma_accumulate(dist, 6);
}

/* ------------------------------------------------------------------ */
/* Synthetic frames: ray-cast the 48 spheres.  Not a   */
/* reference function.  Pixel (u,v) looks along ((u-160)/f,(v-120)/f,1).*/
/* ------------------------------------------------------------------ */
void orc_render_depth(const double *sp, const double *rad, double focal, float *depth_mm)
{
#pragma omp parallel for schedule(static)
    for (int v = 0; v < ORC_H; ++v)
        for (int u = 0; u < ORC_W; ++u) {
            double dx = (u - ORC_W / 2.) / focal, dy = (v - ORC_H / 2.) / focal;
            double a = dx * dx + dy * dy + 1.0;
            double best = INFINITY;
            for (int j = 0; j < ORC_NSPH; ++j) {
                const double *c = &sp[j * 3];
                double b = dx * c[0] + dy * c[1] + c[2];
                double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2] - rad[j] * rad[j];
                double disc = b * b - a * cc;
                if (disc < 0) continue;
                double t = (b - sqrt(disc)) / a;
                if (t > 0 && t < best) best = t;
            }
            depth_mm[v * ORC_W + u] = isfinite(best) ? (float)(best * 10.0) : 0.0f;
        }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
