// hpe_eval.cuh -- the warp-level evaluation of costfunc::cal_cost / cal_cost2
// (src/costfunc.cpp:31-127) for a group of up to HPE_GROUP poses against one frame.
//
// Work split inside a warp:
//   phase 1  lanes = (pose, digit): forward kinematics in FP64 -> sphere centres in shared
//            memory (float4 for the matcher, double for the FP64 terms)
//   phase 2  per pose, lanes = spheres: projection + depth-map / distance-transform gathers
//   phase 3  per pose, lanes = points (8 per lane, registers): 48-sphere float32 sweep with
//            OpenCV's operation order, then the FP64 alignment sum
//   phase 4  (cal_cost2 only) lanes = sphere pairs: self-collision
// The model (DevModel) and the radii are read from shared memory: lanes index them with
// different digits/spheres, which would serialise on the constant bank.
#pragma once
#include "hpe_device.cuh"

// per-warp shared-memory scratch for a group of poses (sized by the launch: slots * 1920 B)
struct WarpScratch {
    float4 *sph4;   // [slot][48]    (x,-y,-z,0): the float32 centres cv::BFMatcher sees
    double *sphd;   // [slot][48][3] un-negated FP64 centres
};
#define HPE_WS_BYTES_PER_SLOT (HPE_NSPH * 16 + HPE_NSPH * 24)

__device__ __forceinline__ WarpScratch warp_scratch(unsigned char *base, int warp, int slots)
{
    // layout: all float4 blocks first (16-byte aligned), then the doubles
    WarpScratch w;
    unsigned char *p = base + (size_t)warp * slots * HPE_WS_BYTES_PER_SLOT;
    w.sph4 = reinterpret_cast<float4 *>(p);
    w.sphd = reinterpret_cast<double *>(p + (size_t)slots * HPE_NSPH * 16);
    return w;
}

// the lane's 8 observed points, float32 (conv_to<fmat>, src/costfunc.cpp:322)
struct LanePoints {
    float x[HPE_PPL], y[HPE_PPL], z[HPE_PPL];
};

__device__ __forceinline__ void load_lane_points(const FrameView &F, int lane, LanePoints &P)
{
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        int k = lane + 32 * m;
        bool ok = k < F.n_pts;
        P.x[m] = ok ? (float)F.cloud[k] : 0.f;
        P.y[m] = ok ? (float)F.cloud[F.n_pts + k] : 0.f;
        P.z[m] = ok ? (float)F.cloud[2 * F.n_pts + k] : 0.f;
    }
}

// Exact restatement of the matcher for ONE point (slow path, used when the fast sweep
// cannot prove its winner): first minimum of sqrtf(d) over the 48 spheres.
static __device__ __noinline__ int match_point_exact(const float4 *sph, float px, float py, float pz)
{
    float best = __int_as_float(0x7f800000);
    int bi = 0;
    for (int j = 0; j < HPE_NSPH; ++j) {
        float4 c = sph[j];
        float t0 = __fsub_rn(px, c.x), t1 = __fsub_rn(py, c.y), t2 = __fsub_rn(pz, c.z);
        float d = __fmul_rn(t0, t0);
        d = __fadd_rn(d, __fmul_rn(t1, t1));
        d = __fadd_rn(d, __fmul_rn(t2, t2));
        float w = __fsqrt_rn(d);
        if (w < best) { best = w; bi = j; }
    }
    return bi;
}

// phase 3a: nearest-centre index for the lane's 8 points (compute_correspondences,
// src/costfunc.cpp:306-343).  Fast sweep compares squared distances and tracks the
// runner-up; sqrtf can only merge values that are within a few ulp, so whenever
// sqrtf(best) == sqrtf(second) the point is re-done with the exact restatement.
__device__ __forceinline__ void match_sweep(const float4 *sph, const LanePoints &P, int n_pts, int lane,
                                            int idx[HPE_PPL])
{
    float b1[HPE_PPL], b2[HPE_PPL];
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        b1[m] = __int_as_float(0x7f800000);
        b2[m] = __int_as_float(0x7f800000);
        idx[m] = 0;
    }
#pragma unroll 4
    for (int j = 0; j < HPE_NSPH; ++j) {
        float4 c = sph[j];
#pragma unroll
        for (int m = 0; m < HPE_PPL; ++m) {
            float t0 = __fsub_rn(P.x[m], c.x), t1 = __fsub_rn(P.y[m], c.y), t2 = __fsub_rn(P.z[m], c.z);
            float d = __fmul_rn(t0, t0);
            d = __fadd_rn(d, __fmul_rn(t1, t1));
            d = __fadd_rn(d, __fmul_rn(t2, t2));
            b2[m] = fminf(b2[m], fmaxf(b1[m], d));
            if (d < b1[m]) { b1[m] = d; idx[m] = j; }
        }
    }
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        if (lane + 32 * m < n_pts && __fsqrt_rn(b1[m]) == __fsqrt_rn(b2[m])) idx[m] = match_point_exact(sph, P.x[m], P.y[m], P.z[m]);
    }
}

// phase 3b: align_models (src/costfunc.cpp:346-377) for the lane's points, FP64.
// cloud_s: shared-memory SoA copy of the frame's cloud (X | -Y | -Z), stride n_pts.
__device__ __forceinline__ double align_lane(const double *rad_s, const double *sphd,
                                             const double *cloud_s, int n_pts, int lane,
                                             const int idx[HPE_PPL])
{
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        int k = lane + 32 * m;
        if (k < n_pts) {
            int j = idx[m];
            // negated-frame centre = (cx, -cy, -cz)
            double a = cloud_s[k] - sphd[j * 3 + 0];
            double b = cloud_s[n_pts + k] + sphd[j * 3 + 1];
            double e = cloud_s[2 * n_pts + k] + sphd[j * 3 + 2];
            double nor = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(e, e)));
            double df = fabs(nor - rad_s[j]);
            acc += __dmul_rn(df, df);
        }
    }
    return acc;
}

// phase 4: self_collision_penalty (src/costfunc.cpp:130-197): 4 adjacent digit pairs x 36
__device__ __forceinline__ double collision_lane(const double *rad_s, const double *sphd, int lane)
{
    double acc = 0.0;
    for (int q = lane; q < 144; q += 32) {
        int pair = q / 36, r = q % 36, a = r / 6, b = r % 6;
        int ia = 2 + 10 * pair + a, ib = 12 + 10 * pair + b;
        double d0 = sphd[ib * 3] - sphd[ia * 3], d1 = sphd[ib * 3 + 1] - sphd[ia * 3 + 1], d2 = sphd[ib * 3 + 2] - sphd[ia * 3 + 2];
        double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
        dist = rad_s[ib] + rad_s[ia] - dist;
        if (dist > 0) acc += __dmul_rn(dist, dist);
    }
    return acc;
}

// phase 1 for up to HPE_GROUP poses; theta(g) returns a pointer to pose g's 26 doubles
template <class ThetaAt>
__device__ __forceinline__ void group_fk(const DevModel &M, ThetaAt theta, int n, const WarpScratch &ws, int lane)
{
    int g = lane / 5, d = lane - 5 * g;
    if (g < n) {
        double J[5][3], S[10][3];
        fk_digit(M, d, theta(g), J);
        int cnt = digit_spheres(d, J, S);
        int first = digit_first_sphere(d);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            if (i < cnt) {
                double *sd = ws.sphd + ((size_t)g * HPE_NSPH + first + i) * 3;
                sd[0] = S[i][0];
                sd[1] = S[i][1];
                sd[2] = S[i][2];
                ws.sph4[g * HPE_NSPH + first + i] =
                    make_float4((float)S[i][0], (float)(-S[i][1]), (float)(-S[i][2]), 0.f);
            }
        }
    }
    __syncwarp();
}

// phase 2: depth_penalty (src/costfunc.cpp:227-304) of pose g, lanes = spheres.  Valid on all lanes.
__device__ __forceinline__ double depth_lanes(const double *rad_s, const FrameView &F, const WarpScratch &ws,
                                              int g, int lane)
{
    const double *sd = ws.sphd + (size_t)g * HPE_NSPH * 3;
    double dv = depth_term(F, sd[lane * 3], sd[lane * 3 + 1], sd[lane * 3 + 2], rad_s[lane]);
    if (lane < HPE_NSPH - 32)
        dv += depth_term(F, sd[(lane + 32) * 3], sd[(lane + 32) * 3 + 1], sd[(lane + 32) * 3 + 2], rad_s[lane + 32]);
    return warp_sum(dv);
}

// phase 3b reduced over the warp and scaled by lambda = |M|/|P| (src/costfunc.cpp:372-373)
__device__ __forceinline__ double align_lanes(const double *rad_s, const WarpScratch &ws, int g,
                                              const double *cloud_s, int n_pts, int lane, const int idx[HPE_PPL])
{
    double a = align_lane(rad_s, ws.sphd + (size_t)g * HPE_NSPH * 3, cloud_s, n_pts, lane, idx);
    return __dmul_rn(warp_sum(a), (double)HPE_NSPH / (double)n_pts);
}
