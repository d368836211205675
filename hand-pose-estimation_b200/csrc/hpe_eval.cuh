// hpe_eval.cuh -- the warp-level evaluation of costfunc::cal_cost / cal_cost2
// (src/costfunc.cpp:31-127) for a group of up to HPE_GROUP poses against one frame.
//
// Work split inside a warp:
//   phase 1  lanes = (pose, digit): forward kinematics in FP64 -> FP64 sphere centres in shared memory
//   phase 2  per pose, lanes = spheres: projection + depth-map / distance-transform gathers; the same
//            lanes stage the pose's float32 centres for the matcher
//   phase 3  per pose, lanes = points (8 per lane, registers): nearest-centre search in FP32, then the
//            FP64 alignment sum
//   phase 4  (cal_cost2 only) lanes = sphere pairs: self-collision
// The model (DevModel) and the radii are read from shared memory: lanes index them with
// different digits/spheres, which would serialise on the constant bank.
//
// The nearest-centre search (compute_correspondences, src/costfunc.cpp:306-343) must return the
// index cv::BFMatcher returns: first minimum over j of sqrtf(d_j), d_j = ((t0^2 + t1^2) + t2^2) in
// float32, t = p - c_j.  Doing that directly costs 8 FP32-pipe + 5 ALU-pipe instructions per
// (point, sphere) pair.  match_sweep() gets the same index with ~5.7:
//   sweep A  s_j = |c'_j|^2 - 2 p'.c'_j (3 FFMA, coordinates recentred on the cloud centroid) for
//            all 48 spheres, reduced to the minimum per block of 4 spheres, keeping the best
//            block, its value M1 and the runner-up block value M2;
//   certify  s_j + |p'|^2 equals the exact squared distance D_j up to E_j <= u(104|p'|^2 + 13|s_j|),
//            u = 2^-24 (forward error of the 3 FFMAs plus the rounding of p', -2c', |c'|^2), and the
//            reference's d_j, sqrtf(d_j) order two spheres like D_j once D differs by > 24u D.
//            Hence if M2 - M1 > tau = u(260|p'|^2 + 52|M1|) no sphere outside the best block can be
//            the reference's answer;
//   sweep B  the 4 spheres of the best block once more (same FFMAs, bit-identical s): winner index
//            and the runner-up r inside the block; the certificate becomes min(M2, r) - M1 > tau, i.e.
//            EVERY other sphere is provably farther in the reference's own arithmetic (ties and
//            sqrtf merges included, they need differences below 24u D);
//   fallback if the certificate fails the point is redone with the exact restatement of the
//            reference's arithmetic over all 48 spheres (match_point_exact).
#pragma once
#include "hpe_device.cuh"

// diagnostics: points sent to the exact fallback / points matched (hpe_debug_counters)
// (one copy per translation unit; the host getters below are instantiated next to the kernels)
static __device__ unsigned long long g_hpe_fallback_points;
#define HPE_DEFINE_FALLBACK_GETTER(name)                                                          \
    unsigned long long name(int reset)                                                            \
    {                                                                                             \
        unsigned long long v = 0, z = 0;                                                          \
        cudaMemcpyFromSymbol(&v, g_hpe_fallback_points, sizeof v);                                \
        if (reset) cudaMemcpyToSymbol(g_hpe_fallback_points, &z, sizeof z);                       \
        return v;                                                                                 \
    }

#define HPE_BLK 4                        // spheres per sweep-A block
#define HPE_NBLK (HPE_NSPH / HPE_BLK)    // 12 blocks

// per-warp shared-memory scratch
struct WarpScratch {
    double *sphd;   // [slot][48][3] un-negated FP64 centres of the group's poses
    float4 *cf;     // [48] (x,-y,-z,0): the float32 centres cv::BFMatcher sees, pose being matched
    float4 *ex;     // [48] (-2c'x, -2c'y, -2c'z, |c'|^2), c' = cf - q, pose being matched
};
#define HPE_WS_BYTES(slots) ((slots) * HPE_NSPH * 24 + 2 * HPE_NSPH * 16)

__device__ __forceinline__ WarpScratch warp_scratch(unsigned char *base, int warp, int slots)
{
    WarpScratch w;
    unsigned char *p = base + (size_t)warp * HPE_WS_BYTES(slots);
    w.cf = reinterpret_cast<float4 *>(p);
    w.ex = w.cf + HPE_NSPH;
    w.sphd = reinterpret_cast<double *>(p + 2 * HPE_NSPH * 16);
    return w;
}

// model + frame data every CTA stages once
struct __align__(16) FrameSmem {
    DevModel M;
    double rad[HPE_NSPH];
    double cloud[3 * HPE_MAXPTS];    // FP64 cloud, SoA with stride n_pts (X | -Y | -Z)
    float cloudf[3 * HPE_MAXPTS];    // its float32 conversion, SoA with stride HPE_MAXPTS
};

// all threads of the CTA; caller syncs afterwards
__device__ __forceinline__ void stage_model(FrameSmem &fs, const DevModel &m)
{
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 8); i += blockDim.x)
        reinterpret_cast<double *>(&fs.M)[i] = reinterpret_cast<const double *>(&m)[i];
    if (threadIdx.x < HPE_NSPH) fs.rad[threadIdx.x] = m.rad[threadIdx.x];
}

__device__ __forceinline__ void stage_frame(FrameSmem &fs, const FrameView &F)
{
    for (int i = threadIdx.x; i < 3 * HPE_MAXPTS; i += blockDim.x) {
        int c = i / HPE_MAXPTS, k = i - c * HPE_MAXPTS;
        double v = k < F.n_pts ? F.cloud[c * F.n_pts + k] : 0.0;
        if (k < F.n_pts) fs.cloud[c * F.n_pts + k] = v;
        fs.cloudf[i] = (float)v;
    }
}

// the lane's 8 observed points: float32 coordinates recentred on q (the cloud centroid) and
// the certificate scale 260 u |p'|^2
struct LanePoints {
    float x[HPE_PPL], y[HPE_PPL], z[HPE_PPL], tau0[HPE_PPL];
    float qx, qy, qz;
};

// cloudf_s: shared-memory float32 copy of the cloud (conv_to<fmat>, src/costfunc.cpp:322), SoA
// with stride HPE_MAXPTS.  Must be called by all 32 lanes.
__device__ __forceinline__ void load_lane_points(const float *cloudf_s, int n_pts, int lane, LanePoints &P)
{
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        int k = lane + 32 * m;
        bool ok = k < n_pts;
        P.x[m] = ok ? cloudf_s[k] : 0.f;
        P.y[m] = ok ? cloudf_s[HPE_MAXPTS + k] : 0.f;
        P.z[m] = ok ? cloudf_s[2 * HPE_MAXPTS + k] : 0.f;
        sx += P.x[m]; sy += P.y[m]; sz += P.z[m];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    P.qx = sx / (float)n_pts; P.qy = sy / (float)n_pts; P.qz = sz / (float)n_pts;
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        P.x[m] = __fsub_rn(P.x[m], P.qx);
        P.y[m] = __fsub_rn(P.y[m], P.qy);
        P.z[m] = __fsub_rn(P.z[m], P.qz);
        float pp = P.x[m] * P.x[m] + P.y[m] * P.y[m] + P.z[m] * P.z[m];
        P.tau0[m] = 1.56e-5f * pp;      // 260 * 2^-24 = 1.5497e-5, rounded up
    }
}

// Exact restatement of the matcher for ONE point: first minimum of sqrtf(d) over the 48 spheres.
static __device__ __noinline__ int match_point_exact(const float4 *cf, float px, float py, float pz)
{
    float best = __int_as_float(0x7f800000);
    int bi = 0;
    for (int j = 0; j < HPE_NSPH; ++j) {
        float4 c = cf[j];
        float t0 = __fsub_rn(px, c.x), t1 = __fsub_rn(py, c.y), t2 = __fsub_rn(pz, c.z);
        float d = __fmul_rn(t0, t0);
        d = __fadd_rn(d, __fmul_rn(t1, t1));
        d = __fadd_rn(d, __fmul_rn(t2, t2));
        float w = __fsqrt_rn(d);
        if (w < best) { best = w; bi = j; }
    }
    return bi;
}

// The same for one point with the whole warp: lane handles spheres lane and lane+32, then a
// lexicographic (sqrtf(d), index) minimum over the warp = first minimum.  All 32 lanes call it.
static __device__ __noinline__ int match_point_exact_warp(const float4 *cf, float px, float py, float pz, int lane)
{
    float w = __int_as_float(0x7f800000);
    int bi = lane;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int j = lane + 32 * h;
        if (j < HPE_NSPH) {
            float4 c = cf[j];
            float t0 = __fsub_rn(px, c.x), t1 = __fsub_rn(py, c.y), t2 = __fsub_rn(pz, c.z);
            float d = __fmul_rn(t0, t0);
            d = __fadd_rn(d, __fmul_rn(t1, t1));
            d = __fadd_rn(d, __fmul_rn(t2, t2));
            float wj = __fsqrt_rn(d);
            if (wj < w) { w = wj; bi = j; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ow = __shfl_xor_sync(0xffffffffu, w, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ow < w || (ow == w && oi < bi)) { w = ow; bi = oi; }
    }
    return bi;
}

__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ unsigned long long pack_indices(const int idx[HPE_PPL]);

// Direct form over all 48 spheres (the reference's arithmetic on every pair): kept as the
// cross-check of match_sweep (hpe_set_matcher(1)).
__device__ __forceinline__ unsigned long long match_sweep_direct(const WarpScratch &ws, const float *cloudf_s,
                                                                 int n_pts, int lane)
{
    int idx[HPE_PPL];
    float px[HPE_PPL], py[HPE_PPL], pz[HPE_PPL], b1[HPE_PPL], b2[HPE_PPL];
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        int k = min(lane + 32 * m, HPE_MAXPTS - 1);
        px[m] = cloudf_s[k]; py[m] = cloudf_s[HPE_MAXPTS + k]; pz[m] = cloudf_s[2 * HPE_MAXPTS + k];
        b1[m] = __int_as_float(0x7f800000);
        b2[m] = __int_as_float(0x7f800000);
        idx[m] = 0;
    }
#pragma unroll 4
    for (int j = 0; j < HPE_NSPH; ++j) {
        float4 c = ws.cf[j];
#pragma unroll
        for (int m = 0; m < HPE_PPL; ++m) {
            float t0 = __fsub_rn(px[m], c.x), t1 = __fsub_rn(py[m], c.y), t2 = __fsub_rn(pz[m], c.z);
            float d = __fmul_rn(t0, t0);
            d = __fadd_rn(d, __fmul_rn(t1, t1));
            d = __fadd_rn(d, __fmul_rn(t2, t2));
            b2[m] = fminf(b2[m], fmaxf(b1[m], d));
            if (d < b1[m]) { b1[m] = d; idx[m] = j; }
        }
    }
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m)
        if (lane + 32 * m < n_pts && __fsqrt_rn(b1[m]) == __fsqrt_rn(b2[m]))
            idx[m] = match_point_exact(ws.cf, px[m], py[m], pz[m]);
    return pack_indices(idx);
}

// phase 3a (see the header comment).  ws.cf / ws.ex must hold the pose being matched.
// Returns the 8 indices packed 8 bits each (point m of the lane in bits [8m, 8m+8)).
__device__ __forceinline__ unsigned long long match_sweep(const WarpScratch &ws, const LanePoints &P,
                                                          const float *cloudf_s, int n_pts, int lane)
{
    const float INF = __int_as_float(0x7f800000);
    float M1[HPE_PPL], M2[HPE_PPL];
    int g1[HPE_PPL];
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) { M1[m] = INF; M2[m] = INF; g1[m] = 0; }

    // ---- sweep A: block minima of s_j over all 48 spheres ----
#pragma unroll 2
    for (int g = 0; g < HPE_NBLK; ++g) {
        const float4 e0 = ws.ex[g * HPE_BLK + 0], e1 = ws.ex[g * HPE_BLK + 1];
        const float4 e2 = ws.ex[g * HPE_BLK + 2], e3 = ws.ex[g * HPE_BLK + 3];
#pragma unroll
        for (int m = 0; m < HPE_PPL; ++m) {
            float s0 = fmaf(P.z[m], e0.z, fmaf(P.y[m], e0.y, fmaf(P.x[m], e0.x, e0.w)));
            float s1 = fmaf(P.z[m], e1.z, fmaf(P.y[m], e1.y, fmaf(P.x[m], e1.x, e1.w)));
            float s2 = fmaf(P.z[m], e2.z, fmaf(P.y[m], e2.y, fmaf(P.x[m], e2.x, e2.w)));
            float s3 = fmaf(P.z[m], e3.z, fmaf(P.y[m], e3.y, fmaf(P.x[m], e3.x, e3.w)));
            float gm = fmin3(fminf(s0, s1), s2, s3);
            M2[m] = fminf(M2[m], fmaxf(M1[m], gm));
            bool better = gm < M1[m];
            M1[m] = fminf(M1[m], gm);
            g1[m] = better ? g : g1[m];
        }
    }

    // ---- sweep B: the 4 spheres of the best block again (bit-identical s), winner + runner-up,
    //      certificate against every other sphere ----
    unsigned long long packed = 0ull;
    unsigned slow = 0u;
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) {
        const float4 *blk = ws.ex + g1[m] * HPE_BLK;
        const float4 e0 = blk[0], e1 = blk[1], e2 = blk[2], e3 = blk[3];
        float s0 = fmaf(P.z[m], e0.z, fmaf(P.y[m], e0.y, fmaf(P.x[m], e0.x, e0.w)));
        float s1 = fmaf(P.z[m], e1.z, fmaf(P.y[m], e1.y, fmaf(P.x[m], e1.x, e1.w)));
        float s2 = fmaf(P.z[m], e2.z, fmaf(P.y[m], e2.y, fmaf(P.x[m], e2.x, e2.w)));
        float s3 = fmaf(P.z[m], e3.z, fmaf(P.y[m], e3.y, fmaf(P.x[m], e3.x, e3.w)));
        float lo01 = fminf(s0, s1), hi01 = fmaxf(s0, s1), lo23 = fminf(s2, s3), hi23 = fmaxf(s2, s3);
        float r = fmin3(fmaxf(lo01, lo23), hi01, hi23);                   // runner-up inside the block
        const float tau = fmaf(3.2e-6f, fabsf(M1[m]), P.tau0[m]);         // 52 * 2^-24 = 3.0994e-6
        bool cert = fminf(M2[m], r) - M1[m] > tau;                        // false on NaN
        int i = s1 < s0 ? 1 : 0, j = s3 < s2 ? 3 : 2;
        int bi = lo23 < lo01 ? j : i;
        packed |= (unsigned long long)(g1[m] * HPE_BLK + bi) << (8 * m);
        slow |= (cert ? 0u : 1u) << m;
    }
    // ---- fallback: the reference's arithmetic on all 48 spheres, one flagged point at a time with
    //      the whole warp on it (lanes = spheres) ----
    if (__any_sync(0xffffffffu, slow != 0u)) {
        for (int m = 0; m < HPE_PPL; ++m) {
            unsigned todo = __ballot_sync(0xffffffffu, ((slow >> m) & 1u) && lane + 32 * m < n_pts);
            if (todo && lane == 0) atomicAdd(&g_hpe_fallback_points, (unsigned long long)__popc(todo));
            while (todo) {
                int src = __ffs(todo) - 1;
                todo &= todo - 1;
                int k = src + 32 * m;
                int e = match_point_exact_warp(ws.cf, cloudf_s[k], cloudf_s[HPE_MAXPTS + k],
                                               cloudf_s[2 * HPE_MAXPTS + k], lane);
                if (lane == src) packed = (packed & ~(0xffull << (8 * m))) | ((unsigned long long)e << (8 * m));
            }
        }
    }
    return packed;
}

__device__ __forceinline__ unsigned long long pack_indices(const int idx[HPE_PPL])
{
    unsigned long long packed = 0ull;
#pragma unroll
    for (int m = 0; m < HPE_PPL; ++m) packed |= (unsigned long long)(idx[m] & 255) << (8 * m);
    return packed;
}

// phase 3b: align_models (src/costfunc.cpp:346-377) for the lane's points, FP64.
// cloud_s: shared-memory SoA copy of the frame's cloud (X | -Y | -Z), stride n_pts.
__device__ __forceinline__ double align_lane(const double *rad_s, const double *sphd,
                                             const double *cloud_s, int n_pts, int lane,
                                             unsigned long long packed)
{
    double acc = 0.0;
#pragma unroll 2
    for (int m = 0; m < HPE_PPL; ++m) {
        int k = lane + 32 * m;
        if (k < n_pts) {
            int j = (int)((packed >> (8 * m)) & 255ull);
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
    __syncwarp();
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
            }
        }
    }
    __syncwarp();
}

// phase 2: depth_penalty (src/costfunc.cpp:227-304) of pose g, lanes = spheres; the same lanes
// stage the pose's float32 centres (cf) and sweep-A operands (ex).  Result valid on all lanes.
__device__ __forceinline__ double depth_and_stage(const double *rad_s, const FrameView &F, const WarpScratch &ws,
                                                  const LanePoints &P, int g, int lane)
{
    const double *sd = ws.sphd + (size_t)g * HPE_NSPH * 3;
    double dv = 0.0;
    __syncwarp();      // previous pose's matcher is done with cf/ex
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        int j = lane + 32 * h;
        if (j < HPE_NSPH) {
            double x = sd[j * 3], y = sd[j * 3 + 1], z = sd[j * 3 + 2];
            dv += depth_term(F, x, y, z, rad_s[j]);
            float4 c = make_float4((float)x, (float)(-y), (float)(-z), 0.f);   // conv_to<fmat>, negated frame
            // c' = cf - q exactly (FP64), then one rounding each for -2c' and |c'|^2
            double cx = (double)c.x - (double)P.qx, cy = (double)c.y - (double)P.qy, cz = (double)c.z - (double)P.qz;
            ws.cf[j] = c;
            ws.ex[j] = make_float4((float)(-2.0 * cx), (float)(-2.0 * cy), (float)(-2.0 * cz),
                                   (float)(cx * cx + cy * cy + cz * cz));
        }
    }
    __syncwarp();
    return warp_sum(dv);
}

// phase 3b reduced over the warp and scaled by lambda = |M|/|P| (src/costfunc.cpp:372-373)
__device__ __forceinline__ double align_lanes(const double *rad_s, const WarpScratch &ws, int g,
                                              const double *cloud_s, int n_pts, int lane, unsigned long long packed)
{
    double a = align_lane(rad_s, ws.sphd + (size_t)g * HPE_NSPH * 3, cloud_s, n_pts, lane, packed);
    return __dmul_rn(warp_sum(a), (double)HPE_NSPH / (double)n_pts);
}
