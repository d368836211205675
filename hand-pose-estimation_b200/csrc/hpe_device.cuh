// hpe_device.cuh -- device-side building blocks shared by the kernels.
//
// Precision policy ("reference precision"): the reference computes everything in double
// except the point->sphere matcher, which runs in float32 inside cv::BFMatcher
// (src/costfunc.cpp:321-334).  We keep exactly that split: forward kinematics,
// projection/depth term and the alignment sum in FP64 (a few % of the instructions), the
// 250x48 distance sweep in FP32 with OpenCV's operation order, so match indices are
// bit-identical and costs agree with the reference arithmetic to ~1e-13 relative.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HPE_NDOF 26
#define HPE_NSPH 48
#define HPE_NJOINT 21
#define HPE_W 320
#define HPE_H 240
#ifndef HPE_GROUP
#define HPE_GROUP 5
#endif                       // poses evaluated together by one warp (5 poses x 5 digits = 25 FK lanes);
                             // 6 would fill 30 lanes but its scratch leaves room for 2 CTAs per SM, not 3
#define HPE_MAXPTS 256       // 8 points per lane
#define HPE_PPL 8
// staging image of a frame: float4 pre[256] | float4 cbox[8] | float4 chalf[8] | double cloud[3][256] | vq[64] | perm[256] | q, n_pts
#define HPE_IMAGE_BYTES (HPE_MAXPTS * 16 + 2 * (HPE_MAXPTS / 32) * 16 + 3 * HPE_MAXPTS * 8 + HPE_MAXPTS / 4 + HPE_MAXPTS + 16)

struct DevModel {
    double base[5][12];      // pose-independent CMC link (finger T01 / thumb Trf), 3x4 row-major
    double backt[5][2];      // translation (x,y) of the reverse link T10 (its z is 0)
    double geo[20];          // segment lengths, cm
    double rad[HPE_NSPH];
    float  radf[HPE_NSPH];
    double tc, ts;           // cos/sin of the thumb's pCMC = CMC + pi (src/thumbmodel.cpp:149)
};

// One observed frame (observedmodel after next_frame), device pointers.
struct FrameView {
    const float  *depth_mm;  // [240][320]
    const float  *dt_px;     // [240][320] FOLDED map: distance transform (px) on the background, minus the depth (mm) on hand pixels
    const double *cloud;     // [3][n_pts] SoA: X, -Y, -Z
    const unsigned char *perm;  // [HPE_MAXPTS] cluster order of the cloud (cluster_kernel): sorted position -> point index
    const uint4 *image;      // staging image of the frame (FrameSmem's frame-dependent head, written by cluster_kernel)
    double scale;            // cm per pixel
    double dt_max;           // disttrans.max()
    double focal, cx, cy;
    int    n_pts;
};

// ---------------------------------------------------------------------------------------
// Forward kinematics of one digit (src/fingermodel.cpp:70-182,270-317,
// src/thumbmodel.cpp:76-195,276-318).  All lanes run the same code: the thumb's tilted
// second link degenerates to the finger's planar link with (cp,sp) = (1,0).
// ---------------------------------------------------------------------------------------
struct Aff {   // 3x4 affine map, rotation r (row-major) + translation t
    double r00, r01, r02, r10, r11, r12, r20, r21, r22, tx, ty, tz;
};

__device__ __forceinline__ double hpe_deg2rad(double d)
{
    // reference: angle/180.0*PI (src/fingermodel.cpp:203-205); one multiply differs from it
    // by <= 1 ulp of the angle
    return d * 0.017453292519943295769236907684886;
}

// cur <- cur * [[c,-s,0,L c],[s,c,0,L s],[0,0,1,0]]
__device__ __forceinline__ void aff_planar(Aff &a, double c, double s, double L)
{
    double n00 = a.r00 * c + a.r01 * s, n01 = a.r01 * c - a.r00 * s;
    double n10 = a.r10 * c + a.r11 * s, n11 = a.r11 * c - a.r10 * s;
    double n20 = a.r20 * c + a.r21 * s, n21 = a.r21 * c - a.r20 * s;
    a.r00 = n00; a.r01 = n01; a.r10 = n10; a.r11 = n11; a.r20 = n20; a.r21 = n21;
    a.tx += L * n00; a.ty += L * n10; a.tz += L * n20;
}

// sines and cosines a digit needs: sc[0..2] = global rotation about z (angle + 180 deg), y, x; sc[3..6] = its
// four joint angles; [k][0] = sin, [k][1] = cos
struct DigitTrig {
    double s[7], c[7];
};

// the angle (degrees) whose sine / cosine is DigitTrig entry k of digit d
__device__ __forceinline__ double trig_angle_deg(const double *th, int d, int k)
{
    return k == 0 ? th[0] + 180.0 : k < 3 ? th[k] : th[6 + 4 * d + (k - 3)];
}

// Global pose T00 * Rz(t0+180) * Ry(t1) * Rx(t2)  (src/fingermodel.cpp:157-180, :289)
__device__ __forceinline__ void aff_global(Aff &g, const double pos[3], const DigitTrig &t)
{
    const double sz = t.s[0], cz = t.c[0], sy = t.s[1], cy = t.c[1], sx = t.s[2], cx = t.c[2];
    // Rz*Ry = [[cz cy, -sz, cz sy],[sz cy, cz, sz sy],[-sy, 0, cy]]
    double a00 = cz * cy, a01 = -sz, a02 = cz * sy;
    double a10 = sz * cy, a11 = cz, a12 = sz * sy;
    double a20 = -sy, a22 = cy;
    // (Rz*Ry)*Rx, Rx = [[1,0,0],[0,cx,-sx],[0,sx,cx]]
    g.r00 = a00; g.r01 = a01 * cx + a02 * sx; g.r02 = a02 * cx - a01 * sx;
    g.r10 = a10; g.r11 = a11 * cx + a12 * sx; g.r12 = a12 * cx - a11 * sx;
    g.r20 = a20; g.r21 = a22 * sx;            g.r22 = a22 * cx;
    g.tx = pos[0]; g.ty = pos[1]; g.tz = pos[2];
}

// wa * a + wb * b with a fixed contraction (one multiply, one FMA), so that every code path that interpolates sphere
// centres produces the same bits whatever the compiler would have fused
__device__ __forceinline__ double mix2(double wa, double a, double wb, double b) { return __fma_rn(wa, a, __dmul_rn(wb, b)); }

// The forward-kinematics chain of digit d (0 = thumb) from the global position pos = theta[3:6]: TRIG hands out the
// sine / cosine of entry k of the digit's seven angles when the chain gets to it (aff_global takes 0..2 together, the
// twist link 3 and 4, then 5, then 6), SINK receives the joints as they become final, in the order 1, 0, 2, 3, 4.
// One body for every caller, so all paths produce the same bits.
template <class TRIG, class SINK>
__device__ __forceinline__ void fk_chain(const DevModel &M, int d, const double pos[3], TRIG &&trig, SINK &&sink)
{
    Aff cur;
    {
        DigitTrig t;
        trig(0, t.s[0], t.c[0]); trig(1, t.s[1], t.c[1]); trig(2, t.s[2], t.c[2]);
        aff_global(cur, pos, t);
    }
    // cur <- cur * base[d]   (general 3x4 product; base has r[2][*] = (0,0,1), t.z = 0)
    {
        const double *b = M.base[d];
        double c = b[0], s = b[4];          // base = planar(CMC, L): [[c,-s,0,Lc],[s,c,0,Ls]]
        double n00 = cur.r00 * c + cur.r01 * s, n01 = cur.r01 * c - cur.r00 * s;
        double n10 = cur.r10 * c + cur.r11 * s, n11 = cur.r11 * c - cur.r10 * s;
        double n20 = cur.r20 * c + cur.r21 * s, n21 = cur.r21 * c - cur.r20 * s;
        cur.tx += cur.r00 * b[3] + cur.r01 * b[7];
        cur.ty += cur.r10 * b[3] + cur.r11 * b[7];
        cur.tz += cur.r20 * b[3] + cur.r21 * b[7];
        cur.r00 = n00; cur.r01 = n01; cur.r10 = n10; cur.r11 = n11; cur.r20 = n20; cur.r21 = n21;
    }
    sink(1, cur.tx, cur.ty, cur.tz);
    {   // joint 0 = (cur * T10).col(3)
        double bx = M.backt[d][0], by = M.backt[d][1];
        sink(0, cur.r00 * bx + cur.r01 * by + cur.tx, cur.r10 * bx + cur.r11 * by + cur.ty, cur.r20 * bx + cur.r21 * by + cur.tz);
    }
    const double *L = M.geo + 4 * d;
    double cp = d == 0 ? M.tc : 1.0, sp = d == 0 ? M.ts : 0.0;
    {   // S = twist(a0) * X(a1), X = [[c2,-s2 cp, s2 sp, L c2],[s2, c2 cp, -c2 sp, L s2],[0, sp, cp, 0]]
        double s1, c1, s2, c2;
        trig(3, s1, c1);
        trig(4, s2, c2);
        double x00 = c2, x01 = -s2 * cp, x02 = s2 * sp, x03 = L[1] * c2;
        double x10 = s2, x11 = c2 * cp, x12 = -c2 * sp, x13 = L[1] * s2;
        double x21 = sp, x22 = cp;
        // twist = [[c1,0,-s1],[s1,0,c1],[0,-1,0]]
        double s00 = c1 * x00, s01 = c1 * x01 - s1 * x21, s02 = c1 * x02 - s1 * x22, s03 = c1 * x03;
        double s10 = s1 * x00, s11 = s1 * x01 + c1 * x21, s12 = s1 * x02 + c1 * x22, s13 = s1 * x03;
        double s20 = -x10, s21 = -x11, s22 = -x12, s23 = -x13;
        Aff n;
        n.r00 = cur.r00 * s00 + cur.r01 * s10 + cur.r02 * s20;
        n.r01 = cur.r00 * s01 + cur.r01 * s11 + cur.r02 * s21;
        n.r02 = cur.r00 * s02 + cur.r01 * s12 + cur.r02 * s22;
        n.tx  = cur.r00 * s03 + cur.r01 * s13 + cur.r02 * s23 + cur.tx;
        n.r10 = cur.r10 * s00 + cur.r11 * s10 + cur.r12 * s20;
        n.r11 = cur.r10 * s01 + cur.r11 * s11 + cur.r12 * s21;
        n.r12 = cur.r10 * s02 + cur.r11 * s12 + cur.r12 * s22;
        n.ty  = cur.r10 * s03 + cur.r11 * s13 + cur.r12 * s23 + cur.ty;
        n.r20 = cur.r20 * s00 + cur.r21 * s10 + cur.r22 * s20;
        n.r21 = cur.r20 * s01 + cur.r21 * s11 + cur.r22 * s21;
        n.r22 = cur.r20 * s02 + cur.r21 * s12 + cur.r22 * s22;
        n.tz  = cur.r20 * s03 + cur.r21 * s13 + cur.r22 * s23 + cur.tz;
        cur = n;
    }
    sink(2, cur.tx, cur.ty, cur.tz);
    {
        double s3, c3;
        trig(5, s3, c3);
        aff_planar(cur, c3, s3, L[2]);
    }
    sink(3, cur.tx, cur.ty, cur.tz);
    {   // last link: only the translation is needed
        double s4, c4;
        trig(6, s4, c4);
        sink(4, cur.tx + L[3] * (cur.r00 * c4 + cur.r01 * s4), cur.ty + L[3] * (cur.r10 * c4 + cur.r11 * s4),
             cur.tz + L[3] * (cur.r20 * c4 + cur.r21 * s4));
    }
}

// joints J[0..4] of digit d from precomputed sines / cosines
__device__ __forceinline__ void fk_digit_trig(const DevModel &M, int d, const double pos[3], const DigitTrig &t,
                                              double J[5][3])
{
    fk_chain(M, d, pos, [&](int k, double &sn, double &cs) { sn = t.s[k]; cs = t.c[k]; },
             [&](int i, double x, double y, double z) { J[i][0] = x; J[i][1] = y; J[i][2] = z; });
}

// sine and cosine of an angle in degrees.  One out-of-line copy for the seven (or 23) call sites of a kernel: the
// evaluation kernels are instruction-cache bound in their straight-line FP64 phases (ncu: no_instruction stalls), and
// seven inlined expansions of sincos are ~450 SASS instructions.  Measured: the out-of-line copy (HPE_SINCOS_INLINE=0)
// shrinks cost_kernel from 5,030 to 4,440 instructions but its calls cost more than the misses they save (-2 %).
#ifndef HPE_SINCOS_INLINE
#define HPE_SINCOS_INLINE 1
#endif
#if HPE_SINCOS_INLINE
__device__ __forceinline__ void sincos_deg(double deg, double *sn, double *cs) { sincos(hpe_deg2rad(deg), sn, cs); }
#else
static __device__ __noinline__ void sincos_deg(double deg, double *sn, double *cs) { sincos(hpe_deg2rad(deg), sn, cs); }
#endif

// joints J[0..4] of digit d (0 = thumb) for pose th[26]: the lane computes its seven sincos itself
__device__ __forceinline__ void fk_digit(const DevModel &M, int d, const double *th, double J[5][3])
{
    double pos[3] = {th[3], th[4], th[5]};
    fk_chain(M, d, pos, [&](int k, double &sn, double &cs) { sincos_deg(trig_angle_deg(th, d, k), &sn, &cs); },
             [&](int i, double x, double y, double z) { J[i][0] = x; J[i][1] = y; J[i][2] = z; });
}

// The same chain writing the digit's sphere centres straight to sd[48][3] (un-negated frame) as soon as the two joints
// of a segment are final -- no joint or sphere arrays stay live, the sincos are taken when the chain needs them.
// Same interpolation arithmetic as digit_spheres().
struct SphereSink {
    double *sd;       // [48][3] of the pose
    int d;
    double jx, jy, jz;            // previous joint of the chain (joint 1 while joint 0 is pending)
    __device__ __forceinline__ void put(int i, double x, double y, double z) const
    {
        double *p = sd + 3 * i;
        p[0] = x; p[1] = y; p[2] = z;
    }
    __device__ __forceinline__ void operator()(int i, double x, double y, double z)
    {
        const int first = d == 0 ? 0 : 8 + 10 * (d - 1);
        if (i == 1) { jx = x; jy = y; jz = z; return; }
        if (i == 0) {             // segment 0 runs from joint 0 (x, y, z) to joint 1 (jx, jy, jz)
            if (d == 0) {
                put(first, mix2(0.5, x, 0.5, jx), mix2(0.5, y, 0.5, jy), mix2(0.5, z, 0.5, jz));
                put(first + 1, jx, jy, jz);
            } else {
                const double t = 1. / 3.;
                put(first, x, y, z);
                put(first + 1, mix2(1. - t, x, t, jx), mix2(1. - t, y, t, jy), mix2(1. - t, z, t, jz));
                put(first + 2, mix2(1. - t * 2, x, t * 2, jx), mix2(1. - t * 2, y, t * 2, jy), mix2(1. - t * 2, z, t * 2, jz));
                put(first + 3, jx, jy, jz);
            }
            return;
        }
        // segment i - 1 -> i (i = 2, 3, 4): its midpoint and its end
        const int at = first + (d == 0 ? 2 * (i - 1) : 2 + 2 * (i - 1));
        put(at, mix2(0.5, jx, 0.5, x), mix2(0.5, jy, 0.5, y), mix2(0.5, jz, 0.5, z));
        put(at + 1, x, y, z);
        jx = x; jy = y; jz = z;
    }
};

template <bool CG>
__device__ __forceinline__ void fk_digit_emit(const DevModel &M, int d, const double *th, double *sd)
{
    double pos[3];
    pos[0] = CG ? __ldcg(th + 3) : th[3]; pos[1] = CG ? __ldcg(th + 4) : th[4]; pos[2] = CG ? __ldcg(th + 5) : th[5];
    SphereSink sink{sd, d, 0.0, 0.0, 0.0};
    fk_chain(M, d, pos,
             [&](int k, double &sn, double &cs) {
                 const double *p = k < 3 ? th + k : th + 6 + 4 * d + (k - 3);
                 double deg = CG ? __ldcg(p) : *p;
                 if (k == 0) deg += 180.0;
                 sincos_deg(deg, &sn, &cs);
             },
             sink);
}

// the same for a pose that may have been written while this kernel runs (chunks landing by DMA under a
// persistent kernel): loads go to L2 (ld.global.cg), never through the non-coherent path
__device__ __forceinline__ void fk_digit_cg(const DevModel &M, int d, const double *th, double J[5][3])
{
    double loc[10];                                  // theta[0:6] and the digit's four angles
#pragma unroll
    for (int i = 0; i < 6; ++i) loc[i] = __ldcg(th + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) loc[6 + k] = __ldcg(th + 6 + 4 * d + k);
    DigitTrig t;
    double pos[3] = {loc[3], loc[4], loc[5]};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double deg = k == 0 ? loc[0] + 180.0 : k < 3 ? loc[k] : loc[6 + (k - 3)];
        sincos_deg(deg, &t.s[k], &t.c[k]);
    }
    fk_digit_trig(M, d, pos, t, J);
}

// One pose evaluated by a whole warp (the one-particle-per-warp latency paths): the 23 distinct sincos of the
// pose are computed by 23 lanes at once -- lane l < 3: global angle l; lane 3 + 4d + k: joint k of digit d --
// and handed to the five digit lanes by shuffles: one sincos latency instead of seven in a row.  Same values
// (sincos of the same argument), same chain arithmetic, so the result is bit-identical to fk_digit.
__device__ __forceinline__ void fk_digit_warp(const DevModel &M, const double *th, int lane, double J[5][3])
{
    double sn = 0.0, cs = 0.0;
    if (lane < 23) {
        const double deg = lane == 0 ? th[0] + 180.0 : lane < 3 ? th[lane] : th[6 + (lane - 3)];
        sincos_deg(deg, &sn, &cs);
    }
    const int d = lane < 5 ? lane : 0;
    DigitTrig t;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int src = k < 3 ? k : 3 + 4 * d + (k - 3);
        t.s[k] = __shfl_sync(0xffffffffu, sn, src);
        t.c[k] = __shfl_sync(0xffffffffu, cs, src);
    }
    if (lane < 5) {
        double pos[3] = {th[3], th[4], th[5]};
        fk_digit_trig(M, d, pos, t, J);
    }
}

// Sphere centres of a digit from its joints (src/fingermodel.cpp:208-267,
// src/thumbmodel.cpp:227-274), un-negated frame.  Returns the count (8 thumb / 10 finger);
// first sphere index of digit d is d == 0 ? 0 : 8 + 10*(d-1)  (src/handmodel.cpp:282-286).
__device__ __forceinline__ int digit_spheres(int d, const double J[5][3], double S[10][3])
{
    const double t = 1. / 3.;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (d == 0) {
            S[0][k] = mix2(0.5, J[0][k], 0.5, J[1][k]);
            S[1][k] = J[1][k];
        } else {
            S[0][k] = J[0][k];
            S[1][k] = mix2(1. - t, J[0][k], t, J[1][k]);
            S[2][k] = mix2(1. - t * 2, J[0][k], t * 2, J[1][k]);
            S[3][k] = J[1][k];
        }
    }
#pragma unroll
    for (int i = 1; i < 4; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double mid = mix2(0.5, J[i][k], 0.5, J[i + 1][k]);
            if (d == 0) { S[2 * i][k] = mid; S[2 * i + 1][k] = J[i + 1][k]; }
            else        { S[2 + 2 * i][k] = mid; S[3 + 2 * i][k] = J[i + 1][k]; }
        }
    }
    return d == 0 ? 8 : 10;
}

__device__ __forceinline__ int digit_first_sphere(int d) { return d == 0 ? 0 : 8 + 10 * (d - 1); }

// ---------------------------------------------------------------------------------------
// depth term of one sphere (src/costfunc.cpp:251-300), un-negated centre (x,y,z), radius r
// ---------------------------------------------------------------------------------------
// RN(mm / 10.0) for a float mm without the division: q0 = mm*0.1, one FMA residual correction.
// Equal to __ddiv_rn((double)mm, 10.0) for EVERY finite float (hpe_selftest_div10 sweeps all 2^32
// bit patterns on the device; tests/test_gpu_parity.py asserts zero mismatches).
__device__ __forceinline__ double div10_exact(float mm)
{
    double a = (double)mm;
    double q = __dmul_rn(a, 0.1);
    double r = __fma_rn(-q, 10.0, a);
    return __fma_rn(r, 0.1, q);
}

// RN(a / f) for a divisor f whose correctly rounded reciprocal rf = RN(1/f) is at hand (Markstein): q0 = RN(a rf)
// is within one ulp of the quotient, r = a - q0 f is exact in one FMA, and q0 + r rf rounds to the IEEE
// quotient.  hpe_selftest_div_const sweeps it against __ddiv_rn for the camera's focal length.
__device__ __forceinline__ double div_const_exact(double a, double f, double rf)
{
    double q = __dmul_rn(a, rf);
    double r = __fma_rn(-q, f, a);
    return __fma_rn(r, rf, q);
}

// floor(RN(a / z)) given rz = RN(1/z): q = a*rz is within 2 ulp of a/z, so the two floors can differ
// only when a/z lies within 2^-51 |a/z| of an integer; anything closer than 1e-6 to an integer (and
// NaN/inf) takes the IEEE division instead.
__device__ __forceinline__ double floor_div(double a, double z, double rz)
{
    double q = __dmul_rn(a, rz);
    double f = floor(q);
    double frac = q - f;
    if (!(frac > 1e-6 && frac < 1.0 - 1e-6) || !(fabs(q) < 1e9)) f = floor(__ddiv_rn(a, z));
    return f;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (DESIGN.md "RNG"); counter = (elem_pair, particle,
// generation, stream), key = 64-bit seed.  Replayed bit-for-bit by the CPU oracle.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                           uint32_t c3, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo)
{
    uint64_t k = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return (double)k * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void uniform2(uint64_t seed, uint32_t elem_pair, uint32_t particle,
                                         uint32_t gen, uint32_t stream, double u[2])
{
    uint32_t w[4];
    philox4x32(seed, elem_pair, particle, gen, stream, w);
    u[0] = u53(w[0], w[1]);
    u[1] = u53(w[2], w[3]);
}

__device__ __forceinline__ void normal2(uint64_t seed, uint32_t elem_pair, uint32_t particle,
                                        uint32_t gen, double z[2])
{
    double u[2];
    uniform2(seed, elem_pair, particle, gen, 0u, u);
    double u1 = u[0] + 1.0 / 9007199254740992.0;
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincos(6.283185307179586476925286766559 * u[1], &s, &c);
    z[0] = r * c;
    z[1] = r * s;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
