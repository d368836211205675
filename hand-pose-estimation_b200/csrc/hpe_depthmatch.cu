// hpe_depthmatch.cu -- costfunc::depthMatch_penalty (src/costfunc.cpp:439-473), SURVEY 8(f) rank 4: the observed depth map
// (cm) with the OUTLINE of every projected sphere drawn into it -- centre = trunc(K c / c_z), radius = trunc(r / scale),
// cv::circle(out_img, centre, radius, Scalar(c_z), thickness = 2) -- in sphere order (a later outline overwrites an
// earlier one).  The image is what the reference's debug views show.
//
// cv::circle with thickness > 1 is not the midpoint walk of the filled disc (hpe_bincomp.cu) but OpenCV's polygonal path
// (imgproc/drawing.cpp): EllipseEx -> ellipse2Poly (a vertex every 90 / 30 / 18 / 5 degrees by radius, in double on a
// table of float sines of whole degrees) -> vertices rounded to Q16 fixed point -> PolyLine -> per segment ThickLine: the
// segment widened to a quadrilateral (FillConvexPoly: its four edges as Q16 DDA lines after clipLine, then the span
// between the two active edges of every scan line) plus a filled circle of radius 1 at the capped ends.  All of it is
// integer work after a handful of correctly rounded FP64 operations, restated here operation by operation, so the raster
// equals the reference's bit for bit; the value drawn is the sphere's depth, an FK output (tests/test_gpu_depthmatch.py: against the CPU restatement, which the CPU tests pin
// against cv2.circle and against the reference's own depthMatch_penalty in the test build of its sources).
//
// One CTA per pose: warp 0 runs the forward kinematics, all threads copy the depth map, then the 48 spheres in order;
// within a sphere every segment of its polygon is drawn by its own thread (all segments write the same value, so their
// order does not matter), spheres are separated by a CTA barrier.
#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace {

constexpr int DM_THREADS = 96;
constexpr int DM_SHIFT = 16;
constexpr long long DM_ONE = 1ll << DM_SHIFT;
constexpr int DM_MAXV = 80;               // 360 / 5 + 1 vertices at most

// sin of whole degrees 0..450 as OpenCV tabulates it (7 decimals, float)
__device__ const float dm_sin[451] = {
    0.0000000f, 0.0174524f, 0.0348995f, 0.0523360f, 0.0697565f, 0.0871557f, 0.1045285f, 0.1218693f,
    0.1391731f, 0.1564345f, 0.1736482f, 0.1908090f, 0.2079117f, 0.2249511f, 0.2419219f, 0.2588190f,
    0.2756374f, 0.2923717f, 0.3090170f, 0.3255682f, 0.3420201f, 0.3583679f, 0.3746066f, 0.3907311f,
    0.4067366f, 0.4226183f, 0.4383711f, 0.4539905f, 0.4694716f, 0.4848096f, 0.5000000f, 0.5150381f,
    0.5299193f, 0.5446390f, 0.5591929f, 0.5735764f, 0.5877853f, 0.6018150f, 0.6156615f, 0.6293204f,
    0.6427876f, 0.6560590f, 0.6691306f, 0.6819984f, 0.6946584f, 0.7071068f, 0.7193398f, 0.7313537f,
    0.7431448f, 0.7547096f, 0.7660444f, 0.7771460f, 0.7880108f, 0.7986355f, 0.8090170f, 0.8191520f,
    0.8290376f, 0.8386706f, 0.8480481f, 0.8571673f, 0.8660254f, 0.8746197f, 0.8829476f, 0.8910065f,
    0.8987940f, 0.9063078f, 0.9135455f, 0.9205049f, 0.9271839f, 0.9335804f, 0.9396926f, 0.9455186f,
    0.9510565f, 0.9563048f, 0.9612617f, 0.9659258f, 0.9702957f, 0.9743701f, 0.9781476f, 0.9816272f,
    0.9848078f, 0.9876883f, 0.9902681f, 0.9925462f, 0.9945219f, 0.9961947f, 0.9975641f, 0.9986295f,
    0.9993908f, 0.9998477f, 1.0000000f, 0.9998477f, 0.9993908f, 0.9986295f, 0.9975641f, 0.9961947f,
    0.9945219f, 0.9925462f, 0.9902681f, 0.9876883f, 0.9848078f, 0.9816272f, 0.9781476f, 0.9743701f,
    0.9702957f, 0.9659258f, 0.9612617f, 0.9563048f, 0.9510565f, 0.9455186f, 0.9396926f, 0.9335804f,
    0.9271839f, 0.9205049f, 0.9135455f, 0.9063078f, 0.8987940f, 0.8910065f, 0.8829476f, 0.8746197f,
    0.8660254f, 0.8571673f, 0.8480481f, 0.8386706f, 0.8290376f, 0.8191520f, 0.8090170f, 0.7986355f,
    0.7880108f, 0.7771460f, 0.7660444f, 0.7547096f, 0.7431448f, 0.7313537f, 0.7193398f, 0.7071068f,
    0.6946584f, 0.6819984f, 0.6691306f, 0.6560590f, 0.6427876f, 0.6293204f, 0.6156615f, 0.6018150f,
    0.5877853f, 0.5735764f, 0.5591929f, 0.5446390f, 0.5299193f, 0.5150381f, 0.5000000f, 0.4848096f,
    0.4694716f, 0.4539905f, 0.4383711f, 0.4226183f, 0.4067366f, 0.3907311f, 0.3746066f, 0.3583679f,
    0.3420201f, 0.3255682f, 0.3090170f, 0.2923717f, 0.2756374f, 0.2588190f, 0.2419219f, 0.2249511f,
    0.2079117f, 0.1908090f, 0.1736482f, 0.1564345f, 0.1391731f, 0.1218693f, 0.1045285f, 0.0871557f,
    0.0697565f, 0.0523360f, 0.0348995f, 0.0174524f, 0.0000000f, -0.0174524f, -0.0348995f, -0.0523360f,
    -0.0697565f, -0.0871557f, -0.1045285f, -0.1218693f, -0.1391731f, -0.1564345f, -0.1736482f, -0.1908090f,
    -0.2079117f, -0.2249511f, -0.2419219f, -0.2588190f, -0.2756374f, -0.2923717f, -0.3090170f, -0.3255682f,
    -0.3420201f, -0.3583679f, -0.3746066f, -0.3907311f, -0.4067366f, -0.4226183f, -0.4383711f, -0.4539905f,
    -0.4694716f, -0.4848096f, -0.5000000f, -0.5150381f, -0.5299193f, -0.5446390f, -0.5591929f, -0.5735764f,
    -0.5877853f, -0.6018150f, -0.6156615f, -0.6293204f, -0.6427876f, -0.6560590f, -0.6691306f, -0.6819984f,
    -0.6946584f, -0.7071068f, -0.7193398f, -0.7313537f, -0.7431448f, -0.7547096f, -0.7660444f, -0.7771460f,
    -0.7880108f, -0.7986355f, -0.8090170f, -0.8191520f, -0.8290376f, -0.8386706f, -0.8480481f, -0.8571673f,
    -0.8660254f, -0.8746197f, -0.8829476f, -0.8910065f, -0.8987940f, -0.9063078f, -0.9135455f, -0.9205049f,
    -0.9271839f, -0.9335804f, -0.9396926f, -0.9455186f, -0.9510565f, -0.9563048f, -0.9612617f, -0.9659258f,
    -0.9702957f, -0.9743701f, -0.9781476f, -0.9816272f, -0.9848078f, -0.9876883f, -0.9902681f, -0.9925462f,
    -0.9945219f, -0.9961947f, -0.9975641f, -0.9986295f, -0.9993908f, -0.9998477f, -1.0000000f, -0.9998477f,
    -0.9993908f, -0.9986295f, -0.9975641f, -0.9961947f, -0.9945219f, -0.9925462f, -0.9902681f, -0.9876883f,
    -0.9848078f, -0.9816272f, -0.9781476f, -0.9743701f, -0.9702957f, -0.9659258f, -0.9612617f, -0.9563048f,
    -0.9510565f, -0.9455186f, -0.9396926f, -0.9335804f, -0.9271839f, -0.9205049f, -0.9135455f, -0.9063078f,
    -0.8987940f, -0.8910065f, -0.8829476f, -0.8746197f, -0.8660254f, -0.8571673f, -0.8480481f, -0.8386706f,
    -0.8290376f, -0.8191520f, -0.8090170f, -0.7986355f, -0.7880108f, -0.7771460f, -0.7660444f, -0.7547096f,
    -0.7431448f, -0.7313537f, -0.7193398f, -0.7071068f, -0.6946584f, -0.6819984f, -0.6691306f, -0.6560590f,
    -0.6427876f, -0.6293204f, -0.6156615f, -0.6018150f, -0.5877853f, -0.5735764f, -0.5591929f, -0.5446390f,
    -0.5299193f, -0.5150381f, -0.5000000f, -0.4848096f, -0.4694716f, -0.4539905f, -0.4383711f, -0.4226183f,
    -0.4067366f, -0.3907311f, -0.3746066f, -0.3583679f, -0.3420201f, -0.3255682f, -0.3090170f, -0.2923717f,
    -0.2756374f, -0.2588190f, -0.2419219f, -0.2249511f, -0.2079117f, -0.1908090f, -0.1736482f, -0.1564345f,
    -0.1391731f, -0.1218693f, -0.1045285f, -0.0871557f, -0.0697565f, -0.0523360f, -0.0348995f, -0.0174524f,
    -0.0000000f, 0.0174524f, 0.0348995f, 0.0523360f, 0.0697565f, 0.0871557f, 0.1045285f, 0.1218693f,
    0.1391731f, 0.1564345f, 0.1736482f, 0.1908090f, 0.2079117f, 0.2249511f, 0.2419219f, 0.2588190f,
    0.2756374f, 0.2923717f, 0.3090170f, 0.3255682f, 0.3420201f, 0.3583679f, 0.3746066f, 0.3907311f,
    0.4067366f, 0.4226183f, 0.4383711f, 0.4539905f, 0.4694716f, 0.4848096f, 0.5000000f, 0.5150381f,
    0.5299193f, 0.5446390f, 0.5591929f, 0.5735764f, 0.5877853f, 0.6018150f, 0.6156615f, 0.6293204f,
    0.6427876f, 0.6560590f, 0.6691306f, 0.6819984f, 0.6946584f, 0.7071068f, 0.7193398f, 0.7313537f,
    0.7431448f, 0.7547096f, 0.7660444f, 0.7771460f, 0.7880108f, 0.7986355f, 0.8090170f, 0.8191520f,
    0.8290376f, 0.8386706f, 0.8480481f, 0.8571673f, 0.8660254f, 0.8746197f, 0.8829476f, 0.8910065f,
    0.8987940f, 0.9063078f, 0.9135455f, 0.9205049f, 0.9271839f, 0.9335804f, 0.9396926f, 0.9455186f,
    0.9510565f, 0.9563048f, 0.9612617f, 0.9659258f, 0.9702957f, 0.9743701f, 0.9781476f, 0.9816272f,
    0.9848078f, 0.9876883f, 0.9902681f, 0.9925462f, 0.9945219f, 0.9961947f, 0.9975641f, 0.9986295f,
    0.9993908f, 0.9998477f, 1.0000000f,
};

struct Pt { long long x, y; };

struct Canvas {
    double *img;          // [240][320]
    double value;
    __device__ __forceinline__ void run(int y, int x1, int x2) const
    {
        double *row = img + (size_t)y * HPE_W;
        for (int x = x1; x <= x2; ++x) row[x] = value;
    }
    __device__ __forceinline__ void point(long long x, long long y) const
    {
        if (0 <= x && x < HPE_W && 0 <= y && y < HPE_H) img[(size_t)y * HPE_W + x] = value;
    }
};

// Circle(img, centre, radius, colour, fill = 1): the midpoint walk, four clipped runs per step
__device__ void fill_circle(const Canvas &cv, int cx, int cy, int radius)
{
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = k == 0 ? cy - dy : k == 1 ? cy + dy : k == 2 ? cy - dx : cy + dx, hw = k < 2 ? dx : dy;
            if (y < 0 || y >= HPE_H) continue;
            const int xa = max(cx - hw, 0), xb = min(cx + hw, HPE_W - 1);
            if (xa <= xb) cv.run(y, xa, xb);
        }
        dy++;
        err += plus;
        plus += 2;
        if (err > 0) { err -= minus; dx -= 1; minus -= 2; }
    }
}

// cv::clipLine(Size2l, Point2l&, Point2l&): the intersections in double (multiply, then divide), truncated
__device__ __forceinline__ long long clip_step(long long a, long long b, long long c)
{
    return __double2ll_rz(__ddiv_rn(__dmul_rn((double)a, (double)b), (double)c));
}

__device__ bool clip_line(long long W, long long H, Pt &p1, Pt &p2)
{
    const long long right = W - 1, bottom = H - 1;
    long long x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += clip_step(a - y1, x2 - x1, y2 - y1);
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += clip_step(a - y2, x2 - x1, y2 - y1);
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += clip_step(a - x1, y2 - y1, x2 - x1);
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += clip_step(a - x2, y2 - y1, x2 - x1);
                x2 = a;
                c2 = 0;
            }
        }
    }
    p1.x = x1; p1.y = y1; p2.x = x2; p2.y = y2;
    return (c1 | c2) == 0;
}

// Line2: 8-connected line between Q16 points (a DDA along the longer axis)
__device__ void line2(const Canvas &cv, Pt pt1, Pt pt2)
{
    if (!clip_line((long long)HPE_W << DM_SHIFT, (long long)HPE_H << DM_SHIFT, pt1, pt2)) return;
    long long dx = pt2.x - pt1.x, dy = pt2.y - pt1.y;
    const long long j = dx < 0 ? -1 : 0, i = dy < 0 ? -1 : 0;
    const long long ax = (dx ^ j) - j, ay = (dy ^ i) - i;
    long long x_step, y_step;
    int ecount;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        if (j) { Pt s = pt1; pt1 = pt2; pt2 = s; }
        x_step = DM_ONE;
        y_step = (dy * DM_ONE) / (ax | 1);
        ecount = (int)((pt2.x - pt1.x) >> DM_SHIFT);
    } else {
        dx = (dx ^ i) - i;
        if (i) { Pt s = pt1; pt1 = pt2; pt2 = s; }
        x_step = (dx * DM_ONE) / (ay | 1);
        y_step = DM_ONE;
        ecount = (int)((pt2.y - pt1.y) >> DM_SHIFT);
    }
    pt1.x += DM_ONE >> 1;
    pt1.y += DM_ONE >> 1;
    cv.point((pt2.x + (DM_ONE >> 1)) >> DM_SHIFT, (pt2.y + (DM_ONE >> 1)) >> DM_SHIFT);
    if (ax > ay) {
        pt1.x >>= DM_SHIFT;
        for (; ecount >= 0; --ecount) {
            cv.point(pt1.x, pt1.y >> DM_SHIFT);
            pt1.x++;
            pt1.y += y_step;
        }
    } else {
        pt1.y >>= DM_SHIFT;
        for (; ecount >= 0; --ecount) {
            cv.point(pt1.x >> DM_SHIFT, pt1.y);
            pt1.x += x_step;
            pt1.y++;
        }
    }
}

// FillConvexPoly(img, v, 4, colour, LINE_8, shift = 16)
__device__ void fill_quad(const Canvas &cv, const Pt (&v)[4])
{
    constexpr int npts = 4, delta = 1 << DM_SHIFT >> 1;
    int idx_[2], di_[2], ye_[2];
    long long x_[2], dx_[2];
    int imin = 0, edges = npts;
    long long xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    Pt p0 = v[npts - 1];
#pragma unroll
    for (int i = 0; i < npts; i++) {
        const Pt p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        ymax = max(ymax, p.y);
        xmax = max(xmax, p.x);
        xmin = min(xmin, p.x);
        line2(cv, p0, p);
        p0 = p;
    }
    xmin = (xmin + delta) >> DM_SHIFT; xmax = (xmax + delta) >> DM_SHIFT;
    ymin = (ymin + delta) >> DM_SHIFT; ymax = (ymax + delta) >> DM_SHIFT;
    if ((int)xmax < 0 || (int)ymax < 0 || (int)xmin >= HPE_W || (int)ymin >= HPE_H) return;
    ymax = min(ymax, (long long)HPE_H - 1);
    int y = (int)ymin;
    idx_[0] = idx_[1] = imin;
    ye_[0] = ye_[1] = y;
    di_[0] = 1; di_[1] = npts - 1;
    x_[0] = x_[1] = -DM_ONE;
    dx_[0] = dx_[1] = 0;
    do {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (y >= ye_[i]) {
                int idx0 = idx_[i], di = di_[i], idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    const int ty = (int)((v[idx].y + delta) >> DM_SHIFT);
                    if (ty > y) {
                        const long long xs = v[idx0].x, xe = v[idx].x;
                        ye_[i] = ty;
                        dx_[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                        x_[i] = xs;
                        idx_[i] = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            const bool swap = x_[0] > x_[1];
            const long long xl = swap ? x_[1] : x_[0], xr = swap ? x_[0] : x_[1];
            int xx1 = (int)((xl + (DM_ONE >> 1)) >> DM_SHIFT), xx2 = (int)((xr + (DM_ONE >> 1)) >> DM_SHIFT);
            if (xx2 >= 0 && xx1 < HPE_W) {
                xx1 = max(xx1, 0);
                xx2 = min(xx2, HPE_W - 1);
                if (xx1 <= xx2) cv.run(y, xx1, xx2);
            }
        }
        x_[0] += dx_[0];
        x_[1] += dx_[1];
    } while (++y <= (int)ymax);
}

// ThickLine(img, p0, p1, colour, thickness, LINE_8, flags, shift = 16), thickness > 1
__device__ void thick_line(const Canvas &cv, Pt p0, Pt p1, int thickness, int flags)
{
    const double inv = 1. / (double)DM_ONE;
    const double dx = __dmul_rn((double)(p0.x - p1.x), inv), dy = __dmul_rn((double)(p1.y - p0.y), inv);
    double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const int odd = thickness & 1;
    thickness <<= DM_SHIFT - 1;
    if (fabs(r) > 2.2204460492503131e-16) {
        r = __ddiv_rn(__dadd_rn((double)thickness, __dmul_rn((double)(odd * DM_ONE), 0.5)), __dsqrt_rn(r));
        const long long dpx = __double2ll_rn(__dmul_rn(dy, r)), dpy = __double2ll_rn(__dmul_rn(dx, r));
        const Pt quad[4] = {{p0.x + dpx, p0.y + dpy}, {p0.x - dpx, p0.y - dpy}, {p1.x - dpx, p1.y - dpy}, {p1.x + dpx, p1.y + dpy}};
        fill_quad(cv, quad);
    }
    for (int i = 0; i < 2; i++) {
        if (flags & 1)
            fill_circle(cv, (int)((p0.x + (DM_ONE >> 1)) >> DM_SHIFT), (int)((p0.y + (DM_ONE >> 1)) >> DM_SHIFT),
                        (int)((thickness + (DM_ONE >> 1)) >> DM_SHIFT));
        p0 = p1;
        flags >>= 1;
    }
}

// double -> int as the reference's implicit conversions (truncation); what no image can hold is clamped
__device__ __forceinline__ int dm_trunc_to_int(double v)
{
    if (!(v > -1048576.0)) return -1048576;
    if (v > 1048576.0) return 1048576;
    return __double2int_rz(v);
}

struct DmSmem {
    DevModel M;
    double sphd[HPE_NSPH * 3];
    Pt v[DM_MAXV];
    int nv;
};

__global__ void __launch_bounds__(DM_THREADS)
depthmatch_kernel(DevModel m, ObsView o, int slot, const double *__restrict__ theta, double *__restrict__ out, int thickness)
{
    __shared__ DmSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p = blockIdx.x;
    for (int i = tid; i < (int)(sizeof(DevModel) / 8); i += DM_THREADS)
        reinterpret_cast<double *>(&sm.M)[i] = reinterpret_cast<const double *>(&m)[i];
    __syncthreads();
    if (warp == 0) {
        WarpScratch ws;
        ws.dep = nullptr; ws.cf = nullptr; ws.ex = nullptr; ws.sphd = sm.sphd;
        group_fk<false, true>(sm.M, [&](int) { return theta + p * HPE_NDOF; }, 1, ws, lane);
    }
    // out_img = depthmap (cm): the file value / 10 (src/observedmodel.cpp:300-303)
    double *img = out + (size_t)p * HPE_W * HPE_H;
    const float *depth = o.depth_mm + (size_t)slot * HPE_W * HPE_H;
    for (int i = tid; i < HPE_W * HPE_H; i += DM_THREADS) img[i] = div10_exact(__ldg(depth + i));
    const double scale = o.meta[slot].scale;
    __syncthreads();

    for (int s = 0; s < HPE_NSPH; ++s) {
        // cens = K * c / c_z in the un-negated frame (:449-455), no FMA contraction; (int) conversions truncate
        const double x = sm.sphd[s * 3], y = sm.sphd[s * 3 + 1], z = sm.sphd[s * 3 + 2];
        const int cx = dm_trunc_to_int(__ddiv_rn(__dadd_rn(__dmul_rn(o.focal, x), __dmul_rn(o.cx, z)), z));
        const int cy = dm_trunc_to_int(__ddiv_rn(__dadd_rn(__dmul_rn(o.focal, y), __dmul_rn(o.cy, z)), z));
        const int radius = dm_trunc_to_int(__ddiv_rn(sm.M.rad[s], scale));
        if (radius < 0) continue;                          // cv::circle asserts radius >= 0 (uniform over the CTA)
        // EllipseEx(centre << 16, axes = radius << 16, 0, 0, 360)
        const Pt center = {(long long)cx * DM_ONE, (long long)cy * DM_ONE};
        const long long axis = (long long)radius * DM_ONE;
        int delta = (int)((axis + (DM_ONE >> 1)) >> DM_SHIFT);
        delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;
        if (tid == 0) {
            // ellipse2Poly in double (angle 0: alpha = cos 0 = 1, beta = sin 0 = 0), vertices rounded to Q16, repeats dropped
            const double alpha = (double)dm_sin[450], beta = (double)dm_sin[0];
            int n = 0;
            Pt prev = {0, 0};
            for (int i = 0; i < 360 + delta; i += delta) {
                const int a = i > 360 ? 360 : i;
                const double ex = __dmul_rn((double)axis, (double)dm_sin[450 - a]), ey = __dmul_rn((double)axis, (double)dm_sin[a]);
                const double px = __dadd_rn(__dadd_rn((double)center.x, __dmul_rn(ex, alpha)), -__dmul_rn(ey, beta));
                const double py = __dadd_rn(__dadd_rn((double)center.y, __dmul_rn(ex, beta)), __dmul_rn(ey, alpha));
                Pt pt;
                pt.x = __double2ll_rn(__dmul_rn(px, 1. / (double)DM_ONE)) * DM_ONE;
                pt.y = __double2ll_rn(__dmul_rn(py, 1. / (double)DM_ONE)) * DM_ONE;
                pt.x += __double2ll_rn(__dadd_rn(px, -(double)pt.x));
                pt.y += __double2ll_rn(__dadd_rn(py, -(double)pt.y));
                if (n == 0 || pt.x != prev.x || pt.y != prev.y) { sm.v[n++] = pt; prev = pt; }
            }
            if (n == 1) { sm.v[0] = center; sm.v[1] = center; n = 2; }
            sm.nv = n;
        }
        __syncthreads();
        const Canvas cv = {img, z};
        // PolyLine (open): segment k = (v[k], v[k+1]); the first caps both ends, the others their far end
        for (int k = tid; k + 1 < sm.nv; k += DM_THREADS) thick_line(cv, sm.v[k], sm.v[k + 1], thickness, k == 0 ? 3 : 2);
        __syncthreads();
    }
}

}  // namespace

cudaError_t hpe_launch_depthmatch(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int64_t n,
                                  double *out_dev, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    depthmatch_kernel<<<(unsigned)n, DM_THREADS, 0, st>>>(m, o, slot, theta_dev, out_dev, 2);
    return cudaGetLastError();
}
