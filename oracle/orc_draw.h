/* orc_draw.h -- OpenCV's outlined circle, restated: cv::circle(img, centre, radius, colour, thickness > 1, LINE_8, shift 0)
 * as costfunc::depthMatch_penalty draws it (src/costfunc.cpp:468, thickness 2).  Test infrastructure (oracle/), shared by
 * the C oracle (hpe_oracle.c) and the stand-in OpenCV header of oracle/_ref (shim/opencv2/opencv_shim.hpp).
 *
 * The algorithm lives in a third-party dependency that is absent from /root/reference: OpenCV ("OpenCV 3.0" in the
 * reference's README.md:7; modules/imgproc/src/drawing.cpp).  Restated from its published algorithm and PINNED bit for
 * bit against the cv2 4.13 Python module on thousands of circles -- clipped, outside the image, radius 0, thickness 2
 * and 3 -- in tests/test_oracle_pins.py:
 *   circle()        thickness > 1: EllipseEx(centre << 16, axes = radius << 16, 0, 0, 360)
 *   EllipseEx()     delta = 90 / 30 / 18 / 5 degrees by radius; ellipse2Poly in double on a table of float sines of whole
 *                   degrees (7 decimals); points rounded to Q16 fixed point, repeats dropped; an open PolyLine
 *   PolyLine()      ThickLine per segment; the first one caps both ends, the others their far end
 *   ThickLine()     thickness > 1: the segment widened to a quadrilateral (normal of length thickness / 2 rounded to
 *                   Q16) -> FillConvexPoly, and a FILLED circle of radius (thickness / 2 rounded) at the capped ends
 *   FillConvexPoly  every edge with Line2 (a Q16 DDA after clipLine), then the span between the two active edges of
 *                   every scan line, rounded to nearest
 * Pixels go to two callbacks (a horizontal run and a single point), so any image type can sit behind it. */
#ifndef ORC_DRAW_H
#define ORC_DRAW_H
#include <math.h>
#include <stdint.h>

#define ORCD_SHIFT 16
#define ORCD_ONE (1 << ORCD_SHIFT)

typedef struct {
    int w, h;
    void *ctx;
    void (*run)(void *ctx, int y, int x1, int x2);      /* pixels [x1, x2] of row y, inside the image */
} orcd_target;

typedef struct { int64_t x, y; } orcd_pt;

/* sin of whole degrees, 0..450, as OpenCV tabulates it: 7 decimals, float */
static const float orcd_sin[451] = {
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

static inline void orcd_point(const orcd_target *t, int64_t x, int64_t y)
{
    if (0 <= x && x < t->w && 0 <= y && y < t->h) t->run(t->ctx, (int)y, (int)x, (int)x);
}

/* Circle(img, centre, radius, colour, fill = 1): midpoint walk, four clipped runs per step */
static inline void orcd_fill_circle(const orcd_target *t, int cx, int cy, int radius)
{
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        const int ys[4] = {cy - dy, cy + dy, cy - dx, cy + dx};
        const int hw[4] = {dx, dx, dy, dy};
        for (int k = 0; k < 4; ++k) {
            if (ys[k] < 0 || ys[k] >= t->h) continue;
            int xa = cx - hw[k], xb = cx + hw[k];
            if (xa < 0) xa = 0;
            if (xb > t->w - 1) xb = t->w - 1;
            if (xa <= xb) t->run(t->ctx, ys[k], xa, xb);
        }
        dy++;
        err += plus;
        plus += 2;
        if (err > 0) { err -= minus; dx -= 1; minus -= 2; }
    }
}

/* cv::clipLine(Size2l, Point2l&, Point2l&) */
static inline int orcd_clip_line(int64_t W, int64_t H, orcd_pt *p1, orcd_pt *p2)
{
    const int64_t right = W - 1, bottom = H - 1;
    if (W <= 0 || H <= 0) return 0;
    int64_t x1 = p1->x, y1 = p1->y, x2 = p2->x, y2 = p2->y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    p1->x = x1; p1->y = y1; p2->x = x2; p2->y = y2;
    return (c1 | c2) == 0;
}

/* Line2: 8-connected line between Q16 points */
static inline void orcd_line2(const orcd_target *t, orcd_pt pt1, orcd_pt pt2)
{
    if (!orcd_clip_line((int64_t)t->w << ORCD_SHIFT, (int64_t)t->h << ORCD_SHIFT, &pt1, &pt2)) return;
    int64_t dx = pt2.x - pt1.x, dy = pt2.y - pt1.y;
    const int64_t j = dx < 0 ? -1 : 0, i = dy < 0 ? -1 : 0;
    const int64_t ax = (dx ^ j) - j, ay = (dy ^ i) - i;
    int64_t x_step, y_step;
    int ecount;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        if (j) { orcd_pt s = pt1; pt1 = pt2; pt2 = s; }
        x_step = ORCD_ONE;
        y_step = (dy * ORCD_ONE) / (ax | 1);
        ecount = (int)((pt2.x - pt1.x) >> ORCD_SHIFT);
    } else {
        dx = (dx ^ i) - i;
        if (i) { orcd_pt s = pt1; pt1 = pt2; pt2 = s; }
        x_step = (dx * ORCD_ONE) / (ay | 1);
        y_step = ORCD_ONE;
        ecount = (int)((pt2.y - pt1.y) >> ORCD_SHIFT);
    }
    pt1.x += ORCD_ONE >> 1;
    pt1.y += ORCD_ONE >> 1;
    orcd_point(t, (pt2.x + (ORCD_ONE >> 1)) >> ORCD_SHIFT, (pt2.y + (ORCD_ONE >> 1)) >> ORCD_SHIFT);
    if (ax > ay) {
        pt1.x >>= ORCD_SHIFT;
        while (ecount >= 0) {
            orcd_point(t, pt1.x, pt1.y >> ORCD_SHIFT);
            pt1.x++;
            pt1.y += y_step;
            ecount--;
        }
    } else {
        pt1.y >>= ORCD_SHIFT;
        while (ecount >= 0) {
            orcd_point(t, pt1.x >> ORCD_SHIFT, pt1.y);
            pt1.x += x_step;
            pt1.y++;
            ecount--;
        }
    }
    (void)x_step;
}

/* FillConvexPoly(img, v, npts, colour, LINE_8, shift = 16) */
static inline void orcd_fill_convex(const orcd_target *t, const orcd_pt *v, int npts)
{
    struct { int idx, di; int64_t x, dx; int ye; } edge[2];
    const int shift = ORCD_SHIFT, delta = 1 << shift >> 1, delta1 = ORCD_ONE >> 1, delta2 = ORCD_ONE >> 1;
    int i, y, imin = 0, edges = npts;
    int64_t xmin, xmax, ymin, ymax;
    orcd_pt p0 = v[npts - 1];
    xmin = xmax = v[0].x;
    ymin = ymax = v[0].y;
    for (i = 0; i < npts; i++) {
        orcd_pt p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        if (p.y > ymax) ymax = p.y;
        if (p.x > xmax) xmax = p.x;
        if (p.x < xmin) xmin = p.x;
        orcd_line2(t, p0, p);
        p0 = p;
    }
    xmin = (xmin + delta) >> shift; xmax = (xmax + delta) >> shift;
    ymin = (ymin + delta) >> shift; ymax = (ymax + delta) >> shift;
    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= t->w || (int)ymin >= t->h) return;
    if (ymax > t->h - 1) ymax = t->h - 1;
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y = (int)ymin;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -ORCD_ONE;
    edge[0].dx = edge[1].dx = 0;
    do {
        for (i = 0; i < 2; i++) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di, idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    const int ty = (int)((v[idx].y + delta) >> shift);
                    if (ty > y) {
                        const int64_t xs = v[idx0].x, xe = v[idx].x;
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
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
            int left = 0, right = 1;
            if (edge[0].x > edge[1].x) { left = 1; right = 0; }
            int xx1 = (int)((edge[left].x + delta1) >> ORCD_SHIFT), xx2 = (int)((edge[right].x + delta2) >> ORCD_SHIFT);
            if (xx2 >= 0 && xx1 < t->w) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= t->w) xx2 = t->w - 1;
                if (xx1 <= xx2) t->run(t->ctx, y, xx1, xx2);
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

/* cvRound: to nearest, ties to even (SSE2 cvtsd2si under the default rounding mode) */
static inline int64_t orcd_round(double v) { return (int64_t)nearbyint(v); }

/* ThickLine(img, p0, p1, colour, thickness > 1, LINE_8, flags, shift = 16) */
static inline void orcd_thick_line(const orcd_target *t, orcd_pt p0, orcd_pt p1, int thickness, int flags)
{
    const double inv = 1. / ORCD_ONE;
    const double dx = (double)(p0.x - p1.x) * inv, dy = (double)(p1.y - p0.y) * inv;
    double r = dx * dx + dy * dy;
    const int odd = thickness & 1;
    thickness <<= ORCD_SHIFT - 1;
    if (fabs(r) > 2.2204460492503131e-16) {
        r = (thickness + odd * ORCD_ONE * 0.5) / sqrt(r);
        const int64_t dpx = orcd_round(dy * r), dpy = orcd_round(dx * r);
        orcd_pt pt[4] = {{p0.x + dpx, p0.y + dpy}, {p0.x - dpx, p0.y - dpy}, {p1.x - dpx, p1.y - dpy}, {p1.x + dpx, p1.y + dpy}};
        orcd_fill_convex(t, pt, 4);
    }
    for (int i = 0; i < 2; i++) {
        if (flags & 1)
            orcd_fill_circle(t, (int)((p0.x + (ORCD_ONE >> 1)) >> ORCD_SHIFT), (int)((p0.y + (ORCD_ONE >> 1)) >> ORCD_SHIFT),
                             (thickness + (ORCD_ONE >> 1)) >> ORCD_SHIFT);
        p0 = p1;
        flags >>= 1;
    }
}

/* cv::circle(img, (cx, cy), radius, colour, thickness >= 2, LINE_8, 0) */
static inline void orcd_circle_outline(const orcd_target *t, int cx, int cy, int radius, int thickness)
{
    const orcd_pt center = {(int64_t)cx * ORCD_ONE, (int64_t)cy * ORCD_ONE};
    const int64_t axis = (int64_t)(radius < 0 ? -radius : radius) * ORCD_ONE;
    int delta = (int)((axis + (ORCD_ONE >> 1)) >> ORCD_SHIFT);
    delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;
    /* ellipse2Poly(centre, axes, angle 0, 0, 360, delta) in double: alpha = cos 0 = 1, beta = sin 0 = 0 */
    const double alpha = (double)orcd_sin[450], beta = (double)orcd_sin[0];
    orcd_pt v[80];
    int n = 0;
    orcd_pt prev = {-1, -1};
    int have_prev = 0, raw = 0;
    for (int i = 0; i < 360 + delta; i += delta) {
        const int a = i > 360 ? 360 : i;
        const double x = (double)axis * (double)orcd_sin[450 - a], y = (double)axis * (double)orcd_sin[a];
        const double px = (double)center.x + x * alpha - y * beta, py = (double)center.y + x * beta + y * alpha;
        orcd_pt pt;
        pt.x = orcd_round(px / ORCD_ONE) * ORCD_ONE;
        pt.y = orcd_round(py / ORCD_ONE) * ORCD_ONE;
        pt.x += orcd_round(px - (double)pt.x);
        pt.y += orcd_round(py - (double)pt.y);
        raw++;
        if (!have_prev || pt.x != prev.x || pt.y != prev.y) { v[n++] = pt; prev = pt; have_prev = 1; }
    }
    (void)raw;
    if (n == 1) { v[0] = center; v[1] = center; n = 2; }
    int flags = 2 + 1;                      /* PolyLine, not closed */
    orcd_pt p0 = v[0];
    for (int i = 1; i < n; i++) {
        orcd_thick_line(t, p0, v[i], thickness, flags);
        p0 = v[i];
        flags = 2;
    }
}
#endif
