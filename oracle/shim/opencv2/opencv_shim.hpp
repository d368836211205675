// oracle/shim/opencv2/opencv_shim.hpp -- a MINIMAL stand-in for the OpenCV C++ API (TEST
// INFRASTRUCTURE), just enough for the reference's src/costfunc.cpp and src/observedmodel.cpp to
// compile unmodified (oracle/Makefile, target _ref).  Original code.  Two functions carry arithmetic:
//   cv::BFMatcher(NORM_L2)::match  -- brute force, float32: d = sum_k (q_k - t_k)^2 accumulated in
//       index order, distance = sqrt(d), first strictly smaller distance wins;
//   cv::distanceTransform(src, dst, CV_DIST_L2, 5) -- the two-pass 5x5 chamfer transform in 16-bit
//       fixed point with weights (1, 1.4, 2.1969), result * 2^-16 as float32.
// Both are checked bit-for-bit against the cv2 4.13 Python module in tests/test_oracle_pins.py (via
// the oracle's own restatements, which tests/test_ref_build.py compares with this build).  The GUI
// and drawing entry points the reference never reaches on the hot path abort if called.
#ifndef HPE_OPENCV_SHIM
#define HPE_OPENCV_SHIM

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../orc_draw.h"      // OpenCV's outlined circle, restated (shared with the C oracle)

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_DIST_L2 2
#define CV_DIST_MASK_5 5
#define CV_RETR_TREE 3
#define CV_CHAIN_APPROX_SIMPLE 2
#define CV_WINDOW_AUTOSIZE 1

namespace cv {

typedef unsigned char uchar;
enum { NORM_L2 = 4, WINDOW_AUTOSIZE = 1 };

[[noreturn]] inline void shim_unreachable(const char *what)
{
    std::fprintf(stderr, "opencv shim: %s is not implemented (GUI / drawing path, unreachable from the hot path)\n", what);
    std::abort();
}

struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
struct Vec4i {
    int val[4];
    int &operator[](int i) { return val[i]; }
};

inline int shim_elem_size(int type)
{
    switch (type) {
    case CV_8UC1: return 1;
    case CV_8UC3: return 3;
    case CV_32F: return 4;
    case CV_64F: return 8;
    }
    shim_unreachable("this matrix type");
}

// dense 2-D matrix, row-major, continuous; shares its buffer on copy like cv::Mat
class Mat {
public:
    int rows, cols;
    int mtype;
    uchar *data;
    std::shared_ptr<std::vector<uchar> > own;

    Mat() : rows(0), cols(0), mtype(CV_8UC1), data(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), mtype(type), data(0) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext) : rows(r), cols(c), mtype(type), data((uchar *)ext) {}   // no copy
    void create(int r, int c, int type)
    {
        rows = r; cols = c; mtype = type;
        own.reset(new std::vector<uchar>((size_t)r * c * shim_elem_size(type), 0));
        data = own->data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type); }
    int type() const { return mtype; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == 0 || rows * cols == 0; }
    template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data) + (size_t)r * cols; }
    template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data) + (size_t)r * cols; }
    template <class T> T &at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T &at(int r, int c) const { return ptr<T>(r)[c]; }

    Mat t() const
    {
        Mat o(cols, rows, mtype);
        const int es = shim_elem_size(mtype);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c)
                std::memcpy(o.data + ((size_t)c * rows + r) * es, data + ((size_t)r * cols + c) * es, es);
        return o;
    }
    Mat clone() const
    {
        Mat o(rows, cols, mtype);
        std::memcpy(o.data, data, (size_t)rows * cols * shim_elem_size(mtype));
        return o;
    }
    // saturate_cast<uchar>(double): round to nearest even, clamp to [0, 255]
    void convertTo(Mat &dst, int type) const
    {
        Mat o(rows, cols, type);
        const size_t n = (size_t)rows * cols;
        if (mtype == CV_64F && type == CV_8UC1) {
            const double *s = reinterpret_cast<const double *>(data);
            for (size_t i = 0; i < n; ++i) {
                long v = std::lrint(s[i]);
                o.data[i] = (uchar)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        } else if (mtype == CV_32F && type == CV_64F) {
            const float *s = reinterpret_cast<const float *>(data);
            double *d = reinterpret_cast<double *>(o.data);
            for (size_t i = 0; i < n; ++i) d[i] = s[i];
        } else if (mtype == type) {
            std::memcpy(o.data, data, n * shim_elem_size(type));
        } else {
            shim_unreachable("convertTo for this type pair");
        }
        dst = o;
    }
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0.f) {}
};

class BFMatcher {
    int norm;
public:
    explicit BFMatcher(int normType = NORM_L2, bool crossCheck = false) : norm(normType)
    {
        if (normType != NORM_L2 || crossCheck) shim_unreachable("BFMatcher other than NORM_L2 without cross-check");
    }
    void match(const Mat &query, const Mat &train, std::vector<DMatch> &matches) const
    {
        if (query.type() != CV_32F || train.type() != CV_32F || query.cols != train.cols) shim_unreachable("BFMatcher on non-float32 descriptors");
        matches.clear();
        matches.reserve(query.rows);
        const int dim = query.cols;
        for (int q = 0; q < query.rows; ++q) {
            const float *a = query.ptr<float>(q);
            DMatch best;
            best.queryIdx = q; best.imgIdx = 0;
            float bestd = 0.f;
            for (int t = 0; t < train.rows; ++t) {
                const float *b = train.ptr<float>(t);
                float d = 0.f;
                for (int k = 0; k < dim; ++k) { const float e = a[k] - b[k]; d += e * e; }
                const float w = std::sqrt(d);
                if (best.trainIdx < 0 || w < bestd) { bestd = w; best.trainIdx = t; }
            }
            best.distance = bestd;
            matches.push_back(best);
        }
    }
};

// 5x5 chamfer distance transform to the nearest zero pixel, 16-bit fixed point
inline void distanceTransform(const Mat &src, Mat &dst, int distanceType, int maskSize)
{
    if (src.type() != CV_8UC1 || distanceType != CV_DIST_L2 || maskSize != 5) shim_unreachable("distanceTransform other than (8UC1, DIST_L2, 5)");
    const int H = src.rows, W = src.cols, B = 2, P = W + 2 * B;
    const unsigned HV = 65536u;                                  // 1.0
    const unsigned DG = (unsigned)std::lrint(1.4 * 65536.0);      // 91750
    const unsigned LG = (unsigned)std::lrint(2.1969 * 65536.0);   // 143976
    const unsigned BIG = 0x7fffffffu >> 1;                       // larger than any reachable distance, no overflow when a weight is added
    std::vector<unsigned> buf((size_t)(H + 2 * B) * P, BIG);
    unsigned *t = buf.data() + (size_t)B * P + B;
    for (int y = 0; y < H; ++y) {
        const uchar *s = src.ptr<uchar>(y);
        unsigned *r = t + (size_t)y * P;
        for (int x = 0; x < W; ++x) {
            if (!s[x]) { r[x] = 0; continue; }
            unsigned m = r[x - 2 * P - 1] + LG, v;
            v = r[x - 2 * P + 1] + LG; if (v < m) m = v;
            v = r[x - P - 2] + LG; if (v < m) m = v;
            v = r[x - P - 1] + DG; if (v < m) m = v;
            v = r[x - P] + HV; if (v < m) m = v;
            v = r[x - P + 1] + DG; if (v < m) m = v;
            v = r[x - P + 2] + LG; if (v < m) m = v;
            v = r[x - 1] + HV; if (v < m) m = v;
            r[x] = m < BIG ? m : BIG;
        }
    }
    dst.create(H, W, CV_32F);
    const float scale = 1.f / 65536.f;
    for (int y = H - 1; y >= 0; --y) {
        unsigned *r = t + (size_t)y * P;
        float *d = dst.ptr<float>(y);
        for (int x = W - 1; x >= 0; --x) {
            unsigned m = r[x];
            if (m > HV) {
                unsigned v;
                v = r[x + 2 * P + 1] + LG; if (v < m) m = v;
                v = r[x + 2 * P - 1] + LG; if (v < m) m = v;
                v = r[x + P + 2] + LG; if (v < m) m = v;
                v = r[x + P + 1] + DG; if (v < m) m = v;
                v = r[x + P] + HV; if (v < m) m = v;
                v = r[x + P - 1] + DG; if (v < m) m = v;
                v = r[x + P - 2] + LG; if (v < m) m = v;
                v = r[x + 1] + HV; if (v < m) m = v;
                r[x] = m;
            }
            d[x] = scale * (float)m;
        }
    }
}

// ---- GUI / drawing: compile-only ----
inline void namedWindow(const std::string &, int = 1) { shim_unreachable("namedWindow"); }
inline void imshow(const std::string &, const Mat &) { shim_unreachable("imshow"); }
inline int waitKey(int = 0) { shim_unreachable("waitKey"); }
// cv::circle on a CV_64F image, line_type 8, shift 0:
//  * thickness < 0, the filled disc costfunc::bincomp_penalty draws (src/costfunc.cpp:423): OpenCV's midpoint walk
//    `Circle(..., fill)` of imgproc/drawing.cpp, the same restatement as oracle/hpe_oracle.c:orc_fill_circle;
//  * thickness >= 2, the outline costfunc::depthMatch_penalty draws (src/costfunc.cpp:468): OpenCV's EllipseEx / PolyLine /
//    ThickLine / FillConvexPoly path, restated in oracle/orc_draw.h.
// Both are pinned bit for bit against cv2.circle (4.13) in tests/test_oracle_pins.py.  Thin outlines (thickness 0, 1: the
// viewer) stay unreachable.
struct shim_circle_ctx { Mat *img; double value; };
inline void shim_circle_run(void *ctx, int y, int x1, int x2)
{
    shim_circle_ctx *c = static_cast<shim_circle_ctx *>(ctx);
    double *line = c->img->ptr<double>(y);
    for (int x = x1; x <= x2; ++x) line[x] = c->value;
}
inline void circle(Mat &img, Point c, int radius, const Scalar &colour, int thickness = 1, int line_type = 8, int shift = 0)
{
    if (line_type != 8 || shift != 0 || img.type() != CV_64F || radius < 0 || thickness == 0 || thickness == 1)
        shim_unreachable("circle (thin outline, other image types)");
    if (thickness >= 2) {
        shim_circle_ctx ctx = {&img, colour.val[0]};
        orcd_target t = {img.cols, img.rows, &ctx, shim_circle_run};
        orcd_circle_outline(&t, c.x, c.y, radius, thickness);
        return;
    }
    int e = 0, dx = radius, dy = 0, up = 1, down = 2 * radius - 1;
    for (; dx >= dy; ++dy) {
        for (int pass = 0; pass < 4; ++pass) {
            const int row = pass == 0 ? c.y - dy : pass == 1 ? c.y + dy : pass == 2 ? c.y - dx : c.y + dx;
            const int half = pass < 2 ? dx : dy;
            if (row < 0 || row >= img.rows) continue;
            double *line = img.ptr<double>(row);
            for (int x = std::max(c.x - half, 0); x <= std::min(c.x + half, img.cols - 1); ++x) line[x] = colour.val[0];
        }
        e += up;
        up += 2;
        if (e > 0) { e -= down; --dx; down -= 2; }
    }
}
inline void findContours(Mat &, std::vector<std::vector<Point> > &, std::vector<Vec4i> &, int, int, Point = Point())
{
    shim_unreachable("findContours");
}
inline void drawContours(Mat &, const std::vector<std::vector<Point> > &, int, const Scalar &, int = 1, int = 8,
                         const std::vector<Vec4i> & = std::vector<Vec4i>(), int = 0, Point = Point())
{
    shim_unreachable("drawContours");
}

}  // namespace cv
#endif
