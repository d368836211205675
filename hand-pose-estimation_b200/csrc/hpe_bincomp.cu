// hpe_bincomp.cu -- costfunc::bincomp_penalty (src/costfunc.cpp:384-436), SURVEY 8(f) rank 4: the 48
// spheres of a pose drawn as filled discs -- centre = trunc(K c / c_z), radius = trunc(r / scale),
// cv::circle(..., thickness = -1) -- into an empty 240x320 image; the penalty is the number of pixels
// where that silhouette and the observed one (depth != 0) differ.
//
// Silhouettes are bit masks (one bit per pixel, 9.6 KB) in shared memory: a CTA builds the frame's
// mask once (ballots over the depth map), then for each of its poses warp 0 runs the forward kinematics,
// 48 lanes walk OpenCV's midpoint circle (drawing.cpp `Circle`, fill) setting whole runs of bits with
// atomicOr, and all threads count popc(drawn ^ observed).  Integer work after the two FP64 divisions
// of the projection, so the result equals the reference's bit for bit.
#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace {

constexpr int BC_THREADS = 128;
constexpr int BC_WORDS = HPE_W * HPE_H / 32;       // 2400
constexpr int BC_ROW = HPE_W / 32;                 // 10 words per image row

struct __align__(16) BincompSmem {
    FrameSmem fs;
    __align__(16) unsigned char ws[HPE_WS_BYTES(1)];
    unsigned observed[BC_WORDS];
    unsigned drawn[BC_WORDS];
    int part[BC_THREADS / 32];
};

// double -> int as the reference's implicit conversions (truncation); what no image can hold is clamped
// so that the disc falls outside the image instead of invoking the undefined conversion
__device__ __forceinline__ int trunc_to_int(double v)
{
    if (!(v > -1048576.0)) return -1048576;
    if (v > 1048576.0) return 1048576;
    return __double2int_rz(v);
}

// bits [xa, xb] of image row y
__device__ __forceinline__ void set_run(unsigned *mask, int y, int xa, int xb)
{
    if (y < 0 || y >= HPE_H) return;
    xa = max(xa, 0);
    xb = min(xb, HPE_W - 1);
    if (xa > xb) return;
    unsigned *row = mask + y * BC_ROW;
    const int wa = xa >> 5, wb = xb >> 5;
    for (int w = wa; w <= wb; ++w) {
        unsigned bits = 0xffffffffu;
        if (w == wa) bits &= 0xffffffffu << (xa & 31);
        if (w == wb) bits &= 0xffffffffu >> (31 - (xb & 31));
        atomicOr(row + w, bits);
    }
}

__global__ void __launch_bounds__(BC_THREADS)
bincomp_kernel(DevModel m, ObsView o, int slot, const double *__restrict__ theta, int64_t n, double *__restrict__ out,
               unsigned *__restrict__ silhouette_out)
{
    __shared__ BincompSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WarpScratch ws = warp_scratch(sm.ws, 0, 1);
    stage_model(sm.fs, m);
    __syncthreads();
    stage_frame(sm.fs, frame_view(o, slot));          // ends synchronised
    // observed silhouette: depth != 0 (dmap.elem(find(dmap)).ones(), :430-431)
    for (int w = warp; w < BC_WORDS; w += BC_THREADS / 32) {
        unsigned bits = __ballot_sync(0xffffffffu, sm.fs.depth_mm[w * 32 + lane] != 0.0f);
        if (lane == 0) sm.observed[w] = bits;
    }
    __syncthreads();

    for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
        for (int w = tid; w < BC_WORDS; w += BC_THREADS) sm.drawn[w] = 0u;
        if (warp == 0) group_fk<false, true>(sm.fs.M, [&](int) { return theta + p * HPE_NDOF; }, 1, ws, lane);
        __syncthreads();
        if (tid < HPE_NSPH) {
            // cens = K * c / c_z in the un-negated frame (:400-405), no FMA contraction
            const double x = ws.sphd[tid * 3], y = ws.sphd[tid * 3 + 1], z = ws.sphd[tid * 3 + 2];
            const int cx = trunc_to_int(__ddiv_rn(__dadd_rn(__dmul_rn(sm.fs.focal, x), __dmul_rn(sm.fs.cx, z)), z));
            const int cy = trunc_to_int(__ddiv_rn(__dadd_rn(__dmul_rn(sm.fs.focal, y), __dmul_rn(sm.fs.cy, z)), z));
            const int radius = trunc_to_int(__ddiv_rn(sm.fs.rad[tid], sm.fs.scale));      // R /= img_scale, (int) R(i)
            if (radius >= 0) {
                int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
                while (dx >= dy) {
                    set_run(sm.drawn, cy - dy, cx - dx, cx + dx);
                    set_run(sm.drawn, cy + dy, cx - dx, cx + dx);
                    set_run(sm.drawn, cy - dx, cx - dy, cx + dy);
                    set_run(sm.drawn, cy + dx, cx - dy, cx + dy);
                    dy++;
                    err += plus;
                    plus += 2;
                    if (err > 0) { err -= minus; dx -= 1; minus -= 2; }
                }
            }
        }
        __syncthreads();
        int cnt = 0;
        for (int w = tid; w < BC_WORDS; w += BC_THREADS) {
            const unsigned d = sm.drawn[w];
            cnt += __popc(d ^ sm.observed[w]);
            if (silhouette_out) silhouette_out[p * BC_WORDS + w] = d;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        if (lane == 0) sm.part[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < BC_THREADS / 32; ++w) tot += sm.part[w];
            out[p] = (double)tot;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t hpe_launch_bincomp(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int64_t n,
                               double *out_dev, unsigned *silhouette_dev, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int sms = hpe_sm_count();
    // every CTA rebuilds the frame's mask (307 KB from L2), so no more CTAs than keep the SMs busy
    int64_t grid = n < 4 * (int64_t)sms ? n : 4 * (int64_t)sms;
    bincomp_kernel<<<(int)grid, BC_THREADS, 0, st>>>(m, o, slot, theta_dev, n, out_dev, silhouette_dev);
    return cudaGetLastError();
}
