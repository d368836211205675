// hpe_obs.cu -- observedmodel::next_frame on the device (src/observedmodel.cpp:420-430):
//   cloud_kernel   load_data + depth_to_ptncloud (:272-310, :110-219): raster-order compaction of
//                  the non-zero pixels, back-projection, FP64 scale estimate, stride subsample
//   chamfer_kernel invert_depthmap + dist_transform (:313-369): OpenCV's 5x5 chamfer DIST_L2
// One CTA per frame; frames are independent.
#include <cstdlib>

#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace {

constexpr int OBS_THREADS = 1024;
constexpr int NPIX = HPE_W * HPE_H;            // 76,800

// Chunks of 32 consecutive pixels are dealt to the warps round-robin (chunk c -> warp c % 32), so the ~12 % of
// the image that is hand -- a compact blob of rows -- is spread over all 32 warps instead of keeping a few of
// them busy with the FP64 work; raster-order ranks come from a prefix sum over the 2,400 chunk counts.
constexpr int N_CHUNKS = NPIX / 32;            // 2,400

__global__ void __launch_bounds__(OBS_THREADS) cloud_kernel(ObsView o, int first)
{
    const int slot = first + blockIdx.x;
    const float *__restrict__ depth = o.depth_mm + (size_t)slot * NPIX;
    double *__restrict__ cloud = o.cloud + (size_t)slot * 3 * o.n_pts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int chunk_base[N_CHUNKS];        // count, then exclusive prefix = rank of the chunk's first hand pixel
    __shared__ unsigned chunk_bits[N_CHUNKS];
    __shared__ int warp_tot[32];
    __shared__ double red_sum[32];
    __shared__ int red_cnt[32];
    __shared__ int n_total;

    // pass 1: hand-pixel bit masks and counts per chunk; 15 loads in flight per lane (the frame comes from HBM)
    constexpr int P1_BATCH = 15;                // 75 chunks per warp = 5 batches
    for (int c0 = warp; c0 < N_CHUNKS; c0 += 32 * P1_BATCH) {
        float v[P1_BATCH];
#pragma unroll
        for (int k = 0; k < P1_BATCH; ++k) v[k] = __ldg(depth + (c0 + 32 * k) * 32 + lane);
#pragma unroll
        for (int k = 0; k < P1_BATCH; ++k) {
            unsigned b = __ballot_sync(0xffffffffu, v[k] != 0.0f);
            if (lane == 0) { chunk_bits[c0 + 32 * k] = b; chunk_base[c0 + 32 * k] = __popc(b); }
        }
    }
    // 2 / dist of the scale estimate for the distances that occur (small integers): exact quotients, computed once
    __shared__ double two_over[256];
    if (threadIdx.x < 256) two_over[threadIdx.x] = threadIdx.x ? __ddiv_rn(2.0, (double)threadIdx.x) : 0.0;
    __syncthreads();
    // the hand-pixel bit mask goes to global memory: the distance transform reads these 9.6 KB instead of the
    // 307 KB frame
    {
        unsigned *gmask = o.mask + (size_t)slot * N_CHUNKS;
        for (int c = threadIdx.x; c < N_CHUNKS; c += OBS_THREADS) gmask[c] = chunk_bits[c];
    }
    // exclusive prefix sum over the chunk counts: thread t owns chunks 3t..3t+2 (800 threads), then a block scan
    {
        int a0 = 0, a1 = 0, a2 = 0;
        const int c3 = threadIdx.x * 3;
        if (c3 < N_CHUNKS) { a0 = chunk_base[c3]; a1 = chunk_base[c3 + 1]; a2 = chunk_base[c3 + 2]; }
        int v = a0 + a1 + a2, sc = v;
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, sc, d);
            if (lane >= d) sc += t;
        }
        if (lane == 31) warp_tot[warp] = sc;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane], ws = w;
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += t;
            }
            warp_tot[lane] = ws - w;
            if (lane == 31) n_total = ws;
        }
        __syncthreads();
        const int base = warp_tot[warp] + sc - v;
        if (c3 < N_CHUNKS) { chunk_base[c3] = base; chunk_base[c3 + 1] = base + a0; chunk_base[c3 + 2] = base + a0 + a1; }
    }
    __syncthreads();
    const int n = n_total;
    if (n == 0) {
        if (threadIdx.x == 0) {
            o.meta[slot].n_hand = 0; o.meta[slot].ok = 0;
            o.meta[slot].scale = 0.0; o.meta[slot].dt_max = 0.0;
        }
        return;
    }
    // :204-217  sample_f = n / num_samples (integer), idx = k * sample_f
    const int sample_f = n / o.n_pts;
    const double rad = 2.0;
    const double rfocal = __drcp_rn(o.focal);
    double ssum = 0.0;
    int scnt = 0;
    for (int c = warp; c < N_CHUNKS; c += 32) {
        const unsigned b = chunk_bits[c];
        if (b == 0u) continue;                                                // warp-uniform
        const bool nz = (b >> lane) & 1u;
        if (nz) {
            const int pix = c * 32 + lane;
            const int rank = chunk_base[c] + __popc(b & ((1u << lane) - 1u));
            const float mm = __ldg(depth + pix);
            int v = pix / HPE_W, u = pix - v * HPE_W;
            // every quotient below is the IEEE division of the reference, obtained without a divide:
            // mm/10 (div10_exact), x/focal (div_const_exact), floor(x/Z) (floor_div, exact fallback near integers)
            double Z = div10_exact(mm);                                        // :300-303 mm -> cm
            double X = div_const_exact(__dmul_rn((double)u - o.cx, Z), o.focal, rfocal);     // :139-140
            // scale estimate :171-202, K*(X,Y,Z) with row0 = f*X + cx*Z, no FMA contraction
            const double rz = __drcp_rn(Z);
            double cu = floor_div(__dadd_rn(__dmul_rn(o.focal, X), __dmul_rn(o.cx, Z)), Z, rz);
            double eu = floor_div(__dadd_rn(__dmul_rn(o.focal, __dadd_rn(X, rad)), __dmul_rn(o.cx, Z)), Z, rz);
            double dist = fabs(eu - cu);   // the row-1 difference is identically zero
            if (dist != 0.0) { ssum += dist < 256.0 ? two_over[(int)dist] : __ddiv_rn(rad, dist); scnt++; }
            if (sample_f > 0) {
                int k = rank / sample_f;
                if (k * sample_f == rank && k < o.n_pts) {
                    double Y = div_const_exact(__dmul_rn((double)v - o.cy, Z), o.focal, rfocal);
                    cloud[k] = X; cloud[o.n_pts + k] = -Y; cloud[2 * o.n_pts + k] = -Z;
                }
            } else if (rank == 0) {
                double Y = div_const_exact(__dmul_rn((double)v - o.cy, Z), o.focal, rfocal);
                for (int k = 0; k < o.n_pts; ++k) {
                    cloud[k] = X; cloud[o.n_pts + k] = -Y; cloud[2 * o.n_pts + k] = -Z;
                }
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        ssum += __shfl_xor_sync(0xffffffffu, ssum, d);
        scnt += __shfl_xor_sync(0xffffffffu, scnt, d);
    }
    if (lane == 0) { red_sum[warp] = ssum; red_cnt[warp] = scnt; }
    __syncthreads();
    if (warp == 0) {
        double s = red_sum[lane];
        int c = red_cnt[lane];
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            c += __shfl_xor_sync(0xffffffffu, c, d);
        }
        if (lane == 0) {
            o.meta[slot].scale = s / (double)c;   // arma::mean(cmPerPixel), :202
            o.meta[slot].n_hand = n;
            o.meta[slot].ok = 1;
        }
    }
}

// Cluster order of a frame's cloud (consumed by stage_frame / match_pass, hpe_eval.cuh): the points are arranged so
// that every run of 32 consecutive positions is spatially compact -- a k-d split three levels deep: widest axis of the
// segment's bounding box, full sort along it (rank by counting, ties by point index), cut at a multiple of 32 points.
// perm[position] = point index; positions >= n_pts repeat the last point.  One CTA of 256 threads per frame.
__device__ __forceinline__ unsigned sortable_bits(float v)
{
    unsigned b = __float_as_uint(v);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);       // total order, also over -0 / NaN payloads
}

__global__ void __launch_bounds__(HPE_MAXPTS) cluster_kernel(ObsView o, int first)
{
    const int slot = first + blockIdx.x;
    unsigned char *perm = o.perm + (size_t)slot * HPE_MAXPTS;
    const int n = o.n_pts, t = threadIdx.x;
    __shared__ float P[3][HPE_MAXPTS];
    __shared__ unsigned char ord[2][HPE_MAXPTS];
    if (!o.meta[slot].ok) {                                     // empty frame (uniform): identity, never consumed
        perm[t] = (unsigned char)min(t, n - 1);
        return;
    }
    const double *cloud = o.cloud + (size_t)slot * 3 * n;
    if (t < n) {
        P[0][t] = (float)cloud[t]; P[1][t] = (float)cloud[n + t]; P[2][t] = (float)cloud[2 * n + t];
    }
    ord[0][t] = (unsigned char)t;
    __syncthreads();
    int lo = 0, hi = n, cur = 0;
    for (int level = 0; level < 3; ++level) {
        const int id = ord[cur][t < n ? t : 0];
        int dst = t;
        const bool split = t < n && hi - lo > 32;
        if (split) {
            float mn0 = P[0][id], mx0 = mn0, mn1 = P[1][id], mx1 = mn1, mn2 = P[2][id], mx2 = mn2;
            for (int u = lo; u < hi; ++u) {
                const int i = ord[cur][u];
                mn0 = fminf(mn0, P[0][i]); mx0 = fmaxf(mx0, P[0][i]);
                mn1 = fminf(mn1, P[1][i]); mx1 = fmaxf(mx1, P[1][i]);
                mn2 = fminf(mn2, P[2][i]); mx2 = fmaxf(mx2, P[2][i]);
            }
            const float e0 = mx0 - mn0, e1 = mx1 - mn1, e2 = mx2 - mn2;
            const int axis = (e0 >= e1 && e0 >= e2) ? 0 : (e1 >= e2 ? 1 : 2);
            const unsigned key = sortable_bits(P[axis][id]);
            int rank = 0;
            for (int u = lo; u < hi; ++u) {
                const int i = ord[cur][u];
                const unsigned ku = sortable_bits(P[axis][i]);
                rank += (ku < key || (ku == key && i < id)) ? 1 : 0;
            }
            dst = lo + rank;
        }
        if (t < n) ord[cur ^ 1][dst] = (unsigned char)id;
        __syncthreads();
        cur ^= 1;
        if (split) {
            const int ncl = (hi - lo + 31) >> 5;
            const int mid = lo + ((ncl + 1) >> 1) * 32;
            if (t < mid) hi = mid; else lo = mid;
        }
    }
    const unsigned char mine = ord[cur][t < n ? t : n - 1];
    perm[t] = mine;
    // the frame's staging image (hpe_eval.cuh): built once here with the code every consumer used to run per CTA, so the
    // bits (centroid, recentred points, cluster boxes) are the ones round 2's kernels computed
    __shared__ FrameSmem fs;
    __shared__ unsigned char final_perm[HPE_MAXPTS];
    final_perm[t] = mine;
    FrameView F = frame_view(o, slot);
    F.perm = final_perm;                                        // generic pointer into shared memory
    stage_frame_build(fs, F);                                   // starts and ends with a barrier
    uint4 *img = reinterpret_cast<uint4 *>(o.image + (size_t)slot * HPE_IMAGE_BYTES);
    const uint4 *src = reinterpret_cast<const uint4 *>(&fs);
    for (int i = t; i < HPE_IMAGE_BYTES / 16; i += HPE_MAXPTS) img[i] = src[i];
}

// 5x5 chamfer metric in OpenCV's Q16 fixed point: a = 65536 (axis), b = 91750 (diagonal),
// c = 143976 (knight).  Distance between two pixels with |dx|,|dy| = closed form below; the
// two-pass raster algorithm of cv::distanceTransform computes exactly min over feature
// pixels of this metric (checked against cv2 in tests/).
__device__ __forceinline__ unsigned chamfer_metric(int dx, int dy)
{
    int mx = max(dx, dy), mn = min(dx, dy);
    return 2 * mn <= mx ? (unsigned)(mn * 143976 + (mx - 2 * mn) * 65536)
                        : (unsigned)((mx - mn) * 143976 + (2 * mn - mx) * 91750);
}

// Row decomposition: h[i][j] = |j - j'| to the nearest hand pixel of row i; the metric is
// monotone in |dx| for fixed |dy|, so DT(i,j) = min_i' metric(h[i'][j], |i-i'|), searched
// outwards from row i and cut off once |i-i'| * a >= best.
__global__ void __launch_bounds__(OBS_THREADS) chamfer_kernel(ObsView o, int first)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned short *h = reinterpret_cast<unsigned short *>(smem_raw);   // [240][320]
    __shared__ float red_max[32];
    const int slot = first + blockIdx.x;
    if (!o.meta[slot].ok) return;
    const float *__restrict__ depth = o.depth_mm + (size_t)slot * NPIX;
    float *__restrict__ dt = o.dt_px + (size_t)slot * NPIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // 1-D distance along each row: a warp owns rows warp, warp+32, ...; lanes own 10-pixel runs
    for (int i = warp; i < HPE_H; i += 32) {
        const float *row = depth + i * HPE_W;
        // nearest hand pixel at or left of each position: scan over the row in 10 chunks of 32
        int carry = -100000;   // column of the last hand pixel seen so far
        for (int c = 0; c < HPE_W / 32; ++c) {
            int j = c * 32 + lane;
            bool nz = row[j] != 0.0f;
            unsigned b = __ballot_sync(0xffffffffu, nz);
            unsigned le = b & (0xffffffffu >> (31 - lane));        // hand pixels at lanes <= mine
            int left = le ? c * 32 + (31 - __clz(le)) : carry;
            int dl = j - left;
            h[i * HPE_W + j] = (unsigned short)min(dl, 65535);
            if (b) carry = c * 32 + (31 - __clz(b));
        }
        __syncwarp();
        carry = 100000;
        for (int c = HPE_W / 32 - 1; c >= 0; --c) {
            int j = c * 32 + lane;
            bool nz = row[j] != 0.0f;
            unsigned b = __ballot_sync(0xffffffffu, nz);
            unsigned ge = b & (0xffffffffu << lane);
            int right = ge ? c * 32 + (__ffs(ge) - 1) : carry;
            int dr = right - j;
            int cur = h[i * HPE_W + j];
            h[i * HPE_W + j] = (unsigned short)min(min(dr, cur), 65535);
            if (b) carry = c * 32 + (__ffs(b) - 1);
        }
    }
    __syncthreads();

    float vmax = 0.f;
    for (int pix = threadIdx.x; pix < NPIX; pix += OBS_THREADS) {
        int i = pix / HPE_W, j = pix - i * HPE_W;
        unsigned best = 0xffffffffu;
        int h0 = h[pix];
        if (h0 == 0) {
            best = 0;
        } else {
            if (h0 < 65535) best = chamfer_metric(h0, 0);
            for (int dy = 1; dy < HPE_H; ++dy) {
                if ((unsigned)dy * 65536u >= best) break;
                int up = i - dy, dn = i + dy;
                if (up < 0 && dn >= HPE_H) break;
                if (up >= 0) {
                    int hh = h[up * HPE_W + j];
                    if (hh < 65535) best = min(best, chamfer_metric(hh, dy));
                }
                if (dn < HPE_H) {
                    int hh = h[dn * HPE_W + j];
                    if (hh < 65535) best = min(best, chamfer_metric(hh, dy));
                }
            }
        }
        // OpenCV: d[j] = (float)(t0 * scale), unsigned -> float (RN) then * 2^-16
        float v = __uint2float_rn(best) * (1.f / 65536.f);
        dt[pix] = v;
        vmax = fmaxf(vmax, v);
    }
    for (int d = 16; d > 0; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    if (lane == 0) red_max[warp] = vmax;
    __syncthreads();
    if (warp == 0) {
        float v = red_max[lane];
        for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
        if (lane == 0) o.meta[slot].dt_max = (double)v;
    }
}

// ---------------------------------------------------------------------------------------
// The same transform as OpenCV runs it: forward and backward raster pass with the 5x5 mask,
// parallelised as a skewed wavefront.  Pixel (i,j) of the forward pass needs (i-1, j-2..j+2),
// (i-2, j-1), (i-2, j+1), (i, j-1): with lane r on row i0+r and column t - 3r at step t all of
// them are final.  One warp walks a frame in bands of 32 rows; the band lives in shared memory
// (row pitch 324 words -> the 32 lanes hit 32 different banks); the forward result of the whole
// frame is parked in the dt_px output buffer (as uint32) and overwritten by the backward pass.
// ---------------------------------------------------------------------------------------
constexpr int WF_PITCH = HPE_W + 4;
constexpr int WF_ROWS = 34;                       // 32 band rows + 2 halo rows
constexpr unsigned WF_MAX = 0x7fffffffu >> 2;     // OpenCV's DIST_MAX
constexpr unsigned WF_A = 65536u, WF_B = 91750u, WF_C = 143976u;
constexpr int WF_SMEM_PER_WARP = WF_ROWS * WF_PITCH * 4 + 32 * 10 * 4;
constexpr int WF_WARPS = 5;                       // 5 x 45,344 B = 226,720 B of the 227 KB a CTA may use

__global__ void __launch_bounds__(WF_WARPS * 32) chamfer_wavefront_kernel(ObsView o, int first, int count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = blockIdx.x * WF_WARPS + warp;
    if (fr >= count) return;
    const int slot = first + fr;
    if (!o.meta[slot].ok) return;
    unsigned *S = reinterpret_cast<unsigned *>(smem_raw + (size_t)warp * WF_SMEM_PER_WARP);
    unsigned *mask = S + WF_ROWS * WF_PITCH;                     // [32][10] bit rows: 1 = hand pixel
    const float *__restrict__ depth = o.depth_mm + (size_t)slot * NPIX;
    unsigned *tmp = reinterpret_cast<unsigned *>(o.dt_px + (size_t)slot * NPIX);
    float *out = o.dt_px + (size_t)slot * NPIX;
    const int NB = (HPE_H + 31) / 32;

    // ---------------- forward pass, bands top to bottom ----------------
    for (int b = 0; b < NB; ++b) {
        const int i0 = b * 32;
        // halo columns and halo rows
        for (int rr = lane; rr < WF_ROWS; rr += 32) {
            S[rr * WF_PITCH + 0] = WF_MAX; S[rr * WF_PITCH + 1] = WF_MAX;
            S[rr * WF_PITCH + HPE_W + 2] = WF_MAX; S[rr * WF_PITCH + HPE_W + 3] = WF_MAX;
        }
        for (int hr = 0; hr < 2; ++hr) {
            int gi = i0 - 2 + hr;
            for (int c = lane; c < HPE_W; c += 32) S[hr * WF_PITCH + 2 + c] = gi >= 0 ? tmp[gi * HPE_W + c] : WF_MAX;
        }
        for (int rr = 0; rr < 32; ++rr) {
            int gi = i0 + rr;
            for (int c = 0; c < HPE_W / 32; ++c) {
                bool hand = gi < HPE_H && depth[gi * HPE_W + c * 32 + lane] != 0.0f;
                unsigned bits = __ballot_sync(0xffffffffu, hand);
                if (lane == 0) mask[rr * 10 + c] = bits;
            }
        }
        __syncwarp();
        const bool live = i0 + lane < HPE_H;
        unsigned *row = S + (lane + 2) * WF_PITCH + 2;           // my row, column 0
        unsigned mbits = mask[lane * 10];
        for (int t = 0; t < HPE_W + 3 * 31; ++t) {
            int j = t - 3 * lane;
            if (live && j >= 0 && j < HPE_W) {
                // all eight neighbours are final: load them together, reduce as a tree
                const unsigned *p = row + j;
                unsigned a0 = p[-2 * WF_PITCH - 1], a1 = p[-2 * WF_PITCH + 1], a2 = p[-WF_PITCH - 2], a3 = p[-WF_PITCH - 1];
                unsigned a4 = p[-WF_PITCH], a5 = p[-WF_PITCH + 1], a6 = p[-WF_PITCH + 2], a7 = p[-1];
                unsigned m01 = min(a0, a1), m26 = min(a2, a6), m35 = min(a3, a5), m47 = min(a4, a7);
                unsigned v = min(min(min(m01, m26) + WF_C, m35 + WF_B), m47 + WF_A);
                bool hand = (mbits >> (j & 31)) & 1u;
                row[j] = hand ? 0u : v;
                if ((j & 31) == 31) mbits = mask[lane * 10 + min((j >> 5) + 1, 9)];
            }
            __syncwarp();
        }
        for (int rr = 0; rr < 32 && i0 + rr < HPE_H; ++rr)
            for (int c = lane; c < HPE_W; c += 32) tmp[(i0 + rr) * HPE_W + c] = S[(rr + 2) * WF_PITCH + 2 + c];
        __syncwarp();
    }
    __threadfence_block();
    // ---------------- backward pass, bands bottom to top ----------------
    float vmax = 0.f;
    for (int b = NB - 1; b >= 0; --b) {
        const int i0 = b * 32;
        const int rows = min(32, HPE_H - i0);
        // smem row rr (0..31) = image row i0 + rr; rows 32,33 = halo below (i0+32, i0+33)
        for (int rr = lane; rr < WF_ROWS; rr += 32) {
            S[rr * WF_PITCH + 0] = WF_MAX; S[rr * WF_PITCH + 1] = WF_MAX;
            S[rr * WF_PITCH + HPE_W + 2] = WF_MAX; S[rr * WF_PITCH + HPE_W + 3] = WF_MAX;
        }
        for (int rr = 0; rr < WF_ROWS; ++rr) {
            int gi = i0 + rr;
            // rows below this band already hold final FLOAT values: convert back to Q16 exactly
            // (a final value is k * 2^-16 with k < 2^25; k survives the float only when k < 2^24,
            // so the backward pass keeps its own exact copy instead, see `fin` below)
            for (int c = lane; c < HPE_W; c += 32)
                S[rr * WF_PITCH + 2 + c] = (gi < HPE_H && rr < rows) ? tmp[gi * HPE_W + c] : WF_MAX;
        }
        __syncwarp();
        // halo rows below: exact Q16 values of the previous (lower) band, kept in global scratch rows
        if (b < NB - 1) {
            for (int hr = 0; hr < 2; ++hr)
                for (int c = lane; c < HPE_W; c += 32)
                    S[(rows + hr) * WF_PITCH + 2 + c] = o.wf_halo[((size_t)slot * 2 + hr) * HPE_W + c];
        }
        __syncwarp();
        const bool live = lane < rows;
        unsigned *row = S + (rows - 1 - lane) * WF_PITCH + 2;    // lane 0 = bottom row of the band
        for (int t = 0; t < HPE_W + 3 * 31; ++t) {
            int j = HPE_W - 1 - (t - 3 * lane);
            if (live && j >= 0 && j < HPE_W) {
                unsigned *p = row + j;
                unsigned a0 = p[2 * WF_PITCH + 1], a1 = p[2 * WF_PITCH - 1], a2 = p[WF_PITCH + 2], a3 = p[WF_PITCH + 1];
                unsigned a4 = p[WF_PITCH], a5 = p[WF_PITCH - 1], a6 = p[WF_PITCH - 2], a7 = p[1], own = p[0];
                unsigned m01 = min(a0, a1), m26 = min(a2, a6), m35 = min(a3, a5), m47 = min(a4, a7);
                p[0] = min(own, min(min(min(m01, m26) + WF_C, m35 + WF_B), m47 + WF_A));
            }
            __syncwarp();
        }
        // the two top rows of this band are the halo of the band above
        for (int hr = 0; hr < 2; ++hr)
            for (int c = lane; c < HPE_W; c += 32)
                o.wf_halo[((size_t)slot * 2 + hr) * HPE_W + c] = S[hr * WF_PITCH + 2 + c];
        for (int rr = 0; rr < rows; ++rr)
            for (int c = lane; c < HPE_W; c += 32) {
                unsigned v = min(S[rr * WF_PITCH + 2 + c], WF_MAX);
                float f = __uint2float_rn(v) * (1.f / 65536.f);   // OpenCV: (float)(t0 * scale)
                out[(i0 + rr) * HPE_W + c] = f;
                vmax = fmaxf(vmax, f);
            }
        __syncwarp();
    }
    for (int d = 16; d > 0; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    if (lane == 0) o.meta[slot].dt_max = (double)vmax;
}

// ---------------------------------------------------------------------------------------
// The same two raster passes, one ROW at a time with the whole row in flight (default path).
// Forward pass, row i:  F(i,j) = min(U(i,j), F(i,j-1) + a)  where U gathers the seven mask
// neighbours of rows i-1 and i-2, which are final.  Unrolled along the row this is
//   F(i,j) = min_{k <= j} ( U(i,k) + (j-k) a ) = j a + prefix_min_k ( U(i,k) - k a ),
// a prefix minimum -- so a warp owns a frame, a lane owns 10 consecutive columns, and a row costs a
// short in-lane scan plus a 5-step shuffle scan instead of 320 dependent steps.  The backward pass
// is the mirror image (suffix minimum, rows bottom-up).  The previous two rows live in registers;
// nothing is staged in shared memory, so 16+ frames are in flight per SM and the kernel streams
// 3 x 307 KB per frame (forward result out and back in, float result out) plus the 9.6 KB hand-pixel mask
// cloud_kernel leaves behind -- the frame itself is read once, by cloud_kernel.
// Values that no feature pixel reaches stay at OpenCV's DIST_MAX-like sentinel (clamped to WF_MAX
// per row); every such value is replaced by the backward pass when the frame has a hand pixel, so
// the result is bit-identical to cv::distanceTransform (tests compare against cv2 and the oracle).
// ---------------------------------------------------------------------------------------
constexpr int SC_WARPS = 8;
constexpr int SC_CPL = HPE_W / 32;               // 10 columns per lane

// U(i,j) for the lane's 10 columns from the two previous rows (+2 / +1 halo columns on each side)
__device__ __forceinline__ void sc_gather(const unsigned (&p1)[SC_CPL], const unsigned (&p2)[SC_CPL], unsigned l1a, unsigned l1b,
                                          unsigned r1a, unsigned r1b, unsigned l2, unsigned r2, unsigned (&u)[SC_CPL])
{
    // e1[k] = row-1 value at local column k-2 (k = 0..13), e2[k] = row-2 value at local column k-1 (k = 0..11)
    unsigned e1[SC_CPL + 4], e2[SC_CPL + 2];
    e1[0] = l1a; e1[1] = l1b; e1[SC_CPL + 2] = r1a; e1[SC_CPL + 3] = r1b;
    e2[0] = l2; e2[SC_CPL + 1] = r2;
#pragma unroll
    for (int k = 0; k < SC_CPL; ++k) { e1[k + 2] = p1[k]; e2[k + 1] = p2[k]; }
#pragma unroll
    for (int k = 0; k < SC_CPL; ++k) {
        unsigned kn = min(min(e2[k], e2[k + 2]), min(e1[k], e1[k + 4])) + WF_C;     // knight moves
        unsigned dg = min(e1[k + 1], e1[k + 3]) + WF_B;                              // diagonals
        unsigned ax = e1[k + 2] + WF_A;                                              // axial
        u[k] = min(min(kn, dg), ax);
    }
}

__global__ void __launch_bounds__(SC_WARPS * 32) chamfer_scan_kernel(ObsView o, int first, int count)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = blockIdx.x * (blockDim.x >> 5) + warp;
    if (fr >= count) return;
    const int slot = first + fr;
    if (!o.meta[slot].ok) return;
    unsigned *tmp = reinterpret_cast<unsigned *>(o.dt_px + (size_t)slot * NPIX);
    float *out = o.dt_px + (size_t)slot * NPIX;
    const int c0 = lane * SC_CPL;
    const int A = (int)WF_A;

    // ---------------- forward pass, rows top to bottom ----------------
    unsigned p1[SC_CPL], p2[SC_CPL];
#pragma unroll
    for (int k = 0; k < SC_CPL; ++k) { p1[k] = WF_MAX; p2[k] = WF_MAX; }
    // hand bits of the lane's 10 pixels of a row: bits 10 lane .. 10 lane + 9 of the row's 320-bit mask (10 words
    // written by cloud_kernel), i.e. a funnel shift over two neighbouring words
    const unsigned *__restrict__ gmask = o.mask + (size_t)slot * (NPIX / 32);
    const int mw = c0 >> 5, msh = c0 & 31;
    unsigned wlo = __ldg(gmask + mw), whi = __ldg(gmask + min(mw + 1, 9));
    for (int i = 0; i < HPE_H; ++i) {
        const unsigned hand = __funnelshift_r(wlo, whi, msh) & 0x3ffu;
        if (i + 1 < HPE_H) {
            wlo = __ldg(gmask + (i + 1) * 10 + mw);
            whi = __ldg(gmask + (i + 1) * 10 + min(mw + 1, 9));
        }
        // halo columns of the two previous rows from the neighbouring lanes (image border: sentinel)
        unsigned l1a = __shfl_up_sync(0xffffffffu, p1[SC_CPL - 2], 1), l1b = __shfl_up_sync(0xffffffffu, p1[SC_CPL - 1], 1);
        unsigned r1a = __shfl_down_sync(0xffffffffu, p1[0], 1), r1b = __shfl_down_sync(0xffffffffu, p1[1], 1);
        unsigned l2 = __shfl_up_sync(0xffffffffu, p2[SC_CPL - 1], 1), r2 = __shfl_down_sync(0xffffffffu, p2[0], 1);
        if (lane == 0) { l1a = WF_MAX; l1b = WF_MAX; l2 = WF_MAX; }
        if (lane == 31) { r1a = WF_MAX; r1b = WF_MAX; r2 = WF_MAX; }
        unsigned u[SC_CPL];
        sc_gather(p1, p2, l1a, l1b, r1a, r1b, l2, r2, u);
#pragma unroll
        for (int k = 0; k < SC_CPL; ++k)
            if ((hand >> k) & 1u) u[k] = 0u;
        // in-lane scan with no carry, then the carry from the lanes to the left
#pragma unroll
        for (int k = 1; k < SC_CPL; ++k) u[k] = min(u[k], u[k - 1] + WF_A);
        int s = (int)u[SC_CPL - 1] - SC_CPL * (lane + 1) * A;          // segment result in prefix-min coordinates
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s = min(s, t);
        }
        int carry = __shfl_up_sync(0xffffffffu, s, 1);                  // exclusive: minimum over the lanes to the left
        if (lane > 0) {
            carry += SC_CPL * lane * A;                                  // = F(i, c0 - 1)
#pragma unroll
            for (int k = 0; k < SC_CPL; ++k) u[k] = min(u[k], (unsigned)(carry + (k + 1) * A));
        }
#pragma unroll
        for (int k = 0; k < SC_CPL; ++k) {
            u[k] = min(u[k], WF_MAX);
            p2[k] = p1[k];
            p1[k] = u[k];
        }
        uint2 *dst = reinterpret_cast<uint2 *>(tmp + (size_t)i * HPE_W + c0);
#pragma unroll
        for (int k = 0; k < SC_CPL / 2; ++k) dst[k] = make_uint2(u[2 * k], u[2 * k + 1]);
    }
    // ---------------- backward pass, rows bottom to top, columns right to left ----------------
    float vmax = 0.f;
#pragma unroll
    for (int k = 0; k < SC_CPL; ++k) { p1[k] = WF_MAX; p2[k] = WF_MAX; }
    uint2 fn[SC_CPL / 2];
    {
        const uint2 *src = reinterpret_cast<const uint2 *>(tmp + (size_t)(HPE_H - 1) * HPE_W + c0);
#pragma unroll
        for (int k = 0; k < SC_CPL / 2; ++k) fn[k] = src[k];
    }
    for (int i = HPE_H - 1; i >= 0; --i) {
        unsigned f[SC_CPL];
#pragma unroll
        for (int k = 0; k < SC_CPL / 2; ++k) { f[2 * k] = fn[k].x; f[2 * k + 1] = fn[k].y; }
        if (i > 0) {
            const uint2 *src = reinterpret_cast<const uint2 *>(tmp + (size_t)(i - 1) * HPE_W + c0);
#pragma unroll
            for (int k = 0; k < SC_CPL / 2; ++k) fn[k] = src[k];      // own columns only: never written by another lane
        }
        unsigned l1a = __shfl_up_sync(0xffffffffu, p1[SC_CPL - 2], 1), l1b = __shfl_up_sync(0xffffffffu, p1[SC_CPL - 1], 1);
        unsigned r1a = __shfl_down_sync(0xffffffffu, p1[0], 1), r1b = __shfl_down_sync(0xffffffffu, p1[1], 1);
        unsigned l2 = __shfl_up_sync(0xffffffffu, p2[SC_CPL - 1], 1), r2 = __shfl_down_sync(0xffffffffu, p2[0], 1);
        if (lane == 0) { l1a = WF_MAX; l1b = WF_MAX; l2 = WF_MAX; }
        if (lane == 31) { r1a = WF_MAX; r1b = WF_MAX; r2 = WF_MAX; }
        unsigned u[SC_CPL];
        sc_gather(p1, p2, l1a, l1b, r1a, r1b, l2, r2, u);              // the mask is symmetric: same gather on the rows below
#pragma unroll
        for (int k = 0; k < SC_CPL; ++k) u[k] = min(u[k], f[k]);
        // suffix scan: B(i,j) = min(V(i,j), B(i,j+1) + a)
#pragma unroll
        for (int k = SC_CPL - 2; k >= 0; --k) u[k] = min(u[k], u[k + 1] + WF_A);
        int s = (int)u[0] - SC_CPL * (32 - lane) * A;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_down_sync(0xffffffffu, s, d);
            if (lane + d < 32) s = min(s, t);
        }
        int carry = __shfl_down_sync(0xffffffffu, s, 1);                // minimum over the lanes to the right
        if (lane < 31) {
            carry += SC_CPL * (31 - lane) * A;                           // = B(i, c0 + 10)
#pragma unroll
            for (int k = 0; k < SC_CPL; ++k) u[k] = min(u[k], (unsigned)(carry + (SC_CPL - k) * A));
        }
        float fo[SC_CPL];
#pragma unroll
        for (int k = 0; k < SC_CPL; ++k) {
            u[k] = min(u[k], WF_MAX);
            p2[k] = p1[k];
            p1[k] = u[k];
            fo[k] = __uint2float_rn(u[k]) * (1.f / 65536.f);            // OpenCV: (float)(t0 * scale)
            vmax = fmaxf(vmax, fo[k]);
        }
        float2 *dst = reinterpret_cast<float2 *>(out + (size_t)i * HPE_W + c0);
#pragma unroll
        for (int k = 0; k < SC_CPL / 2; ++k) dst[k] = make_float2(fo[2 * k], fo[2 * k + 1]);
    }
    for (int d = 16; d > 0; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    if (lane == 0) o.meta[slot].dt_max = (double)vmax;
}

// Folds the depth map into the distance transform: dt_px[p] = -depth_mm[p] on hand pixels (where the transform is 0),
// the transform elsewhere.  The depth term then needs ONE gather per projected sphere (v < 0: the observed depth is -v,
// otherwise v is the distance to the silhouette) instead of two from two 307 KB maps.  Depth values are >= 0 (the file
// format is unsigned millimetres).  A warp per mask word: only the ~12 % of the words that hold hand pixels are touched.
__global__ void __launch_bounds__(256) fold_depth_kernel(ObsView o, int first)
{
    const int slot = first + blockIdx.x;
    const unsigned *__restrict__ gmask = o.mask + (size_t)slot * (NPIX / 32);
    const float *__restrict__ depth = o.depth_mm + (size_t)slot * NPIX;
    float *__restrict__ dt = o.dt_px + (size_t)slot * NPIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 2,400 mask words = 75 chunks of 32: a warp loads its chunks' words at once (lane = word), then visits the non-empty
    // words of a chunk (ballot) with lane = pixel
    constexpr int CH = (NPIX / 32) / 32;       // 75
    unsigned mine[(CH + 7) / 8];
#pragma unroll
    for (int i = 0; i < (CH + 7) / 8; ++i) {
        const int ch = warp + 8 * i;
        mine[i] = ch < CH ? __ldg(gmask + ch * 32 + lane) : 0u;
    }
#pragma unroll
    for (int i = 0; i < (CH + 7) / 8; ++i) {
        const int ch = warp + 8 * i;
        unsigned todo = __ballot_sync(0xffffffffu, mine[i] != 0u);
        while (todo) {                          // four words per round: the loads of a round are in flight together
            int pp[4];
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int src = todo ? __ffs(todo) - 1 : 0;
                const unsigned w = todo ? __shfl_sync(0xffffffffu, mine[i], src) : 0u;
                todo &= todo - 1u;              // 0 stays 0
                pp[u] = ((w >> lane) & 1u) ? (ch * 32 + src) * 32 + lane : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = pp[u] >= 0 ? __ldg(depth + pp[u]) : 0.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (pp[u] >= 0) dt[pp[u]] = -v[u];
        }
    }
}
// the same for one slot whose maps were injected by the caller (no mask)
__global__ void __launch_bounds__(256) fold_depth_scan_kernel(ObsView o, int slot)
{
    const float *__restrict__ depth = o.depth_mm + (size_t)slot * NPIX;
    float *__restrict__ dt = o.dt_px + (size_t)slot * NPIX;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < NPIX; p += gridDim.x * blockDim.x)
        if (depth[p] != 0.0f) dt[p] = -depth[p];
}

// disttrans.max() for a slot whose distance transform was injected by the caller
__global__ void __launch_bounds__(1024) dtmax_kernel(ObsView o, int slot)
{
    __shared__ float red_max[32];
    const float *dt = o.dt_px + (size_t)slot * NPIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float vmax = dt[0];
    for (int pix = threadIdx.x; pix < NPIX; pix += blockDim.x) vmax = fmaxf(vmax, dt[pix]);
    for (int d = 16; d > 0; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    if (lane == 0) red_max[warp] = vmax;
    __syncthreads();
    if (warp == 0) {
        float v = red_max[lane];
        for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
        if (lane == 0) o.meta[slot].dt_max = (double)v;
    }
}

}  // namespace

size_t hpe_observe_scratch_bytes(int) { return 0; }
int g_hpe_chamfer = 0;   // 0 = row-parallel two-pass (prefix-min scans, default), 1 = row-search closed form, 2 = skewed wavefront (cross-checks)

cudaError_t hpe_launch_observe(const ObsView &o, int first, int count, void *, cudaStream_t st)
{
    if (count <= 0) return cudaSuccess;
    static PerDevice<bool> attr_pd;
    bool &attr = attr_pd.get();
    const int smem = HPE_W * HPE_H * (int)sizeof(unsigned short);
    if (!attr) {
        cudaFuncSetAttribute(chamfer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr = true;
    }
    cloud_kernel<<<count, OBS_THREADS, 0, st>>>(o, first);
    cluster_kernel<<<count, HPE_MAXPTS, 0, st>>>(o, first);
    if (g_hpe_chamfer == 1) {
        chamfer_kernel<<<count, OBS_THREADS, smem, st>>>(o, first);
    } else if (g_hpe_chamfer == 0) {
        // a warp owns a frame; every warp slot of the GPU is used.  Tried and dropped (tools/observe_probe.py, 2,368 frames,
        // 2.04M frames/s as is): capping the frames in flight so that the forward rows stay in the L2 until the backward
        // pass reads them (2 / 4 / 8 warps per SM: 1.09M / 1.45M / 1.82M -- too little latency hiding), and running the
        // transform of one chunk of the batch under the cloud kernel of the next on a second stream (1.31M: the chunks
        // are too small to fill the GPU)
        chamfer_scan_kernel<<<(count + SC_WARPS - 1) / SC_WARPS, SC_WARPS * 32, 0, st>>>(o, first, count);
    } else {
        static PerDevice<bool> attr2_pd;
        bool &attr2 = attr2_pd.get();
        if (!attr2) {
            cudaFuncSetAttribute(chamfer_wavefront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 WF_WARPS * WF_SMEM_PER_WARP);
            attr2 = true;
        }
        chamfer_wavefront_kernel<<<(count + WF_WARPS - 1) / WF_WARPS, WF_WARPS * 32, WF_WARPS * WF_SMEM_PER_WARP, st>>>(
            o, first, count);
    }
    fold_depth_kernel<<<count, 256, 0, st>>>(o, first);
    return cudaGetLastError();
}

cudaError_t hpe_launch_fold_depth(const ObsView &o, int slot, cudaStream_t st)
{
    fold_depth_scan_kernel<<<75, 256, 0, st>>>(o, slot);
    return cudaGetLastError();
}

cudaError_t hpe_launch_cluster(const ObsView &o, int first, int count, cudaStream_t st)
{
    if (count <= 0) return cudaSuccess;
    cluster_kernel<<<count, HPE_MAXPTS, 0, st>>>(o, first);
    return cudaGetLastError();
}

cudaError_t hpe_launch_dtmax(const ObsView &o, int slot, cudaStream_t st)
{
    dtmax_kernel<<<1, 1024, 0, st>>>(o, slot);
    return cudaGetLastError();
}
