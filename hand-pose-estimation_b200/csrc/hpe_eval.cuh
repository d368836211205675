// hpe_eval.cuh -- the warp-level evaluation of costfunc::cal_cost / cal_cost2
// (src/costfunc.cpp:31-127) for a group of poses (up to HPE_GROUP = 5; six in the scoring kernel) against one frame.
//
// Work split inside a warp:
//   phase 1  lanes = (pose, digit): forward kinematics in FP64 -> FP64 sphere centres in shared memory
//   phase 2  per pose, lanes = spheres: projection + ONE gather per pixel from the folded distance / depth map; the
//            same lanes stage the pose's float32 centres for the matcher
//   phase 3  per pose: which sphere blocks each cluster of 32 points has to look at (prune_blocks), clusters ranked by
//            list length and dealt to the (pass, quarter) slots; then lanes = points (two passes of 4 points per lane):
//            nearest-centre search in FP32 and the FP64 alignment sum of those points
//   phase 4  (cal_cost2 only) lanes = sphere pairs: self-collision
// Everything a warp needs between phases lives in shared memory (model, frame scalars and the frame's staging image:
// cloud, recentred points, cluster boxes), so the kernels stay at 80 registers and 24 warps fit per SM (3 CTAs of 8, or
// the scoring kernel's one CTA of 24): the FP64 / gather phases of some warps hide under the FP32 sweeps of the others.
//
// The nearest-centre search (compute_correspondences, src/costfunc.cpp:306-343) must return the
// index cv::BFMatcher returns: first minimum over j of sqrtf(d_j), d_j = ((t0^2 + t1^2) + t2^2) in
// float32, t = p - c_j.  Doing that directly costs 8 FP32-pipe + 5 ALU-pipe instructions per
// (point, sphere) pair.  match_pass() gets the same index with ~3:
//   sweep A  s_j = |c'_j|^2 - 2 p'.c'_j (3 FMAs, coordinates recentred on the cloud centroid q) for
//            all 48 spheres, issued as packed fma.rn.f32x2 on sphere PAIRS (the same IEEE fma in each
//            half; 64-bit operands are fetched without the register-bank conflicts that make the scalar
//            three-register FFMAs of this loop average 1.5 issue cycles), reduced to the minimum per block
//            of 4 spheres, keeping the best block's value M1 and the runner-up block value M2.  The block
//            index rides in the 4 low mantissa bits of the block minimum (no compare + select; a
//            perturbation below 16 ulp) and two blocks per step share the runner-up update;
//   sweep B  the 4 spheres of the best block once more (same FMAs, bit-identical s): winner index,
//            the exact block minimum m1 and the runner-up r inside the block;
//   certify  s_j + |p'|^2 equals the exact squared distance D_j up to E_j <= u(104|p'|^2 + 13|s_j|),
//            u = 2^-24 (forward error of the 3 FMAs plus the rounding of p', -2c', |c'|^2), and the
//            reference's d_j / sqrtf(d_j) order two spheres like D_j once D differs by more than 24u D
//            (ties and sqrtf merges need smaller differences); M2 carries an index in its low bits,
//            |M2 - true| < 32u|M2|.  Hence if min(M2, r) - m1 > tau = u(260|p'|^2 + 116|m1|) EVERY
//            other sphere is provably farther in the reference's own arithmetic and the winner is the
//            reference's answer;
//   fallback otherwise (5-6 points in 10,000) the point is redone with the exact restatement of the
//            reference's arithmetic over all 48 spheres, the whole warp on one point.
// Compile-time switches kept for cross-checks and measurements: HPE_FFMA2=0 (scalar FFMA sweeps),
// HPE_PACKED_INDEX=0 (compare + select for the best block), HPE_PAIR_BLOCKS=0.
#pragma once
#include <cstddef>
#include "hpe_device.cuh"

// diagnostics: points sent to the exact fallback (hpe_debug_fallback_points); one copy per
// translation unit, the host getters are instantiated next to the kernels
static __device__ unsigned long long g_hpe_fallback_points;
#define HPE_DEFINE_FALLBACK_GETTER(name)                                                          \
    unsigned long long name(int reset)                                                            \
    {                                                                                             \
        unsigned long long v = 0, z = 0;                                                          \
        cudaMemcpyFromSymbol(&v, g_hpe_fallback_points, sizeof v);                                \
        if (reset) cudaMemcpyToSymbol(g_hpe_fallback_points, &z, sizeof z);                       \
        return v;                                                                                 \
    }

#ifndef HPE_SWEEP_STAGED
#define HPE_SWEEP_STAGED 1
#endif
#ifndef HPE_PACKED_INDEX
#define HPE_PACKED_INDEX 1   // 1: block index packed in the low mantissa bits of the block minimum (+1.4 % with the packed sweeps)
#endif
#ifndef HPE_FFMA2
#define HPE_FFMA2 1          // 1: sweeps on packed fma.rn.f32x2 (two spheres per instruction, conflict-free operand fetch)
#endif
#ifndef HPE_PAIR_BLOCKS
#define HPE_PAIR_BLOCKS 1      // 1: two sweep-A blocks per step share the runner-up update (needs HPE_PACKED_INDEX)
#endif
#ifndef HPE_SWEEP_UNROLL
#define HPE_SWEEP_UNROLL 2   // sweep-A steps per loop body of the unpruned variants and, by default, of the pruned sweep (UNROLL of
                             // match_align: the throughput kernel asks for 1 -- the ranked lists of pass 1 are 2-3 steps long and a
                             // two-step body plus its remainder copy measured 1.6 % slower there; a warp that has its scheduler
                             // to itself, as in the one-particle-per-warp kernels, wants the two independent steps)
#endif
#define HPE_STR2(x) #x
#define HPE_STR(x) HPE_STR2(x)
#define HPE_UNROLL_SWEEP _Pragma(HPE_STR(unroll HPE_SWEEP_UNROLL))
#ifndef HPE_MINBLOCKS
#define HPE_MINBLOCKS 3                  // CTAs of 8 warps per SM the kernels are compiled for
#endif
#define HPE_BLK 4                        // spheres per sweep-A block
#define HPE_NBLK (HPE_NSPH / HPE_BLK)    // 12 blocks
#ifndef HPE_PASS
#define HPE_PASS 4
#endif                                   // points per lane and pass (2 passes cover 8 x 32 = 256 points)

// per-warp shared-memory scratch
struct WarpScratch {
    double *dep;    // [8] depth terms of the group's poses (group_depth -> the per-pose loop); [6], [7]: the cost kernel's loop variables
    unsigned *slots; // [8] (cluster << 12 | block mask) of the cluster matched in slot 4 pass + quarter (match_align)
    double *sphd;   // [slot][48][3] un-negated FP64 centres of the group's poses
    float4 *cf;     // [48] (x,-y,-z,0): the float32 centres cv::BFMatcher sees, pose being matched
    float4 *ex;     // [12][5] (-2c'x, -2c'y, -2c'z, |c'|^2), c' = cf - q, pose being matched; blocks of 4 spheres
                    // are 80 B apart (one pad entry) so that sweep B's per-lane block reads spread over the banks
};
#define HPE_EX_PITCH (HPE_BLK + 1)
#define HPE_EX_AT(j) (((j) / HPE_BLK) * HPE_EX_PITCH + ((j) % HPE_BLK))
#define HPE_EX_BLOCKS (HPE_NBLK + 1)      // + the "no block" entry a lane sweeps once its cluster's list is exhausted
#define HPE_WS_BYTES(slots) ((slots) * HPE_NSPH * 24 + HPE_NSPH * 16 + HPE_EX_BLOCKS * HPE_EX_PITCH * 16 + 64 + 32)

__device__ __forceinline__ WarpScratch warp_scratch(unsigned char *base, int warp, int slots)
{
    WarpScratch w;
    unsigned char *p = base + (size_t)warp * HPE_WS_BYTES(slots);
    w.cf = reinterpret_cast<float4 *>(p);
    w.ex = w.cf + HPE_NSPH;
    w.dep = reinterpret_cast<double *>(p + HPE_NSPH * 16 + HPE_EX_BLOCKS * HPE_EX_PITCH * 16);
    w.slots = reinterpret_cast<unsigned *>(w.dep + 8);
    w.sphd = w.dep + 12;
    return w;
}

// Sphere slots of the matcher: the 48 centres are regrouped into 12 spatially compact blocks of 4 (the sweep's unit of
// work and of pruning) -- thumb 0-3 | 4-7, the 2 x 2 palm patches of (index, middle) and (ring, little), each finger's
// proximal + middle pair, the fingertips of (index, middle) and (ring, little).  Any grouping returns the reference's
// index: exact ties never pass the certificate and are decided by the exact fallback in index order.
#define HPE_NCL (HPE_MAXPTS / 32)        // point clusters of 32
// Storage position of member i of cluster c.  A warp matches 128 positions per pass, position k0 + 32 m with
// k0 = lane + 128 pass for m = 0..3; cluster c = 4 pass + q is given to the eight lanes 8q..8q+7 of that pass (four
// members each), so that every quarter of the warp sweeps the block list of ONE cluster.
__device__ __forceinline__ int cluster_storage(int c, int i)
{
    return 128 * (c >> 2) + 32 * (i >> 3) + 8 * (c & 3) + (i & 7);
}
__device__ __forceinline__ int sphere_slot(int j)
{
    if (j < 8) return j;
    const int f = (j - 8) / 10, i = (j - 8) - 10 * f;
    if (i < 4) return 4 * (2 + (f >> 1) * 2 + (i >> 1)) + (f & 1) * 2 + (i & 1);
    if (i < 8) return 4 * (6 + f) + (i - 4);
    return 4 * (10 + (f >> 1)) + (f & 1) * 2 + (i - 8);
}

// model + frame data every CTA stages once.  The members up to n_pts are the frame's STAGING IMAGE (HPE_IMAGE_BYTES):
// they depend on the cloud only, so cluster_kernel builds them once per observation (stage_frame_build) and every consumer
// CTA copies them (stage_frame) instead of redoing the cluster gather, the centroid and the bounding boxes behind five barriers.
struct __align__(16) FrameSmem {
    float4 pre[HPE_MAXPTS];          // (260u|p'|^2, x', y', z'): float32 points recentred on q, CLUSTER ORDER
    float4 cbox[HPE_NCL];            // per cluster of 32 points: centre of the bounding box of p', w = 2 max|p'|^2
    float4 chalf[HPE_NCL];           // ... its half extents (rounded outwards)
    double2 cxy[HPE_MAXPTS];         // FP64 cloud (X, -Y), CLUSTER ORDER: one 128-bit load per point, which the eight lanes of a quarter
                                     // fetch in one wavefront whatever cluster the quarter holds (three 64-bit arrays cost a two-way
                                     // bank conflict whenever the two quarters of a half-warp hold clusters of the same parity)
    double cz[HPE_MAXPTS];           // ... -Z
    unsigned char vq[HPE_MAXPTS / 4]; // [8 c + i] bit m: storage position cluster_storage(c, i + 8 m) holds a point (not a repeat): the four
                                     // positions k0 + 32 m of a lane in one byte
    unsigned char perm[HPE_MAXPTS];  // storage position -> point index of the observation
    float qx, qy, qz;
    int n_pts;
    // ---- end of the staging image ----
    DevModel M;
    double rad[HPE_NSPH];
    double scale, dt_max, focal, cx, cy, lambda;
    const float *depth_mm, *dt_px;
    unsigned char slot_of[HPE_NSPH]; // sphere -> matcher slot
    unsigned char sph_of[HPE_NSPH];  // matcher slot -> sphere
};
static_assert(offsetof(FrameSmem, M) == HPE_IMAGE_BYTES && HPE_IMAGE_BYTES % 16 == 0, "staging image layout");

// all threads of the CTA; caller syncs afterwards
__device__ __forceinline__ void stage_model(FrameSmem &fs, const DevModel &m)
{
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 8); i += blockDim.x)
        reinterpret_cast<double *>(&fs.M)[i] = reinterpret_cast<const double *>(&m)[i];
    if (threadIdx.x < HPE_NSPH) {
        fs.rad[threadIdx.x] = m.rad[threadIdx.x];
        const int sl = sphere_slot(threadIdx.x);
        fs.slot_of[threadIdx.x] = (unsigned char)sl;
        fs.sph_of[sl] = (unsigned char)threadIdx.x;
    }
}

// all threads of the CTA (blockDim >= 64, a multiple of 32); contains __syncthreads, ends synchronised.
// The cloud is staged in the CLUSTER ORDER of the observation (F.perm, cluster_kernel): member i of cluster c is point
// perm[32 c + i] and lives at storage position cluster_storage(c, i); members beyond n_pts repeat the last point, so all
// HPE_NCL clusters are well defined (fs.valid tells the alignment which positions count).
__device__ __forceinline__ void stage_frame_build(FrameSmem &fs, const FrameView &F)
{
    __syncthreads();
    for (int t = threadIdx.x; t < HPE_MAXPTS; t += blockDim.x) {
        const int k = cluster_storage(t >> 5, t & 31), src = F.perm[t];
        fs.perm[k] = (unsigned char)src;
        fs.cxy[k] = make_double2(F.cloud[src], F.cloud[F.n_pts + src]);
        fs.cz[k] = F.cloud[2 * F.n_pts + src];
    }
    if (threadIdx.x < HPE_MAXPTS / 4) {
        const int c = threadIdx.x >> 3, i = threadIdx.x & 7;
        unsigned v = 0u;
        for (int mm = 0; mm < 4; ++mm) v |= (32 * c + i + 8 * mm < F.n_pts ? 1u : 0u) << mm;
        fs.vq[threadIdx.x] = (unsigned char)v;
    }
    if (threadIdx.x == 0) {
        fs.scale = F.scale; fs.dt_max = F.dt_max; fs.focal = F.focal; fs.cx = F.cx; fs.cy = F.cy;
        fs.lambda = (double)HPE_NSPH / (double)F.n_pts;       // |M|/|P|, src/costfunc.cpp:372
        fs.depth_mm = F.depth_mm; fs.dt_px = F.dt_px; fs.n_pts = F.n_pts;
    }
    __syncthreads();
    if (threadIdx.x < 32) {     // centroid of the float32 cloud (any float origin is valid)
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int t = threadIdx.x; t < F.n_pts; t += 32) {
            const int k = cluster_storage(t >> 5, t & 31);
            sx += (float)fs.cxy[k].x; sy += (float)fs.cxy[k].y; sz += (float)fs.cz[k];
        }
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
            sz += __shfl_xor_sync(0xffffffffu, sz, o);
        }
        if (threadIdx.x == 0) { fs.qx = sx / (float)F.n_pts; fs.qy = sy / (float)F.n_pts; fs.qz = sz / (float)F.n_pts; }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < HPE_MAXPTS; k += blockDim.x) {
        // conv_to<fmat> (src/costfunc.cpp:322), then p' = fl(pf - q)
        float x = __fsub_rn((float)fs.cxy[k].x, fs.qx), y = __fsub_rn((float)fs.cxy[k].y, fs.qy),
              z = __fsub_rn((float)fs.cz[k], fs.qz);
        float pp = x * x + y * y + z * z;
        fs.pre[k] = make_float4(1.56e-5f * pp, x, y, z);      // (260 * 2^-24 |p'|^2 rounded up, x', y', z'): the coordinates sit in
                                                              // register banks 1,2,3, the sphere operands they meet in 0,1,2
    }
    __syncthreads();
    // bounding boxes of the clusters in p' space (prune_blocks): a warp per cluster
    for (int c = threadIdx.x >> 5; c < HPE_NCL; c += blockDim.x >> 5) {
        const float4 q = fs.pre[cluster_storage(c, threadIdx.x & 31)];
        float lx = q.y, hx = q.y, ly = q.z, hy = q.z, lz = q.w, hz = q.w, pp = q.y * q.y + q.z * q.z + q.w * q.w;
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
            ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
            lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
            pp = fmaxf(pp, __shfl_xor_sync(0xffffffffu, pp, o));
        }
        if ((threadIdx.x & 31) == 0) {
            const float bx = 0.5f * (lx + hx), by = 0.5f * (ly + hy), bz = 0.5f * (lz + hz);
            fs.cbox[c] = make_float4(bx, by, bz, 2.0f * pp);
            // half extents rounded outwards: [b - h, b + h] contains [lo, hi] whatever the rounding of b
            fs.chalf[c] = make_float4(fmaf(0.5f * (hx - lx), 1.00001f, 1e-6f * fabsf(bx) + 1e-30f),
                                      fmaf(0.5f * (hy - ly), 1.00001f, 1e-6f * fabsf(by) + 1e-30f),
                                      fmaf(0.5f * (hz - lz), 1.00001f, 1e-6f * fabsf(bz) + 1e-30f), 0.f);
        }
    }
    __syncthreads();
}

// all threads of the CTA; contains __syncthreads, ends synchronised.  Copies the frame's staging image (cluster_kernel)
// and sets the per-frame scalars, which the observation kernels may have written after the image (dt_max).
__device__ __forceinline__ void stage_frame(FrameSmem &fs, const FrameView &F)
{
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(&fs);
    for (int i = threadIdx.x; i < HPE_IMAGE_BYTES / 16; i += blockDim.x) dst[i] = __ldg(F.image + i);
    if (threadIdx.x == 0) {
        fs.scale = F.scale; fs.dt_max = F.dt_max; fs.focal = F.focal; fs.cx = F.cx; fs.cy = F.cy;
        fs.lambda = (double)HPE_NSPH / (double)F.n_pts;       // |M|/|P|, src/costfunc.cpp:372
        fs.depth_mm = F.depth_mm; fs.dt_px = F.dt_px;
    }
    __syncthreads();
}

// Exact restatement of the matcher for one point with the whole warp: lane handles spheres lane
// and lane+32, then a lexicographic (sqrtf(d), index) minimum over the warp = first minimum.
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

__device__ __forceinline__ float expanded(const float4 &p, const float4 &e)
{
    return fmaf(p.w, e.z, fmaf(p.z, e.y, fmaf(p.y, e.x, e.w)));
}

// packed pair of FP32 FMAs (same IEEE fma in each half): the three-register scalar FFMAs of the sweep average ~1.5
// issue cycles because two of their operands often sit in the same register bank; the 64-bit operands of FFMA2
// are fetched conflict-free (profiles/r1_issue_microbench.txt)
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long dup32(float v)
{
    unsigned long long d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(v));
    return d;
}
__device__ __forceinline__ void split32(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// s of the two spheres of a pair for one point: fma(z', ez, fma(y', ey, fma(x', ex, ew))) in each half
__device__ __forceinline__ unsigned long long expanded2(unsigned long long px, unsigned long long py, unsigned long long pz,
                                                         const ulonglong2 &xy, const ulonglong2 &zw)
{
    return ffma2(pz, zw.x, ffma2(py, xy.y, ffma2(px, xy.x, zw.y)));
}

// the reference's arithmetic on every pair, 4 points of the lane (cross-check: hpe_set_matcher(1))
__device__ __forceinline__ void match_pass_direct(const FrameSmem &fs, const WarpScratch &ws, int k0, int lane,
                                                  int idx[HPE_PASS])
{
    const float INF = __int_as_float(0x7f800000);
    float px[HPE_PASS], py[HPE_PASS], pz[HPE_PASS], b1[HPE_PASS];
#pragma unroll
    for (int m = 0; m < HPE_PASS; ++m) {
        int k = k0 + 32 * m;
        px[m] = (float)fs.cxy[k].x; py[m] = (float)fs.cxy[k].y; pz[m] = (float)fs.cz[k];
        b1[m] = INF;
        idx[m] = 0;
    }
#pragma unroll 2
    for (int j = 0; j < HPE_NSPH; ++j) {
        float4 c = ws.cf[j];
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            float t0 = __fsub_rn(px[m], c.x), t1 = __fsub_rn(py[m], c.y), t2 = __fsub_rn(pz[m], c.z);
            float d = __fmul_rn(t0, t0);
            d = __fadd_rn(d, __fmul_rn(t1, t1));
            d = __fadd_rn(d, __fmul_rn(t2, t2));
            float w = __fsqrt_rn(d);
            if (w < b1[m]) { b1[m] = w; idx[m] = j; }
        }
    }
}

// packed FP32 add (same IEEE add in each half)
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

#ifndef HPE_PRUNE
#define HPE_PRUNE 1          // 1: sweep A visits only the sphere blocks prune_blocks() could not exclude for the cluster
#endif

// Which sphere blocks can hold the nearest centre of SOME point of a cluster?  Lane = (cluster c = lane / 4,
// quarter = lane % 4); returns the 12-bit block mask of cluster c on the four lanes of its quad.
//   s_j(p') = w_j + e_j . p' is linear in p', so against a reference sphere k (the one nearest to the centre b of the
//   cluster's bounding box) s_j - s_k >= [s_j(b) - s_k(b)] - sum_i |e_j,i - e_k,i| h_i everywhere in the box (h = its
//   half extents).  If that bound exceeds the margin below, sphere j is farther than sphere k from EVERY point of the
//   cluster, in the reference's own float32 arithmetic, by more than the certificate of match_pass asks for: with
//   u = 2^-24, |s_j + |p'|^2 - D_j| <= u(104|p'|^2 + 13|s_j|) (header comment), |s_j| <= 2 w_j + |p'|^2 (because
//   |e_j . p'| = 2|c'_j . p'| <= w_j + |p'|^2), the reference orders D_j, D_k correctly beyond 24u D, the rounding of the
//   bound itself is below 16u(w_j + w_k + |p'|^2): in total < 100u(w_j + w_k) + 400u max|p'|^2, and the margin is
//   4e-5 (w_j + w_k + 2 max|p'|^2) = 671u(...).  A block is skipped when all four of its spheres are excluded; the block
//   of k itself never is, and the true nearest sphere of a point cannot be excluded (it is not farther than k).
//   NaN / inf operands keep the blocks of the lane that meets them.
__device__ __forceinline__ unsigned prune_blocks(const FrameSmem &fs, const WarpScratch &ws, int lane)
{
    const int c = lane >> 2, sub = lane & 3;
    const float4 bc = fs.cbox[c], hc = fs.chalf[c];
    const unsigned long long bx = dup32(bc.x), by = dup32(bc.y), bz = dup32(bc.z);
    // s_j(b) for the 12 slots of blocks 3 sub .. 3 sub + 2, slot index in the 6 low mantissa bits (any sphere is a valid
    // reference, so the perturbation only affects which near-minimal one is taken)
    unsigned long long sv[6];
    float best = __int_as_float(0x7f800000);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const ulonglong2 *blk = reinterpret_cast<const ulonglong2 *>(ws.ex + (3 * sub + t) * HPE_EX_PITCH);
        sv[2 * t] = expanded2(bx, by, bz, blk[0], blk[1]);
        sv[2 * t + 1] = expanded2(bx, by, bz, blk[2], blk[3]);
        float s0, s1, s2, s3;
        split32(sv[2 * t], s0, s1);
        split32(sv[2 * t + 1], s2, s3);
        const unsigned base = (unsigned)(4 * (3 * sub + t));
        s0 = __uint_as_float((__float_as_uint(s0) & 0xffffffc0u) | base);
        s1 = __uint_as_float((__float_as_uint(s1) & 0xffffffc0u) | (base + 1));
        s2 = __uint_as_float((__float_as_uint(s2) & 0xffffffc0u) | (base + 2));
        s3 = __uint_as_float((__float_as_uint(s3) & 0xffffffc0u) | (base + 3));
        best = fmin3(best, fminf(s0, s1), fminf(s2, s3));
    }
    best = fminf(best, __shfl_xor_sync(0xffffffffu, best, 1));
    best = fminf(best, __shfl_xor_sync(0xffffffffu, best, 2));
    const int ks = min((int)(__float_as_uint(best) & 63u), HPE_NSPH - 1);
    const float *rf = reinterpret_cast<const float *>(ws.ex + (ks >> 2) * HPE_EX_PITCH) + ((ks >> 1) & 1) * 8 + (ks & 1);
    const float rx = rf[0], ry = rf[2], rz = rf[4], rw = rf[6];
    const float sref = fmaf(bc.z, rz, fmaf(bc.y, ry, fmaf(bc.x, rx, rw)));
    const unsigned long long nrx = dup32(-rx), nry = dup32(-ry), nrz = dup32(-rz), nsref = dup32(-sref),
                             wk = dup32(rw + bc.w), nmrg = dup32(-4e-5f);
    unsigned mask = 0u;
    float poison = 0.0f;       // NaN as soon as one bound of this lane is NaN or infinite (0 * x accumulated on the FMA pipe)
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const ulonglong2 *blk = reinterpret_cast<const ulonglong2 *>(ws.ex + (3 * sub + t) * HPE_EX_PITCH);
        float tmin = __int_as_float(0x7f800000);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const ulonglong2 xy = blk[2 * q], zw = blk[2 * q + 1];
            // [s_j(b) - s_k(b)] - 4e-5 (w_j + w_k + 2 max|p'|^2), both spheres of the pair
            unsigned long long acc = ffma2(nmrg, fadd2(zw.y, wk), fadd2(sv[2 * t + q], nsref));
            float a0, a1, dx0, dx1, dy0, dy1, dz0, dz1;
            split32(acc, a0, a1);
            split32(fadd2(xy.x, nrx), dx0, dx1);
            split32(fadd2(xy.y, nry), dy0, dy1);
            split32(fadd2(zw.x, nrz), dz0, dz1);
            a0 = fmaf(-fabsf(dz0), hc.z, fmaf(-fabsf(dy0), hc.y, fmaf(-fabsf(dx0), hc.x, a0)));
            a1 = fmaf(-fabsf(dz1), hc.z, fmaf(-fabsf(dy1), hc.y, fmaf(-fabsf(dx1), hc.x, a1)));
            poison = fmaf(a0, 0.0f, fmaf(a1, 0.0f, poison));
            tmin = fmin3(tmin, a0, a1);
        }
        if (!(tmin > 0.0f)) mask |= 1u << (3 * sub + t);
    }
    if (!(poison == poison)) mask = 7u << (3 * sub);        // keeps all three blocks of the lane: conservative
    mask |= __shfl_xor_sync(0xffffffffu, mask, 1);
    mask |= __shfl_xor_sync(0xffffffffu, mask, 2);
    return mask;
}

// phase 3a for 4 points of the lane, k = k0 + 32 m (see the header comment).  ws.cf / ws.ex must
// hold the pose being matched.
template <int UNROLL = HPE_SWEEP_UNROLL>
__device__ __forceinline__ void match_pass(const FrameSmem &fs, const WarpScratch &ws, int k0, int lane,
                                           unsigned mk, int idx[HPE_PASS])
{
    const float INF = __int_as_float(0x7f800000);
#if HPE_FFMA2
    // ---- packed variant: spheres are staged in pairs (see depth_and_stage), point coordinates duplicated ----
    unsigned long long px[HPE_PASS], py[HPE_PASS], pz[HPE_PASS];
    float pxf[HPE_PASS], pyf[HPE_PASS], pzf[HPE_PASS];
    float M1[HPE_PASS], M2[HPE_PASS];
    int g1[HPE_PASS];
#pragma unroll
    for (int m = 0; m < HPE_PASS; ++m) {
        const float4 q = fs.pre[k0 + 32 * m];
        pxf[m] = q.y; pyf[m] = q.z; pzf[m] = q.w;
        px[m] = dup32(q.y); py[m] = dup32(q.z); pz[m] = dup32(q.w);
        M1[m] = INF; M2[m] = INF; g1[m] = 0;
    }
#if HPE_PRUNE && HPE_PACKED_INDEX
    // pruned sweep A: the eight lanes of a quarter hold the 32 points of ONE cluster (cluster_storage; match_align says
    // which) and mk is that cluster's block mask.  Every quarter walks its own list -- the block address differs per
    // quarter, the control flow does not: all lanes run max(list length) steps, a lane whose list is exhausted sweeps
    // the "no block" entry (s = 3e38, changes neither M1 nor a certified M2).
    const int nit = __reduce_max_sync(0xffffffffu, __popc(mk));
#ifdef HPE_COUNT_ITERS      // measurement build: the fallback counter counts sweep steps instead (tools/fallback_rate.py)
    if (lane == 0) atomicAdd(&g_hpe_fallback_points, (unsigned long long)nit);
#endif
    mk |= 1u << HPE_NBLK;          // sentinel above the 12 block bits: an exhausted list yields the "no block" entry
#pragma unroll UNROLL
    for (int it = 0; it < nit; ++it) {
        const int g = __ffs(mk) - 1;
        mk = (mk & (mk - 1u)) | 1u << HPE_NBLK;
        const ulonglong2 *blk = reinterpret_cast<const ulonglong2 *>(ws.ex + g * HPE_EX_PITCH);
        const ulonglong2 a0 = blk[0], a1 = blk[1], b0 = blk[2], b1 = blk[3];
        unsigned long long sa[HPE_PASS], sb[HPE_PASS];
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(px[m], a0.x, a1.y); sb[m] = ffma2(px[m], b0.x, b1.y); }
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(py[m], a0.y, sa[m]); sb[m] = ffma2(py[m], b0.y, sb[m]); }
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(pz[m], a1.x, sa[m]); sb[m] = ffma2(pz[m], b1.x, sb[m]); }
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            float s0, s1, s2, s3;
            split32(sa[m], s0, s1);
            split32(sb[m], s2, s3);
            float gm = fmin3(fminf(s0, s1), s2, s3);
            gm = __uint_as_float((__float_as_uint(gm) & 0xfffffff0u) | (unsigned)g);
            M2[m] = fminf(M2[m], fmaxf(M1[m], gm));
            M1[m] = fminf(M1[m], gm);
        }
    }
#elif HPE_PACKED_INDEX && HPE_PAIR_BLOCKS
    // two blocks per step: their (index-carrying) minima are ordered once, so the runner-up update costs one
    // fmax + one three-input min for both (5.5 instead of 6 ALU instructions per point and block)
    HPE_UNROLL_SWEEP
    for (int g = 0; g < HPE_NBLK; g += 2) {
        float gma[HPE_PASS];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const ulonglong2 *blk = reinterpret_cast<const ulonglong2 *>(ws.ex + (g + half) * HPE_EX_PITCH);
            const ulonglong2 a0 = blk[0], a1 = blk[1], b0 = blk[2], b1 = blk[3];
            unsigned long long sa[HPE_PASS], sb[HPE_PASS];
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(px[m], a0.x, a1.y); sb[m] = ffma2(px[m], b0.x, b1.y); }
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(py[m], a0.y, sa[m]); sb[m] = ffma2(py[m], b0.y, sb[m]); }
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(pz[m], a1.x, sa[m]); sb[m] = ffma2(pz[m], b1.x, sb[m]); }
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) {
                float s0, s1, s2, s3;
                split32(sa[m], s0, s1);
                split32(sb[m], s2, s3);
                float gm = fmin3(fminf(s0, s1), s2, s3);
                gm = __uint_as_float((__float_as_uint(gm) & 0xfffffff0u) | (unsigned)(g + half));
                if (half == 0) {
                    gma[m] = gm;
                } else {
                    float lo = fminf(gma[m], gm), hi = fmaxf(gma[m], gm);
                    M2[m] = fmin3(M2[m], hi, fmaxf(M1[m], lo));
                    M1[m] = fminf(M1[m], lo);
                }
            }
        }
    }
#else
    HPE_UNROLL_SWEEP
    for (int g = 0; g < HPE_NBLK; ++g) {
        const ulonglong2 *blk = reinterpret_cast<const ulonglong2 *>(ws.ex + g * HPE_EX_PITCH);
        const ulonglong2 a0 = blk[0], a1 = blk[1], b0 = blk[2], b1 = blk[3];
        unsigned long long sa[HPE_PASS], sb[HPE_PASS];
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(px[m], a0.x, a1.y); sb[m] = ffma2(px[m], b0.x, b1.y); }
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(py[m], a0.y, sa[m]); sb[m] = ffma2(py[m], b0.y, sb[m]); }
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) { sa[m] = ffma2(pz[m], a1.x, sa[m]); sb[m] = ffma2(pz[m], b1.x, sb[m]); }
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            float s0, s1, s2, s3;
            split32(sa[m], s0, s1);
            split32(sb[m], s2, s3);
            float gm = fmin3(fminf(s0, s1), s2, s3);
#if HPE_PACKED_INDEX
            // the block index rides in the 4 low mantissa bits of the block minimum (a perturbation below 16 ulp,
            // paid for in tau): no compare + select to remember the best block
            gm = __uint_as_float((__float_as_uint(gm) & 0xfffffff0u) | (unsigned)g);
            M2[m] = fminf(M2[m], fmaxf(M1[m], gm));
            M1[m] = fminf(M1[m], gm);
#else
            M2[m] = fminf(M2[m], fmaxf(M1[m], gm));
            bool better = gm < M1[m];
            M1[m] = fminf(M1[m], gm);
            g1[m] = better ? g : g1[m];
#endif
        }
    }
#endif
#if HPE_PACKED_INDEX
#pragma unroll
    for (int m = 0; m < HPE_PASS; ++m) g1[m] = (int)(__float_as_uint(M1[m]) & 15u);
#endif
    unsigned slow = 0u;
#pragma unroll
    for (int m = 0; m < HPE_PASS; ++m) {
        const ulonglong2 *blk = reinterpret_cast<const ulonglong2 *>(ws.ex + g1[m] * HPE_EX_PITCH);
        float s0, s1, s2, s3;
        split32(expanded2(px[m], py[m], pz[m], blk[0], blk[1]), s0, s1);
        split32(expanded2(px[m], py[m], pz[m], blk[2], blk[3]), s2, s3);
        float lo01 = fminf(s0, s1), hi01 = fmaxf(s0, s1), lo23 = fminf(s2, s3), hi23 = fmaxf(s2, s3);
        float r = fmin3(fmaxf(lo01, lo23), hi01, hi23);                   // runner-up inside the block
#if HPE_PACKED_INDEX
        // exact minimum of the best block; M2 carries an index in its low bits: |M2 - true| < 32u|M2|
        const float m1 = fminf(lo01, lo23);
        // (52 + 2*32) * 2^-24 = 6.9e-6; 260u|p'|^2 <= 1.56e-5 |p'|^2 from the coordinates in registers (a load of the staged
        // value would cost four shared-memory wavefronts per pass: the float4 rows sit 16 bytes apart)
        const float tau = fmaf(7.2e-6f, fabsf(m1), 1.56e-5f * fmaf(pzf[m], pzf[m], fmaf(pyf[m], pyf[m], pxf[m] * pxf[m])));
        bool cert = fminf(M2[m], r) - m1 > tau;                           // false on NaN
#else
        const float tau = fmaf(3.2e-6f, fabsf(M1[m]), fs.pre[k0 + 32 * m].x);   // 52 * 2^-24 = 3.0994e-6
        bool cert = fminf(M2[m], r) - M1[m] > tau;                        // false on NaN
#endif
        int i = s1 < s0 ? 1 : 0, j = s3 < s2 ? 3 : 2;
        idx[m] = fs.sph_of[g1[m] * HPE_BLK + (lo23 < lo01 ? j : i)];     // matcher slot -> sphere
        slow |= (cert ? 0u : 1u) << m;
    }
#else
    float4 p[HPE_PASS];
    float M1[HPE_PASS], M2[HPE_PASS];
    int g1[HPE_PASS];
#pragma unroll
    for (int m = 0; m < HPE_PASS; ++m) {
        p[m] = fs.pre[k0 + 32 * m];
        M1[m] = INF; M2[m] = INF; g1[m] = 0;
    }
#if HPE_PACKED_INDEX
    // ---- sweep A, two blocks per step; the block index rides in the 4 low mantissa bits of the block
    //      minimum (a perturbation below 32u|s|, paid for in tau), so best / runner-up tracking is
    //      5.5 instead of 7 ALU instructions per (point, block) ----
#pragma unroll 2
    for (int g = 0; g < HPE_NBLK; g += 2) {
        float4 e[2 * HPE_BLK];
#pragma unroll
        for (int j = 0; j < 2 * HPE_BLK; ++j) e[j] = ws.ex[HPE_EX_AT(g * HPE_BLK + j)];
        float s[HPE_PASS][2 * HPE_BLK];
#pragma unroll
        for (int j = 0; j < 2 * HPE_BLK; ++j)
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) s[m][j] = fmaf(p[m].y, e[j].x, e[j].w);
#pragma unroll
        for (int j = 0; j < 2 * HPE_BLK; ++j)
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) s[m][j] = fmaf(p[m].z, e[j].y, s[m][j]);
#pragma unroll
        for (int j = 0; j < 2 * HPE_BLK; ++j)
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) s[m][j] = fmaf(p[m].w, e[j].z, s[m][j]);
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            float ga = fmin3(fminf(s[m][0], s[m][1]), s[m][2], s[m][3]);
            float gb = fmin3(fminf(s[m][4], s[m][5]), s[m][6], s[m][7]);
            ga = __uint_as_float((__float_as_uint(ga) & 0xfffffff0u) | (unsigned)g);
            gb = __uint_as_float((__float_as_uint(gb) & 0xfffffff0u) | (unsigned)(g + 1));
            float lo = fminf(ga, gb), hi = fmaxf(ga, gb);
            M2[m] = fmin3(M2[m], hi, fmaxf(M1[m], lo));
            M1[m] = fminf(M1[m], lo);
        }
    }
#pragma unroll
    for (int m = 0; m < HPE_PASS; ++m) g1[m] = (int)(__float_as_uint(M1[m]) & 15u);
#else
    // ---- sweep A: block minima of s_j over all 48 spheres ----
#pragma unroll 3
    for (int g = 0; g < HPE_NBLK; ++g) {
        float4 e[HPE_BLK];
#pragma unroll
        for (int j = 0; j < HPE_BLK; ++j) e[j] = ws.ex[g * HPE_EX_PITCH + j];
#if HPE_SWEEP_STAGED
        // 16 independent accumulators, one coordinate at a time: consecutive FFMAs share the sphere
        // operands (operand-reuse cache) and nothing waits on a 4-cycle dependency
        float s[HPE_PASS][HPE_BLK];
#pragma unroll
        for (int j = 0; j < HPE_BLK; ++j)
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) s[m][j] = fmaf(p[m].y, e[j].x, e[j].w);
#pragma unroll
        for (int j = 0; j < HPE_BLK; ++j)
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) s[m][j] = fmaf(p[m].z, e[j].y, s[m][j]);
#pragma unroll
        for (int j = 0; j < HPE_BLK; ++j)
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) s[m][j] = fmaf(p[m].w, e[j].z, s[m][j]);
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            float gm = fmin3(fminf(s[m][0], s[m][1]), s[m][2], s[m][3]);
#else
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            float gm = fmin3(fminf(expanded(p[m], e[0]), expanded(p[m], e[1])), expanded(p[m], e[2]), expanded(p[m], e[3]));
#endif
            M2[m] = fminf(M2[m], fmaxf(M1[m], gm));
            bool better = gm < M1[m];
            M1[m] = fminf(M1[m], gm);
            g1[m] = better ? g : g1[m];
        }
    }
#endif
    // ---- sweep B + certificate ----
    unsigned slow = 0u;
#pragma unroll
    for (int m = 0; m < HPE_PASS; ++m) {
        const float4 *blk = ws.ex + g1[m] * HPE_EX_PITCH;
        float s0 = expanded(p[m], blk[0]), s1 = expanded(p[m], blk[1]);
        float s2 = expanded(p[m], blk[2]), s3 = expanded(p[m], blk[3]);
        float lo01 = fminf(s0, s1), hi01 = fmaxf(s0, s1), lo23 = fminf(s2, s3), hi23 = fmaxf(s2, s3);
        float r = fmin3(fmaxf(lo01, lo23), hi01, hi23);                   // runner-up inside the block
#if HPE_PACKED_INDEX
        // exact minimum of the best block; M2 carries an index in its low bits: |M2 - true| < 32u|M2|
        const float m1 = fminf(lo01, lo23);
        const float tau = fmaf(7.2e-6f, fabsf(m1), p[m].x);               // (52 + 2*32) * 2^-24 = 6.9e-6
        bool cert = fminf(M2[m], r) - m1 > tau;                           // false on NaN
#else
        const float tau = fmaf(3.2e-6f, fabsf(M1[m]), p[m].x);            // 52 * 2^-24 = 3.0994e-6
        bool cert = fminf(M2[m], r) - M1[m] > tau;                        // false on NaN
#endif
        int i = s1 < s0 ? 1 : 0, j = s3 < s2 ? 3 : 2;
        idx[m] = fs.sph_of[g1[m] * HPE_BLK + (lo23 < lo01 ? j : i)];     // matcher slot -> sphere
        slow |= (cert ? 0u : 1u) << m;
    }
#endif  // HPE_FFMA2
    // ---- fallback ----
    if (__any_sync(0xffffffffu, slow != 0u)) {
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            const int km = k0 + 32 * m;
            // valid bit of position km = cluster_storage(c, i + 8 m): c = 4 (km >> 7) + ((km >> 3) & 3), i = km & 7
            unsigned todo = __ballot_sync(0xffffffffu, ((slow >> m) & 1u) &&
                                          (fs.vq[8 * (4 * (km >> 7) + ((km >> 3) & 3)) + (km & 7)] >> m & 1u));
#ifndef HPE_COUNT_ITERS
            if (todo && lane == 0) atomicAdd(&g_hpe_fallback_points, (unsigned long long)__popc(todo));
#endif
            while (todo) {
                int src = __ffs(todo) - 1;
                todo &= todo - 1;
                int k = __shfl_sync(0xffffffffu, km, src);
                int e = match_point_exact_warp(ws.cf, (float)fs.cxy[k].x, (float)fs.cxy[k].y, (float)fs.cz[k], lane);
                if (lane == src) idx[m] = e;
            }
        }
    }
}

// phase 3 of pose g: both passes; returns the warp-reduced align_models value
// (src/costfunc.cpp:346-377), valid on all lanes.  frozen: optional fixed correspondences
// (n_pts ints); match_row: optional output row (n_pts ints).
// sqrt of a non-negative finite double without the library's special-case branch: rsqrt seed (the upper word of the
// reciprocal root, relative error < 2^-20), ONE coupled Newton step (g ~ sqrt x and h ~ 1 / (2 sqrt x) to ~2^-39) and the
// FMA-residual correction g + (x - g^2) h, whose only rounding is the last one -- within one ulp of the rounded root and
// equal to it except about once in 10^7 arguments (checked on the CPU with a perturbed, truncated seed; a second Newton
// step, round 2's form, removes those cases and costs 3 DFMA per point).  0 and inputs below 1e-280 map to 0.  The
// alignment sums agree with the oracle to ~1e-15 relative either way (summation order).
#ifndef HPE_SQRT_STEPS
#define HPE_SQRT_STEPS 1
#endif
__device__ __forceinline__ double sqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = __dmul_rn(x, y), h = __dmul_rn(0.5, y);
    double r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
#if HPE_SQRT_STEPS > 1
    r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
#endif
    const double d = __fma_rn(-g, g, x);
    g = __fma_rn(d, h, g);
    return x > 1e-280 ? g : 0.0;      // 0 (seed = inf) and denormal-range inputs: the root is below 1e-140
}

#ifndef HPE_SORT_SLOTS
#define HPE_SORT_SLOTS 1     // 1: the clusters with the four longest block lists are matched in pass 0, the other four in pass 1
#endif
template <int MATCHER, bool AUX = true, int UNROLL = HPE_SWEEP_UNROLL>
__device__ __forceinline__ double match_align(const FrameSmem &fs, const WarpScratch &ws, int g, int lane,
                                              const int32_t *frozen, int32_t *match_row)
{
    const double *sd = ws.sphd + (size_t)g * HPE_NSPH * 3;
    double acc = 0.0;
#if HPE_PRUNE && HPE_FFMA2 && HPE_PACKED_INDEX
    const bool pruned = !(AUX && frozen) && MATCHER == 0;
    if (pruned) {
        // Which cluster goes to which (pass, quarter) slot.  A pass costs max(list length) over its four quarters, so the
        // clusters are ranked by list length (ties by index) and slot r = 4 pass + quarter takes the cluster of rank r:
        // the two passes then cost l(1) + l(5) steps, the minimum over all assignments (11.1 instead of 12.7 steps per
        // evaluation on the benchmark's poses, tools/prune_balance_sim.py).  Which lane sums which point depends on the pose only,
        // so every kernel still returns the same bits for the same pose and frame.
        const unsigned qmask = prune_blocks(fs, ws, lane);        // mask of cluster lane / 4 on the lanes of its quad
        const int c = lane >> 2, sub = lane & 3;
#if HPE_SORT_SLOTS
        const int cnt = __popc(qmask);
        int rank = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {                             // the quad shares the seven comparisons
            const int d = 2 * sub + t;
            const int other = __shfl_sync(0xffffffffu, cnt, (lane + 4 * d) & 31), c2 = (c + d) & 7;
            rank += (d != 0 && (other > cnt || (other == cnt && c2 < c))) ? 1 : 0;
        }
        rank += __shfl_xor_sync(0xffffffffu, rank, 1);
        rank += __shfl_xor_sync(0xffffffffu, rank, 2);
#else
        const int rank = c;
#endif
        __syncwarp();
        if (sub == 0) ws.slots[rank] = qmask | (unsigned)c << 12;
        __syncwarp();
    }
#else
    const bool pruned = false;
#endif
#pragma unroll 1
    for (int pass = 0; pass < HPE_PPL / HPE_PASS; ++pass) {
        // storage positions k0 + 32 m of the lane's four points (stage_frame): member (lane & 7) + 8 m of the quarter's cluster
        int k0 = lane + 32 * HPE_PASS * pass, vidx = lane + 32 * pass;      // vidx = 8 cluster + member: the lane's byte of fs.vq
        unsigned mk = 0xfffu;
        if (pruned) {
            const unsigned sl = ws.slots[4 * pass + (lane >> 3)];
            const int c = (int)(sl >> 12);
            mk = sl & 0xfffu;
            k0 = 128 * (c >> 2) + 8 * (c & 3) + (lane & 7);
            vidx = 8 * c + (lane & 7);
        }
        int idx[HPE_PASS];
        if (AUX && frozen) {
#pragma unroll
            for (int m = 0; m < HPE_PASS; ++m) idx[m] = frozen[fs.perm[k0 + 32 * m]];
        } else if (MATCHER == 1) {
            match_pass_direct(fs, ws, k0, lane, idx);
        } else {
            match_pass<UNROLL>(fs, ws, k0, lane, mk, idx);
        }
        // align_models for these 4 points, FP64, the reference's operation order.  Branch-free: positions that only
        // repeat a point (fs.vq bit clear) are computed like the others and dropped by a select, the square root is the
        // straight-line Newton sequence below, so the four chains interleave.
        const unsigned vbits = fs.vq[vidx];
        double n2[HPE_PASS], rj[HPE_PASS];
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            const int k = k0 + 32 * m, j = idx[m];
            if (AUX && match_row && (vbits >> m & 1u)) match_row[fs.perm[k]] = j;
            // negated-frame centre = (cx, -cy, -cz)
            const double2 pxy = fs.cxy[k];
            const double a = pxy.x - sd[j * 3 + 0];
            const double b = pxy.y + sd[j * 3 + 1];
            const double e = fs.cz[k] + sd[j * 3 + 2];
            n2[m] = __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(e, e));
            rj[m] = fs.rad[j];
        }
#pragma unroll
        for (int m = 0; m < HPE_PASS; ++m) {
            const double df = sqrt_pos(n2[m]) - rj[m];
            const double d2 = __dmul_rn(df, df);
            acc += (vbits >> m & 1u) ? d2 : 0.0;
        }
    }
    return __dmul_rn(warp_sum(acc), fs.lambda);
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

// phase 1 for up to HPE_GROUP poses; theta(g) returns a pointer to pose g's 26 doubles.  CG: read the
// pose through L2 (poses that land by DMA while the kernel runs)
// WARP1: take the whole-warp path when n == 1 (only the latency-bound callers ask for it: the extra code costs
// the throughput kernels ~4 %)
template <bool CG = false, bool WARP1 = false, class ThetaAt>
__device__ __forceinline__ void group_fk(const DevModel &M, ThetaAt theta, int n, const WarpScratch &ws, int lane)
{
    int g = lane / 5, d = lane - 5 * g;
    __syncwarp();
    if (WARP1 && n == 1) {
        // a single pose for the whole warp (warp-uniform branch): sincos spread over 23 lanes
        double J[5][3], S[10][3];
        fk_digit_warp(M, theta(0), lane, J);
        if (lane < 5) {
            int cnt = digit_spheres(lane, J, S);
            int first = digit_first_sphere(lane);
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                if (i < cnt) {
                    double *sd = ws.sphd + (size_t)(first + i) * 3;
                    sd[0] = S[i][0];
                    sd[1] = S[i][1];
                    sd[2] = S[i][2];
                }
            }
        }
        __syncwarp();
        return;
    }
    if (g < n) fk_digit_emit<CG>(M, d, theta(g), ws.sphd + (size_t)g * HPE_NSPH * 3);
    __syncwarp();
}

// 1/z without the library's special-case branch: seed + two Newton steps (full precision for normal z; z = 0, denormals
// and non-finite z give inf / NaN, which floor_div sends to the exact division).  floor_div's result does not depend on
// the last bits of the reciprocal.
__device__ __forceinline__ double rcp_pos(double z)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(z));
    double e = __fma_rn(-z, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-z, y, 1.0);
    return __fma_rn(y, e, y);
}

// pixel (dy * 320 + dx) the un-negated centre (x, y, z) projects to (src/costfunc.cpp:251-255), -1 outside the image.
// K*c without FMA contraction: row0 = f*x + cx*z, row1 = f*y + cy*z, row2 = z; then floor(row / z).  One rare branch
// for both coordinates (a quotient within 1e-6 of an integer takes the IEEE division).
__device__ __forceinline__ int project_pixel(const FrameSmem &fs, double x, double y, double z)
{
    const double rz = rcp_pos(z);
    const double ax = __dadd_rn(__dmul_rn(fs.focal, x), __dmul_rn(fs.cx, z));
    const double ay = __dadd_rn(__dmul_rn(fs.focal, y), __dmul_rn(fs.cy, z));
    const double qx = __dmul_rn(ax, rz), qy = __dmul_rn(ay, rz);
    double dx = floor(qx), dy = floor(qy);
    const double fx = qx - dx, fy = qy - dy;
    const bool safe = fx > 1e-6 && fx < 1.0 - 1e-6 && fy > 1e-6 && fy < 1.0 - 1e-6 && fabs(qx) < 1e9 && fabs(qy) < 1e9;
    if (!safe) { dx = floor(__ddiv_rn(ax, z)); dy = floor(__ddiv_rn(ay, z)); }
    const bool inside = dx >= 0.0 && dx < (double)HPE_W && dy >= 0.0 && dy < (double)HPE_H;
    return inside ? (int)dy * HPE_W + (int)dx : -1;
}

// depth_penalty of one sphere from its pixel and the two gathered values (src/costfunc.cpp:274-300), squared
__device__ __forceinline__ double depth_value(const FrameSmem &fs, int at, float mm, float dtv, double z, double r)
{
    double v;
    if (at >= 0) {
        if (mm != 0.0f) v = fmax(0.0, div10_exact(mm) - z);             // depthmp is the file value / 10
        else v = __dadd_rn(__dmul_rn((double)dtv, fs.scale), r);
    } else {
        v = __dadd_rn(__dmul_rn(fs.dt_max, fs.scale), r);
    }
    return __dmul_rn(v, v);
}

// phase 2b: the float32 centres (cf) and sweep operands (ex) of pose g, lanes = spheres (j = lane, lane + 32)
__device__ __forceinline__ void stage_pose(const FrameSmem &fs, const WarpScratch &ws, int g, int lane)
{
    const double *sd = ws.sphd + (size_t)g * HPE_NSPH * 3;
    __syncwarp();      // previous pose's matcher is done with cf/ex
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        if (j < HPE_NSPH) {
            const double x = sd[j * 3], y = sd[j * 3 + 1], z = sd[j * 3 + 2];
            float4 c = make_float4((float)x, (float)(-y), (float)(-z), 0.f);   // conv_to<fmat>, negated frame
            // c' = cf - q exactly (FP64), then one rounding each for -2c' and |c'|^2
            double cx = (double)c.x - (double)fs.qx, cy = (double)c.y - (double)fs.qy, cz = (double)c.z - (double)fs.qz;
            ws.cf[j] = c;
            const float4 e = make_float4((float)(-2.0 * cx), (float)(-2.0 * cy), (float)(-2.0 * cz),
                                         (float)(cx * cx + cy * cy + cz * cz));
            const int sl = fs.slot_of[j];                  // matcher slot of sphere j (sphere_slot)
#if HPE_FFMA2
            // slots 4g+2q and 4g+2q+1 share 64-bit words: block g = | x01 y01 | z01 w01 | x23 y23 | z23 w23 | pad |
            float *fb = reinterpret_cast<float *>(ws.ex + (sl / HPE_BLK) * HPE_EX_PITCH) + ((sl % HPE_BLK) / 2) * 8 + (sl & 1);
            fb[0] = e.x; fb[2] = e.y; fb[4] = e.z; fb[6] = e.w;
#else
            ws.ex[HPE_EX_AT(sl)] = e;
#endif
        }
    }
#if HPE_FFMA2
    // the "no block" entry of the pruned sweep: e = 0, w = 3e38 for its four slots
    if (lane < 4) ws.ex[HPE_NBLK * HPE_EX_PITCH + lane] = (lane & 1) ? make_float4(0.f, 0.f, 3e38f, 3e38f) : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
    __syncwarp();
}

// phase 2a for one pose: depth_penalty (src/costfunc.cpp:227-304), lanes = spheres j = lane and lane + 32, in two halves
// so that a caller can put independent work between them: depth_issue projects the lane's two centres and issues all four
// gathers (depth map and distance transform of both pixels), depth_finish turns them into the penalty.  The value is
// DEFINED as the butterfly sum over lanes of (v_lane^2 + v_{lane+32}^2); group_depth reproduces exactly that order.
struct DepthInFlight {
    int at[2];
    float mm[2], dtv[2];
};

__device__ __forceinline__ DepthInFlight depth_issue(const FrameSmem &fs, const WarpScratch &ws, int g, int lane)
{
    const double *sd = ws.sphd + (size_t)g * HPE_NSPH * 3;
    DepthInFlight f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h, jj = j < HPE_NSPH ? j : 0;
        f.at[h] = project_pixel(fs, sd[jj * 3], sd[jj * 3 + 1], sd[jj * 3 + 2]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        f.mm[h] = 0.f; f.dtv[h] = 0.f;
        // ONE gather per projection: the observation folds the depth map into the distance transform (fold_depth_kernel:
        // hand pixel -> minus its depth in mm, background -> its distance), which halves the sectors an evaluation pulls
        // through the L2 and the footprint of a frame in it
        if (f.at[h] >= 0) { const float v = __ldg(fs.dt_px + f.at[h]); f.mm[h] = v < 0.0f ? -v : 0.0f; f.dtv[h] = v; }
    }
    return f;
}

// result valid on all lanes
__device__ __forceinline__ double depth_finish(const FrameSmem &fs, const WarpScratch &ws, int g, int lane, const DepthInFlight &f)
{
    const double *sd = ws.sphd + (size_t)g * HPE_NSPH * 3;
    double dv = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        if (j < HPE_NSPH) dv += depth_value(fs, f.at[h], f.mm[h], f.dtv[h], sd[j * 3 + 2], fs.rad[j]);
    }
    return warp_sum(dv);
}

__device__ __forceinline__ double depth_pose(const FrameSmem &fs, const WarpScratch &ws, int g, int lane)
{
    const DepthInFlight f = depth_issue(fs, ws, g, lane);
    return depth_finish(fs, ws, g, lane, f);
}

// phases 2a + 2b of pose g (the one-pose-at-a-time callers)
__device__ __forceinline__ double depth_and_stage(const FrameSmem &fs, const WarpScratch &ws, int g, int lane)
{
    const double d = depth_pose(fs, ws, g, lane);
    stage_pose(fs, ws, g, lane);
    return d;
}

// phase 2a for the cnt <= GROUP poses of a group at once: items (pose, sphere) = lane + 32 r fill every lane (7.5 rounds
// for 5 poses instead of 10) and all gathers of the group are in flight together.  Sphere j of pose g sits on lane
// (16 g + j) mod 32 -- the lane depth_pose uses, rotated by 16 for odd g, which the xor butterfly does not see -- and
// lane-local partial sums add sphere j before sphere j + 32, so ws.dep[g] has the bits depth_pose returns.
template <int GROUP>
__device__ __forceinline__ void group_depth(const FrameSmem &fs, const WarpScratch &ws, int cnt, int lane)
{
    static_assert(GROUP <= 8, "ws.dep holds 8 values");
    constexpr int ROUNDS = (GROUP * HPE_NSPH + 31) / 32;
    const int n_items = cnt * HPE_NSPH;
    int at[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int item = lane + 32 * r, it = item < n_items ? item : 0;
        at[r] = project_pixel(fs, ws.sphd[it * 3], ws.sphd[it * 3 + 1], ws.sphd[it * 3 + 2]);
    }
    float mm[ROUNDS], dtv[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        mm[r] = 0.f; dtv[r] = 0.f;
        if (at[r] >= 0 && lane + 32 * r < n_items) { const float v = __ldg(fs.dt_px + at[r]); mm[r] = v < 0.0f ? -v : 0.0f; dtv[r] = v; }
    }
    double acc[GROUP];
#pragma unroll
    for (int g = 0; g < GROUP; ++g) acc[g] = 0.0;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int item = lane + 32 * r, it = item < n_items ? item : 0;
        const int g_lo = (32 * r) / HPE_NSPH, g_hi = (32 * r + 31) / HPE_NSPH;      // compile-time after unrolling
        const bool hi = item >= HPE_NSPH * (g_lo + 1);
        const int j = it - HPE_NSPH * (hi ? g_hi : g_lo);
        double v2 = depth_value(fs, at[r], mm[r], dtv[r], ws.sphd[it * 3 + 2], fs.rad[j]);
        if (item >= n_items) v2 = 0.0;
        if (g_lo < GROUP) acc[g_lo < GROUP ? g_lo : 0] += hi ? 0.0 : v2;
        if (g_hi != g_lo && g_hi < GROUP) acc[g_hi < GROUP ? g_hi : 0] += hi ? v2 : 0.0;
    }
#pragma unroll
    for (int g = 0; g < GROUP; ++g) {
        if (g < cnt) {                      // warp-uniform
            const double d = warp_sum(acc[g]);
            if (lane == 0) ws.dep[g] = d;
        }
    }
    __syncwarp();
}
