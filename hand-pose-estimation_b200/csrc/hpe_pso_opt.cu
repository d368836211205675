// hpe_pso_opt.cu -- PSO::pso_optimise (src/PSO.cpp:539-712), the reference's hybrid of per-particle
// coordinate descent and canonical (w, c1, c2, gbest) PSO -- SURVEY 8(f) rank 3.
//
// Per iteration every particle first runs graditer = 10 descent steps (:592-626): parameter
// permu(m) = randi[0,25]; f_k = cal_cost2 (src/costfunc.cpp:31-86; fresh correspondences for m == 0,
// the particle's frozen ones afterwards); cal_gradient (:380-405), a central difference on that one
// parameter; the Goldstein search (:438-480); the step; f_k again (for m == 0 this recomputes the
// correspondences at the moved pose, cal_corr is still true at :616); pbest update.  Then the swarm
// moves with v = w v + c1 rp (pbest - x) + c2 rg (gbest - x) (:664-666), check_constraints, cal_cost,
// pbest / gbest.  gbest_pos is the CURRENT position of the particle with the smallest pcost (:637,
// :697), as written.
//
// One CTA of 8 warps per swarm (swarms = frames are independent and fill the GPU); in the descent
// phase a warp owns a particle and runs its ~70-340 sequential evaluations with the shared evaluation
// core (hpe_eval.cuh), the particle's correspondences in the warp's shared-memory row; the swarm phase
// is the CTA-wide update + one evaluation per particle.  Positions, velocities and personal bests live
// in L2-resident global scratch as in pso_kernel.  Arithmetic that feeds a comparison is written with
// round-to-nearest intrinsics in the reference's operation order (no FMA contraction).  Random draws
// are the counter-based streams of DESIGN "RNG": 0 initial normals, 1 rp, 2 rg, 4 permu.
#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace {

constexpr int PO_WARPS = 8;
constexpr int PO_GRADITER = 10;      // src/PSO.cpp:573
#ifndef HPE_OPT_MINBLOCKS
#define HPE_OPT_MINBLOCKS 4     // measured on B200, 1,184 swarms of 16 x 4: 2 -> 37.5k, 3 -> 37.9k, 4 -> 40.8k frames/s
#endif

struct __align__(16) OptSmemHead {
    FrameSmem fs;
    __align__(16) unsigned char ws[PO_WARPS * HPE_WS_BYTES(1)];
    int32_t match[PO_WARPS][HPE_MAXPTS];
    double th[PO_WARPS][HPE_NDOF];       // ctheta of the warp's particle
    double cand[PO_WARPS][HPE_NDOF];     // probe / line-search pose
    double gbest[HPE_NDOF];
    double x0[HPE_NDOF];
    double gbest_cost;
    unsigned long long evals;
    int fmin_id;
    int pad;
    // followed by double pcost[num_p]
};

__host__ __device__ inline size_t opt_smem_bytes(int num_p) { return sizeof(OptSmemHead) + (size_t)num_p * 8 + 16; }

// costfunc::cal_cost2 of one pose by one warp (collision included, src/costfunc.cpp:81); frozen == nullptr:
// compute the correspondences and store them in match_out
__device__ __noinline__ double opt_cost2(OptSmemHead &sm, const WarpScratch &ws, const double *theta, int lane,
                                            const int32_t *frozen, int32_t *match_out)
{
    group_fk<false, true>(sm.fs.M, [&](int) { return theta; }, 1, ws, lane);
    double depth = depth_and_stage(sm.fs, ws, 0, lane);
    double align = match_align<0>(sm.fs, ws, 0, lane, frozen, match_out);
    double colli = warp_sum(collision_lane(sm.fs.rad, ws.sphd, lane));
    __syncwarp();
    return __shfl_sync(0xffffffffu, (align + depth) + colli, 0);
}

// costfunc::cal_cost (src/costfunc.cpp:89-127): fresh correspondences, no collision term
__device__ __noinline__ double opt_cost(OptSmemHead &sm, const WarpScratch &ws, const double *theta, int lane)
{
    group_fk<false, true>(sm.fs.M, [&](int) { return theta; }, 1, ws, lane);
    double depth = depth_and_stage(sm.fs, ws, 0, lane);
    double align = match_align<0>(sm.fs, ws, 0, lane, nullptr, nullptr);
    __syncwarp();
    return __shfl_sync(0xffffffffu, align + depth, 0);
}

// first minimum of pcost by warp 0 (pcost.min(fmin_id), :635, :693); gbest test :637-641 / :696-703
__device__ __forceinline__ void opt_gbest(OptSmemHead &sm, const double *pcost, int num_p, int lane)
{
    double c = __longlong_as_double(0x7ff0000000000000LL);
    int ci = 0x7fffffff;
    for (int i = lane; i < num_p; i += 32)
        if (pcost[i] < c) { c = pcost[i]; ci = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double oc = __shfl_xor_sync(0xffffffffu, c, o);
        int oi = __shfl_xor_sync(0xffffffffu, ci, o);
        if (oc < c || (oc == c && oi < ci)) { c = oc; ci = oi; }
    }
    if (lane == 0) {
        if (ci == 0x7fffffff) ci = 0;
        sm.fmin_id = -1;
        if (pcost[ci] < sm.gbest_cost) {
            sm.gbest_cost = pcost[ci];
            sm.fmin_id = ci;               // copy the position
        }
    }
}

__global__ void __launch_bounds__(PO_WARPS * 32, HPE_OPT_MINBLOCKS)
pso_optimise_kernel(DevModel m, ObsView o, PsoDev prm, double w, double c1, double c2,
                    const double *__restrict__ x0_in, int first, int count_swarms, const uint64_t *__restrict__ seeds,
                    double *__restrict__ state_scratch, double *__restrict__ bestp_out, double *__restrict__ best_cost_out,
                    long long *__restrict__ evals_out, double *__restrict__ final_x, double *__restrict__ final_pcost)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OptSmemHead &sm = *reinterpret_cast<OptSmemHead *>(smem_raw);
    double *pcost = reinterpret_cast<double *>(smem_raw + sizeof(OptSmemHead));
    const int num_p = prm.num_p, D = HPE_NDOF;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const WarpScratch ws = warp_scratch(sm.ws, warp, 1);
    double *x = state_scratch + (size_t)blockIdx.x * 3 * num_p * D;
    double *v = x + (size_t)num_p * D;
    double *pb = v + (size_t)num_p * D;
    double *th = sm.th[warp], *cand = sm.cand[warp];
    int32_t *match = sm.match[warp];

    stage_model(sm.fs, m);
    __syncthreads();

    for (int b = blockIdx.x; b < count_swarms; b += gridDim.x) {
        const int slot = first + b;
        const uint64_t seed = seeds ? seeds[b] : 1000ull;
        stage_frame(sm.fs, frame_view(o, slot));     // ends synchronised
        if (tid < D) sm.x0[tid] = x0_in[(size_t)b * D + tid];
        if (tid == 0) { sm.gbest_cost = 1e100; sm.evals = 0ull; }
        __syncthreads();
        const bool frame_ok = o.meta[slot].ok != 0;

        // generate_particles: x0 + randn % std, v = 0 (:56-74, :545-552)
        for (int it = tid; it < num_p * (D / 2); it += nthr) {
            int i = it / (D / 2), dp = it - i * (D / 2);
            double z[2];
            normal2(seed, (uint32_t)dp, (uint32_t)i, 0u, z);
            x[i * D + 2 * dp] = __dadd_rn(sm.x0[2 * dp], __dmul_rn(z[0], prm.std[2 * dp]));
            x[i * D + 2 * dp + 1] = __dadd_rn(sm.x0[2 * dp + 1], __dmul_rn(z[1], prm.std[2 * dp + 1]));
            v[i * D + 2 * dp] = 0.0;
            v[i * D + 2 * dp + 1] = 0.0;
        }
        __syncthreads();
        // initial evaluation (:555-568): pcost, pbest, gbest = first minimum
        for (int i = warp; i < num_p; i += PO_WARPS) {
            double f = opt_cost(sm, ws, x + (size_t)i * D, lane);
            if (lane == 0) pcost[i] = f;
            if (lane < D) pb[i * D + lane] = x[i * D + lane];
        }
        __syncthreads();
        if (warp == 0) opt_gbest(sm, pcost, num_p, lane);
        __syncthreads();
        if (sm.fmin_id >= 0 && tid < D) sm.gbest[tid] = x[sm.fmin_id * D + tid];
        __syncthreads();

        for (int iter = 2; iter <= prm.maxiter; ++iter) {
            const uint32_t gen = (uint32_t)(iter - 1);
            // ---- coordinate descent, a warp per particle (:586-633) ----
            unsigned long long my_evals = 0ull;
            for (int i = warp; i < num_p; i += PO_WARPS) {
                if (lane < D) th[lane] = x[i * D + lane];
                double pc = pcost[i];
                __syncwarp();
                for (int mm = 0; mm < PO_GRADITER; ++mm) {
                    double u[2];
                    uniform2(seed, (uint32_t)(mm >> 1), (uint32_t)i, gen, 4u, u);
                    int param = (int)floor(__dmul_rn(u[mm & 1], (double)D));
                    if (param > D - 1) param = D - 1;
                    const bool cal_corr = mm == 0;
                    double f_k = opt_cost2(sm, ws, th, lane, cal_corr ? nullptr : match, cal_corr ? match : nullptr);
                    // cal_gradient (:380-405): [f(x+h) - f(x-h)] / [2h] on the frozen correspondences
                    const double eps = 1e-5;
                    if (lane < D) cand[lane] = lane == param ? __dadd_rn(th[lane], eps) : th[lane];
                    __syncwarp();
                    const double f_xph = opt_cost2(sm, ws, cand, lane, match, nullptr);
                    if (lane < D) cand[lane] = lane == param ? __dsub_rn(th[lane], eps) : th[lane];
                    __syncwarp();
                    const double f_xmh = opt_cost2(sm, ws, cand, lane, match, nullptr);
                    const double g = __ddiv_rn(__dsub_rn(f_xph, f_xmh), 2 * eps);
                    // goldstein (:438-480); g_k is zero off `param`, so dot(g_k, p_k) = g * (-g)
                    const double dot_gp = __dmul_rn(g, -g);
                    double a = 0.0, bb = 1e100, alpha = 0.5, tk = 0.0;
                    int ls = 0;
                    while (ls < 30) {
                        ls += 1;
                        // theta1 = theta + alpha * p_k, p_k = -1 * g_k
                        if (lane < D) cand[lane] = __dadd_rn(th[lane], __dmul_rn(alpha, lane == param ? -g : -0.0));
                        __syncwarp();
                        const double f1 = opt_cost2(sm, ws, cand, lane, match, nullptr);
                        const double armijo = __dadd_rn(f_k, __dmul_rn(__dmul_rn(0.25, alpha), dot_gp));
                        const double goldst = __dadd_rn(f_k, __dmul_rn(__dmul_rn(1 - 0.25, alpha), dot_gp));
                        if (f1 <= armijo) {
                            if (f1 >= goldst) { tk = alpha; break; }
                            a = alpha;
                            const double t2 = __dmul_rn(2.0, alpha), hm = __dmul_rn(0.5, __dadd_rn(alpha, bb));
                            alpha = hm < t2 ? hm : t2;          // std::min(t*alpha, 0.5*(alpha+b))
                        } else {
                            bb = alpha;
                            alpha = __dmul_rn(0.5, __dadd_rn(a, alpha));
                        }
                    }
                    // ctheta = ctheta - tk * g_k (:614)
                    __syncwarp();
                    if (lane < D) th[lane] = __dsub_rn(th[lane], __dmul_rn(tk, lane == param ? g : 0.0));
                    __syncwarp();
                    f_k = opt_cost2(sm, ws, th, lane, cal_corr ? nullptr : match, cal_corr ? match : nullptr);
                    my_evals += 4 + ls;
                    if (f_k < pc) {                               // :619-625
                        pc = f_k;
                        if (lane < D) pb[i * D + lane] = th[lane];
                    }
                }
                if (lane == 0) pcost[i] = pc;
                // check_constraints(ctheta, cveloc) (:628-630, :358-377)
                if (lane < D) {
                    double t = th[lane], vel = v[i * D + lane];
                    if (t < prm.lb[lane] || t > prm.ub[lane]) { t = prm.lb[lane]; vel = 0.0; }
                    x[i * D + lane] = t;
                    v[i * D + lane] = vel;
                }
                __syncwarp();
            }
            if (lane == 0) atomicAdd(&sm.evals, my_evals);
            __syncthreads();
            // ---- gbest after the descent (:635-642) ----
            if (warp == 0) opt_gbest(sm, pcost, num_p, lane);
            __syncthreads();
            if (sm.fmin_id >= 0 && tid < D) sm.gbest[tid] = x[sm.fmin_id * D + tid];
            __syncthreads();
            // ---- canonical update + check_constraints (:644-677) ----
            for (int it = tid; it < num_p * (D / 2); it += nthr) {
                int i = it / (D / 2), dp = it - i * (D / 2);
                double rp[2], rg[2];
                uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 1u, rp);
                uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 2u, rg);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int d = 2 * dp + e, at = i * D + d;
                    double xi = x[at];
                    double nv = __dadd_rn(__dadd_rn(__dmul_rn(w, v[at]), __dmul_rn(__dmul_rn(c1, rp[e]), __dsub_rn(pb[at], xi))),
                                          __dmul_rn(__dmul_rn(c2, rg[e]), __dsub_rn(sm.gbest[d], xi)));
                    double nx = __dadd_rn(xi, nv);
                    if (nx < prm.lb[d] || nx > prm.ub[d]) { nx = prm.lb[d]; nv = 0.0; }
                    x[at] = nx;
                    v[at] = nv;
                }
            }
            __syncthreads();
            // ---- cal_cost of every particle, pbest (:679-690) ----
            for (int i = warp; i < num_p; i += PO_WARPS) {
                double f = opt_cost(sm, ws, x + (size_t)i * D, lane);
                const double pci = pcost[i];
                __syncwarp();
                if (f < pci) {
                    if (lane < D) pb[i * D + lane] = x[i * D + lane];
                    if (lane == 0) pcost[i] = f;
                }
                __syncwarp();
            }
            __syncthreads();
            // ---- gbest (:693-704) ----
            if (warp == 0) opt_gbest(sm, pcost, num_p, lane);
            __syncthreads();
            if (sm.fmin_id >= 0 && tid < D) sm.gbest[tid] = x[sm.fmin_id * D + tid];
            __syncthreads();
        }
        if (tid < D) bestp_out[(size_t)b * D + tid] = frame_ok ? sm.gbest[tid] : sm.x0[tid];
        if (tid == 0) {
            best_cost_out[b] = frame_ok ? sm.gbest_cost : __longlong_as_double(0x7ff8000000000000LL);
            if (evals_out)
                evals_out[b] = (long long)sm.evals + (long long)num_p * (prm.maxiter > 1 ? prm.maxiter : 1);
        }
        if (final_x)
            for (int it = tid; it < num_p * D; it += nthr) final_x[(size_t)b * num_p * D + it] = x[it];
        if (final_pcost)
            for (int i = tid; i < num_p; i += nthr) final_pcost[(size_t)b * num_p + i] = pcost[i];
        __syncthreads();
    }
}

}  // namespace

HPE_DEFINE_FALLBACK_GETTER(hpe_fallbacks_pso_opt)

cudaError_t hpe_launch_pso_optimise(const DevModel &m, const ObsView &o, const PsoDev &p, double w, double c1, double c2,
                                    const double *x0_dev, int first, int count, const uint64_t *seeds_dev,
                                    double *bestp_dev, double *best_cost_dev, long long *evals_dev, double *final_x_dev,
                                    double *final_pcost_dev, cudaStream_t st)
{
    if (count <= 0) return cudaSuccess;
    size_t smem = opt_smem_bytes(p.num_p);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static PerDevice<size_t> attr_set_pd;
    size_t &attr_set = attr_set_pd.get();
    if (smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pso_optimise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = smem;
    }
    const int sms = hpe_sm_count();
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, pso_optimise_kernel, PO_WARPS * 32, smem);
    if (resident < 1) resident = 1;
    int blocks = count < resident * sms ? count : resident * sms;
    // positions, velocities, personal bests of the resident swarms: grow-only scratch (L2-resident)
    static PerDevice<double *> scratch_pd;
    static PerDevice<size_t> scratch_bytes_pd;
    double *&scratch = scratch_pd.get();
    size_t &scratch_bytes = scratch_bytes_pd.get();
    size_t per = (size_t)p.num_p * HPE_NDOF * sizeof(double);
    if (3 * per * blocks > scratch_bytes) {
        if (scratch) cudaFree(scratch);
        scratch = nullptr; scratch_bytes = 0;
        cudaError_t e = cudaMalloc(&scratch, 3 * per * blocks);
        if (e != cudaSuccess) return e;
        scratch_bytes = 3 * per * blocks;
    }
    pso_optimise_kernel<<<blocks, PO_WARPS * 32, smem, st>>>(m, o, p, w, c1, c2, x0_dev, first, count, seeds_dev, scratch,
                                                             bestp_dev, best_cost_dev, evals_dev, final_x_dev,
                                                             final_pcost_dev);
    return cudaGetLastError();
}
