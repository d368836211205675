// hpe_refine.cu -- PSO::refine_init_pose (src/PSO.cpp:216-266) with PSO::cal_grad (:183-214) and
// PSO::goldstein (:438-480): gradient descent on the global rotation, then on the global position,
// every evaluation a costfunc::cal_cost2 (src/costfunc.cpp:31-86) -- fresh correspondences for f_k,
// frozen ones for the six gradient probes and the line search.  Runs before every pso_evolve in the
// reference's tracking loop (src/testmodel.cpp:126).
//
// One CTA of 8 warps per frame; a warp evaluates one candidate pose with the shared evaluation core
// (hpe_eval.cuh), so the sequential chain of the reference (1 + 6 + up to 30 evaluations per gradient
// step) becomes a few rounds:
//   round A   warp 0: f_k = cal_cost2(x, fresh correspondences)  -> match[] in shared memory
//   round B   warps 0-5: the six central-difference probes x +- 1e-5 e_d           -> grad
//   rounds C  the Goldstein search is a binary decision tree (accept / grow / shrink); warps 0-6
//             evaluate the 7 nodes of the next three levels speculatively, thread 0 then walks the
//             tree with exactly the reference's comparisons.  The step lengths visited, the accepted
//             tk and the iteration count are those of the sequential search.
// Arithmetic that feeds a comparison is written with explicit round-to-nearest intrinsics in the
// reference's operation order (no FMA contraction), like the rest of the FP64 path.
#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace {

constexpr int RF_WARPS = 8;
constexpr int RF_NODES = 7;          // three levels of the line-search tree per round

struct __align__(16) RefineSmem {
    FrameSmem fs;
    __align__(16) unsigned char ws[RF_WARPS * HPE_WS_BYTES(1)];
    int32_t match[HPE_MAXPTS];
    double cand[RF_WARPS][HPE_NDOF];
    double fval[RF_WARPS];
    double x[HPE_NDOF], grad[HPE_NDOF];
    double node_alpha[RF_NODES + 1], node_a[RF_NODES + 1], node_b[RF_NODES + 1];   // 1-based heap order
    double f_k, dot_gp, tk, tol;
    int ls_iter, ls_done, evals, cnt, iter;
};

// cal_cost2 of one pose by one warp; frozen == nullptr: compute correspondences and store them
__device__ __forceinline__ double eval_cost2(RefineSmem &sm, const WarpScratch &ws, const double *theta, int lane,
                                             const int32_t *frozen, int32_t *match_out)
{
    group_fk<false, true>(sm.fs.M, [&](int) { return theta; }, 1, ws, lane);
    double depth = depth_and_stage(sm.fs, ws, 0, lane);
    double align = match_align<0>(sm.fs, ws, 0, lane, frozen, match_out);
    double colli = warp_sum(collision_lane(sm.fs.rad, ws.sphd, lane));
    __syncwarp();
    return (align + depth) + colli;      // align_cost + depth_cost + colli_cost, src/costfunc.cpp:81
}

// children of a line-search state (src/PSO.cpp:462-473)
__device__ __forceinline__ void ls_grow(double a, double b, double alpha, double &na, double &nb, double &nalpha)
{
    (void)a;
    na = alpha;
    nb = b;
    const double x = __dmul_rn(2.0, alpha), y = __dmul_rn(0.5, __dadd_rn(alpha, b));
    nalpha = y < x ? y : x;              // std::min(t*alpha, 0.5*(alpha+b))
}
__device__ __forceinline__ void ls_shrink(double a, double b, double alpha, double &na, double &nb, double &nalpha)
{
    (void)b;
    na = a;
    nb = alpha;
    nalpha = __dmul_rn(0.5, __dadd_rn(a, alpha));
}

__global__ void __launch_bounds__(RF_WARPS * 32, 2)
refine_kernel(DevModel m, ObsView o, int first, int count, double *__restrict__ x_io, int *__restrict__ evals_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RefineSmem &sm = *reinterpret_cast<RefineSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, D = HPE_NDOF;
    const WarpScratch ws = warp_scratch(sm.ws, warp, 1);
    stage_model(sm.fs, m);
    __syncthreads();

    for (int b = blockIdx.x; b < count; b += gridDim.x) {
        const int slot = first + b;
        stage_frame(sm.fs, frame_view(o, slot));          // ends synchronised
        if (tid < D) sm.x[tid] = x_io[(size_t)b * D + tid];
        if (tid == 0) sm.evals = 0;
        __syncthreads();
        if (o.meta[slot].ok == 0) continue;               // empty frame: pose unchanged (uniform branch)

        for (int grp = 0; grp < 2; ++grp) {               // start_idx = {0, 3}, end_idx = {2, 5}
            const int d0 = 3 * grp;
            if (tid == 0) { sm.tol = 1.0; sm.cnt = 0; sm.iter = 0; }
            __syncthreads();
            while (sm.tol > 1e-6 && sm.iter < 15 && sm.cnt < 1) {
                // ---- round A: f_k with fresh correspondences (src/PSO.cpp:245) ----
                if (warp == 0) {
                    double f = eval_cost2(sm, ws, sm.x, lane, nullptr, sm.match);
                    if (lane == 0) sm.f_k = f;
                }
                __syncthreads();
                // ---- round B: central differences on the frozen correspondences (:246, :183-214) ----
                if (warp < 6) {
                    const int d = d0 + (warp >> 1);
                    if (lane < D) {
                        double v = sm.x[lane];
                        if (lane == d) v = (warp & 1) ? __dsub_rn(v, 1e-5) : __dadd_rn(v, 1e-5);
                        sm.cand[warp][lane] = v;
                    }
                    __syncwarp();
                    double f = eval_cost2(sm, ws, sm.cand[warp], lane, sm.match, nullptr);
                    if (lane == 0) sm.fval[warp] = f;
                }
                __syncthreads();
                if (tid == 0) {
                    for (int d = 0; d < D; ++d) sm.grad[d] = 0.0;
                    for (int k = 0; k < 3; ++k)
                        sm.grad[d0 + k] = __ddiv_rn(__dsub_rn(sm.fval[2 * k], sm.fval[2 * k + 1]), 2 * 1e-5);
                    // dot(g_k, p_k), p_k = -1*g_k: op_dot::direct_dot_arma, two partial sums
                    double v1 = 0.0, v2 = 0.0;
                    int i, j;
                    for (i = 0, j = 1; j < D; i += 2, j += 2) {
                        v1 = __dadd_rn(v1, __dmul_rn(sm.grad[i], -sm.grad[i]));
                        v2 = __dadd_rn(v2, __dmul_rn(sm.grad[j], -sm.grad[j]));
                    }
                    sm.dot_gp = __dadd_rn(v1, v2);
                    sm.node_a[1] = 0.0; sm.node_b[1] = 1e100; sm.node_alpha[1] = 0.5;
                    sm.ls_iter = 0; sm.ls_done = 0; sm.tk = 0.0;
                    sm.evals += 7;
                }
                __syncthreads();
                // ---- rounds C: Goldstein line search, three tree levels per round (:438-480) ----
                while (!sm.ls_done) {
                    if (tid == 0) {
                        for (int nd = 1; nd <= 3; ++nd) {
                            ls_grow(sm.node_a[nd], sm.node_b[nd], sm.node_alpha[nd], sm.node_a[2 * nd], sm.node_b[2 * nd],
                                    sm.node_alpha[2 * nd]);
                            ls_shrink(sm.node_a[nd], sm.node_b[nd], sm.node_alpha[nd], sm.node_a[2 * nd + 1],
                                      sm.node_b[2 * nd + 1], sm.node_alpha[2 * nd + 1]);
                        }
                    }
                    __syncthreads();
                    if (warp < RF_NODES) {
                        const double alpha = sm.node_alpha[warp + 1];
                        // theta1 = theta + alpha * p_k, p_k = -1 * g_k
                        if (lane < D) sm.cand[warp][lane] = __dadd_rn(sm.x[lane], __dmul_rn(alpha, -sm.grad[lane]));
                        __syncwarp();
                        double f = eval_cost2(sm, ws, sm.cand[warp], lane, sm.match, nullptr);
                        if (lane == 0) sm.fval[warp] = f;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        int nd = 1;
                        double na = 0, nb = 0, nalpha = 0;
                        for (int level = 0; level < 3; ++level) {
                            sm.ls_iter += 1;
                            const double alpha = sm.node_alpha[nd], f1 = sm.fval[nd - 1];
                            const double armijo = __dadd_rn(sm.f_k, __dmul_rn(__dmul_rn(0.25, alpha), sm.dot_gp));
                            const double goldst = __dadd_rn(sm.f_k, __dmul_rn(__dmul_rn(1 - 0.25, alpha), sm.dot_gp));
                            bool grow = false;
                            if (f1 <= armijo) {
                                if (f1 >= goldst) { sm.tk = alpha; sm.ls_done = 1; break; }
                                grow = true;
                            }
                            if (sm.ls_iter >= 30) { sm.tk = 0.0; sm.ls_done = 1; break; }
                            if (level == 2) {
                                // state of the node the next round starts from
                                if (grow) ls_grow(sm.node_a[nd], sm.node_b[nd], alpha, na, nb, nalpha);
                                else ls_shrink(sm.node_a[nd], sm.node_b[nd], alpha, na, nb, nalpha);
                            }
                            nd = grow ? 2 * nd : 2 * nd + 1;
                        }
                        if (!sm.ls_done) { sm.node_a[1] = na; sm.node_b[1] = nb; sm.node_alpha[1] = nalpha; }
                    }
                    __syncthreads();
                }
                // ---- x0 = x0 - tk*grad; tol = |grad| (:255-262) ----
                if (tid == 0) {
                    sm.evals += sm.ls_iter;
                    if (sm.tk == 0.0) sm.cnt += 1;
                    double a1 = 0.0, a2 = 0.0;          // sum(grad % grad, 0): arrayops::accumulate
                    int i, j;
                    for (i = 0, j = 1; j < D; i += 2, j += 2) {
                        a1 = __dadd_rn(a1, __dmul_rn(sm.grad[i], sm.grad[i]));
                        a2 = __dadd_rn(a2, __dmul_rn(sm.grad[j], sm.grad[j]));
                    }
                    sm.tol = sqrt(__dadd_rn(a1, a2));
                    sm.iter += 1;
                }
                if (tid < D) sm.x[tid] = __dsub_rn(sm.x[tid], __dmul_rn(sm.tk, sm.grad[tid]));
                __syncthreads();
            }
        }
        if (tid < D) x_io[(size_t)b * D + tid] = sm.x[tid];
        if (tid == 0 && evals_out) evals_out[b] = sm.evals;
        __syncthreads();
    }
}

}  // namespace

HPE_DEFINE_FALLBACK_GETTER(hpe_fallbacks_refine)

cudaError_t hpe_launch_refine(const DevModel &m, const ObsView &o, int first, int count, double *x_dev, int *evals_dev,
                              cudaStream_t st)
{
    if (count <= 0) return cudaSuccess;
    const int sms = hpe_sm_count();
    static PerDevice<bool> attr_pd;
    if (bool &attr = attr_pd.get(); !attr) {
        cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RefineSmem));
        attr = true;
    }
    int grid = count < 2 * sms ? count : 2 * sms;
    refine_kernel<<<grid, RF_WARPS * 32, sizeof(RefineSmem), st>>>(m, o, first, count, x_dev, evals_dev);
    return cudaGetLastError();
}
