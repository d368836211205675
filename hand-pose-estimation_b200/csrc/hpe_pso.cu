// hpe_pso.cu -- PSO::pso_evolve (src/PSO.cpp:717-886) as one persistent CTA per swarm: all
// generations, the evaluation of every particle (hpe_eval.cuh), the pbest/gbest bookkeeping
// and the informant topology run inside one kernel launch; swarms (frames) are independent
// and fill the GPU.  In chain mode a single CTA walks a frame sequence, seeding frame k+1
// with the best pose of frame k (src/testmodel.cpp:117-138).
#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace {

#ifndef HPE_PSO_UNROLL
#define HPE_PSO_UNROLL 2      // sweep-A steps per loop body in pso_kernel
#endif
#ifndef HPE_PSO_DYN
#define HPE_PSO_DYN 1         // 1: the warps of a CTA take the particle groups of a generation from a counter in shared memory
                              // (a warp whose particles were cheap takes the next group instead of waiting at the barrier)
#endif
#ifndef HPE_PSO_WARPS
#define HPE_PSO_WARPS 8
#endif
constexpr int PSO_WARPS = HPE_PSO_WARPS;
constexpr int PSO_NK = 3;            // informants drawn per particle, src/PSO.cpp:769

struct __align__(16) PsoSmemHead {
    FrameSmem fs;
    double gbest[HPE_NDOF];
    double x0[HPE_NDOF];
    double gbest_cost;
    int count;
    int fmin_id;
    int next_group;      // first particle of the next group to hand out (HPE_PSO_DYN)
    // followed by: warp scratch [PSO_WARPS][gsz] slots; double pcost[num_p];
    //              double fx[num_p]; int informant[num_p]; int R[num_p][3];
};

__host__ __device__ inline size_t pso_ws_bytes(int gsz) { return (size_t)PSO_WARPS * HPE_WS_BYTES(gsz); }
__host__ __device__ inline size_t pso_smem_bytes(int num_p, int gsz)
{
    return sizeof(PsoSmemHead) + pso_ws_bytes(gsz) + (size_t)num_p * (8 + 8 + 4 + PSO_NK * 4) + 16;
}

// lexicographic (cost, index) minimum over the warp; valid on all lanes
__device__ __forceinline__ void warp_argmin(double &c, int &i)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double oc = __shfl_xor_sync(0xffffffffu, c, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (oc < c || (oc == c && oi < i)) { c = oc; i = oi; }
    }
}

__global__ void __launch_bounds__(PSO_WARPS * 32, HPE_MINBLOCKS)
pso_kernel(DevModel m, ObsView o, PsoDev prm, const double *__restrict__ x0_in, int first, int count_swarms,
           const uint64_t *__restrict__ seeds, double *__restrict__ state_scratch,
           double *__restrict__ bestp_out, double *__restrict__ best_cost_out, double *__restrict__ trace_out,
           int chain, int seq_stride, int frame_stride, int gsz, double W1, double C1)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PsoSmemHead &sm = *reinterpret_cast<PsoSmemHead *>(smem_raw);
    const int num_p = prm.num_p, D = HPE_NDOF;
    unsigned char *ws_base = smem_raw + sizeof(PsoSmemHead);
    double *pcost = reinterpret_cast<double *>(ws_base + pso_ws_bytes(gsz));
    double *fx = pcost + num_p;
    int *informant = reinterpret_cast<int *>(fx + num_p);
    int *R = informant + num_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    // positions, velocities and personal bests of this CTA's swarm: global scratch (81 B per
    // particle and array, L2-resident); only the costs and the topology live in shared memory
    double *x = state_scratch + (size_t)blockIdx.x * 3 * num_p * D;
    double *v = x + (size_t)num_p * D;
    double *pb = v + (size_t)num_p * D;
    const double C2 = C1;

    stage_model(sm.fs, m);
    __syncthreads();

    // independent swarms: CTA-strided; chain mode: one CTA walks the sequence
    // sequences (chain = frames per sequence, count_swarms = sequences x chain): this CTA walks sequences q = blockIdx.x,
    // blockIdx.x + gridDim.x, ...; frame k of sequence q is swarm b = q * chain + k on slot
    // first + q * seq_stride + k * frame_stride and starts from the best pose of frame k - 1
    for (int q = chain ? (int)blockIdx.x : 0; q < (chain ? count_swarms / chain : 1); q += chain ? (int)gridDim.x : 1)
    for (int b = chain ? q * chain : (int)blockIdx.x; b < (chain ? (q + 1) * chain : count_swarms); b += chain ? 1 : (int)gridDim.x) {
        const bool head = !chain || b == q * chain;
        const int slot = first + (chain ? q * seq_stride + (b - q * chain) * frame_stride : b);
        const uint64_t seed = seeds ? seeds[b] : 1000ull;
        stage_frame(sm.fs, frame_view(o, slot));     // ends synchronised
        if (tid < D) {
            // chain mode: x0 = bestp of the previous frame (src/testmodel.cpp:138)
            sm.x0[tid] = head ? x0_in[(size_t)(chain ? q : b) * D + tid] : sm.gbest[tid];
        }
        __syncthreads();
        const bool frame_ok = o.meta[slot].ok != 0;

        // generate_particles: x0 + randn % std, v = 0 (src/PSO.cpp:56-74, :726)
        for (int it = tid; it < num_p * (D / 2); it += nthr) {
            int i = it / (D / 2), dp = it - i * (D / 2);
            double z[2];
            normal2(seed, (uint32_t)dp, (uint32_t)i, 0u, z);
            x[i * D + 2 * dp] = __dadd_rn(sm.x0[2 * dp], __dmul_rn(z[0], prm.std[2 * dp]));
            x[i * D + 2 * dp + 1] = __dadd_rn(sm.x0[2 * dp + 1], __dmul_rn(z[1], prm.std[2 * dp + 1]));
            v[i * D + 2 * dp] = 0.0;
            v[i * D + 2 * dp + 1] = 0.0;
        }
        if (tid == 0) { sm.gbest_cost = 1e100; sm.count = 100; sm.next_group = 0; }
        __syncthreads();

        for (int iter = 1;; ++iter) {
            // ---- evaluate every particle: cal_cost (src/PSO.cpp:748-752, 848-852) ----
#if HPE_PSO_DYN
            for (;;) {
                int p0 = 0;
                if (lane == 0) p0 = atomicAdd(&sm.next_group, gsz);
                p0 = __shfl_sync(0xffffffffu, p0, 0);
                if (p0 >= num_p) break;
                int cnt = min(gsz, num_p - p0);
#else
            const int per_round = PSO_WARPS * gsz;
            for (int p0 = warp * gsz; p0 < num_p; p0 += per_round) {
                int cnt = min(gsz, num_p - p0);
#endif
                const WarpScratch ws = warp_scratch(ws_base, warp, gsz);
                group_fk(sm.fs.M, [&](int g) { return x + (size_t)(p0 + g) * D; }, cnt, ws, lane);
                for (int g = 0; g < cnt; ++g) {
                    double depth = depth_and_stage(sm.fs, ws, g, lane);
                    double align = match_align<0, true, HPE_PSO_UNROLL>(sm.fs, ws, g, lane, nullptr, nullptr);
                    if (lane == 0) fx[p0 + g] = align + depth;
                }
                __syncwarp();
            }
            __syncthreads();
            if (tid == 0) sm.next_group = 0;       // every warp has left the loop above; the next fetch is several barriers away
            // ---- pbest (:753, :855-860) ----
            for (int it = tid; it < num_p * D; it += nthr) {
                int i = it / D;
                if (iter == 1 || fx[i] < pcost[i]) pb[it] = x[it];
            }
            __syncthreads();
            for (int i = tid; i < num_p; i += nthr)
                if (iter == 1 || fx[i] < pcost[i]) pcost[i] = fx[i];
            __syncthreads();
            // ---- gbest (:755-760, :864-876): first minimum of pcost ----
            if (warp == 0) {
                double c = __longlong_as_double(0x7ff0000000000000LL);
                int ci = 0x7fffffff;
                for (int i = lane; i < num_p; i += 32)
                    if (pcost[i] < c) { c = pcost[i]; ci = i; }
                warp_argmin(c, ci);
                if (lane == 0) {
                    if (ci == 0x7fffffff) ci = 0;      // all-NaN swarm
                    sm.fmin_id = ci;
                    if (pcost[ci] < sm.gbest_cost) {
                        sm.gbest_cost = pcost[ci];
                        sm.count = iter == 1 ? 100 : 0;
                        sm.fmin_id = ci | 0x40000000;  // flag: copy the position
                    } else if (iter > 1) {
                        sm.count += 1;
                    }
                }
            }
            __syncthreads();
            if ((sm.fmin_id & 0x40000000) && tid < D) sm.gbest[tid] = x[(sm.fmin_id & 0x3fffffff) * D + tid];
            if (tid == 0 && iter > 1 && trace_out)
                trace_out[(size_t)b * (prm.maxiter - 1) + (iter - 2)] = sm.gbest_cost;
            if (iter >= prm.maxiter) break;
            const uint32_t gen = (uint32_t)iter;       // generation number of the coming update
            __syncthreads();
            // ---- topology (:790-803): L = eye; L(s, R[s][k]) = 1 ----
            if (sm.count > 0) {
                for (int it = tid; it < num_p * PSO_NK; it += nthr) {
                    int s = it / PSO_NK, k = it - s * PSO_NK;
                    double u[2];
                    uniform2(seed, (uint32_t)(k >> 1), (uint32_t)s, gen, 3u, u);
                    R[it] = (int)floor(__dadd_rn(__dmul_rn(u[k & 1], (double)(num_p - 1)), 0.5));
                }
            }
            __syncthreads();
            // ---- informant = first minimum of pcost over {s : L(s,i) = 1} (:810-812) ----
            // Four threads per particle, each scanning every fourth s (round 2: one thread per particle, i.e. two warps of
            // a 64-particle swarm worked while six waited at the barrier -- 6 % of the kernel's stall samples).  The
            // reference's scan takes the first linked s whatever its cost and replaces it only by a strictly smaller cost:
            // that is the first linked s if its cost is NaN (or if no linked cost compares below +inf), else the
            // lexicographic minimum of (cost, s) over the linked s -- which is what the quad combines.
            for (int base = 0; base < num_p; base += nthr >> 2) {
                const int i = base + (tid >> 2), sub = tid & 3;
                int first = 0x7fffffff, best = 0x7fffffff;
                double bc = __longlong_as_double(0x7ff0000000000000LL);
                if (i < num_p) {
                    for (int s = sub; s < num_p; s += 4) {
                        const bool linked = (s == i) | (R[s * PSO_NK] == i) | (R[s * PSO_NK + 1] == i) | (R[s * PSO_NK + 2] == i);
                        if (linked) {
                            first = min(first, s);
                            const double c = pcost[s];
                            if (c < bc) { bc = c; best = s; }
                        }
                    }
                }
#pragma unroll
                for (int o = 1; o < 4; o <<= 1) {
                    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
                    const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
                    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
                    if (oc < bc || (oc == bc && ob < best)) { bc = oc; best = ob; }
                }
                if (i < num_p && sub == 0) {
                    const double cf0 = pcost[first];
                    informant[i] = (cf0 != cf0 || best == 0x7fffffff) ? first : best;
                }
            }
            __syncthreads();
            // ---- velocity, position, check_constraints (:824-840, :358-377) ----
            for (int it = tid; it < num_p * (D / 2); it += nthr) {
                int i = it / (D / 2), dp = it - i * (D / 2);
                int inf = informant[i];
                double rp[2], rg[2];
                uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 1u, rp);
                uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 2u, rg);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int d = 2 * dp + e, at = i * D + d;
                    double xi = x[at];
                    double nv = __dadd_rn(__dmul_rn(v[at], W1), __dmul_rn(__dmul_rn(rp[e], C1), __dadd_rn(pb[at], -xi)));
                    if (inf != i)
                        nv = __dadd_rn(nv, __dmul_rn(__dmul_rn(rg[e], C2), __dadd_rn(pb[inf * D + d], -xi)));
                    double nx = __dadd_rn(xi, nv);
                    // an upper-bound violation is reset to the LOWER bound (src/PSO.cpp:372)
                    if (nx < prm.lb[d] || nx > prm.ub[d]) { nx = prm.lb[d]; nv = 0.0; }
                    x[at] = nx;
                    v[at] = nv;
                }
            }
            __syncthreads();
        }
        __syncthreads();
        if (tid < D) bestp_out[(size_t)b * D + tid] = frame_ok ? sm.gbest[tid] : sm.x0[tid];
        if (tid == 0) best_cost_out[b] = frame_ok ? sm.gbest_cost : __longlong_as_double(0x7ff8000000000000LL);
    }
}

}  // namespace

HPE_DEFINE_FALLBACK_GETTER(hpe_fallbacks_pso)

cudaError_t hpe_launch_pso(const DevModel &m, const ObsView &o, const PsoDev &p, const double *x0_dev,
                           int first, int count, const uint64_t *seeds_dev, double *bestp_dev,
                           double *best_cost_dev, double *trace_dev, int chain, cudaStream_t st, int seq_stride, int frame_stride,
                           int scratch_id)
{
    if (scratch_id < 0 || scratch_id > 2) return cudaErrorInvalidValue;
    if (count <= 0) return cudaSuccess;
    if (chain < 0 || (chain > 0 && count % chain != 0)) return cudaErrorInvalidValue;
    if (seq_stride == 0) seq_stride = chain;
    const int units = chain ? count / chain : count;         // sequences or independent swarms
    // balance: rounds of PSO_WARPS groups, group size <= HPE_GROUP
    int rounds = (p.num_p + PSO_WARPS * HPE_GROUP - 1) / (PSO_WARPS * HPE_GROUP);
    int gsz = (p.num_p + PSO_WARPS * rounds - 1) / (PSO_WARPS * rounds);
    if (gsz < 1) gsz = 1;
    if (gsz > HPE_GROUP) gsz = HPE_GROUP;
    size_t smem = pso_smem_bytes(p.num_p, gsz);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static PerDevice<size_t> attr_set_pd;
    size_t &attr_set = attr_set_pd.get();
    if (smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pso_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = smem;
    }
    const int sms = hpe_sm_count();
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, pso_kernel, PSO_WARPS * 32, smem);
    if (resident < 1) resident = 1;
    int blocks = units < resident * sms ? units : resident * sms;
    // velocity + pbest of the resident swarms: grow-only scratch (L2-resident, 27 KB per swarm)
    static PerDevice<double *> scratch_pd[3];
    static PerDevice<size_t> scratch_bytes_pd[3];
    double *&scratch = scratch_pd[scratch_id].get();
    size_t &scratch_bytes = scratch_bytes_pd[scratch_id].get();
    size_t per = (size_t)p.num_p * HPE_NDOF * sizeof(double);
    if (3 * per * blocks > scratch_bytes) {
        if (scratch) cudaFree(scratch);
        scratch = nullptr; scratch_bytes = 0;
        cudaError_t e = cudaMalloc(&scratch, 3 * per * blocks);
        if (e != cudaSuccess) return e;
        scratch_bytes = 3 * per * blocks;
    }
    cudaError_t e;
    const double W1 = 1. / (2 * log(2.0));       // src/PSO.cpp:772
    const double C1 = 0.5 + log(2.0);            // src/PSO.cpp:773-774
    pso_kernel<<<blocks, PSO_WARPS * 32, smem, st>>>(m, o, p, x0_dev, first, count, seeds_dev, scratch, bestp_dev,
                                                     best_cost_dev, trace_dev, chain, seq_stride, frame_stride, gsz, W1, C1);
    e = cudaGetLastError();
    return e;
}
