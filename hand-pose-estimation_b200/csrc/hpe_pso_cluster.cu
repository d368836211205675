// hpe_pso_cluster.cu -- PSO::pso_evolve (src/PSO.cpp:717-886) for ONE swarm spread over a thread-block
// cluster of 8 CTAs (64 warps): the latency path.  A single 64 x 30 swarm is a chain of 30 dependent
// generations; on one CTA (hpe_pso.cu) each warp evaluates 8 particles per generation back to back.
// Here every warp owns one particle per round, the 8 CTAs of the cluster meet at two hardware cluster
// barriers per generation, and the swarm state they share (x, v, pbest, pcost, fx) lives in global
// memory (L2).  Everything that is cheap and needs the whole swarm -- the first-minimum gbest, the
// `count` bookkeeping, the topology redraw and the informants -- is recomputed identically by every
// CTA in its own shared memory, so no broadcast is needed.  Same arithmetic, same Philox counters:
// results are bit-identical to pso_kernel (tests/test_gpu_swarm.py).
// Used for frame sequences (frame k+1 starts from frame k's best pose, src/testmodel.cpp:117-138) and for
// small batches of swarms; clusters stride over the batch.
#include <cooperative_groups.h>

#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int CL_WARPS = 8;     // warps per CTA
constexpr int CL_SIZE = 8;      // CTAs per cluster (portable maximum)
constexpr int CL_NK = 3;        // informants drawn per particle, src/PSO.cpp:769

struct __align__(16) ClusterSmemHead {
    FrameSmem fs;
    double gbest[HPE_NDOF];
    double x0[HPE_NDOF];
    double gbest_cost;
    int count;
    int fmin_id;
    // followed by: warp scratch [CL_WARPS] x 1 slot; int informant[num_p]; int R[num_p][3]
};

__host__ __device__ inline size_t cluster_smem_bytes(int num_p)
{
    return sizeof(ClusterSmemHead) + (size_t)CL_WARPS * HPE_WS_BYTES(1) + (size_t)num_p * (4 + CL_NK * 4) + 16;
}

__device__ __forceinline__ void warp_argmin_d(double &c, int &i)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double oc = __shfl_xor_sync(0xffffffffu, c, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (oc < c || (oc == c && oi < i)) { c = oc; i = oi; }
    }
}

// per-cluster global state: x, v, pb [P][26]; pcost[2][P] (double buffered); fx[P]
__host__ __device__ inline size_t cluster_state_doubles(int num_p) { return (size_t)num_p * (3 * HPE_NDOF + 3); }

__global__ void __launch_bounds__(CL_WARPS * 32, 2)
pso_cluster_kernel(DevModel m, ObsView o, PsoDev prm, const double *__restrict__ x0_in, int first, int count_swarms,
                   const uint64_t *__restrict__ seeds, double *__restrict__ state, double *__restrict__ bestp_out,
                   double *__restrict__ best_cost_out, double *__restrict__ trace_out, int chain, int seq_stride, int frame_stride,
                   double W1, double C1)
{
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClusterSmemHead &sm = *reinterpret_cast<ClusterSmemHead *>(smem_raw);
    const int num_p = prm.num_p, D = HPE_NDOF;
    unsigned char *ws_base = smem_raw + sizeof(ClusterSmemHead);
    int *informant = reinterpret_cast<int *>(ws_base + (size_t)CL_WARPS * HPE_WS_BYTES(1));
    int *R = informant + num_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CL_SIZE, n_clusters = gridDim.x / CL_SIZE;
    const int cwarp = crank * CL_WARPS + warp;                 // warp index inside the cluster, 0..63
    const int ctid = crank * nthr + tid, cthreads = CL_SIZE * nthr;
    double *x = state + (size_t)cluster_id * cluster_state_doubles(num_p);
    double *v = x + (size_t)num_p * D;
    double *pb = v + (size_t)num_p * D;
    double *pcost2 = pb + (size_t)num_p * D;                   // [2][num_p]
    double *fx = pcost2 + 2 * (size_t)num_p;
    const double C2 = C1;
    const WarpScratch ws = warp_scratch(ws_base, warp, 1);

    stage_model(sm.fs, m);
    __syncthreads();

    // independent swarms (chain == 0): swarm b on slot first + b, cluster-strided.  Sequences (chain = frames per sequence,
    // count_swarms = sequences x chain): this cluster walks sequences q = cluster_id, cluster_id + n_clusters, ...; frame k
    // of sequence q is swarm b = q * chain + k on slot first + q * seq_stride + k * frame_stride and starts from the best
    // pose of frame k - 1 (src/testmodel.cpp:138)
    for (int q = chain ? cluster_id : 0; q < (chain ? count_swarms / chain : 1); q += chain ? n_clusters : 1)
    for (int b = chain ? q * chain : cluster_id; b < (chain ? (q + 1) * chain : count_swarms); b += chain ? 1 : n_clusters) {
        const bool head = !chain || b == q * chain;             // first frame of its sequence: x0 comes from the caller
        const int slot = first + (chain ? q * seq_stride + (b - q * chain) * frame_stride : b);
        const uint64_t seed = seeds ? seeds[b] : 1000ull;
        stage_frame(sm.fs, frame_view(o, slot));               // every CTA stages the frame; ends synchronised
        if (tid < D) sm.x0[tid] = head ? x0_in[(size_t)(chain ? q : b) * D + tid] : sm.gbest[tid];
        if (tid == 0) { sm.gbest_cost = 1e100; sm.count = 100; }
        __syncthreads();
        // generate_particles (src/PSO.cpp:56-74): the cluster's threads share the num_p x 13 pairs
        for (int it = ctid; it < num_p * (D / 2); it += cthreads) {
            int i = it / (D / 2), dp = it - i * (D / 2);
            double z[2];
            normal2(seed, (uint32_t)dp, (uint32_t)i, 0u, z);
            x[i * D + 2 * dp] = __dadd_rn(sm.x0[2 * dp], __dmul_rn(z[0], prm.std[2 * dp]));
            x[i * D + 2 * dp + 1] = __dadd_rn(sm.x0[2 * dp + 1], __dmul_rn(z[1], prm.std[2 * dp + 1]));
            v[i * D + 2 * dp] = 0.0;
            v[i * D + 2 * dp + 1] = 0.0;
        }
        cluster.sync();

        int cur = 0;                                           // pcost buffer holding the current values
        for (int iter = 1;; ++iter) {
            // ---- evaluate: one particle per warp and round (src/PSO.cpp:748-752, 848-852) ----
            for (int i = cwarp; i < num_p; i += CL_SIZE * CL_WARPS) {
                group_fk<false, true>(sm.fs.M, [&](int) { return x + (size_t)i * D; }, 1, ws, lane);
                double depth = depth_and_stage(sm.fs, ws, 0, lane);
                double align = match_align<0>(sm.fs, ws, 0, lane, nullptr, nullptr);
                if (lane == 0) fx[i] = align + depth;
            }
            cluster.sync();                                     // (A) every fx is visible
            // ---- pbest, pcost (:753, :855-860); cluster threads share the num_p x 26 elements ----
            const double *pc_old = pcost2 + cur * num_p;
            double *pc_new = pcost2 + (cur ^ 1) * num_p;
            for (int it = ctid; it < num_p * D; it += cthreads) {
                int i = it / D;
                bool take = iter == 1 || __ldcg(fx + i) < __ldcg(pc_old + i);     // data of other CTAs: bypass L1
                if (take) pb[it] = __ldcg(x + it);
                if (it - i * D == 0) pc_new[i] = take ? __ldcg(fx + i) : __ldcg(pc_old + i);
            }
            cur ^= 1;
            cluster.sync();                                     // (B) pbest / pcost are visible
            const double *pcost = pcost2 + cur * num_p;
            // ---- gbest (:755-760, :864-876): first minimum of pcost, recomputed by every CTA ----
            if (warp == 0) {
                double c = __longlong_as_double(0x7ff0000000000000LL);
                int ci = 0x7fffffff;
                for (int i = lane; i < num_p; i += 32)
                    if (__ldcg(pcost + i) < c) { c = __ldcg(pcost + i); ci = i; }
                warp_argmin_d(c, ci);
                if (lane == 0) {
                    if (ci == 0x7fffffff) ci = 0;
                    sm.fmin_id = ci;
                    if (__ldcg(pcost + ci) < sm.gbest_cost) {
                        sm.gbest_cost = __ldcg(pcost + ci);
                        sm.count = iter == 1 ? 100 : 0;
                        sm.fmin_id = ci | 0x40000000;
                    } else if (iter > 1) {
                        sm.count += 1;
                    }
                }
            }
            __syncthreads();
            // gbest_pos = particles.col(fmin_id): that particle has just improved, so x == pbest
            if ((sm.fmin_id & 0x40000000) && tid < D) sm.gbest[tid] = __ldcg(pb + (sm.fmin_id & 0x3fffffff) * D + tid);
            if (crank == 0 && tid == 0 && iter > 1 && trace_out)
                trace_out[(size_t)b * (prm.maxiter - 1) + (iter - 2)] = sm.gbest_cost;
            if (iter >= prm.maxiter) break;
            const uint32_t gen = (uint32_t)iter;
            __syncthreads();
            // ---- topology (:790-803), every CTA draws the same R ----
            if (sm.count > 0) {
                for (int it = tid; it < num_p * CL_NK; it += nthr) {
                    int s = it / CL_NK, k = it - s * CL_NK;
                    double u[2];
                    uniform2(seed, (uint32_t)(k >> 1), (uint32_t)s, gen, 3u, u);
                    R[it] = (int)floor(__dadd_rn(__dmul_rn(u[k & 1], (double)(num_p - 1)), 0.5));
                }
            }
            __syncthreads();
            // ---- informants of this CTA's particles (:810-812) ----
            // a warp looks for the informant of the particle it evaluates: lanes scan the candidates s,
            // then a lexicographic (pcost, s) minimum over the warp = first minimum
            for (int i = cwarp; i < num_p; i += CL_SIZE * CL_WARPS) {
                // sequential semantics being reproduced: best = first linked s, replaced by a later s only
                // when pcost[s] < pcost[best]  (so a NaN cost never replaces, and a NaN first link sticks)
                const double INF = __longlong_as_double(0x7ff0000000000000LL);
                double bc = INF;
                int best = 0x7fffffff, first_linked = 0x7fffffff;
                for (int s = lane; s < num_p; s += 32) {
                    bool linked = (s == i) | (R[s * CL_NK] == i) | (R[s * CL_NK + 1] == i) | (R[s * CL_NK + 2] == i);
                    if (linked) {
                        double c = __ldcg(pcost + s);
                        double key = c == c ? c : INF;
                        first_linked = min(first_linked, s);
                        if (key < bc || best == 0x7fffffff) { bc = key; best = s; }
                    }
                }
                warp_argmin_d(bc, best);
                for (int o = 16; o > 0; o >>= 1) first_linked = min(first_linked, __shfl_xor_sync(0xffffffffu, first_linked, o));
                double cf = __ldcg(pcost + first_linked);
                if (lane == 0) informant[i] = cf == cf ? best : first_linked;
            }
            __syncthreads();
            // ---- velocity, position, check_constraints (:824-840, :358-377) for this CTA's particles ----
            const int rounds = (num_p + CL_SIZE * CL_WARPS - 1) / (CL_SIZE * CL_WARPS);
            for (int it = tid; it < rounds * CL_WARPS * (D / 2); it += nthr) {
                int li = it / (D / 2), dp = it - li * (D / 2);
                int i = crank * CL_WARPS + (li % CL_WARPS) + (li / CL_WARPS) * CL_SIZE * CL_WARPS;
                if (i >= num_p) continue;
                int inf = informant[i];
                double rp[2], rg[2];
                uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 1u, rp);
                uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 2u, rg);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int d = 2 * dp + e, at = i * D + d;
                    double xi = x[at];
                    double nv = __dadd_rn(__dmul_rn(v[at], W1), __dmul_rn(__dmul_rn(rp[e], C1), __dadd_rn(__ldcg(pb + at), -xi)));
                    if (inf != i)
                        nv = __dadd_rn(nv, __dmul_rn(__dmul_rn(rg[e], C2), __dadd_rn(__ldcg(pb + inf * D + d), -xi)));
                    double nx = __dadd_rn(xi, nv);
                    if (nx < prm.lb[d] || nx > prm.ub[d]) { nx = prm.lb[d]; nv = 0.0; }   // :372
                    x[at] = nx;
                    v[at] = nv;
                }
            }
            // a warp evaluates only particles its own CTA has just moved: a CTA barrier is enough
            __syncthreads();
        }
        const bool frame_ok = o.meta[slot].ok != 0;
        __syncthreads();
        if (crank == 0) {
            if (tid < D) bestp_out[(size_t)b * D + tid] = frame_ok ? sm.gbest[tid] : sm.x0[tid];
            if (tid == 0) best_cost_out[b] = frame_ok ? sm.gbest_cost : __longlong_as_double(0x7ff8000000000000LL);
        }
        cluster.sync();      // nobody starts overwriting x for the next swarm while others still read pb
    }
}

// ---------------------------------------------------------------------------------------
// The same swarm-per-cluster scheme with the shared swarm state in DISTRIBUTED SHARED MEMORY (swarms of up to
// CLD_MAXP particles): the warp that evaluates a particle pushes its (fx, position) row into every CTA of the
// cluster (st.shared::cluster through map_shared_rank), so after ONE cluster barrier per generation each CTA
// holds all rows and runs the whole bookkeeping -- pbest / pcost, first-minimum gbest, `count`, topology,
// informants, the move of its own particles -- out of its own shared memory: no L2 round trips, no second
// barrier.  Incoming rows are double buffered by generation parity (a CTA can be at most one barrier ahead of
// a CTA that still reads the previous rows).  Same arithmetic and counters: bit-identical to pso_kernel.
// ---------------------------------------------------------------------------------------
constexpr int CLD_MAXP = 128;

struct __align__(16) ClusterDsmHead {
    FrameSmem fs;
    double gbest[HPE_NDOF];
    double x0[HPE_NDOF];
    double gbest_cost;
    int count;
    int fmin_id;
    int pad;
    // followed by: warp scratch [CLW] x 1 slot; double rows[2][num_p][27] (fx, x) incoming; double pb[num_p][26];
    //              double pcost[num_p]; int informant[num_p]; int R[num_p][3]
};

__host__ __device__ inline size_t cluster_dsm_smem_bytes(int num_p, int warps)
{
    return sizeof(ClusterDsmHead) + (size_t)warps * HPE_WS_BYTES(1) +
           (size_t)num_p * (2 * (HPE_NDOF + 1) * 8 + HPE_NDOF * 8 + 8 + 4 + CL_NK * 4) + 16;
}

// CLS CTAs per cluster, 64 / CLS warps per CTA: 8 x 8 (portable cluster size, two CTAs per SM) for small batches of
// swarms, 16 x 4 (non-portable size, one CTA per SM so that a warp has its SM sub-partition to itself) for a single
// sequence: +15 % frames/s
template <int CLS>
__global__ void __launch_bounds__(64 / CLS * 32, 2)
pso_cluster_dsm_kernel(DevModel m, ObsView o, PsoDev prm, const double *__restrict__ x0_in, int first, int count_swarms,
                       const uint64_t *__restrict__ seeds, double *__restrict__ state, double *__restrict__ bestp_out,
                       double *__restrict__ best_cost_out, double *__restrict__ trace_out, int chain, int seq_stride, int frame_stride,
                   double W1, double C1)
{
    constexpr int CLW = 64 / CLS;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClusterDsmHead &sm = *reinterpret_cast<ClusterDsmHead *>(smem_raw);
    const int num_p = prm.num_p, D = HPE_NDOF, ROWD = HPE_NDOF + 1;
    unsigned char *ws_base = smem_raw + sizeof(ClusterDsmHead);
    double *rows = reinterpret_cast<double *>(ws_base + (size_t)CLW * HPE_WS_BYTES(1));     // [2][num_p][27]
    double *pb = rows + 2 * (size_t)num_p * ROWD;                                                // [num_p][26]
    double *pcost = pb + (size_t)num_p * D;                                                      // [num_p]
    int *informant = reinterpret_cast<int *>(pcost + num_p);
    int *R = informant + num_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / CLS, n_clusters = gridDim.x / CLS;
    const int cwarp = crank * CLW + warp;                 // warp index inside the cluster, 0..63
    // positions and velocities of this CTA's own particles stay in global scratch (only this CTA touches them)
    double *x = state + (size_t)cluster_id * 2 * num_p * D;
    double *v = x + (size_t)num_p * D;
    const double C2 = C1;
    const WarpScratch ws = warp_scratch(ws_base, warp, 1);
    double *peer_rows[CLS];
#pragma unroll
    for (int r = 0; r < CLS; ++r) peer_rows[r] = cluster.map_shared_rank(rows, r);

    stage_model(sm.fs, m);
    __syncthreads();

    // independent swarms (chain == 0): swarm b on slot first + b, cluster-strided.  Sequences (chain = frames per sequence,
    // count_swarms = sequences x chain): this cluster walks sequences q = cluster_id, cluster_id + n_clusters, ...; frame k
    // of sequence q is swarm b = q * chain + k on slot first + q * seq_stride + k * frame_stride and starts from the best
    // pose of frame k - 1 (src/testmodel.cpp:138)
    for (int q = chain ? cluster_id : 0; q < (chain ? count_swarms / chain : 1); q += chain ? n_clusters : 1)
    for (int b = chain ? q * chain : cluster_id; b < (chain ? (q + 1) * chain : count_swarms); b += chain ? 1 : n_clusters) {
        const bool head = !chain || b == q * chain;             // first frame of its sequence: x0 comes from the caller
        const int slot = first + (chain ? q * seq_stride + (b - q * chain) * frame_stride : b);
        const uint64_t seed = seeds ? seeds[b] : 1000ull;
        stage_frame(sm.fs, frame_view(o, slot));               // every CTA stages the frame; ends synchronised
        if (tid < D) sm.x0[tid] = head ? x0_in[(size_t)(chain ? q : b) * D + tid] : sm.gbest[tid];
        if (tid == 0) { sm.gbest_cost = 1e100; sm.count = 100; }
        __syncthreads();
        // generate_particles (src/PSO.cpp:56-74) for this CTA's own particles
        const int rounds = (num_p + CLS * CLW - 1) / (CLS * CLW);
        for (int it = tid; it < rounds * CLW * (D / 2); it += nthr) {
            int li = it / (D / 2), dp = it - li * (D / 2);
            int i = crank * CLW + (li % CLW) + (li / CLW) * CLS * CLW;
            if (i >= num_p) continue;
            double z[2];
            normal2(seed, (uint32_t)dp, (uint32_t)i, 0u, z);
            x[i * D + 2 * dp] = __dadd_rn(sm.x0[2 * dp], __dmul_rn(z[0], prm.std[2 * dp]));
            x[i * D + 2 * dp + 1] = __dadd_rn(sm.x0[2 * dp + 1], __dmul_rn(z[1], prm.std[2 * dp + 1]));
            v[i * D + 2 * dp] = 0.0;
            v[i * D + 2 * dp + 1] = 0.0;
        }
        __syncthreads();

        for (int iter = 1;; ++iter) {
            double *in = rows + (size_t)(iter & 1) * num_p * ROWD;     // this generation's rows, local copy
            // ---- evaluate: one particle per warp and round (src/PSO.cpp:748-752, 848-852); the row goes to every CTA ----
            for (int i = cwarp; i < num_p; i += CLS * CLW) {
                group_fk<false, true>(sm.fs.M, [&](int) { return x + (size_t)i * D; }, 1, ws, lane);
                double depth = depth_and_stage(sm.fs, ws, 0, lane);
                double align = match_align<0>(sm.fs, ws, 0, lane, nullptr, nullptr);
                if (lane < ROWD) {
                    const double val = lane == 0 ? align + depth : x[(size_t)i * D + lane - 1];
                    const size_t at = ((size_t)(iter & 1) * num_p + i) * ROWD + lane;
#pragma unroll
                    for (int r = 0; r < CLS; ++r) peer_rows[r][at] = val;
                }
            }
            // Cluster barrier in two halves.  Between arriving and waiting -- the time a CTA spends waiting for the slowest
            // particle of the cluster anyway -- every thread takes the random draws of the coming update: they depend on the
            // seed and the generation only, not on this generation's costs (its share of the topology, used only if the
            // swarm turns out not to have improved, and the rp / rg of its (particle, dimension pair) item of the move).
            // The Philox rounds leave the serial part of a generation (8.85 -> 8.77 us on a 16-CTA cluster: +1 % tracked frames/s).
            auto token = cluster.barrier_arrive();
            const uint32_t gen = (uint32_t)iter;                // generation number of the coming update
            constexpr int TOPO_PER_THREAD = 3;                  // num_p <= 128, >= 128 threads: at most 3 of the 3 num_p draws each
            int topo[TOPO_PER_THREAD] = {0, 0, 0};
            double rp[2] = {0.0, 0.0}, rg[2] = {0.0, 0.0};
            if (iter < prm.maxiter) {
#pragma unroll
                for (int k3 = 0; k3 < TOPO_PER_THREAD; ++k3) {
                    const int it = tid + k3 * nthr;
                    if (it < num_p * CL_NK) {
                        int s = it / CL_NK, k = it - s * CL_NK;
                        double u[2];
                        uniform2(seed, (uint32_t)(k >> 1), (uint32_t)s, gen, 3u, u);
                        topo[k3] = (int)floor(__dadd_rn(__dmul_rn(u[k & 1], (double)(num_p - 1)), 0.5));
                    }
                }
                if (tid < rounds * CLW * (D / 2)) {             // the thread's item of the move loop below (at most one)
                    int li = tid / (D / 2), dp = tid - li * (D / 2);
                    int i = crank * CLW + (li % CLW) + (li / CLW) * CLS * CLW;
                    if (i < num_p) {
                        uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 1u, rp);
                        uniform2(seed, (uint32_t)dp, (uint32_t)i, gen, 2u, rg);
                    }
                }
            }
            cluster.barrier_wait(std::move(token));             // every row of this generation is in every CTA
            // ---- pbest, pcost (:753, :855-860), every CTA for all particles, locally ----
            for (int it = tid; it < num_p * D; it += nthr) {
                int i = it / D;
                bool take = iter == 1 || in[i * ROWD] < pcost[i];
                if (take) pb[it] = in[i * ROWD + 1 + (it - i * D)];
            }
            __syncthreads();
            for (int i = tid; i < num_p; i += nthr)
                if (iter == 1 || in[i * ROWD] < pcost[i]) pcost[i] = in[i * ROWD];
            __syncthreads();
            // ---- gbest (:755-760, :864-876): first minimum of pcost ----
            if (warp == 0) {
                double c = __longlong_as_double(0x7ff0000000000000LL);
                int ci = 0x7fffffff;
                for (int i = lane; i < num_p; i += 32)
                    if (pcost[i] < c) { c = pcost[i]; ci = i; }
                warp_argmin_d(c, ci);
                if (lane == 0) {
                    if (ci == 0x7fffffff) ci = 0;
                    sm.fmin_id = ci;
                    if (pcost[ci] < sm.gbest_cost) {
                        sm.gbest_cost = pcost[ci];
                        sm.count = iter == 1 ? 100 : 0;
                        sm.fmin_id = ci | 0x40000000;
                    } else if (iter > 1) {
                        sm.count += 1;
                    }
                }
            }
            __syncthreads();
            // gbest_pos = particles.col(fmin_id): that particle has just improved, so x == pbest
            if ((sm.fmin_id & 0x40000000) && tid < D) sm.gbest[tid] = pb[(sm.fmin_id & 0x3fffffff) * D + tid];
            if (crank == 0 && tid == 0 && iter > 1 && trace_out)
                trace_out[(size_t)b * (prm.maxiter - 1) + (iter - 2)] = sm.gbest_cost;
            if (iter >= prm.maxiter) break;
            __syncthreads();
            // ---- topology (:790-803), every CTA takes the same R: the draws made while waiting at the cluster barrier ----
            if (sm.count > 0) {
#pragma unroll
                for (int k3 = 0; k3 < TOPO_PER_THREAD; ++k3)
                    if (tid + k3 * nthr < num_p * CL_NK) R[tid + k3 * nthr] = topo[k3];
            }
            __syncthreads();
            // ---- informants of this CTA's particles (:810-812) ----
            for (int i = cwarp; i < num_p; i += CLS * CLW) {
                const double INF = __longlong_as_double(0x7ff0000000000000LL);
                double bc = INF;
                int best = 0x7fffffff, first_linked = 0x7fffffff;
                for (int s = lane; s < num_p; s += 32) {
                    bool linked = (s == i) | (R[s * CL_NK] == i) | (R[s * CL_NK + 1] == i) | (R[s * CL_NK + 2] == i);
                    if (linked) {
                        double c = pcost[s];
                        double key = c == c ? c : INF;
                        first_linked = min(first_linked, s);
                        if (key < bc || best == 0x7fffffff) { bc = key; best = s; }
                    }
                }
                warp_argmin_d(bc, best);
                for (int off = 16; off > 0; off >>= 1) first_linked = min(first_linked, __shfl_xor_sync(0xffffffffu, first_linked, off));
                double cf = pcost[first_linked];
                if (lane == 0) informant[i] = cf == cf ? best : first_linked;
            }
            __syncthreads();
            // ---- velocity, position, check_constraints (:824-840, :358-377) for this CTA's particles; the thread's item
            //      is the one whose rp / rg it drew above ----
            if (tid < rounds * CLW * (D / 2)) {
                const int it = tid;
                int li = it / (D / 2), dp = it - li * (D / 2);
                int i = crank * CLW + (li % CLW) + (li / CLW) * CLS * CLW;
                if (i < num_p) {
                int inf = informant[i];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int d = 2 * dp + e, at = i * D + d;
                    double xi = x[at];
                    double nv = __dadd_rn(__dmul_rn(v[at], W1), __dmul_rn(__dmul_rn(rp[e], C1), __dadd_rn(pb[at], -xi)));
                    if (inf != i)
                        nv = __dadd_rn(nv, __dmul_rn(__dmul_rn(rg[e], C2), __dadd_rn(pb[inf * D + d], -xi)));
                    double nx = __dadd_rn(xi, nv);
                    if (nx < prm.lb[d] || nx > prm.ub[d]) { nx = prm.lb[d]; nv = 0.0; }   // :372
                    x[at] = nx;
                    v[at] = nv;
                }
                }
            }
            // a warp evaluates only particles its own CTA has just moved: a CTA barrier is enough
            __syncthreads();
        }
        const bool frame_ok = o.meta[slot].ok != 0;
        __syncthreads();
        if (crank == 0) {
            if (tid < D) bestp_out[(size_t)b * D + tid] = frame_ok ? sm.gbest[tid] : sm.x0[tid];
            if (tid == 0) best_cost_out[b] = frame_ok ? sm.gbest_cost : __longlong_as_double(0x7ff8000000000000LL);
        }
        cluster.sync();      // nobody pushes rows of the next swarm while others still read this one's
    }
}

}  // namespace

HPE_DEFINE_FALLBACK_GETTER(hpe_fallbacks_pso_cluster)

int g_hpe_cluster_dsm = 1;   // hpe_set_pso_path(3): keep the cluster kernel's swarm state in global memory (cross-check)

cudaError_t hpe_launch_pso_cluster(const DevModel &m, const ObsView &o, const PsoDev &p, const double *x0_dev, int first,
                                   int count, const uint64_t *seeds_dev, double *bestp_dev, double *best_cost_dev,
                                   double *trace_dev, int chain, cudaStream_t st, int seq_stride, int frame_stride)
{
    if (count <= 0) return cudaSuccess;
    if (chain < 0 || (chain > 0 && count % chain != 0)) return cudaErrorInvalidValue;
    if (seq_stride == 0) seq_stride = chain;
    const int units = chain ? count / chain : count;         // sequences or independent swarms
    const bool dsm = p.num_p <= CLD_MAXP && g_hpe_cluster_dsm;       // swarm state in distributed shared memory
    const int sms = hpe_sm_count();
    // a sequence (or a handful of swarms) gets 16-CTA clusters, one CTA per SM; more swarms get 8-CTA clusters
    static PerDevice<bool> wide_refused_pd;                  // set if the device refuses 16-CTA clusters
    bool &wide_refused = wide_refused_pd.get();
    // how many 16-CTA clusters the device can hold at once (a cluster lives inside one GPC): asked once per device
    static PerDevice<int> wide_max_pd;
    int &wide_max = wide_max_pd.get();
    if (!wide_max && dsm && !wide_refused) {
        cudaFuncSetAttribute(pso_cluster_dsm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024);
        cudaFuncSetAttribute(pso_cluster_dsm_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(16 * (sms / 16)); q.blockDim = dim3(64 / 16 * 32); q.dynamicSmemBytes = 116 * 1024;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa; q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, pso_cluster_dsm_kernel<16>, &q) != cudaSuccess || n < 1) { cudaGetLastError(); n = -1; }
        wide_max = n;
    }
    // sequences / swarms that all fit at once get 16-CTA clusters (one CTA per SM); more get 8-CTA clusters
    const bool wide = !wide_refused && dsm && wide_max > 0 && units <= wide_max;
    const int cls = wide ? 16 : CL_SIZE;
    size_t smem = dsm ? cluster_dsm_smem_bytes(p.num_p, 64 / cls) : cluster_smem_bytes(p.num_p);
    if (wide && smem < 116 * 1024) smem = 116 * 1024;       // more than half of an SM's shared memory: one CTA per SM
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    struct Attr3 { size_t v[3]; };
    static PerDevice<Attr3> attr_pd;
    size_t *attr_set = attr_pd.get().v;
    const int which = dsm ? (wide ? 2 : 1) : 0;
    if (smem > attr_set[which]) {
        cudaError_t e;
        if (which == 0) e = cudaFuncSetAttribute(pso_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else if (which == 1) e = cudaFuncSetAttribute(pso_cluster_dsm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else {
            e = cudaFuncSetAttribute(pso_cluster_dsm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(pso_cluster_dsm_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        }
        if (e != cudaSuccess) return e;
        attr_set[which] = smem;
    }
    int max_clusters = wide ? wide_max : 2 * sms / cls;
    int n_clusters = units < max_clusters ? units : max_clusters;
    static PerDevice<double *> state_pd;
    static PerDevice<size_t> state_bytes_pd;
    double *&state = state_pd.get();
    size_t &state_bytes = state_bytes_pd.get();
    size_t need = cluster_state_doubles(p.num_p) * sizeof(double) * n_clusters;
    if (need > state_bytes) {
        if (state) cudaFree(state);
        state = nullptr; state_bytes = 0;
        cudaError_t e = cudaMalloc(&state, need);
        if (e != cudaSuccess) return e;
        state_bytes = need;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * cls);
    cfg.blockDim = dim3(64 / cls * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cls;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const double W1 = 1. / (2 * log(2.0));       // src/PSO.cpp:772
    const double C1 = 0.5 + log(2.0);            // src/PSO.cpp:773-774
    if (dsm && wide) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, pso_cluster_dsm_kernel<16>, m, o, p, x0_dev, first, count, seeds_dev, state,
                                           bestp_dev, best_cost_dev, trace_dev, chain, seq_stride, frame_stride, W1, C1);
        if (e == cudaSuccess) return e;
        cudaGetLastError();                                  // e.g. cudaErrorInvalidClusterSize: use 8-CTA clusters from now on
        wide_refused = true;
        return hpe_launch_pso_cluster(m, o, p, x0_dev, first, count, seeds_dev, bestp_dev, best_cost_dev, trace_dev, chain, st,
                                      seq_stride, frame_stride);
    }
    if (dsm)
        return cudaLaunchKernelEx(&cfg, pso_cluster_dsm_kernel<8>, m, o, p, x0_dev, first, count, seeds_dev, state, bestp_dev,
                                  best_cost_dev, trace_dev, chain, seq_stride, frame_stride, W1, C1);
    return cudaLaunchKernelEx(&cfg, pso_cluster_kernel, m, o, p, x0_dev, first, count, seeds_dev, state, bestp_dev,
                              best_cost_dev, trace_dev, chain, seq_stride, frame_stride, W1, C1);
}
