// hpe_cost.cu -- batched handmodel::build_hand_model and costfunc::cal_cost/cal_cost2 kernels,
// plus the sphere-model ray-caster used to synthesise Subject1-format frames.
#include <atomic>
#include <cooperative_groups.h>

#include "hpe_eval.cuh"
#include "hpe_internal.h"

namespace {

constexpr int COST_WARPS = 8;
#ifndef HPE_GROUP_DEPTH
#define HPE_GROUP_DEPTH 0  // 1: depth term of a whole group of poses at once (group_depth: fewer instructions, but its unrolled
                           // rounds cost more in instruction-cache misses than they save -- measured -12 %), 0: pose by pose
#endif
}  // namespace

int g_hpe_matcher = 0;   // 0 = certified two-sweep matcher, 1 = direct form on all 48 spheres (cross-check)

namespace {

// ---------------------------------------------------------------------------------------
// handmodel::build_hand_model for n poses (src/handmodel.cpp:259-298): thread = (pose, digit)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160) fk_kernel(DevModel m, const double *__restrict__ theta, int64_t n,
                                                  double *__restrict__ spheres, double *__restrict__ joints)
{
    __shared__ DevModel Ms;
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 8); i += blockDim.x)
        reinterpret_cast<double *>(&Ms)[i] = reinterpret_cast<const double *>(&m)[i];
    __syncthreads();
    int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t pose = item / 5;
    int d = (int)(item - pose * 5);
    if (pose >= n) return;
    const double *th = theta + pose * HPE_NDOF;
    double J[5][3], S[10][3];
    fk_digit(Ms, d, th, J);
    int cnt = digit_spheres(d, J, S);
    int first = digit_first_sphere(d);
    if (spheres) {
        double *out = spheres + (pose * HPE_NSPH + first) * 3;
        for (int i = 0; i < cnt; ++i) {
            out[i * 3 + 0] = S[i][0];
            out[i * 3 + 1] = -S[i][1];   // sphere_centres.cols(1,2) *= -1, src/handmodel.cpp:288
            out[i * 3 + 2] = -S[i][2];
        }
    }
    if (joints) {
        // hand_joints rows: 0 = g_pos, 1-16 = index..little joints 1..4, 17-20 = thumb (:291-296)
        double *out = joints + pose * HPE_NJOINT * 3;
        if (d == 0) { out[0] = th[3]; out[1] = th[4]; out[2] = th[5]; }
        int row = d == 0 ? 17 : 1 + 4 * (d - 1);
        for (int j = 1; j < 5; ++j)
            for (int k = 0; k < 3; ++k) out[(row + j - 1) * 3 + k] = J[j][k];
    }
}

// ---------------------------------------------------------------------------------------
// cal_cost / cal_cost2 for n poses against one frame.  One CTA of SCORE_WARPS warps per SM; the warps take
// groups of SCORE_GROUP poses (then smaller units) from counters in global memory until the batch is used up.
// ---------------------------------------------------------------------------------------
// Self-scheduling counters of cost_kernel: a ring of slots, one per launch in flight.  The launcher hands out (slot, ticket)
// round-robin; thread 0 of every CTA CLAIMS the slot for the launch's ticket before the CTA touches the counters (a slot
// still owned by an older launch that is running -- possible when a long launch on one stream is lapped by a ring's worth
// of short ones on another -- is waited for: its owner is running), and the last warp of a launch zeroes the counters and
// then releases the slot.
struct SchedSlot { int next, tail, done; unsigned owner; };
constexpr int HPE_SCHED_SLOTS = 4096;
__device__ SchedSlot g_sched[HPE_SCHED_SLOTS];
#ifndef HPE_TAIL_R1
#define HPE_TAIL_R1 3.0f     // poses per warp handed out in units smaller than a group (the warps leave their last group over a window of HPE_GROUP poses)
#endif
#ifndef HPE_TAIL_R2
#define HPE_TAIL_R2 1.0f     // ... of these, poses per warp handed out one by one
#endif
#ifndef HPE_TAIL_C2
#define HPE_TAIL_C2 3
#endif

struct __align__(16) CostSmem {
    FrameSmem fs;
    __align__(16) unsigned char ws[COST_WARPS * HPE_WS_BYTES(HPE_GROUP)];
};
// cost_kernel's own shape (the other kernels of this file keep COST_WARPS x HPE_MINBLOCKS and groups of HPE_GROUP): ONE CTA of
// 24 warps per SM.  With self-scheduling there is nothing to balance between CTAs, one copy of the frame per SM instead of
// three frees 25 KB of shared memory, and that pays for groups of SIX poses: the forward kinematics then runs on 30 of 32 lanes.
// Measured (evaluations/s): 8 x 3, groups of 5: 3.357e8; 12 x 2: 3.365e8; 24 x 1: 3.373e8; 24 x 1, groups of 6: 3.403e8;
// 25 x 1 (the register file would hold it): ptxas caps the kernel at 72 registers, 3.31e8.
#ifndef HPE_SCORE_WARPS
#define HPE_SCORE_WARPS 24
#endif
#ifndef HPE_SCORE_CTAS
#define HPE_SCORE_CTAS 1
#endif
#ifndef HPE_SCORE_GROUP
#define HPE_SCORE_GROUP 6
#endif
constexpr int SCORE_WARPS = HPE_SCORE_WARPS, SCORE_CTAS = HPE_SCORE_CTAS, SCORE_GROUP = HPE_SCORE_GROUP;
static_assert(SCORE_GROUP <= 6, "five lanes per pose, dep[] holds six values");
struct __align__(16) ScoreSmem {
    FrameSmem fs;
    __align__(16) unsigned char ws[SCORE_WARPS * HPE_WS_BYTES(SCORE_GROUP)];
};

// AUX: frozen correspondences in and / or correspondences out (cal_cost2's matchId); the plain scoring kernel is
// compiled without them (fewer live registers and no per-point predicates in the alignment loop)
template <bool WITH_COLLI, int MATCHER, bool AUX>
__global__ void __launch_bounds__(SCORE_WARPS * 32, SCORE_CTAS)
cost_kernel(DevModel m, ObsView o, int slot, const double *__restrict__ theta, int64_t n,
            double *__restrict__ cost, double *__restrict__ terms, int64_t tstride,
            const int32_t *__restrict__ frozen, int32_t *__restrict__ match_out, int sched_slot, unsigned ticket, int n1, int n2, int t2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScoreSmem &sm = *reinterpret_cast<ScoreSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {                      // claim the counters for this launch (ordered before their use by the barriers below)
        unsigned *owner = &g_sched[sched_slot].owner;
        unsigned cur = atomicCAS(owner, 0u, ticket);
        // bounded like every other wait in this library (~30 s): an owner that never releases means a launch died with the
        // slot in its hands; the context is lost then anyway, and a trap is reported where a hang would not be
        for (long long spins = 0; cur != 0u && cur != ticket; ++spins) {
            if (spins > (1ll << 27)) __trap();
            __nanosleep(256);
            cur = atomicCAS(owner, 0u, ticket);
        }
    }
    stage_model(sm.fs, m);
    stage_frame(sm.fs, frame_view(o, slot));     // ends synchronised
    const WarpScratch ws = warp_scratch(sm.ws, warp, SCORE_GROUP);
    const int n_pts = sm.fs.n_pts;

    // Work distribution.  Poses are handed out through two counters in global memory (self-scheduling): groups of HPE_GROUP
    // from [0, n1), then -- once these run out -- units of HPE_TAIL_C2 from [n1, n2) and single poses from [n2, n).  A static split
    // (round 1: n / warps poses each) leaves the warps of a scheduler finishing at very different times: ncu showed 20.2 of
    // 24 warps resident on average and the last sixth of the kernel running with too few warps to hide the FP64 / gather
    // latencies.  With the counters every warp works until the pool is empty; the shrinking units keep the final idle
    // window below one pose (a group costs ~55 us, a pose ~11 us of a 216 us kernel).  32-bit pose indices (a launch
    // scores < 2^31 poses, hpe_launch_cost splits larger batches).  The loop variables live in the warp's shared-memory
    // scratch and are re-read where they are used: every register is taken inside the matching phases, and a loop counter
    // spilled to local memory costs an L1 miss per pose (ncu: 3.5 % of the kernel's stall samples).
    const int total_warps = (int)gridDim.x * SCORE_WARPS;
    SchedSlot *const S = g_sched + sched_slot;
    volatile int *wl = reinterpret_cast<volatile int *>(ws.dep + 6);      // [0] = first pose of the group, [1] = main range exhausted, [2] = group size, [3] = g
    if (lane == 0) wl[1] = 0;
    __syncwarp();
#pragma unroll 1
    for (;;) {
        if (lane == 0) {
            int s = 0, cnt = 0;
            if (!wl[1]) {
                s = atomicAdd(&S->next, SCORE_GROUP);
                cnt = min(SCORE_GROUP, n1 - s);
                if (cnt <= 0) wl[1] = 1;
            }
            if (cnt <= 0) {
                const int t = atomicAdd(&S->tail, 1);
                if (t < t2) { s = n1 + HPE_TAIL_C2 * t; cnt = min(HPE_TAIL_C2, n2 - s); }
                else { s = n2 + (t - t2); cnt = s < (int)n ? 1 : 0; }
            }
            wl[0] = s; wl[2] = cnt; wl[3] = 0;
        }
        __syncwarp();
        if (wl[2] <= 0) break;
        group_fk(sm.fs.M, [&](int g) { return theta + (size_t)(wl[0] + g) * HPE_NDOF; }, wl[2], ws, lane);
#if HPE_GROUP_DEPTH
        group_depth<SCORE_GROUP>(sm.fs, ws, wl[2], lane);
#endif
#pragma unroll 1
        while (wl[3] < wl[2]) {
#if HPE_GROUP_DEPTH
            stage_pose(sm.fs, ws, wl[3], lane);
#else
            {
                // (issuing the depth gathers before stage_pose + prune_blocks and consuming them afterwards -- depth_issue /
                //  depth_finish -- was measured 5 % slower: the values in flight cost registers the matcher needs)
                const double depth = depth_and_stage(sm.fs, ws, wl[3], lane);
                if (lane == 0) reinterpret_cast<volatile double *>(ws.dep)[wl[3]] = depth;     // parked like the loop variables
            }
#endif
            double align = match_align<MATCHER, AUX, 1>(sm.fs, ws, wl[3], lane, AUX ? frozen : nullptr,
                                                     AUX && match_out ? match_out + (size_t)(wl[0] + wl[3]) * n_pts : nullptr);
            double colli = 0.0;
            if (WITH_COLLI) colli = warp_sum(collision_lane(sm.fs.rad, ws.sphd + (size_t)wl[3] * HPE_NSPH * 3, lane));
            __syncwarp();
            if (lane == 0) {
                const int64_t at = wl[0] + wl[3];
                const double depth = reinterpret_cast<volatile double *>(ws.dep)[wl[3]];
                if (cost) cost[at] = WITH_COLLI ? align + depth + colli : align + depth;
                if (terms) {
                    terms[at] = align;           // SoA: align | depth | colli, tstride apart
                    terms[tstride + at] = depth;
                    terms[2 * tstride + at] = colli;
                }
                wl[3] = wl[3] + 1;
            }
            __syncwarp();
        }
    }
    // the last warp out re-arms the counters for the next launch that is given this slot
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(&S->done, 1) == total_warps - 1) {
            S->next = 0; S->tail = 0; S->done = 0;
            __threadfence();
            atomicExch(&S->owner, 0u);           // released only once the counters read zero
        }
    }
}

// ---------------------------------------------------------------------------------------
// cal_cost / cal_cost2 for a batch whose poses are STILL ARRIVING from the host: the kernel is
// launched before the upload; the poses land chunk by chunk (DMA on another stream) and each chunk
// raises arrived[c].  Warps take groups in order (group q = warp, warp + W, ...), wait for the chunk
// of their group and score it, so the evaluation chases the copy instead of waiting for it and
// every warp still walks several groups (no lock-step of the phases).  Opt-in (hpe_set_cost_pipeline(1, 0)):
// it needs the copy engine and the SMs to run concurrently, which profilers and sanitizers prevent.
// ---------------------------------------------------------------------------------------
template <bool WITH_COLLI>
__global__ void __launch_bounds__(COST_WARPS * 32, HPE_MINBLOCKS)
cost_chase_kernel(DevModel m, ObsView o, int slot, const double *theta, int64_t n, double *__restrict__ terms,
                  int64_t tstride, const int32_t *__restrict__ frozen, int32_t *__restrict__ match_out,
                  const unsigned *arrived, ChaseChunks ck, int *timeout_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CostSmem &sm = *reinterpret_cast<CostSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage_model(sm.fs, m);
    stage_frame(sm.fs, frame_view(o, slot));     // ends synchronised
    const WarpScratch ws = warp_scratch(sm.ws, warp, HPE_GROUP);
    const int n_pts = sm.fs.n_pts;
    const int64_t total_warps = (int64_t)gridDim.x * COST_WARPS;
    const int64_t gw = (int64_t)blockIdx.x * COST_WARPS + warp;
    // groups of HPE_GROUP poses, except the last ~one-per-warp poses, which are taken one at a time so that the
    // work left when the final chunk lands is spread over every warp of the grid
    const int64_t tail = min(n, total_warps);
    const int64_t g5 = (n - tail) / HPE_GROUP;
    const int64_t groups = g5 + (n - g5 * HPE_GROUP);
    int seen = -1;                                // chunks [0, seen] are known to have landed
    for (int64_t q = gw; q < groups; q += total_warps) {
        const int64_t p0 = q < g5 ? q * HPE_GROUP : g5 * HPE_GROUP + (q - g5);
        const int cnt = q < g5 ? HPE_GROUP : 1;
        int need = seen < 0 ? 0 : seen;
        while (ck.end[need] < p0 + cnt) ++need;   // chunk holding the last pose of the group
        if (need > seen) {
            if (lane == 0) {
                long long spins = 0;
                const volatile unsigned *f = arrived + need;
                while (*f == 0u) {
                    if (++spins > (1ll << 23)) { *timeout_flag = 1; break; }   // ~2 s: a failed / serialised upload must not hang the GPU
                    __nanosleep(200);
                }
                __threadfence_system();
            }
            __syncwarp();
            seen = need;                          // chunks are uploaded in order on one stream
        }
        group_fk<true>(sm.fs.M, [&](int g) { return theta + (p0 + g) * HPE_NDOF; }, cnt, ws, lane);
        for (int g = 0; g < cnt; ++g) {
            double depth = depth_and_stage(sm.fs, ws, g, lane);
            double align = match_align<0>(sm.fs, ws, g, lane, frozen, match_out ? match_out + (p0 + g) * n_pts : nullptr);
            double colli = 0.0;
            if (WITH_COLLI) colli = warp_sum(collision_lane(sm.fs.rad, ws.sphd + (size_t)g * HPE_NSPH * 3, lane));
            if (lane == 0) {
                terms[p0 + g] = align;
                terms[tstride + p0 + g] = depth;
                terms[2 * tstride + p0 + g] = colli;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// One generation of a large swarm for the particles [lo, lo+own) owned by this GPU, fused:
// velocity/position update with the reference's clamp (src/PSO.cpp:824-840, 358-377), evaluation
// (cal_cost) and the pbest row update (:855-860) -- the warp that evaluates a particle also moves it
// and writes its (pcost, pbest) row.  sw.update = 0 for generation 0 (x already generated).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(COST_WARPS * 32, HPE_MINBLOCKS)
swarm_step_kernel(DevModel m, ObsView o, int slot, SwarmStepArgs sw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CostSmem &sm = *reinterpret_cast<CostSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage_model(sm.fs, m);
    stage_frame(sm.fs, frame_view(o, slot));     // ends synchronised
    const WarpScratch ws = warp_scratch(sm.ws, warp, HPE_GROUP);
    constexpr int D = HPE_NDOF, ROW = HPE_NDOF + 1;
    const int64_t n = sw.own;
    const int64_t total_warps = (int64_t)gridDim.x * COST_WARPS;
    const int64_t gw = (int64_t)blockIdx.x * COST_WARPS + warp;
    const int64_t lo = n * gw / total_warps, hi = n * (gw + 1) / total_warps;
    const int iter = sw.ctl->iter;               // evaluations done so far = generation of this update
    const double *table = sw.table_in;   // rows of the previous generation (complete on every rank)

    for (int64_t p0 = lo; p0 < hi; p0 += HPE_GROUP) {
        int cnt = (int)min((int64_t)HPE_GROUP, hi - p0);
        if (sw.update) {
            for (int it = lane; it < cnt * (D / 2); it += 32) {
                int g = it / (D / 2), dp = it - g * (D / 2);
                int li = (int)(p0 + g), i = sw.lo + li, informant = sw.inf[i];
                const double *pb = table + (size_t)i * ROW + 1, *pbi = table + (size_t)informant * ROW + 1;
                double rp[2], rg[2];
                uniform2(sw.seed, (uint32_t)dp, (uint32_t)i, (uint32_t)iter, 1u, rp);
                uniform2(sw.seed, (uint32_t)dp, (uint32_t)i, (uint32_t)iter, 2u, rg);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int d = 2 * dp + e, at = li * D + d;
                    double xi = sw.x[at];
                    double nv = __dadd_rn(__dmul_rn(sw.v[at], sw.W1), __dmul_rn(__dmul_rn(rp[e], sw.C1), __dadd_rn(pb[d], -xi)));
                    if (informant != i) nv = __dadd_rn(nv, __dmul_rn(__dmul_rn(rg[e], sw.C1), __dadd_rn(pbi[d], -xi)));
                    double nx = __dadd_rn(xi, nv);
                    if (nx < sw.prm.lb[d] || nx > sw.prm.ub[d]) { nx = sw.prm.lb[d]; nv = 0.0; }   // :372
                    sw.x[at] = nx;
                    sw.v[at] = nv;
                }
            }
            __syncwarp();
        }
        group_fk(sm.fs.M, [&](int g) { return sw.x + (p0 + g) * D; }, cnt, ws, lane);
        for (int g = 0; g < cnt; ++g) {
            double depth = depth_and_stage(sm.fs, ws, g, lane);
            double align = match_align<0>(sm.fs, ws, g, lane, nullptr, nullptr);
            double fx = align + depth;
            // pbest row of this particle goes to the other table: informants must keep seeing the previous
            // generation's rows while other warps of this launch are still updating their particles
            const double *row = table + (size_t)(sw.lo + p0 + g) * ROW;
            double *row_out = sw.table_out + (size_t)(sw.lo + p0 + g) * ROW;
            bool take = iter == 0 || fx < row[0];
            if (lane < ROW) row_out[lane] = take ? (lane == 0 ? fx : sw.x[(p0 + g) * D + lane - 1]) : row[lane];
        }
    }
}

// ---------------------------------------------------------------------------------------
// The whole pso_evolve of one large sharded swarm in ONE cooperative launch per GPU (hpe_swarm_run):
// every generation's topology draw, informant search, move + evaluate + pbest, the exchange of the
// owned (pcost, pbest) rows with the other GPUs and the global argmin run inside this kernel.
//   * exchange: the warp that finishes a particle stores its 27-double row into the local table AND
//     straight into every peer's table over NVLink (CUDA IPC mappings).  After the step barrier CTA 0
//     posts this rank's (minimum cost, index) on every peer's board (and its own) and raises the
//     generation flag next to it; EVERY CTA then waits for the flags of all ranks, takes the first
//     minimum over the boards and updates its own copy of (gbest_cost, count) -- all CTAs compute the
//     same values, so no barrier is needed to publish them.  No NCCL call, no host involvement.
//   * informants: ONE pass over the num_p * 4 links: a 128-bit compare-and-swap minimum on
//     (cost bits, index) keeps the first minimum (src/PSO.cpp:810-812) in one go, and only links that
//     end in a particle this rank owns are looked at.  The keys are double buffered by generation
//     parity (a CTA may start the next generation's pass while another still resets the old keys).
//   * rows are double buffered exactly like the stepwise path (hpe_swarm.cu), and a rank enters
//     generation g+1 only after seeing every peer's flag of generation g, so stores into the buffer a
//     peer still reads cannot happen; boards are double buffered by generation parity.
//   * 2 grid barriers per generation (informants | step).
// Counters are keyed by the global particle index: the result is bit-identical to the stepwise
// all-gather path, to any other sharding and to pso_kernel.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long run_cost_bits(double c)
{
    return c == c ? (unsigned long long)__double_as_longlong(c) : 0xfff8000000000000ull;
}

// lexicographic minimum on (hi, lo) at a 16-byte aligned address: atom.cas.b128 loop
__device__ __forceinline__ void atomic_min_u128(unsigned long long *addr, unsigned long long hi, unsigned long long lo)
{
    ulonglong2 cur = __ldcg(reinterpret_cast<const ulonglong2 *>(addr));      // x = lo, y = hi
    while (hi < cur.y || (hi == cur.y && lo < cur.x)) {
        unsigned long long olo, ohi;
        asm volatile("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
                     "atom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\n\tmov.b128 {%0, %1}, o;\n\t}"
                     : "=l"(olo), "=l"(ohi) : "l"(cur.x), "l"(cur.y), "l"(lo), "l"(hi), "l"(addr) : "memory");
        if (olo == cur.x && ohi == cur.y) break;
        cur.x = olo; cur.y = ohi;
    }
}

template <bool WARP1>     // WARP1: at most one particle per warp (small shards): whole-warp forward kinematics
__global__ void __launch_bounds__(COST_WARPS * 32, HPE_MINBLOCKS)
swarm_run_kernel(DevModel m, ObsView o, int slot, SwarmRunArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CostSmem &sm = *reinterpret_cast<CostSmem *>(smem_raw);
    __shared__ unsigned long long skey[COST_WARPS];
    __shared__ int sidx[COST_WARPS];
    __shared__ double s_gbest_cost;             // this CTA's copy of the swarm's control state (identical in every CTA)
    __shared__ int s_count, s_err;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_model(sm.fs, m);
    stage_frame(sm.fs, frame_view(o, slot));     // ends synchronised
    const WarpScratch ws = warp_scratch(sm.ws, warp, HPE_GROUP);
    constexpr int D = HPE_NDOF, ROW = HPE_NDOF + 1, NK = 3;
    const int num_p = a.num_p, own = a.hi - a.lo;
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x, gtid = (int64_t)blockIdx.x * blockDim.x + tid;
    const int64_t total_warps = (int64_t)gridDim.x * COST_WARPS, gw = (int64_t)blockIdx.x * COST_WARPS + warp;
    const int64_t plo = (int64_t)own * gw / total_warps, phi = (int64_t)own * (gw + 1) / total_warps;
    SwarmCtl *ctl = a.ctl;
    ulonglong2 *key2 = reinterpret_cast<ulonglong2 *>(a.key2);          // [2][own]: (x = index, y = cost bits)
    const ulonglong2 KEY_NONE = make_ulonglong2(0x7fffffffull, ~0ull);
    // optional phase timestamps (hpe_swarm_debug_times): 8 slots per generation, written by one thread
#define RUN_STAMP(slot_) do { if (a.times && gtid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.times[g * 8 + (slot_)] = t_; } } while (0)

    // generate_particles for the owned rows (src/PSO.cpp:56-74); informant search state cleared
    for (int64_t it = gtid; it < (int64_t)own * (D / 2); it += gthreads) {
        int li = (int)(it / (D / 2)), dp = (int)(it - (int64_t)li * (D / 2));
        double z[2];
        normal2(a.seed, (uint32_t)dp, (uint32_t)(a.lo + li), 0u, z);
        a.x[li * D + 2 * dp] = __dadd_rn(a.x0[2 * dp], __dmul_rn(z[0], a.prm.std[2 * dp]));
        a.x[li * D + 2 * dp + 1] = __dadd_rn(a.x0[2 * dp + 1], __dmul_rn(z[1], a.prm.std[2 * dp + 1]));
        a.v[li * D + 2 * dp] = 0.0;
        a.v[li * D + 2 * dp + 1] = 0.0;
    }
    for (int64_t i = gtid; i < 2 * (int64_t)own; i += gthreads) key2[i] = KEY_NONE;
    if (gtid == 0) { ctl->gbest_cost = 1e100; ctl->count = 100; ctl->iter = 0; ctl->error = 0; }
    if (tid == 0) { s_gbest_cost = 1e100; s_count = 100; s_err = 0; }
    grid.sync();

    int cur = 0;
    for (int g = 0; g < a.maxiter; ++g) {
        const double *table = a.table[cur];                 // rows of the previous generation, complete
        double *out = g == 0 ? a.table[0] : a.table[cur ^ 1];
        const int ob = g == 0 ? 0 : cur ^ 1;
        const int iter = g;                                 // evaluations done so far
        ulonglong2 *keys = key2 + (size_t)(g & 1) * own;    // this generation's informant keys
        RUN_STAMP(0);
        if (g > 0) {
            // ---- topology (redrawn when count > 0, src/PSO.cpp:790-803) + informants (:810-812): link (s -> r) offers
            //      (pcost[s], s) to r; only targets this rank owns are kept ----
            const bool redraw = s_count > 0;
            for (int64_t it = gtid; it < (int64_t)num_p * (NK + 1); it += gthreads) {
                int s = (int)(it / (NK + 1)), k = (int)(it - (int64_t)s * (NK + 1));
                int r = s;                                  // L = eye + links
                if (k < NK) {
                    if (redraw) {
                        double u[2];
                        uniform2(a.seed, (uint32_t)(k >> 1), (uint32_t)s, (uint32_t)iter, 3u, u);
                        r = (int)floor(__dadd_rn(__dmul_rn(u[k & 1], (double)(num_p - 1)), 0.5));
                        a.R[s * NK + k] = r;
                    } else {
                        r = a.R[s * NK + k];
                    }
                }
                if (r >= a.lo && r < a.hi)
                    atomic_min_u128(reinterpret_cast<unsigned long long *>(keys + (r - a.lo)),
                                    run_cost_bits(__ldcg(table + (size_t)s * ROW)), (unsigned long long)s);
            }
            RUN_STAMP(1);
            grid.sync();
            RUN_STAMP(2);
        }
        // ---- move (g > 0), evaluate, pbest row -> local table and every peer's table ----
        unsigned long long wkey = ~0ull;                     // first minimum of pcost over this warp's particles
        int widx = 0x7fffffff;
        for (int64_t p0 = plo; p0 < phi; p0 += HPE_GROUP) {
            int cnt = (int)min((int64_t)HPE_GROUP, phi - p0);
            if (g > 0) {
                for (int it = lane; it < cnt * (D / 2); it += 32) {
                    int gq = it / (D / 2), dp = it - gq * (D / 2);
                    int li = (int)(p0 + gq), i = a.lo + li, informant = (int)__ldcg(&keys[li].x);
                    const double *pb = table + (size_t)i * ROW + 1, *pbi = table + (size_t)informant * ROW + 1;
                    double rp[2], rg[2];
                    uniform2(a.seed, (uint32_t)dp, (uint32_t)i, (uint32_t)iter, 1u, rp);
                    uniform2(a.seed, (uint32_t)dp, (uint32_t)i, (uint32_t)iter, 2u, rg);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        int d = 2 * dp + e, at = li * D + d;
                        double xi = a.x[at];
                        double nv = __dadd_rn(__dmul_rn(a.v[at], a.W1), __dmul_rn(__dmul_rn(rp[e], a.C1), __dadd_rn(__ldcg(pb + d), -xi)));
                        if (informant != i) nv = __dadd_rn(nv, __dmul_rn(__dmul_rn(rg[e], a.C1), __dadd_rn(__ldcg(pbi + d), -xi)));
                        double nx = __dadd_rn(xi, nv);
                        if (nx < a.prm.lb[d] || nx > a.prm.ub[d]) { nx = a.prm.lb[d]; nv = 0.0; }   // :372
                        a.x[at] = nx;
                        a.v[at] = nv;
                    }
                }
                __syncwarp();
            }
            group_fk<false, WARP1>(sm.fs.M, [&](int gq) { return a.x + (p0 + gq) * D; }, cnt, ws, lane);
            for (int gq = 0; gq < cnt; ++gq) {
                double depth = depth_and_stage(sm.fs, ws, gq, lane);
                double align = match_align<0>(sm.fs, ws, gq, lane, nullptr, nullptr);
                double fx = align + depth;
                const size_t roff = (size_t)(a.lo + p0 + gq) * ROW;
                const double *row = table + roff;
                const double old_cost = g == 0 ? 0.0 : __ldcg(row);
                bool take = g == 0 || fx < old_cost;
                {
                    unsigned long long kb = run_cost_bits(take ? fx : old_cost);
                    if (kb < wkey) { wkey = kb; widx = (int)(a.lo + p0 + gq); }      // ascending index: ties keep the first
                }
                if (lane < ROW) {
                    double val = take ? (lane == 0 ? fx : a.x[(p0 + gq) * D + lane - 1]) : __ldcg(row + lane);
                    out[roff + lane] = val;
                    for (int p = 0; p < a.world; ++p)
                        if (a.peer_table[ob][p]) a.peer_table[ob][p][roff + lane] = val;      // NVLink store
                }
            }
        }
        // release the rows this warp stored into the peers' tables at system scope (the flag that announces them is
        // written by CTA 0 after the grid barrier below; release / acquire order is cumulative across the barrier).
        // Only the warps that stored anything, and only when there are peers: a system-scope fence is slow (it was 30 % of
        // this kernel's stall samples when every thread executed the sequentially consistent form)
        if (a.world > 1 && plo < phi) asm volatile("fence.acq_rel.sys;" ::: "memory");
        // this CTA's minimum -> cta_key / cta_idx (the global argmin is then a reduction over CTAs and ranks)
        if (lane == 0) { skey[warp] = wkey; sidx[warp] = widx; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long kk = skey[0];
            int ii = sidx[0];
            for (int w = 1; w < COST_WARPS; ++w)
                if (skey[w] < kk || (skey[w] == kk && sidx[w] < ii)) { kk = skey[w]; ii = sidx[w]; }
            a.cta_key[blockIdx.x] = kk;
            a.cta_idx[blockIdx.x] = ii;
        }
        RUN_STAMP(3);
        grid.sync();
        RUN_STAMP(4);
        // the keys of this generation have been consumed; its buffer is next used two generations from now
        for (int64_t i = gtid; i < own; i += gthreads) keys[i] = KEY_NONE;
        const int par = g & 1;
        const unsigned want = a.gen_base + (unsigned)g + 1u;
        // ---- CTA 0: this rank's first minimum -> every peer's board, then the flag; wait for the peers' flags; the
        //      global first minimum (src/PSO.cpp:755-760, 864-876) is published to the other CTAs of this GPU ----
        unsigned long long gkey = ~0ull;
        int gidx = 0x7fffffff;
        bool lost = false;
        if (blockIdx.x == 0) {
            unsigned long long key = ~0ull;
            int idx = 0x7fffffff;
            for (int i = tid; i < (int)gridDim.x; i += COST_WARPS * 32) {
                unsigned long long kk = __ldcg(a.cta_key + i);
                int ii = __ldcg(a.cta_idx + i);
                if (kk < key || (kk == key && ii < idx)) { key = kk; idx = ii; }
            }
            for (int off = 16; off > 0; off >>= 1) {
                unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, off);
                int oi = __shfl_xor_sync(0xffffffffu, idx, off);
                if (ok < key || (ok == key && oi < idx)) { key = ok; idx = oi; }
            }
            __syncthreads();                                  // skey / sidx of the step phase have been consumed
            if (lane == 0) { skey[warp] = key; sidx[warp] = idx; }
            __syncthreads();
            if (warp == 0) {
                key = lane < COST_WARPS ? skey[lane] : ~0ull;
                idx = lane < COST_WARPS ? sidx[lane] : 0x7fffffff;
                for (int off = 16; off > 0; off >>= 1) {
                    unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, off);
                    int oi = __shfl_xor_sync(0xffffffffu, idx, off);
                    if (ok < key || (ok == key && oi < idx)) { key = ok; idx = oi; }
                }
                if (lane < a.world && lane != a.rank) {
                    SwarmMin *dst = a.peer_min[lane] + par * HPE_MAX_PEERS + a.rank;
                    dst->key = key;
                    dst->idx = idx;
                    __threadfence_system();
                    a.peer_flags[lane][a.rank] = want;
                    long long spins = 0;
                    while (a.my_flags[lane] < want) {
                        if (++spins > a.spin_limit) { lost = true; break; }   // a lost peer must not hang the GPU
                        __nanosleep(32);
                    }
                    __threadfence_system();
                    const volatile SwarmMin *src = a.my_min + par * HPE_MAX_PEERS + lane;
                    key = src->key;
                    idx = src->idx;
                } else if (lane >= a.world) {
                    key = ~0ull;
                    idx = 0x7fffffff;
                }
                lost = __any_sync(0xffffffffu, lost);
                RUN_STAMP(5);
                for (int off = 16; off > 0; off >>= 1) {
                    unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, off);
                    int oi = __shfl_xor_sync(0xffffffffu, idx, off);
                    if (ok < key || (ok == key && oi < idx)) { key = ok; idx = oi; }
                }
                if (idx == 0x7fffffff) idx = 0;
                gkey = key; gidx = idx;
                if (lane == 0) {                              // release the other CTAs of this GPU
                    volatile SwarmResult *res = a.result + par;
                    res->idx = idx;
                    res->lost = lost ? 1 : 0;
                    __threadfence();
                    res->gen = want;
                }
            }
        } else if (warp == 0) {
            volatile SwarmResult *res = a.result + par;
            if (lane == 0) {
                long long spins = 0;
                while (res->gen != want) {
                    if (++spins > 4 * a.spin_limit) { lost = true; break; }
                    __nanosleep(64);
                }
                __threadfence();
                gidx = res->idx;
                lost = lost || res->lost != 0;
            }
            gidx = __shfl_sync(0xffffffffu, gidx, 0);
            lost = __any_sync(0xffffffffu, lost);
        }
        (void)gkey;
        if (warp == 0) {
            const double fmin = __ldcg(out + (size_t)gidx * ROW);   // the winner's row (remote rows arrived by NVLink stores)
            const bool improved = fmin < s_gbest_cost;
            if (blockIdx.x == 0) {
                if (improved && lane < D) ctl->gbest[lane] = __ldcg(out + (size_t)gidx * ROW + 1 + lane);
                if (lane == 0) {
                    if (improved) ctl->gbest_cost = fmin;
                    ctl->count = improved ? (g == 0 ? 100 : 0) : (g > 0 ? s_count + 1 : s_count);
                    ctl->iter = g + 1;
                    if (lost) ctl->error = 1;
                    if (a.trace && g >= 1) a.trace[g - 1] = improved ? fmin : s_gbest_cost;
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (improved) { s_gbest_cost = fmin; s_count = g == 0 ? 100 : 0; }
                else if (g > 0) s_count += 1;
                if (lost) s_err = 1;
            }
        }
        __syncthreads();
        RUN_STAMP(6);
        RUN_STAMP(7);
        if (g > 0) cur ^= 1;
    }
}

// ---------------------------------------------------------------------------------------
// Synthetic frames: ray-cast the 48 spheres at pose theta[f] into frame f.  One CTA per
// frame; FP64 so that the frame equals the oracle's renderer up to FK rounding.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) render_kernel(DevModel m, const double *__restrict__ theta, double focal,
                                                      float *__restrict__ depth_mm)
{
    __shared__ DevModel Ms;
    __shared__ double S[HPE_NSPH][4];   // x,y,z (un-negated), cc = |c|^2 - r^2
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 8); i += blockDim.x)
        reinterpret_cast<double *>(&Ms)[i] = reinterpret_cast<const double *>(&m)[i];
    __syncthreads();
    const double *th = theta + (size_t)blockIdx.x * HPE_NDOF;
    if (threadIdx.x < 5) {
        double J[5][3], C[10][3];
        int d = threadIdx.x;
        fk_digit(Ms, d, th, J);
        int cnt = digit_spheres(d, J, C);
        int first = digit_first_sphere(d);
        for (int i = 0; i < cnt; ++i) {
            double r = Ms.rad[first + i];
            S[first + i][0] = C[i][0]; S[first + i][1] = C[i][1]; S[first + i][2] = C[i][2];
            S[first + i][3] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(C[i][0], C[i][0]), __dmul_rn(C[i][1], C[i][1])),
                                                  __dmul_rn(C[i][2], C[i][2])), -__dmul_rn(r, r));
        }
    }
    __syncthreads();
    float *out = depth_mm + (size_t)blockIdx.x * HPE_W * HPE_H;
    for (int pix = threadIdx.x; pix < HPE_W * HPE_H; pix += blockDim.x) {
        int v = pix / HPE_W, u = pix - v * HPE_W;
        double dx = __ddiv_rn((double)u - HPE_W / 2., focal), dy = __ddiv_rn((double)v - HPE_H / 2., focal);
        double a = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 1.0);
        double best = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll 4
        for (int j = 0; j < HPE_NSPH; ++j) {
            double b = __dadd_rn(__dadd_rn(__dmul_rn(dx, S[j][0]), __dmul_rn(dy, S[j][1])), S[j][2]);
            double disc = __dadd_rn(__dmul_rn(b, b), -__dmul_rn(a, S[j][3]));
            if (disc >= 0) {
                double t = __ddiv_rn(__dadd_rn(b, -sqrt(disc)), a);
                if (t > 0 && t < best) best = t;
            }
        }
        out[pix] = best < 1e300 ? (float)__dmul_rn(best, 10.0) : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------
// costfunc::gnd_truth_err (src/costfunc.cpp:476-507): joints x 10 (cm -> mm), y and z negated,
// distance to the ground-truth joints of joint.txt for the wrist and the five tips
// (rows 0, 4, 8, 12, 16, 20), summed the way Armadillo sums a vector (two partial sums).
// One thread per pose; joints [n][21][3] as fk_kernel wrote them, gt [n][63].
// ---------------------------------------------------------------------------------------
__global__ void gt_err_kernel(const double *__restrict__ joints, const double *__restrict__ gt, int64_t n,
                              double *__restrict__ err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *J = joints + i * HPE_NJOINT * 3, *G = gt + i * 63;
    double d[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const int j = 4 * q;
        double hx = __dmul_rn(J[j * 3], 10.0), hy = -__dmul_rn(J[j * 3 + 1], 10.0), hz = -__dmul_rn(J[j * 3 + 2], 10.0);
        double dx = G[j * 3] - hx, dy = G[j * 3 + 1] - hy, dz = G[j * 3 + 2] - hz;
        d[q] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    }
    err[i] = __dadd_rn(__dadd_rn(__dadd_rn(d[0], d[2]), d[4]), __dadd_rn(__dadd_rn(d[1], d[3]), d[5]));
}

// ---------------------------------------------------------------------------------------
// diagnostics: latency of ONE evaluation by ONE warp on an otherwise idle SM, per phase (clock64)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) eval_latency_kernel(DevModel m, ObsView o, int slot, const double *theta, int reps,
                                                          long long *cycles, double *cost)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CostSmem &sm = *reinterpret_cast<CostSmem *>(smem_raw);
    const int lane = threadIdx.x & 31;
    stage_model(sm.fs, m);
    stage_frame(sm.fs, frame_view(o, slot));
    const WarpScratch ws = warp_scratch(sm.ws, 0, HPE_GROUP);
    long long t_fk = 0, t_depth = 0, t_match = 0;
    double acc = 0.0;
    for (int r = 0; r < reps; ++r) {
        long long t0 = clock64();
        group_fk<false, true>(sm.fs.M, [&](int) { return theta; }, 1, ws, lane);
        long long t1 = clock64();
        double depth = depth_and_stage(sm.fs, ws, 0, lane);
        long long t2 = clock64();
        double align = match_align<0>(sm.fs, ws, 0, lane, nullptr, nullptr);
        long long t3 = clock64();
        acc += align + depth;
        t_fk += t1 - t0; t_depth += t2 - t1; t_match += t3 - t2;
    }
    if (lane == 0) { cycles[0] = t_fk / reps; cycles[1] = t_depth / reps; cycles[2] = t_match / reps; *cost = acc / reps; }
}

}  // namespace

HPE_DEFINE_FALLBACK_GETTER(hpe_fallbacks_cost)

cudaError_t hpe_launch_fk(const DevModel &m, const double *theta_dev, int64_t n, double *spheres_dev,
                          double *joints_dev, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    int64_t items = n * 5;
    int blocks = (int)((items + 159) / 160);
    fk_kernel<<<blocks, 160, 0, st>>>(m, theta_dev, n, spheres_dev, joints_dev);
    return cudaGetLastError();
}

cudaError_t hpe_launch_cost_chase(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int64_t n,
                                  bool with_colli, double *terms_dev, int64_t terms_stride, const int32_t *frozen_dev,
                                  int32_t *match_out_dev, const unsigned *arrived_dev, const ChaseChunks &chunks, int *timeout_dev,
                                  cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int sms = hpe_sm_count();
    static PerDevice<bool> attr_pd;
    if (bool &attr = attr_pd.get(); !attr) {
        cudaFuncSetAttribute(cost_chase_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem));
        cudaFuncSetAttribute(cost_chase_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem));
        attr = true;
    }
    int64_t groups = (n + HPE_GROUP - 1) / HPE_GROUP;
    int64_t want = (groups + COST_WARPS - 1) / COST_WARPS;
    int blocks = (int)(want < (long long)HPE_MINBLOCKS * sms ? want : (long long)HPE_MINBLOCKS * sms);
    if (with_colli)
        cost_chase_kernel<true><<<blocks, COST_WARPS * 32, sizeof(CostSmem), st>>>(m, o, slot, theta_dev, n, terms_dev, terms_stride,
                                                                                    frozen_dev, match_out_dev, arrived_dev, chunks, timeout_dev);
    else
        cost_chase_kernel<false><<<blocks, COST_WARPS * 32, sizeof(CostSmem), st>>>(m, o, slot, theta_dev, n, terms_dev, terms_stride,
                                                                                     frozen_dev, match_out_dev, arrived_dev, chunks, timeout_dev);
    return cudaGetLastError();
}

cudaError_t hpe_launch_eval_latency(const DevModel &m, const ObsView &o, int slot, const double *theta_dev, int reps,
                                    long long *cycles_dev, double *cost_dev, cudaStream_t st)
{
    cudaFuncSetAttribute(eval_latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem));
    eval_latency_kernel<<<1, 32, sizeof(CostSmem), st>>>(m, o, slot, theta_dev, reps, cycles_dev, cost_dev);
    return cudaGetLastError();
}

cudaError_t hpe_launch_gt_err(const double *joints_dev, const double *gt_dev, int64_t n, double *err_dev, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    gt_err_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(joints_dev, gt_dev, n, err_dev);
    return cudaGetLastError();
}

// Self-scheduling plan of cost_kernel for a batch of n poses on a device with `sms` SMs: groups of SCORE_GROUP from [0, n1),
// units of HPE_TAIL_C2 from [n1, n2) (t2 of them), single poses from [n2, n); batches with fewer than two groups per warp
// are handed out in groups only.  A batch with fewer groups than the device has CTAs is spread over as many SMs as it has
// groups (the warps of a CTA that find the counters empty leave at once).
struct CostPlan { int blocks, n1, n2, t2; };
static CostPlan cost_plan(int64_t n, int sms)
{
    CostPlan pl;
    const int64_t groups = (n + SCORE_GROUP - 1) / SCORE_GROUP;
    pl.blocks = (int)(groups < (long long)SCORE_CTAS * sms ? groups : (long long)SCORE_CTAS * sms);
    const int64_t warps = (int64_t)pl.blocks * SCORE_WARPS;
    int64_t r1 = n >= 2 * SCORE_GROUP * warps ? (int64_t)(HPE_TAIL_R1 * (float)warps) : 0, r2 = (int64_t)(HPE_TAIL_R2 * (float)warps);
    if (r2 > r1) r2 = r1;
    pl.n1 = (int)(n - r1);
    pl.t2 = (int)((r1 - r2 + HPE_TAIL_C2 - 1) / HPE_TAIL_C2);
    pl.n2 = (int)(n - r2);
    return pl;
}

extern "C" int hpe_debug_cost_plan(int64_t n, int sms, int plan[6])
{
    if (n < 0 || n > 0x40000000 || sms < 1 || !plan) return HPE_ERR_ARG;
    const CostPlan pl = cost_plan(n, sms);
    plan[0] = pl.blocks * SCORE_WARPS; plan[1] = SCORE_GROUP; plan[2] = HPE_TAIL_C2;
    plan[3] = pl.n1; plan[4] = pl.n2; plan[5] = pl.t2;
    return HPE_OK;
}

cudaError_t hpe_launch_cost(const DevModel &m, const ObsView &o, int slot, const double *theta_dev,
                            int64_t n, bool with_colli, double *cost_dev, double *terms_dev,
                            const int32_t *frozen_match_dev, int32_t *match_out_dev, cudaStream_t st,
                            int64_t terms_stride)
{
    if (n <= 0) return cudaSuccess;
    if (terms_stride <= 0) terms_stride = n;
    if (n > 0x40000000) {          // the kernel indexes poses with 32 bits: split (2^30 poses are 223 GB, more than a B200 holds)
        const int64_t half = n / 2;
        cudaError_t e = hpe_launch_cost(m, o, slot, theta_dev, half, with_colli, cost_dev, terms_dev, frozen_match_dev,
                                        match_out_dev, st, terms_stride);
        if (e != cudaSuccess) return e;
        return hpe_launch_cost(m, o, slot, theta_dev + half * HPE_NDOF, n - half, with_colli, cost_dev ? cost_dev + half : nullptr,
                               terms_dev ? terms_dev + half : nullptr, frozen_match_dev,
                               match_out_dev ? match_out_dev + half * o.n_pts : nullptr, st, terms_stride);
    }
    const int sms = hpe_sm_count();
    const CostPlan plan = cost_plan(n, sms);
    const int blocks = plan.blocks;
    static PerDevice<bool> attr_pd;
    bool &attr = attr_pd.get();
    if (!attr) {
        const int smem = (int)sizeof(ScoreSmem);
        cudaFuncSetAttribute(cost_kernel<false, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(cost_kernel<true, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(cost_kernel<false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(cost_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(cost_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(cost_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr = true;
    }
    const int n1 = plan.n1, n2 = plan.n2, t2 = plan.t2;
    static std::atomic<unsigned> next_ticket{0};
    unsigned ticket = next_ticket.fetch_add(1, std::memory_order_relaxed) + 1u;
    if (ticket == 0u) ticket = next_ticket.fetch_add(1, std::memory_order_relaxed) + 1u;      // 0 means "free"
    const int sched_slot = (int)(ticket % HPE_SCHED_SLOTS);
    auto launch = [&](auto kern) {
        kern<<<blocks, SCORE_WARPS * 32, sizeof(ScoreSmem), st>>>(m, o, slot, theta_dev, n, cost_dev, terms_dev,
                                                                 terms_stride, frozen_match_dev, match_out_dev, sched_slot, ticket, n1, n2, t2);
    };
    const bool aux = frozen_match_dev || match_out_dev;
    if (g_hpe_matcher == 1) launch(with_colli ? cost_kernel<true, 1, true> : cost_kernel<false, 1, true>);
    else if (aux) launch(with_colli ? cost_kernel<true, 0, true> : cost_kernel<false, 0, true>);
    else launch(with_colli ? cost_kernel<true, 0, false> : cost_kernel<false, 0, false>);
    return cudaGetLastError();
}

cudaError_t hpe_launch_swarm_step(const DevModel &m, const ObsView &o, int slot, const SwarmStepArgs &a, cudaStream_t st)
{
    if (a.own <= 0) return cudaSuccess;
    const int sms = hpe_sm_count();
    static PerDevice<bool> attr_pd;
    bool &attr = attr_pd.get();
    if (!attr) {
        cudaFuncSetAttribute(swarm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem));
        attr = true;
    }
    int64_t groups = ((int64_t)a.own + HPE_GROUP - 1) / HPE_GROUP;
    int64_t want = (groups + COST_WARPS - 1) / COST_WARPS;
    int blocks = (int)(want < (long long)HPE_MINBLOCKS * sms ? want : (long long)HPE_MINBLOCKS * sms);
    swarm_step_kernel<<<blocks, COST_WARPS * 32, sizeof(CostSmem), st>>>(m, o, slot, a);
    return cudaGetLastError();
}

cudaError_t hpe_launch_swarm_run(const DevModel &m, const ObsView &o, int slot, const SwarmRunArgs &a, cudaStream_t st)
{
    const int sms = hpe_sm_count();
    static PerDevice<int> per_sm_pd;
    int &per_sm = per_sm_pd.get();
    if (!per_sm) {
        cudaFuncSetAttribute(swarm_run_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem));
        cudaFuncSetAttribute(swarm_run_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CostSmem));
        int p0 = 0, p1 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p0, swarm_run_kernel<false>, COST_WARPS * 32, sizeof(CostSmem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p1, swarm_run_kernel<true>, COST_WARPS * 32, sizeof(CostSmem));
        per_sm = p0 < p1 ? p0 : p1;
        if (per_sm > HPE_MINBLOCKS) per_sm = HPE_MINBLOCKS;
        if (per_sm < 1) per_sm = 1;
    }
    // all CTAs must be co-resident (grid barriers): one full wave, fewer when the shard is small
    int64_t own = a.hi - a.lo;
    int64_t want = (own + COST_WARPS - 1) / COST_WARPS;                 // one particle per warp is enough
    int blocks = (int)(want < (int64_t)per_sm * sms ? want : (int64_t)per_sm * sms);
    if (blocks < sms) blocks = sms;                                     // the scatter / reset phases want the whole chip
    const bool warp1 = own <= (int64_t)blocks * COST_WARPS;             // no warp gets more than one particle
    DevModel mm = m;
    ObsView oo = o;
    SwarmRunArgs aa = a;
    void *args[] = {&mm, &oo, &slot, &aa};
    return cudaLaunchCooperativeKernel(warp1 ? (void *)swarm_run_kernel<true> : (void *)swarm_run_kernel<false>, dim3(blocks),
                                       dim3(COST_WARPS * 32), args, sizeof(CostSmem), st);
}

cudaError_t hpe_launch_render(const DevModel &m, const double *theta_dev, int64_t n, double focal,
                              float *depth_dev, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    render_kernel<<<(unsigned)n, 256, 0, st>>>(m, theta_dev, focal, depth_dev);
    return cudaGetLastError();
}
