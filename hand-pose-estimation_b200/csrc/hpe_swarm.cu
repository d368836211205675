// hpe_swarm.cu -- PSO::pso_evolve (src/PSO.cpp:717-886) for ONE large swarm whose particles are
// block-partitioned over GPUs (BASELINE config "multi-restart swarm").  A rank owns particles
// [lo,hi); per generation it needs every particle's (pcost, pbest) row because informants are
// drawn over the whole swarm (src/PSO.cpp:790-812), so the caller all-gathers the 27-double rows
// between hpe_swarm_step / hpe_swarm_reduce (NCCL over NVLink in bench.py, gloo in the CPU tests of
// the host logic).  All arithmetic is keyed by the GLOBAL particle index (Philox counters), so the
// sharded run is bit-identical to the single-GPU run and to the one-CTA kernel of hpe_pso.cu.
//
// Launches per generation: links (topology redraw + key init), two scatter-min passes for the
// informants, ONE fused kernel that moves, evaluates and writes the pbest rows of the owned particles
// (swarm_step_kernel, hpe_cost.cu), and after the exchange one single-CTA argmin + gbest/count update.
// The rows are double buffered: a generation reads table[cur] and writes table[cur ^ 1], because an
// informant's row must stay that of the previous generation while other warps of the same launch
// already finish theirs.  Nothing synchronises with the host inside a generation.
#include <cmath>
#include <cstring>
#include <new>

#include "hpe_internal.h"

namespace {

constexpr int NK = 3;                 // informants drawn per particle, src/PSO.cpp:769
constexpr int ROW = 1 + HPE_NDOF;     // table row: pcost, pbest[26]

__device__ __forceinline__ unsigned long long cost_bits(double c)
{
    // costs are >= 0 (or NaN/inf): the IEEE bit pattern of a non-negative double orders like the value;
    // NaN maps above +inf so that it never wins a minimum
    return c == c ? (unsigned long long)__double_as_longlong(c) : 0xfff8000000000000ull;
}

// generate_particles for the owned rows (src/PSO.cpp:56-74): x0 + randn % std, v = 0
__global__ void swarm_generate(PsoDev prm, const double *x0, uint64_t seed, int lo, int hi, double *x, double *v)
{
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (hi - lo) * (HPE_NDOF / 2);
    if (it >= n) return;
    int li = it / (HPE_NDOF / 2), dp = it - li * (HPE_NDOF / 2);
    double z[2];
    normal2(seed, (uint32_t)dp, (uint32_t)(lo + li), 0u, z);
    x[li * HPE_NDOF + 2 * dp] = __dadd_rn(x0[2 * dp], __dmul_rn(z[0], prm.std[2 * dp]));
    x[li * HPE_NDOF + 2 * dp + 1] = __dadd_rn(x0[2 * dp + 1], __dmul_rn(z[1], prm.std[2 * dp + 1]));
    v[li * HPE_NDOF + 2 * dp] = 0.0;
    v[li * HPE_NDOF + 2 * dp + 1] = 0.0;
}

// global first minimum of pcost + gbest/count bookkeeping (src/PSO.cpp:755-760, 864-876), one CTA
__global__ void __launch_bounds__(1024) swarm_reduce_kernel(int num_p, const double *table, SwarmCtl *ctl, double *trace)
{
    __shared__ unsigned long long skey[32];
    __shared__ int sidx[32];
    __shared__ int improved, winner;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long key = ~0ull;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < num_p; i += blockDim.x) {
        unsigned long long k = cost_bits(table[(size_t)i * ROW]);
        if (k < key) { key = k; idx = i; }            // ascending i: the lowest index of a tie stays
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ok < key || (ok == key && oi < idx)) { key = ok; idx = oi; }
    }
    if (lane == 0) { skey[warp] = key; sidx[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        key = skey[lane]; idx = sidx[lane];
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ok < key || (ok == key && oi < idx)) { key = ok; idx = oi; }
        }
        if (lane == 0) {
            if (idx == 0x7fffffff) idx = 0;
            double fmin = table[(size_t)idx * ROW];
            improved = fmin < ctl->gbest_cost;
            winner = idx;
            if (improved) {
                ctl->gbest_cost = fmin;
                ctl->count = ctl->iter == 0 ? 100 : 0;
            } else if (ctl->iter > 0) {
                ctl->count += 1;
            }
            ctl->iter += 1;
            if (trace && ctl->iter >= 2) trace[ctl->iter - 2] = ctl->gbest_cost;
        }
    }
    __syncthreads();
    // gbest_pos = particles.col(fmin_id): that particle has just improved, so its position is its pbest
    if (improved && threadIdx.x < HPE_NDOF) ctl->gbest[threadIdx.x] = table[(size_t)winner * ROW + 1 + threadIdx.x];
}

// topology (src/PSO.cpp:790-803): R[s][k] = floor(U (num_p-1) + 0.5), redrawn only when count > 0;
// and the start values of the informant search: key[i] = own cost (L = eye), inf[i] = unset
__global__ void swarm_links(int num_p, uint64_t seed, const SwarmCtl *ctl, const double *table, int *R,
                            unsigned long long *key, int *inf)
{
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it < num_p) {
        key[it] = cost_bits(table[(size_t)it * ROW]);
        inf[it] = 0x7fffffff;
    }
    if (ctl->count > 0 && it < num_p * NK) {
        int s = it / NK, k = it - s * NK;
        double u[2];
        uniform2(seed, (uint32_t)(k >> 1), (uint32_t)s, (uint32_t)ctl->iter, 3u, u);
        R[it] = (int)floor(__dadd_rn(__dmul_rn(u[k & 1], (double)(num_p - 1)), 0.5));
    }
}

// informant(i) = first minimum of pcost over {s : L(s,i) = 1} (src/PSO.cpp:810-812), as two scatter
// passes over the num_p*(NK+1) links: minimum cost, then lowest index attaining it
__global__ void swarm_inf_cost(int num_p, const double *table, const int *R, unsigned long long *key)
{
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= num_p * NK) return;
    atomicMin(&key[R[it]], cost_bits(table[(size_t)(it / NK) * ROW]));
}

__global__ void swarm_inf_index(int num_p, const double *table, const int *R, const unsigned long long *key, int *inf)
{
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= num_p * (NK + 1)) return;
    int s = it / (NK + 1), k = it - s * (NK + 1);
    int r = k == NK ? s : R[s * NK + k];
    if (cost_bits(table[(size_t)s * ROW]) == key[r]) atomicMin(&inf[r], s);
}

// ---- device-initiated exchange over NVLink peer memory (instead of the caller's all-gather) ----
// Every rank maps the other ranks' tables (CUDA IPC).  After its step kernel a rank stores its owned rows
// straight into every peer's table and then raises its flag in every peer's flag array; the peers' reduce
// kernel waits on those flags.  Generation g reads buffer g%2 and writes (g+1)%2, and a rank starts
// generation g+1 only after it has seen every peer's flag for generation g, i.e. after every peer has
// finished reading buffer g%2 -- so the stores of generation g+1 into that buffer cannot race.
constexpr int MAX_PEERS = HPE_MAX_PEERS;
struct PeerPtrs {
    double *table[MAX_PEERS];          // peer table being written this generation (NULL for self)
    volatile unsigned *flags[MAX_PEERS];
};

__global__ void swarm_push_rows(PeerPtrs peers, int world, const double *table, int lo, int hi)
{
    const size_t n = (size_t)(hi - lo) * ROW, off = (size_t)lo * ROW;
    for (int p = 0; p < world; ++p) {
        double *dst = peers.table[p];
        if (!dst) continue;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
            dst[off + i] = table[off + i];
    }
}

// runs after swarm_push_rows in stream order: publish "my rows of generation `gen` are in your table"
__global__ void swarm_signal(PeerPtrs peers, int world, int rank, unsigned gen)
{
    __threadfence_system();
    if (threadIdx.x < world && peers.flags[threadIdx.x]) peers.flags[threadIdx.x][rank] = gen;
    __threadfence_system();
}

// first kernel after the exchange point: wait until every peer has delivered generation `gen`
__global__ void swarm_wait(volatile unsigned *flags, int world, int rank, unsigned gen, SwarmCtl *ctl, long long spin_limit)
{
    if (threadIdx.x < world && threadIdx.x != rank) {
        long long spins = 0;
        while (flags[threadIdx.x] < gen) {
            if (++spins > spin_limit) { ctl->error = 1; break; }   // bounded: a lost peer must not hang the GPU
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

long long g_swarm_spin_limit = 1ll << 28;      // hpe_swarm_set_wait_limit(): polls before a flag wait gives up

inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

}  // namespace

struct hpe_swarm {
    const hpe_model *m;
    const hpe_obs *o;
    int slot, num_p, lo, hi, cur;
    PsoDev prm;
    uint64_t seed;
    double W1, C1;
    double *x0, *x, *v, *table[2], *trace;
    int *R, *inf;
    unsigned long long *key;
    unsigned long long *key2;          // informant keys of hpe_swarm_run: [2][own] x 16 bytes
    SwarmResult *result;               // [2] release slots of hpe_swarm_run
    SwarmCtl *ctl;
    // peer exchange
    void *shared_base;                 // one allocation: table[0] | table[1] | flags
    unsigned *flags;
    int world, rank;
    unsigned gen;                      // generations pushed so far
    void *peer_base[MAX_PEERS];
    unsigned long long *cta_key;       // per-CTA minima of hpe_swarm_run
    int *cta_idx;
    unsigned long long *times;         // debug: phase timestamps of hpe_swarm_run (NULL until requested)
};

static size_t table_bytes(int num_p) { return ((sizeof(double) * (size_t)num_p * ROW + 255) / 256) * 256; }
// one allocation exposed over CUDA IPC: table[0] | table[1] | flags[MAX_PEERS] (padded to 256 B) | min boards [2][MAX_PEERS]
static size_t shared_bytes(int num_p) { return 2 * table_bytes(num_p) + 256 + sizeof(SwarmMin) * 2 * MAX_PEERS; }

// after the owned rows of table[cur] are complete: hand them to the peers and raise the flags
static cudaError_t push_to_peers(hpe_swarm *s, cudaStream_t st)
{
    if (s->world <= 1) return cudaSuccess;
    PeerPtrs pp;
    memset(&pp, 0, sizeof pp);
    for (int p = 0; p < s->world; ++p) {
        if (p == s->rank) continue;
        char *base = (char *)s->peer_base[p];
        pp.table[p] = (double *)(base + (size_t)s->cur * table_bytes(s->num_p));
        pp.flags[p] = (volatile unsigned *)(base + 2 * table_bytes(s->num_p));
    }
    s->gen += 1;
    if (s->hi > s->lo) swarm_push_rows<<<64, 256, 0, st>>>(pp, s->world, s->table[s->cur], s->lo, s->hi);
    swarm_signal<<<1, 32, 0, st>>>(pp, s->world, s->rank, s->gen);
    return cudaGetLastError();
}

static SwarmStepArgs step_args(const hpe_swarm *s, int update, const double *in, double *out)
{
    SwarmStepArgs a;
    a.prm = s->prm; a.seed = s->seed; a.table_in = in; a.table_out = out; a.inf = s->inf; a.ctl = s->ctl;
    a.x = s->x; a.v = s->v; a.W1 = s->W1; a.C1 = s->C1; a.lo = s->lo; a.own = s->hi - s->lo; a.update = update;
    return a;
}

extern "C" {

int hpe_swarm_destroy(hpe_swarm *s)
{
    if (!s) return HPE_OK;
    for (int p = 0; p < MAX_PEERS; ++p)
        if (s->peer_base[p]) cudaIpcCloseMemHandle(s->peer_base[p]);
    cudaFree(s->x0); cudaFree(s->x); cudaFree(s->v); cudaFree(s->shared_base); cudaFree(s->trace);
    cudaFree(s->R); cudaFree(s->inf); cudaFree(s->key); cudaFree(s->key2); cudaFree(s->result); cudaFree(s->ctl); cudaFree(s->times); cudaFree(s->cta_key); cudaFree(s->cta_idx);
    delete s;
    return HPE_OK;
}

int hpe_swarm_create(const hpe_model *m, const hpe_obs *o, int slot, const hpe_pso_params *p, int num_particles,
                     int lo, int hi, const double x0[HPE_NDOF], uint64_t seed, hpe_swarm **out)
{
    if (!m || !o || !p || !x0 || !out || num_particles < 1 || lo < 0 || hi > num_particles || lo > hi || p->maxiter < 1)
        return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_create: bad argument");
    if (slot < 0 || slot >= o->v.capacity) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_create: slot out of range");
    hpe_swarm *s = new (std::nothrow) hpe_swarm();
    if (!s) return hpe_set_error(HPE_ERR_ARG, "out of host memory");
    s->m = m; s->o = o; s->slot = slot; s->num_p = num_particles; s->lo = lo; s->hi = hi; s->seed = seed; s->cur = 0;
    memcpy(s->prm.lb, p->lb, sizeof s->prm.lb);
    memcpy(s->prm.ub, p->ub, sizeof s->prm.ub);
    memcpy(s->prm.std, p->std, sizeof s->prm.std);
    s->prm.maxiter = p->maxiter;
    s->prm.num_p = num_particles;
    s->W1 = 1. / (2 * log(2.0));      // src/PSO.cpp:772
    s->C1 = 0.5 + log(2.0);           // src/PSO.cpp:773-774
    const size_t own = (size_t)(hi - lo > 0 ? hi - lo : 1), np = (size_t)num_particles;
    const int tl = p->maxiter > 1 ? p->maxiter - 1 : 1;
    cudaError_t e = cudaMalloc(&s->x0, sizeof(double) * HPE_NDOF);
    if (e == cudaSuccess) e = cudaMalloc(&s->x, sizeof(double) * own * HPE_NDOF);
    if (e == cudaSuccess) e = cudaMalloc(&s->v, sizeof(double) * own * HPE_NDOF);
    // both row buffers and the peer flags in one allocation, so that one IPC handle exposes them
    const size_t tb = table_bytes(num_particles);
    if (e == cudaSuccess) e = cudaMalloc(&s->shared_base, shared_bytes(num_particles));
    if (e == cudaSuccess) e = cudaMemset(s->shared_base, 0, shared_bytes(num_particles));
    if (e == cudaSuccess) e = cudaMalloc(&s->cta_key, sizeof(unsigned long long) * 2048);
    if (e == cudaSuccess) e = cudaMalloc(&s->cta_idx, sizeof(int) * 2048);
    if (e == cudaSuccess) {
        s->table[0] = (double *)s->shared_base;
        s->table[1] = (double *)((char *)s->shared_base + tb);
        s->flags = (unsigned *)((char *)s->shared_base + 2 * tb);
    }
    s->world = 1; s->rank = 0; s->gen = 0;
    if (e == cudaSuccess) e = cudaMalloc(&s->trace, sizeof(double) * tl);
    if (e == cudaSuccess) e = cudaMalloc(&s->R, sizeof(int) * np * NK);
    if (e == cudaSuccess) e = cudaMalloc(&s->inf, sizeof(int) * np);
    if (e == cudaSuccess) e = cudaMalloc(&s->key, sizeof(unsigned long long) * np);
    if (e == cudaSuccess) e = cudaMalloc(&s->key2, 2 * own * 16);
    if (e == cudaSuccess) e = cudaMalloc(&s->result, 2 * sizeof(SwarmResult));
    if (e == cudaSuccess) e = cudaMemset(s->result, 0, 2 * sizeof(SwarmResult));
    if (e == cudaSuccess) e = cudaMalloc(&s->ctl, sizeof(SwarmCtl));
    if (e == cudaSuccess) e = cudaMemcpy(s->x0, x0, sizeof(double) * HPE_NDOF, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the cleared flags must be in place before a peer can attach
    if (e != cudaSuccess) {
        hpe_swarm_destroy(s);
        return hpe_set_error(HPE_ERR_CUDA, "hpe_swarm_create: %s", cudaGetErrorString(e));
    }
    *out = s;
    return HPE_OK;
}

// the table the last hpe_swarm_init / hpe_swarm_step wrote its owned rows into (it alternates)
double *hpe_swarm_table(hpe_swarm *s) { return s ? s->table[s->cur] : nullptr; }

int hpe_swarm_init(hpe_swarm *s, void *stream)
{
    if (!s) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_init: NULL swarm");
    cudaStream_t st = (cudaStream_t)stream;
    const int own = s->hi - s->lo;
    SwarmCtl c;
    memset(&c, 0, sizeof c);
    c.gbest_cost = 1e100; c.count = 100; c.iter = 0;
    HPE_CUDA(cudaMemcpyAsync(s->ctl, &c, sizeof c, cudaMemcpyHostToDevice, st));
    s->cur = 0;
    if (own > 0) {
        swarm_generate<<<nblk((long long)own * (HPE_NDOF / 2), 256), 256, 0, st>>>(s->prm, s->x0, s->seed, s->lo, s->hi, s->x, s->v);
        HPE_CUDA(hpe_launch_swarm_step(s->m->dm, s->o->v, s->slot, step_args(s, 0, s->table[0], s->table[0]), st));
    }
    HPE_CUDA(push_to_peers(s, st));
    HPE_CUDA(cudaGetLastError());
    return HPE_OK;
}

int hpe_swarm_reduce(hpe_swarm *s, void *stream)
{
    if (!s) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_reduce: NULL swarm");
    if (s->world > 1) swarm_wait<<<1, 32, 0, (cudaStream_t)stream>>>(s->flags, s->world, s->rank, s->gen, s->ctl, g_swarm_spin_limit * 8);
    swarm_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(s->num_p, s->table[s->cur], s->ctl, s->trace);
    HPE_CUDA(cudaGetLastError());
    return HPE_OK;
}

int hpe_swarm_step(hpe_swarm *s, void *stream)
{
    if (!s) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_step: NULL swarm");
    cudaStream_t st = (cudaStream_t)stream;
    const int np = s->num_p;
    const double *in = s->table[s->cur];
    double *out = s->table[s->cur ^ 1];
    swarm_links<<<nblk((long long)np * NK, 256), 256, 0, st>>>(np, s->seed, s->ctl, in, s->R, s->key, s->inf);
    swarm_inf_cost<<<nblk((long long)np * NK, 256), 256, 0, st>>>(np, in, s->R, s->key);
    swarm_inf_index<<<nblk((long long)np * (NK + 1), 256), 256, 0, st>>>(np, in, s->R, s->key, s->inf);
    HPE_CUDA(hpe_launch_swarm_step(s->m->dm, s->o->v, s->slot, step_args(s, 1, in, out), st));
    s->cur ^= 1;
    HPE_CUDA(push_to_peers(s, st));
    HPE_CUDA(cudaGetLastError());
    return HPE_OK;
}

// ---- peer exchange setup: each rank publishes one 64-byte CUDA IPC handle, then maps the others ----
int hpe_swarm_ipc_handle(hpe_swarm *s, unsigned char handle[64])
{
    if (!s || !handle) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_ipc_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    HPE_CUDA(cudaIpcGetMemHandle(&h, s->shared_base));
    memcpy(handle, &h, 64);
    return HPE_OK;
}

int hpe_swarm_attach_peers(hpe_swarm *s, const unsigned char *handles, int world, int rank)
{
    if (!s || !handles || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world)
        return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_attach_peers: bad argument (at most %d ranks)", MAX_PEERS);
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)p * 64, 64);
        HPE_CUDA(cudaIpcOpenMemHandle(&s->peer_base[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    s->world = world;
    s->rank = rank;
    return HPE_OK;
}

// the whole run (init + maxiter-1 generations, exchanges included) as one cooperative launch
int hpe_swarm_run(hpe_swarm *s, void *stream)
{
    if (!s) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_run: NULL swarm");
    SwarmRunArgs a;
    memset(&a, 0, sizeof a);
    a.prm = s->prm; a.seed = s->seed; a.W1 = s->W1; a.C1 = s->C1;
    a.num_p = s->num_p; a.lo = s->lo; a.hi = s->hi; a.maxiter = s->prm.maxiter;
    a.table[0] = s->table[0]; a.table[1] = s->table[1];
    a.x = s->x; a.v = s->v; a.x0 = s->x0; a.R = s->R; a.inf = s->inf; a.key = s->key; a.key2 = s->key2; a.ctl = s->ctl; a.trace = s->trace;
    a.spin_limit = g_swarm_spin_limit;
    a.result = s->result;
    a.world = s->world; a.rank = s->rank; a.gen_base = s->gen; a.my_flags = s->flags;
    const size_t tb = table_bytes(s->num_p);
    a.my_min = (volatile SwarmMin *)((char *)s->shared_base + 2 * tb + 256);
    a.cta_key = s->cta_key; a.cta_idx = s->cta_idx;
    for (int p = 0; p < s->world && s->world > 1; ++p) {
        if (p == s->rank) continue;
        char *base = (char *)s->peer_base[p];
        a.peer_table[0][p] = (double *)base;
        a.peer_table[1][p] = (double *)(base + tb);
        a.peer_flags[p] = (volatile unsigned *)(base + 2 * tb);
        a.peer_min[p] = (SwarmMin *)(base + 2 * tb + 256);
    }
    a.times = s->times;
    HPE_CUDA(hpe_launch_swarm_run(s->m->dm, s->o->v, s->slot, a, (cudaStream_t)stream));
    s->gen += (unsigned)s->prm.maxiter;
    // the last generation g = maxiter-1 wrote buffer (g == 0 ? 0 : 1 - (g-1)%2 ...): track it like the stepwise path
    s->cur = s->prm.maxiter > 1 ? (s->prm.maxiter - 1) & 1 : 0;
    return HPE_OK;
}

// bound of the peer-flag waits, in polls (~50-100 ns each): default 2^28 (tens of seconds).  Tests lower it to see the
// HPE_ERR_STATE of a peer that never delivers.
int hpe_swarm_set_wait_limit(long long polls)
{
    if (polls < 1) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_set_wait_limit: polls must be positive");
    g_swarm_spin_limit = polls;
    return HPE_OK;
}

// diagnostics: globaltimer stamps (ns) of the phases of the last hpe_swarm_run, 8 per generation: start, after the
// informant cost pass, after the index pass, end of this CTA's step work, after the step barrier, exchange done,
// reduce done, after the closing barrier.  The first call only switches the recording on (out may be NULL).
int hpe_swarm_debug_times(hpe_swarm *s, unsigned long long *out)
{
    if (!s) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_debug_times: NULL swarm");
    const size_t n = (size_t)s->prm.maxiter * 8;
    if (!s->times) {
        HPE_CUDA(cudaMalloc(&s->times, sizeof(unsigned long long) * n));
        HPE_CUDA(cudaMemset(s->times, 0, sizeof(unsigned long long) * n));
        return HPE_OK;
    }
    if (out) HPE_CUDA(cudaMemcpy(out, s->times, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
    return HPE_OK;
}

int hpe_swarm_best(hpe_swarm *s, double bestp[HPE_NDOF], double *best_cost, double *trace)
{
    if (!s) return hpe_set_error(HPE_ERR_ARG, "hpe_swarm_best: NULL swarm");
    SwarmCtl c;
    HPE_CUDA(cudaMemcpy(&c, s->ctl, sizeof c, cudaMemcpyDeviceToHost));
    if (c.error)               // set by the bounded waits of the peer exchange, cleared by hpe_swarm_init / _run
        return hpe_set_error(HPE_ERR_STATE, "sharded swarm: a peer did not deliver its rows in time (rank %d of %d)", s->rank, s->world);
    if (bestp) memcpy(bestp, c.gbest, sizeof c.gbest);
    if (best_cost) *best_cost = c.gbest_cost;
    if (trace && s->prm.maxiter > 1)
        HPE_CUDA(cudaMemcpy(trace, s->trace, sizeof(double) * (s->prm.maxiter - 1), cudaMemcpyDeviceToHost));
    return HPE_OK;
}

}  // extern "C"
