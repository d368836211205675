// Microbenchmarks of B200 issue behaviour for the instruction mixes of the matcher sweep.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// NF scalar FFMA (operands a,b shared -> reuse cache) and NA FMNMX per iteration, all independent chains
template <int NF, int NA>
__global__ void __launch_bounds__(256) mix_kernel(float *out, int iters, float a, float b)
{
    float r[16], m[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { r[k] = threadIdx.x * 1e-3f + k; m[k] = 1e30f - k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NF; ++k) r[k % 16] = fmaf(r[k % 16], a, b);
#pragma unroll
        for (int k = 0; k < NA; ++k) m[k % 16] = fminf(m[k % 16], r[(k + 5) % 16]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += r[k] + m[k];
    if (s == 123.456f) out[0] = s;
}

// NF2 FFMA2 (shared operands) and NA FMNMX
template <int NF2, int NA>
__global__ void __launch_bounds__(256) mix2_kernel(float *out, int iters, unsigned long long a, unsigned long long b)
{
    unsigned long long r[8];
    float m[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = ((unsigned long long)__float_as_uint(threadIdx.x * 1e-3f + k) << 32) | 0x3f800000u;
#pragma unroll
    for (int k = 0; k < 16; ++k) m[k] = 1e30f - k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NF2; ++k) r[k % 8] = ffma2(r[k % 8], a, b);
#pragma unroll
        for (int k = 0; k < NA; ++k) m[k % 16] = fminf(m[k % 16], __uint_as_float((unsigned)(r[(k + 3) % 8] >> ((k & 1) * 32))));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += m[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) s += __uint_as_float((unsigned)r[k]);
    if (s == 123.456f) out[0] = s;
}

template <class F>
static double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

int main()
{
    int sms, khz;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *d;
    cudaMalloc(&d, 1024);
    const int iters = 1 << 12, blocks = sms * 8, threads = 256;
    const double warps_per_smsp = 8.0 * 8 / 4;   // resident warps per scheduler
    auto report = [&](const char *name, double ms, int instr) {
        // cycles per iteration per warp-scheduler slot: time * clock / (iters * warps per SMSP)
        double cyc = ms * 1e-3 * khz * 1e3 / (iters * warps_per_smsp);
        printf("%-28s %8.3f ms  %6.2f SMSP-cycles per iteration (%d instructions -> IPC %.2f)\n", name, ms, cyc, instr,
               instr / cyc);
    };
#define RUN(NF, NA) report("FFMA x" #NF " + FMNMX x" #NA, time_ms([&] { mix_kernel<NF, NA><<<blocks, threads>>>(d, iters, 1.0001f, 1e-7f); }), NF + NA)
#define RUN2(NF, NA) report("FFMA2 x" #NF " + FMNMX x" #NA, time_ms([&] { mix2_kernel<NF, NA><<<blocks, threads>>>(d, iters, 0x3f8000003f800000ull, 0ull); }), NF + NA)
    printf("clock attr %.0f MHz, %d SMs\n", khz * 1e-3, sms);
    RUN(48, 0); RUN(0, 16); RUN(48, 8); RUN(48, 16); RUN(48, 24); RUN(48, 32); RUN(32, 32);
    RUN2(24, 0); RUN2(24, 8); RUN2(24, 16); RUN2(24, 24); RUN2(24, 32);
    cudaFree(d);
    return 0;
}
