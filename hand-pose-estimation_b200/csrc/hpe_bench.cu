// hpe_bench.cu -- measurement helpers: the FP32 (non-tensor) FFMA peak of the device, which
// MEASURED_PEAKS.json does not carry and the cost kernel's roofline is quoted against.
#include "hpe_internal.h"

namespace {

// 16 independent FFMA chains per thread, register operands only.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float *out, int iters, float a, float b)
{
    float r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = fmaf(r[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += r[k];
    if (s == 123.456f) out[0] = s;     // never true; keeps the chains alive
}

}  // namespace

extern "C" int hpe_measure_fp32_peak(double *tflops, double *sm_mhz_est)
{
    if (!tflops) return hpe_set_error(HPE_ERR_ARG, "hpe_measure_fp32_peak: NULL output");
    int dev = 0, sms = 0;
    HPE_CUDA(cudaGetDevice(&dev));
    HPE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *d = nullptr;
    HPE_CUDA(cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    HPE_CUDA(cudaEventCreate(&e0));
    HPE_CUDA(cudaEventCreate(&e1));
    const int iters = 1 << 14, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        HPE_CUDA(cudaEventRecord(e0, 0));
        ffma_peak_kernel<<<blocks, threads>>>(d, iters, 1.000001f, 1e-7f);
        HPE_CUDA(cudaEventRecord(e1, 0));
        HPE_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        HPE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
        double tf = flops / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    if (sm_mhz_est) *sm_mhz_est = best * 1e12 / (2.0 * 128.0 * sms) * 1e-6;   // if all 128 lanes issue
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return HPE_OK;
}
