// la3dm_b200 -- measurement helper: FP32 FMA throughput of the device (the denominator of the fp32-pipe roofline of
// the predict kernels; SURVEY.md section 8d asks for it to be measured on the box, like MEASURED_PEAKS.json's HBM number).
// Not part of the reference's API.
#include "common.cuh"

namespace la3dm_b200 {
namespace {

constexpr int kIters = 4096;
constexpr int kChains = 8;

// 8 independent register-resident FMA chains per thread; explicit __fmaf_rn (the library is built with -fmad=false)
__global__ void __launch_bounds__(256) k_fma_peak(float *out, float a, float b) {
    float v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = (float) (threadIdx.x + c) * 1e-3f;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) v[c] = __fmaf_rn(v[c], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += v[c];
    if (s == 123.456f) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace
}  // namespace la3dm_b200

extern "C" int la3dm_bench_fp32_peak(int device, float *tflops) {
    if (!tflops) return LA3DM_ERR_INVALID;
    *tflops = 0.f;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return LA3DM_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) return LA3DM_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return LA3DM_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LA3DM_ERR_CUDA;
    float *d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return LA3DM_ERR_CUDA;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int ctas = prop.multiProcessorCount * 8 * 4;
    float best = 0.f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        la3dm_b200::k_fma_peak<<<ctas, 256>>>(d, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * (double) ctas * 256.0 * la3dm_b200::kIters * la3dm_b200::kChains;
        const float tf = (float) (flop / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;   // first launch is the warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return LA3DM_ERR_CUDA;
    *tflops = best;
    return LA3DM_OK;
}
