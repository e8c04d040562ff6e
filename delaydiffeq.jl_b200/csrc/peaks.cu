// peaks.cu — measured roofline denominators that MEASURED_PEAKS.json does not carry: the FP64 FMA peak of the
// device (the MethodOfSteps kernel is FP64-issue bound, not HBM bound; see DESIGN.md "Roofline").
#include <cuda_runtime.h>

#include "b200dde.h"

namespace {
__global__ void __launch_bounds__(256) fp64_fma_kernel(double* out, int iters, double x, double y) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
}  // namespace

extern "C" int b2d_measure_fp64_tflops(int32_t device, double* tflops) {
    if (!tflops) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        fp64_fma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-6);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return -1; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = fl / (ms * 1e-3) * 1e-12 > best ? fl / (ms * 1e-3) * 1e-12 : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return 0;
}
