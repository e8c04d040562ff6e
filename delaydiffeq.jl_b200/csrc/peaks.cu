// peaks.cu — measured roofline denominators that MEASURED_PEAKS.json does not carry: the FP64 FMA peak of the
// device (the MethodOfSteps kernel is FP64-issue bound, not HBM bound; see DESIGN.md "Roofline").
#include "detmath.h"
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


// ---- diagnostic: dm_div_try against the division operator ----------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned long long splitmix(unsigned long long& x) {
    unsigned long long z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// operand classes: 0 random bit patterns, 1 random significands with exponents of solver magnitudes, 2 quotients next to a
// rounding tie (a = q b perturbed by one ulp), 3 powers of two / tiny / huge operands
__global__ void division_check_kernel(unsigned long long seed, unsigned long long n, int cls, unsigned long long* mismatches,
                                      unsigned long long* flagged) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0, flag = 0;
    for (; i < n; i += stride) {
        unsigned long long st = seed ^ (i * 0xD1342543DE82EF95ull);
        double a, b;
        if (cls == 0) {
            a = __longlong_as_double((long long)splitmix(st));
            b = __longlong_as_double((long long)splitmix(st));
        } else if (cls == 1) {
            const unsigned long long ma = splitmix(st) >> 12, mb = splitmix(st) >> 12;
            const int ea = (int)(splitmix(st) % 80) - 60, eb = (int)(splitmix(st) % 80) - 60;
            a = ldexp(1.0 + (double)ma * 2.220446049250313e-16, ea) * ((splitmix(st) & 1) ? -1.0 : 1.0);
            b = ldexp(1.0 + (double)mb * 2.220446049250313e-16, eb) * ((splitmix(st) & 1) ? -1.0 : 1.0);
        } else if (cls == 2) {
            const double q = 1.0 + (double)(splitmix(st) >> 12) * 2.220446049250313e-16;
            b = 1.0 + (double)(splitmix(st) >> 12) * 2.220446049250313e-16;
            a = q * b;
            const int k = (int)(splitmix(st) % 5) - 2;
            a = __longlong_as_double(__double_as_longlong(a) + k);
        } else {
            const int ea = (int)(splitmix(st) % 2100) - 1074, eb = (int)(splitmix(st) % 2100) - 1074;
            a = ldexp((splitmix(st) & 1) ? 1.0 : 1.5, ea);
            b = ldexp((splitmix(st) & 3) ? 1.0 : 1.75, eb);
            if ((splitmix(st) % 64) == 0) a = 0.0;
        }
        double q;
        const bool ok = dm_div_try(a, b, dm_rcp_newton(b), &q);
        const double ref = a / b;
        if (ok) {
            if (__double_as_longlong(q) != __double_as_longlong(ref) && !(q != q && ref != ref)) ++bad;
        } else {
            ++flag;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
    if (flag) atomicAdd(flagged, flag);
}
}  // namespace

extern "C" int b2d_check_division(int32_t device, int32_t operand_class, uint64_t n_pairs, uint64_t seed, uint64_t* mismatches,
                                  uint64_t* flagged) {
    if (!mismatches || !flagged || operand_class < 0 || operand_class > 3) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, 16) != cudaSuccess) return -1;
    cudaMemset(d, 0, 16);
    division_check_kernel<<<148 * 8, 256>>>(seed, n_pairs, operand_class, d, d + 1);
    unsigned long long h[2] = {0, 0};
    const bool ok = cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d);
    if (!ok) return -1;
    *mismatches = h[0];
    *flagged = h[1];
    return 0;
}
