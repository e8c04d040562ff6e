// ifetch_rate.cu — issue rate of straight-line code as a function of its size (B200): how many instructions per cycle can a
// sub-partition issue when the loop body no longer fits its L0 instruction cache and streams from the SM's L1.5 / L2?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ifetch_rate tools/ifetch_rate.cu
// The body is N dependent-chain-free FFMAs (8 independent accumulators per thread, 4 warps per sub-partition: never
// latency-bound, FP32 pipe limit 1 warp-instruction per cycle per sub-partition), unrolled so that the loop is N x 16 bytes of
// SASS.  The solver kernels' hot loops are 20-30 KB (profiles/README.md); their issue rate sits at 0.46-0.56 per cycle.
#include <cstdio>
#include <cuda_runtime.h>

template <int N, bool F64>
__global__ void __launch_bounds__(512, 1) body(float* out, long long* cyc, int iters, float a, float b) {
    float x[8];
    double y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = x[i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (F64 && (k % 3 == 0)) y[(k / 3) % 8] = fma(y[(k / 3) % 8], (double)a, (double)b);  // one DFMA in three: the solver's FP64 share
            else x[k % 8] = fmaf(x[k % 8], a, b);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + (float)y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int N, bool F64>
void run(float* d_out, int warps_per_smsp, int ctas) {
    long long *d_c, c = 0;
    cudaMalloc(&d_c, 8);
    const int iters = (1 << 22) / N;
    for (int rep = 0; rep < 2; ++rep) body<N, F64><<<ctas, 128 * warps_per_smsp>>>(d_out, d_c, iters, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_c);
    const double inst = (double)iters * N * warps_per_smsp;  // warp instructions per sub-partition
    printf("%-8s body %6d instr (%5.1f KB)  %d warps/SMSP  %3d CTAs: %.3f warp-instr / cycle / SMSP\n", F64 ? "mixed64" : "fp32", N, N * 16 / 1024.0,
           warps_per_smsp, ctas, inst / (double)c);
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 148 * 512 * 4 * 4);
    printf("# issue rate of unrolled straight-line code vs its size (1 CTA = one SM alone; 148 CTAs = every SM streaming from L2 at once)\n");
#define ROWS(F, W, C) run<128, F>(d_out, W, C); run<256, F>(d_out, W, C); run<384, F>(d_out, W, C); run<512, F>(d_out, W, C); run<768, F>(d_out, W, C); \
    run<1024, F>(d_out, W, C); run<1536, F>(d_out, W, C); run<2048, F>(d_out, W, C); run<3072, F>(d_out, W, C); run<4096, F>(d_out, W, C); run<8192, F>(d_out, W, C);
    ROWS(false, 4, 1)
    ROWS(false, 4, 148)
    ROWS(true, 4, 148)
    ROWS(true, 1, 148)
    cudaFree(d_out);
    return 0;
}
