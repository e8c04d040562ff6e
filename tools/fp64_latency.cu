// fp64_latency.cu — dependent-issue latency and throughput of the FP64 instructions the solver kernel is made of (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/fp64_latency tools/fp64_latency.cu
// One warp per SM sub-partition measures latency (a dependent chain); `ilp` independent chains show how much
// instruction-level parallelism the FP64 pipe needs before it is throughput-bound; W warps per SMSP show the same for
// thread-level parallelism.  Results go to stdout as a table (profiles/r02_fp64_latency.txt).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void ddiv_chain(double* out, long long* cyc, int iters, double b) {
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) x = (x + 1.0) / b;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dsqrt_chain(double* out, long long* cyc, int iters) {
    double x = 2.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) x = sqrt(x + 1.0);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// pointer chase through a buffer of `bytes` (one lane): L1 / L2 hit latency as the kernel sees it
__global__ void chase(const unsigned* next, unsigned* out, long long* cyc, int iters) {
    unsigned i = 0;
    for (int k = 0; k < 1024; ++k) i = __ldcg(next + i);  // warm
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) i = __ldcg(next + i);
    long long t1 = clock64();
    *out = i;
    *cyc = t1 - t0;
}

template <class F>
double run(F launch, int ops) {
    long long* d_c;
    cudaMalloc(&d_c, 8);
    launch(d_c);
    launch(d_c);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_c);
    return (double)c / ops;
}

int main() {
    double* d_out;
    cudaMalloc(&d_out, 1 << 24);
    const int iters = 4096;
    printf("# FP64 dependent-issue latency / throughput, cycles per instruction of ONE warp (clock64), B200\n");
    printf("%-44s %10s\n", "experiment", "cyc/inst");
#define ROW(name, ILPV, warps)                                                                                     \
    printf("%-44s %10.2f\n", name,                                                                                 \
           run([&](long long* c) { dfma_chain<ILPV><<<1, 32 * warps>>>(d_out, c, iters, 0.999, 0.001); }, iters * 8 * ILPV));
    ROW("DFMA chain, 1 warp, ILP 1 (latency)", 1, 1)
    ROW("DFMA, 1 warp, ILP 2", 2, 1)
    ROW("DFMA, 1 warp, ILP 4", 4, 1)
    ROW("DFMA, 1 warp, ILP 8", 8, 1)
    ROW("DFMA, 4 warps (1 per SMSP), ILP 1", 1, 4)
    ROW("DFMA, 16 warps (4 per SMSP), ILP 1", 1, 16)
    ROW("DFMA, 16 warps (4 per SMSP), ILP 2", 2, 16)
    ROW("DFMA, 16 warps (4 per SMSP), ILP 4", 4, 16)
    ROW("DFMA, 20 warps (5 per SMSP), ILP 1", 1, 20)
    ROW("DFMA, 32 warps (8 per SMSP), ILP 1", 1, 32)
    printf("%-44s %10.2f\n", "(x+1)/b chain, 1 warp (cycles per division)",
           run([&](long long* c) { ddiv_chain<<<1, 32>>>(d_out, c, iters, 1.0000001); }, iters * 8));
    printf("%-44s %10.2f\n", "(x+1)/b chain, 16 warps",
           run([&](long long* c) { ddiv_chain<<<1, 512>>>(d_out, c, iters, 1.0000001); }, iters * 8));
    printf("%-44s %10.2f\n", "sqrt(x+1) chain, 1 warp (cycles per sqrt)",
           run([&](long long* c) { dsqrt_chain<<<1, 32>>>(d_out, c, iters); }, iters * 8));
    // pointer chase: 64 KB (L1 bypassed by ld.cg -> L2 hit), 64 MB (L2 hit, both dies), 1 GB (DRAM)
    for (size_t bytes : {(size_t)64 << 10, (size_t)64 << 20, (size_t)1 << 30}) {
        const size_t n = bytes / 4;
        unsigned* h = (unsigned*)malloc(bytes);
        const size_t stride = 4099 * 8;  // words; coprime with n's power of two
        for (size_t i = 0; i < n; ++i) h[i] = (unsigned)((i + stride) % n);
        unsigned *d, *d_o;
        cudaMalloc(&d, bytes);
        cudaMalloc(&d_o, 4);
        cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
        char name[64];
        snprintf(name, sizeof name, "ld.cg pointer chase, %zu KB buffer", bytes >> 10);
        printf("%-44s %10.2f\n", name, run([&](long long* c) { chase<<<1, 1>>>(d, d_o, c, 20000); }, 20000));
        cudaFree(d);
        cudaFree(d_o);
        free(h);
    }
    cudaFree(d_out);
    return 0;
}
