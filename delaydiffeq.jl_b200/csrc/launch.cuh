// launch.cuh — definition of launch_t: occupancy, ring allocation, L2 window and the kernel launch itself.
#ifndef B2D_LAUNCH_CUH
#define B2D_LAUNCH_CUH
#include "host_internal.h"

namespace b2dhost {

template <class M, int ALG, int OUT>
int launch_t(b2d_solver* s, Slot& sl, b2d::SolveParams& P, cudaStream_t stream) {
    using T = b2d::Traj<M, ALG, OUT>;
    auto kern = b2d::mos_kernel<M, ALG, OUT, b2d::LaunchCfg<M>::MINB>;
    static_assert(T::SM_STRIDE == kThreads, "shared-memory rows are strided by the CTA size");
    size_t smem = (size_t)T::SM_ROWS * kThreads * sizeof(double);
    P.save_smem = (P.n_save > 0 && P.n_save <= kSaveatSmemMax && !getenv("B2D_NO_SAVEAT_SMEM")) ? 1 : 0;
    if (P.save_smem) smem += (size_t)P.n_save * sizeof(double);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (const char* e = getenv("B2D_SMEM_CARVEOUT"))  // experiment switch: preferred shared-memory carveout in percent
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    if (occ < 1) occ = 1;
    if (const char* e = getenv("B2D_BLOCKS_PER_SM")) {  // occupancy experiments (profiles/README.md)
        const int lim = atoi(e);
        if (lim >= 1 && lim < occ) occ = lim;
    }
    const size_t slab = (size_t)P.cap * T::F * 32 * sizeof(double);
    occ = cap_ctas_by_l2(s, occ, slab);
    long long tiles = (P.n_traj + 31) / 32;
    long long want_blocks = (tiles + (kThreads / 32) - 1) / (kThreads / 32);
    int blocks = (int)std::min<long long>((long long)s->sm_count * occ, std::max<long long>(want_blocks, 1));
    const size_t need = slab * (size_t)blocks * (kThreads / 32);
    if (need > sl.ring_bytes) {
        if (sl.ring) cudaFree(sl.ring);
        sl.ring = nullptr;
        sl.ring_bytes = 0;
        CK(cudaMalloc((void**)&sl.ring, need));
        sl.ring_bytes = need;
    }
    P.ring = sl.ring;
    P.tile_counter = sl.counter;
    s->resident_traj = (long long)s->sm_count * occ * kThreads;
    CK(cudaMemsetAsync(sl.counter, 0, sizeof(unsigned int), stream));
    set_ring_l2_window(s, stream, sl.ring, need);
    kern<<<blocks, kThreads, smem, stream>>>(P);
    clear_ring_l2_window(s, stream);
    CK(cudaGetLastError());
    s->last_launches += 1;
    return 0;
}


// ---- diagnostic: the model functors alone (b2d_eval_model) -----------------------------------------------------------
// History handle that returns caller-supplied values: the right-hand side, Jacobian, time gradient and lag functions of a
// registry model can then be compared with their CPU restatements point by point, independently of any solve.
template <class M>
struct EvalHandle {
    const double* hv;   // [slots][N] history values (full state vectors), this point
    const double* dhv;  // [slots][N] their time derivatives
    template <int S>
    __device__ __forceinline__ void eval(double, double* out) {
#pragma unroll
        for (int c = 0; c < M::NH; ++c) out[c] = hv[S * M::N + M::hist_comp(c)];
    }
    template <int S>
    __device__ __forceinline__ void deriv(double, double* out) {
#pragma unroll
        for (int c = 0; c < M::NH; ++c) out[c] = dhv[S * M::N + M::hist_comp(c)];
    }
    template <int S>
    __device__ __forceinline__ void evalw(double tq, const double* pp, double* w) {
        double h[M::NH];
        eval<S>(tq, h);
        b2d::HmapOf<M>::apply(pp, h, w);
    }
};
template <class M>
__global__ void eval_model_kernel(int n_pts, const double* u, const double* p, const double* hv, const double* dhv, double t,
                                  const double* lags, double* out_f, double* out_jac, double* out_tgrad, double* out_lag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pts) return;
    constexpr int NSL = (M::NCL + M::NDL) > 0 ? (M::NCL + M::NDL) : 1;
    double lg[4] = {0, 0, 0, 0};
    for (int k = 0; k < M::NCL && k < 4; ++k) lg[k] = lags[k];
    double uu[M::N], pp[M::NP > 0 ? M::NP : 1], du[M::N];
    for (int k = 0; k < M::N; ++k) uu[k] = u[(size_t)i * M::N + k];
    for (int k = 0; k < M::NP; ++k) pp[k] = p[(size_t)i * M::NP + k];
    EvalHandle<M> h{hv + (size_t)i * NSL * M::N, dhv + (size_t)i * NSL * M::N};
    M::f(du, uu, h, pp, t, lg);
    for (int k = 0; k < M::N; ++k) out_f[(size_t)i * M::N + k] = du[k];
    for (int k = 0; k < M::NDL; ++k) out_lag[(size_t)i * (M::NDL > 0 ? M::NDL : 1) + k] = M::dep_lag(k, uu, pp, t);
    if constexpr (M::HAS_JAC) {
        double J[M::N * M::N], dT[M::N];
        M::jac(J, uu, h, pp, t, lg);
        M::tgrad(dT, uu, h, pp, t, lg);
        for (int k = 0; k < M::N * M::N; ++k) out_jac[(size_t)i * M::N * M::N + k] = J[k];
        for (int k = 0; k < M::N; ++k) out_tgrad[(size_t)i * M::N + k] = dT[k];
    }
}
template <class M>
int eval_model_t(int n_pts, const double* d_u, const double* d_p, const double* d_hv, const double* d_dhv, double t,
                 const double* d_lags, double* d_f, double* d_jac, double* d_tgrad, double* d_lag) {
    eval_model_kernel<M><<<(n_pts + 127) / 128, 128>>>(n_pts, d_u, d_p, d_hv, d_dhv, t, d_lags, d_f, d_jac, d_tgrad, d_lag);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

}  // namespace b2dhost
#endif
