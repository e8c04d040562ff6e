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

}  // namespace b2dhost
#endif
