// host_internal.h — types and helpers shared by the host translation units of libb200dde.so (not installed).
// The kernels are instantiated in one translation unit per model (csrc/inst.cu compiled with -DB2D_INST_MODEL=...), so
// that `make -j` builds them in parallel; b200dde.cu only references launch_t<Model, Alg, OutputMode>.
#ifndef B2D_HOST_INTERNAL_H
#define B2D_HOST_INTERNAL_H
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cctype>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b200dde.h"
#include "mos_device.cuh"

namespace b2dhost {

int fail(const char* fmt, ...);  // sets the thread-local error text (b2d_last_error), returns -1
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct ModelInfo {
    int n, np, ncl, ndl, nh, order0;
    bool has_jac;
    int mass_kind;     // 0 identity, 1 non-singular diagonal, 2 singular (DDAE)
    double mass[16];
};
template <class M>
constexpr ModelInfo info_of() {
    ModelInfo mi{M::N, M::NP, M::NCL, M::NDL, M::NH, M::ORDER0, M::HAS_JAC, 0, {}};
    mi.mass_kind = !b2d::MassOf<M>::HAS ? 0 : (b2d::MassOf<M>::singular() ? 2 : 1);
    for (int i = 0; i < M::N && i < 16; ++i) mi.mass[i] = b2d::MassOf<M>::at(i);
    return mi;
}
constexpr int kThreads = B2D_CTA_THREADS;  // 2 warps per CTA
#ifndef B2D_MINBLOCKS
#define B2D_MINBLOCKS 4
#endif
constexpr int kMaxSlots = 4, kDefaultSlots = 3;
constexpr int kMinBlocks = B2D_MINBLOCKS;  // register budget 65536/(kMinBlocks*128) per thread

// One launch slot: stream, ring, tile counter and staging buffers.  Several slots (n_slots) let the device->host copy of one
// wave overlap the kernel of the next.
struct Slot {
    cudaStream_t stream = nullptr;
    double* ring = nullptr;
    size_t ring_bytes = 0;
    unsigned int* counter = nullptr;
    double* d_saveat = nullptr;
    int saveat_cap = 0;
    double *d_u0 = nullptr, *d_p = nullptr, *d_out = nullptr;
    int* d_rc = nullptr;
    long long* d_stats = nullptr;
    size_t cap_u0 = 0, cap_p = 0, cap_out = 0, cap_rc = 0, cap_stats = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // b2d_solve_device: completion of the last launch on the caller's stream, and the grid it left on the device
    cudaEvent_t ev_done = nullptr;
    bool dev_done_valid = false;
    cudaStream_t dev_done_stream = nullptr;
    std::vector<double> saveat_dev;
    bool saveat_on_device = false;
    // pinned staging for the small per-trajectory results: a D2H copy into pageable caller memory is synchronous and
    // would serialise the wave pipeline (kernel of wave i+1 could not be issued under the copy of wave i)
    int* h_rc = nullptr;
    long long* h_stats = nullptr;
    size_t cap_hrc = 0, cap_hstats = 0;
    // where the staged results of the wave in flight on this slot go once its stream has drained
    int* dst_rc = nullptr;
    long long* dst_stats = nullptr;
    size_t pending = 0;
    // pageable caller output (a Julia Array): the wave's block is copied device -> pinned staging asynchronously and from
    // there into the caller's array by host threads while the GPU works on the next wave
    double* h_out = nullptr;
    size_t cap_hout = 0;
    double* dst_out = nullptr;
    size_t pending_out = 0;  // doubles
};

}  // namespace b2dhost

// Kernels of a user model compiled at run time (NVRTC -> cubin -> driver-API module)
struct UserKernels {
    CUmodule mod = nullptr;
    CUfunction fn[4] = {nullptr, nullptr, nullptr, nullptr};  // [output mode: lean saveat, everystep, general saveat, general saveat
                                                              //  without callback / Anderson code]
    int F[4] = {0, 0, 0, 0};                                  // ring fields per node
    int sm_rows[4] = {0, 0, 0, 0};                            // shared-memory rows per thread
};

struct b2d_solver {
    b2d_config cfg;
    b2dhost::ModelInfo mi;
    UserKernels* user = nullptr;
    // optional per-component tolerances and save_idxs (b2d_set_tolerances / b2d_set_save_idxs)
    bool use_tol_v = false;
    double abstol_v[16], reltol_v[16];
    int n_save_idx = 0;
    int save_idx[16];
    // solve(...; tstops, d_discontinuities) (b2d_set_tstops / b2d_set_d_discontinuities) and their device copies, which
    // are filtered to (t0, tf], multiplied by tdir and sorted for the time span of the solve at hand
    std::vector<double> user_tstops, user_disc_t;
    std::vector<int> user_disc_o;
    double *d_uts = nullptr, *d_udd_t = nullptr;
    int* d_udd_o = nullptr;
    int n_uts = 0, n_udd = 0;
    double stops_t0 = 0, stops_tf = 0;
    bool stops_dirty = true;
    // ring capacity chosen by the pilot launch (auto_cap) for this time span / lags, until an option changes
    int cap_hint = 0;
    double hint_t0 = 0, hint_tf = 0, hint_lags[4] = {0, 0, 0, 0};
    double *pilot_u0 = nullptr, *pilot_p = nullptr, *pilot_out = nullptr;
    int* pilot_rc = nullptr;
    long long* pilot_stats = nullptr;
    size_t pilot_cap_u0 = 0, pilot_cap_p = 0, pilot_cap_out = 0, pilot_cap_rc = 0, pilot_cap_stats = 0;
    int sm_count = 0;
    size_t l2_persist_max = 0, l2_window_max = 0;  // cudaDeviceProp::persistingL2CacheMaxSize / accessPolicyMaxWindowSize
    b2dhost::Slot slot[b2dhost::kMaxSlots];
    int n_slots = b2dhost::kDefaultSlots;  // waves in flight of the host-buffer path (experiment switch B2D_SLOTS)
    long long resident_traj = 0;  // trajectories one full grid holds at once (set by the first launch)
    double last_ms = 0.0;
    int last_launches = 0;
    int last_cap = 0;
    bool force_general = false;  // the re-solve of B2D_RC_INEXACT_DIVISION trajectories: general kernel (division operator)
    // per-trajectory constant lags (b2d_set_trajectory_lags): kept on the host; during a solve they ride behind every
    // trajectory's parameters (row width mi.np + p_cols_extra), which is what the general kernels read them from
    std::vector<double> traj_lags;
    long long traj_lags_n = 0, traj_lags_offset = 0;
    int p_cols_extra = 0;
    std::vector<double> saveat_inner;  // scratch
    // multi-process hosts: NCCL communicator for the final gather of the saveat blocks (b2d_comm_init)
    void* comm = nullptr;
    int comm_ranks = 0, comm_rank = 0;
};

namespace b2dhost {

void set_ring_l2_window(const b2d_solver* s, cudaStream_t stream, void* ring, size_t bytes);
void clear_ring_l2_window(const b2d_solver* s, cudaStream_t stream);
int cap_ctas_by_l2(const b2d_solver* s, int occ, size_t slab_bytes);
constexpr int kSaveatSmemMax = 512;  // saveat points staged in shared memory (4 KB per CTA at most)

// One launch of mos_kernel<M, ALG, OUT> for the trajectories described by P on `stream` (defined in launch.cuh,
// instantiated per model in inst.cu).
template <class M, int ALG, int OUT>
int launch_t(b2d_solver* s, Slot& sl, b2d::SolveParams& P, cudaStream_t stream);

template <class M>
int eval_model_t(int n_pts, const double* d_u, const double* d_p, const double* d_hv, const double* d_dhv, double t,
                 const double* d_lags, double* d_f, double* d_jac, double* d_tgrad, double* d_lag);

}  // namespace b2dhost
#endif
