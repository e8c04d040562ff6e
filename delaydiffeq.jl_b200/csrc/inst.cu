// inst.cu — the kernels of ONE model: compiled once per registry model with -DB2D_INST_MODEL=<struct> (and
// -DB2D_INST_JAC=1 for models that provide a Jacobian), see the Makefile.
#include "launch.cuh"

namespace b2dhost {
#define B2D_INSTANTIATE(ALG)                                                                                        \
    template int launch_t<b2d::B2D_INST_MODEL, ALG, 0>(b2d_solver*, Slot&, b2d::SolveParams&, cudaStream_t);        \
    template int launch_t<b2d::B2D_INST_MODEL, ALG, 1>(b2d_solver*, Slot&, b2d::SolveParams&, cudaStream_t);        \
    template int launch_t<b2d::B2D_INST_MODEL, ALG, 2>(b2d_solver*, Slot&, b2d::SolveParams&, cudaStream_t);        \
    template int launch_t<b2d::B2D_INST_MODEL, ALG, 3>(b2d_solver*, Slot&, b2d::SolveParams&, cudaStream_t);
B2D_INSTANTIATE(B2D_ALG_TSIT5)
B2D_INSTANTIATE(B2D_ALG_BS3)
#if B2D_INST_JAC
B2D_INSTANTIATE(B2D_ALG_ROSENBROCK23)
#endif
template int eval_model_t<b2d::B2D_INST_MODEL>(int, const double*, const double*, const double*, const double*, double, const double*,
                                               double*, double*, double*, double*);
}  // namespace b2dhost
