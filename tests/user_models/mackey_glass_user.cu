// A user-supplied model: the Julia right-hand side
//     f(du,u,h,p,t) = (du[1] = p[1]*h(p,t-17)[1]/(1 + h(p,t-17)[1]^10) - 0.1*u[1]);   h(p,t) = p[2]
// transcribed to the device-functor interface of csrc/models.cuh.  Mathematically (and arithmetically) the same as the
// registry's MackeyGlassModel, so the parity test can check it against the CPU oracle.
struct MyMackeyGlass {
    static constexpr int N = 1, NP = 2, NCL = 1, NDL = 0, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ static void f(double* du, const double* u, H& h, const double* p, double t, const double* lags) {
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        const double x = hv[0];
        const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x10 = x8 * x2;
        du[0] = p[0] * x / (1.0 + x10) - 0.1 * u[0];
    }
    __device__ static void h_pre(const double* p, double, double* hv, int deriv) { hv[0] = deriv == 0 ? p[1] : 0.0; }
    __device__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// The same model with a DiscreteCallback of its own: condition(u, t, integrator) = (t == par[0]),
// affect!(integrator) = (integrator.u[1] += par[1]) — the device functors the solver calls after every accepted step
// when the solve is configured with callback = B2D_CB_MODEL (par = b2d_config::cb_par0..3).
struct MyMackeyGlassKick : MyMackeyGlass {
    static constexpr bool HAS_CALLBACK = true;
    __device__ static bool cb_condition(const double*, double t, const double*, const double* par) { return t == par[0]; }
    __device__ static bool cb_affect(double* u, double, const double*, const double* par) {
        u[0] += par[1];
        return false;  // true would be terminate!(integrator)
    }
};

// The same model with a ContinuousCallback of its own: condition(u, t, integrator) = u[1] - par[0] (an event is a sign
// change of it inside a step), affect!(integrator) = (integrator.u[1] *= par[1]) — callback = B2D_CB_CONT_MODEL.
struct MyMackeyGlassThreshold : MyMackeyGlass {
    static constexpr bool HAS_CONT_CALLBACK = true;
    static constexpr bool HAS_CALLBACK = true;
    __device__ static double cb_value(const double* u, double, const double*, const double* par) { return u[0] - par[0]; }
    __device__ static bool cb_condition(const double*, double, const double*, const double*) { return false; }
    __device__ static bool cb_affect(double* u, double, const double*, const double* par) {
        u[0] *= par[1];
        return false;
    }
};
