// User model with a state-dependent lag: u'(t) = -u(t - (0.5 + 0.1|u(t)|)), h = 1
// (test/regression/allocations.jl:185-202 of the reference), dependent_lags = [(u,p,t) -> 0.5 + 0.1|u|].
struct MyStateDep {
    static constexpr int N = 1, NP = 1, NCL = 0, NDL = 1, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ static void f(double* du, const double* u, H& h, const double*, double t, const double*) {
        double hv[NH];
        h.template eval<0>(t - dep_lag(0, u, nullptr, t), hv);
        du[0] = -hv[0];
    }
    __device__ static void h_pre(const double*, double, double* hv, int deriv) { hv[0] = deriv == 0 ? 1.0 : 0.0; }
    __device__ static double dep_lag(int, const double* u, const double*, double) { return 0.5 + 0.1 * fabs(u[0]); }
};
