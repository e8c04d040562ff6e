// models.cuh — device right-hand sides of the ahead-of-time model registry (b200dde.h B2D_MODEL_*).
//
// A Julia closure cannot cross a C ABI, so f(du,u,h,p,t), h(p,t) and the lag functions of the
// DDEProblem are device functors selected by id.  Each functor cites the Julia definition in the
// reference tree it reproduces.  The history handle `h` is the device HistoryFunction
// (mos_device.cuh, reference src/history_function.jl:20-62); `h.template eval<SLOT>(tq, hv)`
// returns only the NH history components the model declares through hist_comp(), and SLOT names
// the lag (constant lags first, then dependent lags) so that each lag keeps its own ring cursor.
//
// Arithmetic: this translation unit is compiled with -fmad=false; expressions are evaluated
// exactly as written (Julia does not contract user code either).
#ifndef B2D_MODELS_CUH
#define B2D_MODELS_CUH

namespace b2d {

// README.md:23-40 bc_model.  p = [p0, q0, v0] per trajectory, remaining constants are the README's.
struct BcModel {
    static constexpr int N = 3, NP = 3, NCL = 1, NDL = 0, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 2; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        const double p0 = p[0], q0 = p[1];
        const double d0 = 5.0, p1 = 0.2, q1 = 0.3, d1 = 1.0, d2 = 1.0;
        double w[NW];
        h.template evalw<0>(t - lags[0], p, w);  // w0 = v0 / (1 + beta0 h3^2), w1 = v1 / (1 + beta1 h3^2): see hmap
        const double w0 = w[0], w1 = w[1];
        du[0] = w0 * (p0 - q0) * u[0] - d0 * u[0];
        du[1] = w0 * (1.0 - p0 + q0) * u[0] + w1 * (p1 - q1) * u[1] - d1 * u[1];
        du[2] = w1 * (1.0 - p1 + q1) * u[1] - d2 * u[2];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? 1.0 : 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    // The part of f that depends on the history value only (README.md:29-32: the two quotients).  The solver evaluates it
    // once per history lookup instead of once per stage — stages 6 and 7 look the history up at the same time — and both
    // quotients share one reciprocal of their common divisor 1 + h3^2 (beta0 = beta1 = 1); dm_div_try is the compiler's
    // own division sequence, so the values are those of `v0 / (...)` and `v1 / (...)`, bit for bit.
    static constexpr int NW = 2;
    // hmap_try: branch-free (lean kernels; false = an operand outside the range of the fast division sequence);
    // hmap: the same quantities with the division operator (general kernels, diagnostics)
    __device__ __forceinline__ static bool hmap_try(const double* p, const double* hv, double* w) {
        const double v0 = p[2], v1 = 1.0, beta0 = 1.0, beta1 = 1.0;
        const double h3 = hv[0];
        const double den0 = 1.0 + beta0 * (h3 * h3), den1 = 1.0 + beta1 * (h3 * h3);
        const double y = dm_rcp_newton(den0);  // den1 is the same number
        const bool ok = dm_div_try(v0, den0, y, &w[0]);
        return dm_div_try(v1, den1, y, &w[1]) && ok;
    }
    __device__ __forceinline__ static void hmap(const double* p, const double* hv, double* w) {
        const double v0 = p[2], v1 = 1.0, beta0 = 1.0, beta1 = 1.0;
        const double h3 = hv[0];
        w[0] = v0 / (1.0 + beta0 * (h3 * h3));
        w[1] = v1 / (1.0 + beta1 * (h3 * h3));
    }
};

// Mackey-Glass, p = [beta, h0], tau = lags[0], exponent 10, decay 0.1.
struct MackeyGlassModel {
    static constexpr int N = 1, NP = 2, NCL = 1, NDL = 0, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        double w[NW];
        h.template evalw<0>(t - lags[0], p, w);  // beta x / (1 + x^10), x = u(t - tau): see hmap
        du[0] = w[0] - 0.1 * u[0];
    }
    __device__ __forceinline__ static void h_pre(const double* p, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? p[1] : 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    // the delayed production term: a function of the history value only, evaluated once per lookup (see BcModel::hmap)
    static constexpr int NW = 1;
    __device__ __forceinline__ static bool hmap_try(const double* p, const double* hv, double* w) {
        const double x = hv[0];
        const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x10 = x8 * x2;
        const double num = p[0] * x, den = 1.0 + x10;
        return dm_div_try(num, den, dm_rcp_newton(den), &w[0]);
    }
    __device__ __forceinline__ static void hmap(const double* p, const double* hv, double* w) {
        const double x = hv[0];
        const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x10 = x8 * x2;
        w[0] = p[0] * x / (1.0 + x10);
    }
};

// test/regression/allocations.jl:185-202: u' = -u(t - (0.5 + 0.1|u|)), h = 1.
struct StateDepModel {
    static constexpr int N = 1, NP = 1, NCL = 0, NDL = 1, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double*, double t,
                                             const double*) {
        const double tau = 0.5 + 0.1 * fabs(u[0]);
        double hv[NH];
        h.template eval<0>(t - tau, hv);
        du[0] = -hv[0];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? 1.0 : 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double* u, const double*, double) {
        return 0.5 + 0.1 * fabs(u[0]);
    }
};

// SURVEY §8d C5 stiff delayed kinetics, p = [lambda], mu = 50, h = (1,1,1).
struct StiffModel {
    static constexpr int N = 3, NP = 1, NCL = 1, NDL = 0, NH = 2, ORDER0 = 0;
    static constexpr bool HAS_JAC = true;
    __host__ __device__ static constexpr int hist_comp(int j) { return j; }  // u1, u2
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        const double lam = p[0], mu = 50.0;
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        du[0] = -lam * u[0] + lam * hv[1];
        du[1] = -u[1] + hv[0] * u[2];
        du[2] = -mu * (u[2] - 1.0) + 0.1 * u[0] * u[1];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? 1.0 : 0.0;
        hv[1] = deriv == 0 ? 1.0 : 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    template <class H>
    __device__ __forceinline__ static void jac(double* J, const double* u, H& h, const double* p, double t,
                                               const double* lags) {
        const double lam = p[0], mu = 50.0;
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        J[0] = -lam; J[1] = 0.0; J[2] = 0.0;
        J[3] = 0.0; J[4] = -1.0; J[5] = hv[0];
        J[6] = 0.1 * u[1]; J[7] = 0.1 * u[0]; J[8] = -mu;
    }
    template <class H>
    __device__ __forceinline__ static void tgrad(double* dT, const double* u, H& h, const double* p, double t,
                                                 const double* lags) {
        const double lam = p[0];
        double dh[NH];
        h.template deriv<0>(t - lags[0], dh);
        dT[0] = lam * dh[1];
        dT[1] = dh[0] * u[2];
        dT[2] = 0.0;
    }
};

// test/interface/parameters.jl:7-22 delayed logistic, h = 0.1, order_discontinuity_t0 = 1.
struct LogisticModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, NH = 1, ORDER0 = 1;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        du[0] = p[0] * u[0] * (1.0 - hv[0]);
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? 0.1 : 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// DDEProblemLibrary prob_dde_constant_1delay_ip (upstream test dependency): u' = -u(t-1), h = 0.
struct OneDelayModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = true;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double*, H& h, const double*, double t,
                                             const double* lags) {
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        du[0] = -hv[0];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int) { hv[0] = 0.0; }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    template <class H>
    __device__ __forceinline__ static void jac(double* J, const double*, H&, const double*, double, const double*) {
        J[0] = 0.0;
    }
    template <class H>
    __device__ __forceinline__ static void tgrad(double* dT, const double*, H& h, const double*, double t,
                                                 const double* lags) {
        double dh[NH];
        h.template deriv<0>(t - lags[0], dh);
        dT[0] = -dh[0];
    }
};

// DDEProblemLibrary prob_dde_constant_2delays_ip: u' = -u(t-1/3) - u(t-1/5), h = 0.
struct TwoDelaysModel {
    static constexpr int N = 1, NP = 1, NCL = 2, NDL = 0, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double*, H& h, const double*, double t,
                                             const double* lags) {
        double h1[NH], h2[NH];
        h.template eval<0>(t - lags[0], h1);
        h.template eval<1>(t - lags[1], h2);
        du[0] = -h1[0] - h2[0];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int) { hv[0] = 0.0; }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// DDEProblemLibrary prob_dde_constant_1delay_long_ip: u' = u - u(t-1/5), h = 0.
struct OneDelayLongModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double*, double t,
                                             const double* lags) {
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        du[0] = u[0] - hv[0];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int) { hv[0] = 0.0; }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// test/interface/dependent_delays.jl:18: 1-delay problem with dependent_lags = ((u,p,t) -> 1,).
struct OneDelayDepModel {
    static constexpr int N = 1, NP = 1, NCL = 0, NDL = 1, NH = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double*, H& h, const double*, double t,
                                             const double*) {
        double hv[NH];
        h.template eval<0>(t - 1.0, hv);
        du[0] = -hv[0];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int) { hv[0] = 0.0; }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 1.0; }
};

// test/interface/backwards.jl:4-12: u' = h(p, t+1), h = 1, reverse time, constant_lags = [-1].
struct BackwardsModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, NH = 1, ORDER0 = 1;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int) { return 0; }
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double*, H& h, const double*, double t,
                                             const double* lags) {
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        du[0] = hv[0];
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? 1.0 : 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// test/interface/mass_matrix.jl:7-19: SIR with delayed recovery, S' = -g I S, I' = g I S - g I(t-tau) S(t-tau),
// R' = g I(t-tau) S(t-tau); p = [gamma], constant_lags = [tau] (4 in the test), h = (1, 0, 0).
struct SirModel {
    static constexpr int N = 3, NP = 1, NCL = 1, NDL = 0, NH = 2, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    __host__ __device__ static constexpr int hist_comp(int j) { return j; }  // S, I
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        const double g = p[0];
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        const double infection = g * u[1] * u[0];
        const double recovery = g * hv[1] * hv[0];
        du[0] = -infection;
        du[1] = infection - recovery;
        du[2] = recovery;
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? 1.0 : 0.0;
        hv[1] = 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// test/interface/mass_matrix.jl:33-58: the same system as a delay differential-algebraic problem, third equation
// replaced by the conservation law 0 = S + I + R - 1, mass_matrix = Diagonal([1, 1, 0]) (singular: isdae, neutral).
struct SirDdaeModel {
    static constexpr int N = 3, NP = 1, NCL = 1, NDL = 0, NH = 2, ORDER0 = 0;
    static constexpr bool HAS_JAC = true;
    static constexpr bool HAS_MASS = true;
    __host__ __device__ static constexpr double mass(int i) { return i < 2 ? 1.0 : 0.0; }
    __host__ __device__ static constexpr int hist_comp(int j) { return j; }  // S, I
    template <class H>
    __device__ __forceinline__ static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        const double g = p[0];
        double hv[NH];
        h.template eval<0>(t - lags[0], hv);
        const double infection = g * u[1] * u[0];
        const double recovery = g * hv[1] * hv[0];
        du[0] = -infection;
        du[1] = infection - recovery;
        du[2] = u[0] + u[1] + u[2] - 1.0;
    }
    __device__ __forceinline__ static void h_pre(const double*, double, double* hv, int deriv) {
        hv[0] = deriv == 0 ? 1.0 : 0.0;
        hv[1] = 0.0;
    }
    __device__ __forceinline__ static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    template <class H>
    __device__ __forceinline__ static void jac(double* J, const double* u, H&, const double* p, double, const double*) {
        const double g = p[0];
        J[0] = -(g * u[1]); J[1] = -(g * u[0]); J[2] = 0.0;
        J[3] = g * u[1]; J[4] = g * u[0]; J[5] = 0.0;
        J[6] = 1.0; J[7] = 1.0; J[8] = 1.0;
    }
    template <class H>
    __device__ __forceinline__ static void tgrad(double* dT, const double*, H& h, const double* p, double t,
                                                 const double* lags) {
        const double g = p[0];
        double hv[NH], dh[NH];
        h.template eval<0>(t - lags[0], hv);
        h.template deriv<0>(t - lags[0], dh);
        dT[0] = 0.0;
        dT[1] = -(g * (dh[1] * hv[0] + hv[1] * dh[0]));
        dT[2] = 0.0;
    }
};

// History map of a model: `hmap(p, hv, w)` turns the NH looked-up history values of one lag into the NW quantities its f
// actually uses (`h.evalw`).  Models without the members use the history values themselves (NW = NH, identity).
template <class M, class = void>
struct HmapOf {
    static constexpr bool HAS = false;
    static constexpr int NW = M::NH;
    __device__ __forceinline__ static void apply(const double*, const double* hv, double* w) {
#pragma unroll
        for (int c = 0; c < M::NH; ++c) w[c] = hv[c];
    }
    __device__ __forceinline__ static bool apply_try(const double* p, const double* hv, double* w) {
        apply(p, hv, w);
        return true;
    }
};
template <class M>
struct HmapOf<M, decltype((void)M::NW)> {
    static constexpr bool HAS = true;
    static constexpr int NW = M::NW;
    __device__ __forceinline__ static void apply(const double* p, const double* hv, double* w) { M::hmap(p, hv, w); }
    __device__ __forceinline__ static bool apply_try(const double* p, const double* hv, double* w) { return M::hmap_try(p, hv, w); }
};

// DiscreteCallback of a model: `HAS_CALLBACK` + `cb_condition(u, t, p, par)` / `cb_affect(u, t, p, par)` (returns true for
// terminate!); `par` are the four numbers b2d_config::cb_par0..3.  Models without the members have none.
template <class M, class = void>
struct CallbackOf {
    static constexpr bool HAS = false;
    __device__ __forceinline__ static bool condition(const double*, double, const double*, const double*) { return false; }
    __device__ __forceinline__ static bool affect(double*, double, const double*, const double*) { return false; }
};
template <class M>
struct CallbackOf<M, decltype((void)M::HAS_CALLBACK)> {
    static constexpr bool HAS = M::HAS_CALLBACK;
    __device__ __forceinline__ static bool condition(const double* u, double t, const double* p, const double* par) {
        return M::cb_condition(u, t, p, par);
    }
    __device__ __forceinline__ static bool affect(double* u, double t, const double* p, const double* par) {
        return M::cb_affect(u, t, p, par);
    }
};

// ContinuousCallback of a model: `HAS_CONT_CALLBACK` + `cb_value(u, t, p, par)`, the real-valued condition whose sign change is
// the event (b2d_config::callback = B2D_CB_CONT_MODEL); the affect! is the model's cb_affect when it has one (HAS_CALLBACK),
// else the preset b2d_config::cb_affect.
template <class M, class = void>
struct ContCallbackOf {
    static constexpr bool HAS = false;
    __device__ __forceinline__ static double value(const double*, double, const double*, const double*) { return 1.0; }
};
template <class M>
struct ContCallbackOf<M, decltype((void)M::HAS_CONT_CALLBACK)> {
    static constexpr bool HAS = M::HAS_CONT_CALLBACK;
    __device__ __forceinline__ static double value(const double* u, double t, const double* p, const double* par) {
        return M::cb_value(u, t, p, par);
    }
};

// Mass matrix of a model: diagonal, compile-time (`HAS_MASS` + `mass(i)`); models without the members have M = I.
template <class M, class = void>
struct MassOf {
    static constexpr bool HAS = false;
    __host__ __device__ static constexpr double at(int) { return 1.0; }
    __host__ __device__ static constexpr bool singular() { return false; }
};
template <class M>
struct MassOf<M, decltype((void)M::HAS_MASS)> {
    static constexpr bool HAS = M::HAS_MASS;
    __host__ __device__ static constexpr double at(int i) { return M::mass(i); }
    __host__ __device__ static constexpr bool singular() {
        for (int i = 0; i < M::N; ++i)
            if (M::mass(i) == 0.0) return true;
        return false;
    }
};

}  // namespace b2d
#endif
