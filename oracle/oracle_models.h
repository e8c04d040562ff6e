// oracle_models.h — CPU restatements of the right-hand sides used by the parity tests.
//
// TEST INFRASTRUCTURE ONLY (see oracle_dde.cpp header).  Each model mirrors one problem that
// the reference defines in Julia; the file:line of the Julia definition is cited per model.
// "upstream" = DDEProblemLibrary 0.1.2 (un-vendored test dependency of the reference,
// test/Project.toml:3,24); those equations are restated from the library's documentation and
// cross-checked through the reference's own known-answer tests that use them.
//
// A model provides
//   N, NP, NCL, NDL          sizes: states, parameters, constant lags, dependent lags
//   ORDER0                   default order_discontinuity_t0 of the DDEProblem
//   f(du,u,h,p,t,lags)       the Julia f(du,u,h,p,t); `h(tq,out)` returns the full history vector
//   h_pre(p,t,out[,deriv])   the Julia h(p,t) for t before t0
//   dep_lag(j,u,p,t)         the j-th entry of dependent_lags
//   jac / tgrad              analytic d f/d u and d f/d t (what ForwardDiff gives Rosenbrock23)
#ifndef ORACLE_MODELS_H
#define ORACLE_MODELS_H
#include <cmath>

namespace oracle {

// README.md:23-40 (the reference's headline example).  p = [p0, q0, v0]; the other constants
// are the README's literals.  h(p,t) = ones(3).
struct BcModel {
    static constexpr int N = 3, NP = 3, NCL = 1, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double* u, H& h, const double* p, double t, const double* lags) {
        const double p0 = p[0], q0 = p[1], v0 = p[2];
        const double d0 = 5.0, p1 = 0.2, q1 = 0.3, v1 = 1.0, d1 = 1.0, d2 = 1.0, beta0 = 1.0, beta1 = 1.0;
        double hv[N];
        h(t - lags[0], hv);
        const double h3 = hv[2];
        const double w0 = v0 / (1.0 + beta0 * (h3 * h3));
        const double w1 = v1 / (1.0 + beta1 * (h3 * h3));
        du[0] = w0 * (p0 - q0) * u[0] - d0 * u[0];
        du[1] = w0 * (1.0 - p0 + q0) * u[0] + w1 * (p1 - q1) * u[1] - d1 * u[1];
        du[2] = w1 * (1.0 - p1 + q1) * u[1] - d2 * u[2];
    }
    static void h_pre(const double*, double, double* out, int deriv = 0) {
        for (int i = 0; i < N; ++i) out[i] = deriv == 0 ? 1.0 : 0.0;
    }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// SURVEY §8d C3 / BASELINE.json configs[2]: Mackey-Glass, tau = lags[0] (17), n = 10, gamma = 0.1,
// p = [beta, h0], h(p,t) = h0.  (The reference only runs the tau=14 library variant,
// test/integrators/verner.jl:74-88.)
struct MackeyGlassModel {
    static constexpr int N = 1, NP = 2, NCL = 1, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double* u, H& h, const double* p, double t, const double* lags) {
        double hv[N];
        h(t - lags[0], hv);
        const double x = hv[0];
        const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x10 = x8 * x2;
        du[0] = p[0] * x / (1.0 + x10) - 0.1 * u[0];
    }
    static void h_pre(const double* p, double, double* out, int deriv = 0) { out[0] = deriv == 0 ? p[1] : 0.0; }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// test/regression/allocations.jl:185-202: u' = -u(t - (0.5 + 0.1|u|)), h = 1, one dependent lag.
struct StateDepModel {
    static constexpr int N = 1, NP = 1, NCL = 0, NDL = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double* u, H& h, const double*, double t, const double*) {
        const double tau = 0.5 + 0.1 * std::fabs(u[0]);
        double hv[N];
        h(t - tau, hv);
        du[0] = -hv[0];
    }
    static void h_pre(const double*, double, double* out, int deriv = 0) { out[0] = deriv == 0 ? 1.0 : 0.0; }
    static double dep_lag(int, const double* u, const double*, double) { return 0.5 + 0.1 * std::fabs(u[0]); }
};

// SURVEY §8d C5 stiff delayed kinetics (no reference analogue; Rosenbrock23 workload).
//   u1' = -lam u1 + lam u2(t-tau);  u2' = -u2 + u1(t-tau) u3;  u3' = -mu (u3 - 1) + 0.1 u1 u2
// p = [lam], mu = 50, h = (1,1,1).
struct StiffModel {
    static constexpr int N = 3, NP = 1, NCL = 1, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = true;
    template <class H>
    static void f(double* du, const double* u, H& h, const double* p, double t, const double* lags) {
        const double lam = p[0], mu = 50.0;
        double hv[N];
        h(t - lags[0], hv);
        du[0] = -lam * u[0] + lam * hv[1];
        du[1] = -u[1] + hv[0] * u[2];
        du[2] = -mu * (u[2] - 1.0) + 0.1 * u[0] * u[1];
    }
    static void h_pre(const double*, double, double* out, int deriv = 0) {
        for (int i = 0; i < N; ++i) out[i] = deriv == 0 ? 1.0 : 0.0;
    }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    // analytic Jacobian (history held fixed) and time gradient (through d/dt h(t - tau))
    template <class H>
    static void jac(double* J /*row-major N*N*/, const double* u, H& h, const double* p, double t, const double* lags) {
        const double lam = p[0], mu = 50.0;
        double hv[N];
        h(t - lags[0], hv);
        J[0] = -lam; J[1] = 0.0; J[2] = 0.0;
        J[3] = 0.0; J[4] = -1.0; J[5] = hv[0];
        J[6] = 0.1 * u[1]; J[7] = 0.1 * u[0]; J[8] = -mu;
    }
    template <class H>
    static void tgrad(double* dT, const double* u, H& h, const double* p, double t, const double* lags) {
        const double lam = p[0];
        double dh[N];
        h.deriv(t - lags[0], dh);
        dT[0] = lam * dh[1];
        dT[1] = dh[0] * u[2];
        dT[2] = 0.0;
    }
};

// test/interface/parameters.jl:7-22: delayed logistic u' = p1 u (1 - u(t-1)), h = 0.1,
// order_discontinuity_t0 = 1.
struct LogisticModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, ORDER0 = 1;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double* u, H& h, const double* p, double t, const double* lags) {
        double hv[N];
        h(t - lags[0], hv);
        du[0] = p[0] * u[0] * (1.0 - hv[0]);
    }
    static void h_pre(const double*, double, double* out, int deriv = 0) { out[0] = deriv == 0 ? 0.1 : 0.0; }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// upstream prob_dde_constant_1delay_ip: u' = -u(t-1), h = 0, u0 = 1, tspan (0,10)
// (used by test/interface/discontinuities.jl:85-96, test/integrators/retcode.jl).
struct OneDelayModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = true;
    template <class H>
    static void f(double* du, const double*, H& h, const double*, double t, const double* lags) {
        double hv[N];
        h(t - lags[0], hv);
        du[0] = -hv[0];
    }
    static void h_pre(const double*, double, double* out, int = 0) { out[0] = 0.0; }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    template <class H>
    static void jac(double* J, const double*, H&, const double*, double, const double*) { J[0] = 0.0; }
    template <class H>
    static void tgrad(double* dT, const double*, H& h, const double*, double t, const double* lags) {
        double dh[N];
        h.deriv(t - lags[0], dh);
        dT[0] = -dh[0];
    }
};

// upstream prob_dde_constant_2delays_ip: u' = -u(t-1/3) - u(t-1/5), h = 0, u0 = 1
// (test/interface/discontinuities.jl:5,28-60; the _long variant in test/interface/fpsolve.jl:5).
struct TwoDelaysModel {
    static constexpr int N = 1, NP = 1, NCL = 2, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double*, H& h, const double*, double t, const double* lags) {
        double h1[N], h2[N];
        h(t - lags[0], h1);
        h(t - lags[1], h2);
        du[0] = -h1[0] - h2[0];
    }
    static void h_pre(const double*, double, double* out, int = 0) { out[0] = 0.0; }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// upstream prob_dde_constant_1delay_long_ip: u' = u - u(t-1/5), h = 0, u0 = 1, tspan (0,100)
// (test/integrators/unique_times.jl:4, test/interface/saveat.jl:4).
struct OneDelayLongModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double* u, H& h, const double*, double t, const double* lags) {
        double hv[N];
        h(t - lags[0], hv);
        du[0] = u[0] - hv[0];
    }
    static void h_pre(const double*, double, double* out, int = 0) { out[0] = 0.0; }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// test/interface/dependent_delays.jl:18: the 1-delay problem with the lag given as the
// dependent lag (u,p,t) -> 1 instead of constant_lags = [1].
struct OneDelayDepModel {
    static constexpr int N = 1, NP = 1, NCL = 0, NDL = 1, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double*, H& h, const double*, double t, const double*) {
        double hv[N];
        h(t - 1.0, hv);
        du[0] = -hv[0];
    }
    static void h_pre(const double*, double, double* out, int = 0) { out[0] = 0.0; }
    static double dep_lag(int, const double*, const double*, double) { return 1.0; }
};

// test/interface/backwards.jl:4-12: f(u,h,p,t) = h(p,t+1), h = 1, tspan (2,0), DDEProblem(f,h,tspan)
// form => u0 = h(p,t0) = 1 and order_discontinuity_t0 = 1; constant_lags = [-1].
struct BackwardsModel {
    static constexpr int N = 1, NP = 1, NCL = 1, NDL = 0, ORDER0 = 1;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double*, H& h, const double*, double t, const double* lags) {
        double hv[N];
        h(t - lags[0], hv);
        du[0] = hv[0];
    }
    static void h_pre(const double*, double, double* out, int deriv = 0) { out[0] = deriv == 0 ? 1.0 : 0.0; }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// test/interface/mass_matrix.jl:7-19: SIR with delayed recovery, S' = -g I S, I' = g I S - g I(t-tau) S(t-tau),
// R' = g I(t-tau) S(t-tau); p = [gamma], constant_lags = [tau] (4 in the test), h = (1, 0, 0).
struct SirModel {
    static constexpr int N = 3, NP = 1, NCL = 1, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = false;
    template <class H>
    static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        const double g = p[0];
        double hv[N];
        h(t - lags[0], hv);
        const double infection = g * u[1] * u[0];
        const double recovery = g * hv[1] * hv[0];
        du[0] = -infection;
        du[1] = infection - recovery;
        du[2] = recovery;
    }
    static void h_pre(const double*, double, double* out, int deriv = 0) {
        out[0] = deriv == 0 ? 1.0 : 0.0;
        out[1] = 0.0;
        out[2] = 0.0;
    }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
};

// test/interface/mass_matrix.jl:33-58: the same system as a delay differential-algebraic problem, third equation
// replaced by the conservation law 0 = S + I + R - 1, mass_matrix = Diagonal([1, 1, 0]) (singular: isdae, neutral).
struct SirDdaeModel {
    static constexpr int N = 3, NP = 1, NCL = 1, NDL = 0, ORDER0 = 0;
    static constexpr bool HAS_JAC = true;
    static constexpr bool HAS_MASS = true;
    static constexpr double mass(int i) { return i < 2 ? 1.0 : 0.0; }
    template <class H>
    static void f(double* du, const double* u, H& h, const double* p, double t,
                                             const double* lags) {
        const double g = p[0];
        double hv[N];
        h(t - lags[0], hv);
        const double infection = g * u[1] * u[0];
        const double recovery = g * hv[1] * hv[0];
        du[0] = -infection;
        du[1] = infection - recovery;
        du[2] = u[0] + u[1] + u[2] - 1.0;
    }
    static void h_pre(const double*, double, double* out, int deriv = 0) {
        out[0] = deriv == 0 ? 1.0 : 0.0;
        out[1] = 0.0;
        out[2] = 0.0;
    }
    static double dep_lag(int, const double*, const double*, double) { return 0.0; }
    template <class H>
    static void jac(double* J, const double* u, H&, const double* p, double, const double*) {
        const double g = p[0];
        J[0] = -(g * u[1]); J[1] = -(g * u[0]); J[2] = 0.0;
        J[3] = g * u[1]; J[4] = g * u[0]; J[5] = 0.0;
        J[6] = 1.0; J[7] = 1.0; J[8] = 1.0;
    }
    template <class H>
    static void tgrad(double* dT, const double*, H& h, const double* p, double t,
                                                 const double* lags) {
        const double g = p[0];
        double hv[N], dh[N];
        h(t - lags[0], hv);
        h.deriv(t - lags[0], dh);
        dT[0] = 0.0;
        dT[1] = -(g * (dh[1] * hv[0] + hv[1] * dh[0]));
        dT[2] = 0.0;
    }
};

// Mass matrix of a model: diagonal, compile-time (`HAS_MASS` + `mass(i)`); models without the members have M = I.
template <class M, class = void>
struct MassOf {
    static constexpr bool HAS = false;
    static constexpr double at(int) { return 1.0; }
    static constexpr bool singular() { return false; }
};
template <class M>
struct MassOf<M, decltype((void)M::HAS_MASS)> {
    static constexpr bool HAS = M::HAS_MASS;
    static constexpr double at(int i) { return M::mass(i); }
    static constexpr bool singular() {
        for (int i = 0; i < M::N; ++i)
            if (M::mass(i) == 0.0) return true;
        return false;
    }
};

}  // namespace oracle
#endif
