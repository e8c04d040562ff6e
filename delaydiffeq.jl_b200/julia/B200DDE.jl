# B200DDE.jl — the Julia-side binding a DelayDiffEq.jl maintainer would add to use libb200dde.so.
#
# STATUS: NOT EXECUTED in this repository's environment — there is no `julia` binary in the build image or on
# the GPU box (SURVEY.md §0.4).  The executable host is the Python ctypes mirror in ../b200dde/, which makes
# exactly the calls below, one for one.  This file is the drop-in stub: it adds an ensemble algorithm type and
# one `__solve` method; nothing in DelayDiffEq's own source changes.
#
# What it replaces: SciMLBase.__solve(::AbstractEnsembleProblem, alg, ::EnsembleThreads) (upstream SciMLBase),
# which calls DelayDiffEq's `__solve(prob::AbstractDDEProblem, alg::MethodOfSteps)` (src/solve.jl:1-9) once per
# trajectory.  Here the whole ensemble goes through one `ccall` of `b2d_solve` (include/b200dde.h).
module B200DDE

using DelayDiffEq, SciMLBase, LinearAlgebra
import SciMLBase: __solve, EnsembleAlgorithm, AbstractEnsembleProblem, EnsembleSolution, build_solution, ReturnCode

const lib = get(ENV, "B200DDE_LIB", joinpath(@__DIR__, "..", "lib", "libb200dde.so"))

# ---- mirror of `b2d_config` (include/b200dde.h), field for field ------------------------------------------------
mutable struct B2DConfig
    struct_size::Int32; alg::Int32; constrained::Int32; fp_max_iter::Int32
    fp_kappa::Float64; fp_div_rule::Int32; controller_pow::Int32
    model_id::Int32; order_discontinuity_t0::Int32; neutral::Int32
    n_state::Int32; n_param::Int32; n_const_lags::Int32; n_dep_lags::Int32
    abstol::Float64; reltol::Float64; dt0::Float64; dtmin::Float64; dtmax::Float64; maxiters::Int64
    gamma::Float64; qmin::Float64; qmax::Float64; qsteady_min::Float64; qsteady_max::Float64
    qoldinit::Float64; failfactor::Float64
    disc_interp_points::Int32; track_theta_mode::Int32; disc_abstol::Float64; disc_reltol::Float64
    rb23_dT_mode::Int32; ring_capacity::Int32; device::Int32; chunk_traj::Int32
    fp_alg::Int32; fp_max_history::Int32; fp_aa_start::Int32; fp_anderson_reset::Int32
    B2DConfig() = new()
end

const RETCODES = Dict(1 => ReturnCode.Success, 2 => ReturnCode.MaxIters, 3 => ReturnCode.DtLessThanMin,
                      4 => ReturnCode.Unstable, 5 => ReturnCode.DtNaN, 6 => ReturnCode.Failure, 7 => ReturnCode.Failure)

"""
    EnsembleB200(; device = 0)

Ensemble algorithm that solves every trajectory of a DDE ensemble on one B200 through libb200dde.so.
"""
struct EnsembleB200 <: EnsembleAlgorithm
    device::Int
end
EnsembleB200(; device = 0) = EnsembleB200(device)

"""
    B200Model(id)

The right-hand side, history and lag functions of a DDEProblem as a device functor of the library's registry
(`B2D_MODEL_*`).  A Julia closure cannot cross a C ABI; the user states which compiled model `prob.f` is by
attaching this tag: `DDEProblem(DDEFunction(f; sys = B200Model(0)), ...)` or by `b200_model(prob) = B200Model(0)`.
"""
struct B200Model
    id::Int32
end
b200_model(prob) = prob.f.sys isa B200Model ? prob.f.sys : error("DDEProblem carries no B200Model tag")

alg_id(::Tsit5) = Int32(0)
alg_id(::Rosenbrock23) = Int32(1)
alg_id(::BS3) = Int32(2)
alg_id(alg) = error("EnsembleB200 supports MethodOfSteps(Tsit5()), MethodOfSteps(BS3()) and MethodOfSteps(Rosenbrock23()), got $(typeof(alg))")

check(rc) = rc == 0 || error(unsafe_string(ccall((:b2d_last_error, lib), Cstring, ())))

function SciMLBase.__solve(ensembleprob::AbstractEnsembleProblem, alg::MethodOfSteps, ensemblealg::EnsembleB200;
        trajectories, saveat, abstol = nothing, reltol = nothing, dt = 0.0, dtmin = nothing, dtmax = nothing,
        maxiters = 1_000_000, save_start = true, save_end = true, kwargs...)
    prob = ensembleprob.prob
    model = b200_model(prob)
    t0, tf = prob.tspan
    n, np = length(prob.u0), length(prob.p)

    # prob_func is evaluated on the host, exactly like solve_batch does, to build the input matrices
    u0 = Matrix{Float64}(undef, n, trajectories)
    p = Matrix{Float64}(undef, np, trajectories)
    for i in 1:trajectories
        pi = ensembleprob.prob_func(deepcopy(prob), i, 1)
        u0[:, i] .= pi.u0
        p[:, i] .= pi.p
    end

    cfg = B2DConfig()
    ccall((:b2d_config_default, lib), Cvoid, (Ref{B2DConfig}, Int32), cfg, alg_id(alg.alg))
    cfg.constrained = DelayDiffEq.isconstrained(alg)          # src/alg_utils.jl:17-19
    cfg.fp_max_iter = alg.fpsolve.max_iter                     # src/fpsolve/utils.jl:35
    cfg.fp_kappa = alg.fpsolve.κ
    if alg.fpsolve isa NLAnderson                              # src/fpsolve/utils.jl:19-31
        alg.fpsolve.droptol === nothing || error("EnsembleB200: NLAnderson(droptol = ...) is not supported")
        cfg.fp_alg = 1; cfg.fp_max_history = alg.fpsolve.max_history; cfg.fp_aa_start = alg.fpsolve.aa_start
    end
    cfg.model_id = model.id
    cfg.order_discontinuity_t0 = prob.order_discontinuity_t0   # src/solve.jl:107
    cfg.neutral = prob.neutral
    # DDEFunction(f; mass_matrix) is part of the compiled model (test/interface/mass_matrix.jl:46-58); prob.neutral already
    # follows from it in DDEProblem.  The tag and the problem must agree:
    mdiag = ones(Float64, n)
    mkind = ccall((:b2d_model_mass_matrix, lib), Cint, (Int32, Ptr{Float64}, Int32), model.id, mdiag, n)
    mkind < 0 && check(mkind)
    (prob.f.mass_matrix === I ? mkind == 0 : prob.f.mass_matrix == Diagonal(mdiag)) ||
        error("EnsembleB200: mass matrix of the DDEProblem differs from the one compiled into B200Model($(model.id))")
    cfg.n_state, cfg.n_param = n, np
    cfg.n_const_lags = prob.constant_lags === nothing ? 0 : length(prob.constant_lags)
    cfg.n_dep_lags = prob.dependent_lags === nothing ? 0 : length(prob.dependent_lags)
    abstol === nothing || (cfg.abstol = abstol)                # src/utils.jl:91-133 defaults otherwise
    reltol === nothing || (cfg.reltol = reltol)
    cfg.dt0 = dt
    dtmin === nothing || (cfg.dtmin = dtmin)
    dtmax === nothing || (cfg.dtmax = dtmax)
    cfg.maxiters = maxiters
    cfg.device = ensemblealg.device

    # the saveat points the reference would push into its heap (src/solve.jl:262): interior points only
    grid = saveat isa Number ? collect((t0 + saveat):saveat:tf) : collect(Float64, saveat)
    grid = filter(s -> t0 < s < tf || tf < s < t0, grid)
    n_out = save_start + length(grid) + save_end
    lags = prob.constant_lags === nothing ? Float64[] : collect(Float64, prob.constant_lags)

    out_u = Array{Float64, 3}(undef, n, n_out, trajectories)
    out_t = Vector{Float64}(undef, n_out)
    retcode = Vector{Int32}(undef, trajectories)
    stats = Matrix{Int64}(undef, 8, trajectories)

    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b2d_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{B2DConfig}), h, cfg))
    # array-valued tolerances (src/utils.jl:91-133) and save_idxs (src/solve.jl:47) go through the two setters:
    #   ccall((:b2d_set_tolerances, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32), h[], abstol, reltol, n)
    #   ccall((:b2d_set_save_idxs,  lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int32), h[], Int32.(save_idxs .- 1), length(save_idxs))
    # the tstops / d_discontinuities keywords (src/solve.jl:45-46):
    #   ccall((:b2d_set_tstops, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32), h[], tstops, length(tstops))
    #   ccall((:b2d_set_d_discontinuities, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Int32), h[],
    #         Float64[d.t for d in ds], Int32[d.order for d in ds], length(ds))
    elapsed = @elapsed GC.@preserve u0 p lags grid out_u out_t retcode stats begin
        check(ccall((:b2d_solve, lib), Cint,
            (Ptr{Cvoid}, Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32,
                Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
            h[], trajectories, t0, tf, u0, p, lags, grid, length(grid), save_start, save_end, out_u, out_t, retcode, stats))
    end
    ccall((:b2d_destroy, lib), Cint, (Ptr{Cvoid},), h[])

    sols = map(1:trajectories) do i
        us = [out_u[:, j, i] for j in 1:n_out]
        sol = build_solution(prob, alg.alg, copy(out_t), us; retcode = RETCODES[Int(retcode[i])], dense = false)
        sol.stats.naccept, sol.stats.nreject, sol.stats.nf = stats[1, i], stats[2, i], stats[3, i]
        sol.stats.nfpiter, sol.stats.nfpconvfail = stats[4, i], stats[5, i]
        ensembleprob.output_func(sol, i)[1]
    end
    return EnsembleSolution(sols, elapsed, all(==(1), retcode))
end

export EnsembleB200, B200Model

end # module
