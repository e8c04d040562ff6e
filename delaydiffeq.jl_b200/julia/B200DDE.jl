# B200DDE.jl — the Julia-side binding a DelayDiffEq.jl maintainer would add to use libb200dde.so.
#
# STATUS: NOT EXECUTED in this repository's environment — there is no `julia` binary in the build image or on
# the GPU box (SURVEY.md §0.4).  The executable host is the Python ctypes mirror in ../b200dde/, which makes
# exactly the calls below, one for one; tests/test_abi.py parses every `ccall` signature of this file and checks
# it against the prototypes of include/b200dde.h, and the field list of `B2DConfig` against `b2d_config`.
# This file is the drop-in stub: it adds an ensemble algorithm type and one `__solve` method; nothing in
# DelayDiffEq's own source changes.
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
    callback::Int32; cb_affect::Int32; cb_idx::Int32; cb_save_positions::Int32
    cb_par0::Float64; cb_par1::Float64; cb_par2::Float64; cb_par3::Float64
    cb_cond_idx::Int32; cb_direction::Int32; cb_interp_points::Int32; cb_reserved::Int32
    B2DConfig() = new()
end

const RETCODES = Dict(1 => ReturnCode.Success, 2 => ReturnCode.MaxIters, 3 => ReturnCode.DtLessThanMin,
                      4 => ReturnCode.Unstable, 5 => ReturnCode.DtNaN, 6 => ReturnCode.Failure, 7 => ReturnCode.Failure,
                      8 => ReturnCode.Terminated, 9 => ReturnCode.Failure)   # 9 never leaves b2d_solve (solved again in the general kernel)

"""
    B200Callback(kind; affect = :none, idx = 1, cond_idx = 1, direction = 0, interp_points = 10,
                 save_positions = (true, true), par = (0.0, 0.0, 0.0, 0.0))

The `callback` keyword of `solve` for EnsembleB200.  A Julia closure cannot cross the C ABI, so condition and affect! are
named: `kind` is `:model` / `:at_time` / `:periodic` (DiscreteCallback shapes: the model's device functors, `t == par[1]`,
every `par[1]`) or `:cont_model` / `:cont_time` / `:cont_state` (ContinuousCallback: the model's `cb_value`, `t - par[1]`,
`u[cond_idx] - par[1]`); `affect` is `:none`, `:terminate`, `:add` (`u[idx] += par[2]`) or `:negate`.  Times of `:at_time` /
`:periodic` callbacks go into `tstops` as DiffEqCallbacks does (src/solve.jl:632-681, src/integrators/interface.jl:566-632).
"""
Base.@kwdef struct B200Callback
    kind::Symbol
    affect::Symbol = :none
    idx::Int = 1
    cond_idx::Int = 1
    direction::Int = 0
    interp_points::Int = 10
    save_positions::Tuple{Bool, Bool} = (true, true)
    par::NTuple{4, Float64} = (0.0, 0.0, 0.0, 0.0)
end
B200Callback(kind::Symbol; kwargs...) = B200Callback(; kind, kwargs...)
const CB_KINDS = Dict(:model => 1, :at_time => 2, :periodic => 3, :cont_model => 4, :cont_time => 5, :cont_state => 6)
const CB_AFFECTS = Dict(:none => 0, :terminate => 1, :add => 2, :negate => 3)
function apply_callback!(cfg::B2DConfig, cb::B200Callback)
    cfg.callback = CB_KINDS[cb.kind]; cfg.cb_affect = CB_AFFECTS[cb.affect]
    cfg.cb_idx = cb.idx - 1; cfg.cb_cond_idx = cb.cond_idx - 1
    cfg.cb_direction = cb.direction; cfg.cb_interp_points = cb.interp_points
    cfg.cb_save_positions = Int32(cb.save_positions[1]) | (Int32(cb.save_positions[2]) << 1)
    cfg.cb_par0, cfg.cb_par1, cfg.cb_par2, cfg.cb_par3 = cb.par
    return cfg
end
callback_tstops(cb::B200Callback, t0, tf) =
    cb.kind === :at_time ? [cb.par[1]] :
    cb.kind === :periodic ? collect((t0 + sign(tf - t0) * cb.par[1]):(sign(tf - t0) * cb.par[1]):tf) : Float64[]

"""
    EnsembleB200(; device = 0, devices = [device])

Ensemble algorithm that solves every trajectory of a DDE ensemble on B200s through libb200dde.so.  With several
`devices` the ensemble is sharded in contiguous ranges over one handle per GPU from this one process
(`b2d_solve_multi`): every GPU copies its shard straight into the result array, there is nothing to gather.
"""
struct EnsembleB200 <: EnsembleAlgorithm
    devices::Vector{Int}
end
EnsembleB200(; device = 0, devices = [device]) = EnsembleB200(collect(Int, devices))

"""
    B200Model(id)          a device functor of the library's registry (`B2D_MODEL_*`)
    B200UserModel(source, struct_name)
                           CUDA C++ source of a struct with the interface of csrc/models.cuh (N, NP, NCL, NDL, NH, ORDER0,
                           HAS_JAC, hist_comp, f, h_pre, dep_lag[, jac, tgrad]); compiled with NVRTC when the handle is made

The right-hand side, history and lag functions of a DDEProblem.  A Julia closure cannot cross a C ABI; the user states
which compiled model `prob.f` is by attaching one of these tags: `DDEProblem(DDEFunction(f; sys = B200Model(0)), ...)`.
"""
struct B200Model
    id::Int32
end
struct B200UserModel
    source::String
    struct_name::String
end
b200_model(prob) = prob.f.sys isa Union{B200Model, B200UserModel} ? prob.f.sys : error("DDEProblem carries no B200Model / B200UserModel tag")

alg_id(::Tsit5) = Int32(0)
alg_id(::Rosenbrock23) = Int32(1)
alg_id(::BS3) = Int32(2)
alg_id(alg) = error("EnsembleB200 supports MethodOfSteps(Tsit5()), MethodOfSteps(BS3()) and MethodOfSteps(Rosenbrock23()), got $(typeof(alg))")

check(rc) = rc == 0 || error(unsafe_string(ccall((:b2d_last_error, lib), Cstring, ())))

# ---- handles are kept across solves ------------------------------------------------------------------------------
# One handle per (configuration, model, device).  (b2d_destroy also parks a handle's streams and buffers for the next
# b2d_create, so even a host that does not keep handles pays the allocations once per process.)
const HANDLES = Dict{Any, Ptr{Cvoid}}()
const HANDLE_LOCK = ReentrantLock()

config_key(cfg::B2DConfig) = ntuple(i -> getfield(cfg, i), fieldcount(B2DConfig))

function handle_for(cfg::B2DConfig, model)
    lock(HANDLE_LOCK) do
        get!(HANDLES, (config_key(cfg), model)) do
            h = Ref{Ptr{Cvoid}}(C_NULL)
            if model isa B200UserModel
                check(ccall((:b2d_create_user_model, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{B2DConfig}, Cstring, Cstring),
                    h, cfg, model.source, model.struct_name))
            else
                check(ccall((:b2d_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ref{B2DConfig}), h, cfg))
            end
            h[]
        end
    end
end

"""
    release_handles()

Destroy every cached handle and free what the library keeps parked (`b2d_destroy`, `b2d_release_cached`).
"""
function release_handles()
    lock(HANDLE_LOCK) do
        for h in values(HANDLES)
            ccall((:b2d_destroy, lib), Cint, (Ptr{Cvoid},), h)
        end
        empty!(HANDLES)
        ccall((:b2d_release_cached, lib), Cint, ())
    end
    return nothing
end

# ---- the result: trajectories are built on demand ------------------------------------------------------------------
# An ensemble of 10^6 trajectories with 101 save points would be 10^8 small vectors if every ODESolution were materialised;
# the block the library wrote is kept as it is and `sol[i]` wraps columns of it when asked.
struct LazyTrajectories{P, A, F} <: AbstractVector{Any}
    prob::P
    alg::A
    output_func::F
    t::Vector{Float64}
    u::Array{Float64, 3}          # [n_saved_components, n_out, trajectories], exactly what b2d_solve filled
    retcode::Vector{Int32}
    stats::Matrix{Int64}          # [8, trajectories]
end
Base.size(l::LazyTrajectories) = (size(l.u, 3),)
function Base.getindex(l::LazyTrajectories, i::Int)
    us = [view(l.u, :, j, i) for j in axes(l.u, 2)]
    sol = build_solution(l.prob, l.alg, l.t, us; retcode = RETCODES[Int(l.retcode[i])], dense = false)
    sol.stats.naccept, sol.stats.nreject, sol.stats.nf = l.stats[1, i], l.stats[2, i], l.stats[3, i]
    sol.stats.nfpiter, sol.stats.nfpconvfail = l.stats[4, i], l.stats[5, i]
    return l.output_func(sol, i)[1]
end

# what prob_func may change from trajectory to trajectory: u0 and p.  Everything else of a DDEProblem (tspan, lags,
# history, neutral, ...) is shared by the ensemble on this path (the reference unpacks it per trajectory, src/solve.jl:159);
# a prob_func that changes it is refused instead of silently ignored.
# The saveat block is by far the largest thing that crosses PCIe (2.4 GB for 10^6 bc_model trajectories).  An ordinary Julia
# Array is pageable: the library then stages every wave through its own pinned buffers and copies it out with host threads
# (measured 3.7e8 instead of 7.9e8 .. 8.7e8 accepted steps/s end to end).  So the result array is allocated page-locked by the
# library and handed to Julia as an Array whose finalizer returns it (falls back to a plain Array if the allocation fails).
function pinned_array(dims::Int...)
    ptr = ccall((:b2d_host_alloc, lib), Ptr{Cvoid}, (UInt64,), UInt64(prod(dims) * sizeof(Float64)))
    ptr == C_NULL && return Array{Float64}(undef, dims...)
    a = unsafe_wrap(Array, Ptr{Float64}(ptr), dims; own = false)
    finalizer(_ -> ccall((:b2d_host_free, lib), Cvoid, (Ptr{Cvoid},), ptr), a)
    return a
end

function check_shared_fields(base, pi, i)
    pi.tspan == base.tspan || error("EnsembleB200: prob_func changed tspan for trajectory $i; only u0 and p may vary")
    ncl(x) = x === nothing ? 0 : length(x)
    ncl(pi.constant_lags) == ncl(base.constant_lags) ||
        error("EnsembleB200: prob_func changed the number of constant lags for trajectory $i; only u0, p and the lag values may vary")
    pi.dependent_lags === base.dependent_lags || error("EnsembleB200: prob_func changed dependent_lags for trajectory $i; only u0 and p may vary")
    pi.h === base.h || error("EnsembleB200: prob_func changed the history function for trajectory $i; only u0 and p may vary")
    (pi.neutral == base.neutral && pi.order_discontinuity_t0 == base.order_discontinuity_t0) ||
        error("EnsembleB200: prob_func changed neutral / order_discontinuity_t0 for trajectory $i; only u0 and p may vary")
    pi.f === base.f || error("EnsembleB200: prob_func changed the right-hand side for trajectory $i; only u0 and p may vary")
    return nothing
end

function SciMLBase.__solve(ensembleprob::AbstractEnsembleProblem, alg::MethodOfSteps, ensemblealg::EnsembleB200;
        trajectories, saveat, abstol = nothing, reltol = nothing, dt = 0.0, dtmin = nothing, dtmax = nothing,
        maxiters = 1_000_000, save_start = true, save_end = true, save_idxs = nothing, tstops = (), d_discontinuities = (),
        callback = nothing, kwargs...)
    isempty(kwargs) || error("EnsembleB200: unsupported keyword arguments $(collect(keys(kwargs)))")
    prob = ensembleprob.prob
    model = b200_model(prob)
    t0, tf = prob.tspan
    n, np = length(prob.u0), length(prob.p)

    # prob_func is evaluated on the host, exactly like solve_batch does, to build the input matrices
    u0 = Matrix{Float64}(undef, n, trajectories)
    p = Matrix{Float64}(undef, np, trajectories)
    nlag = prob.constant_lags === nothing ? 0 : length(prob.constant_lags)
    traj_lags = Matrix{Float64}(undef, nlag, trajectories)      # per-trajectory constant lags (src/solve.jl:159)
    for i in 1:trajectories
        pi = ensembleprob.prob_func(deepcopy(prob), i, 1)
        check_shared_fields(prob, pi, i)
        u0[:, i] .= pi.u0
        p[:, i] .= pi.p
        nlag > 0 && (traj_lags[:, i] .= pi.constant_lags)
    end
    lags_vary = nlag > 0 && any(traj_lags .!= collect(Float64, prob.constant_lags))

    cfg = B2DConfig()
    ccall((:b2d_config_default, lib), Cvoid, (Ref{B2DConfig}, Int32), cfg, alg_id(alg.alg))
    cfg.constrained = DelayDiffEq.isconstrained(alg)          # src/alg_utils.jl:17-19
    cfg.fp_max_iter = alg.fpsolve.max_iter                     # src/fpsolve/utils.jl:35
    cfg.fp_kappa = alg.fpsolve.κ
    if alg.fpsolve isa NLAnderson                              # src/fpsolve/utils.jl:19-31
        alg.fpsolve.droptol === nothing || error("EnsembleB200: NLAnderson(droptol = ...) is not supported")
        cfg.fp_alg = 1; cfg.fp_max_history = alg.fpsolve.max_history; cfg.fp_aa_start = alg.fpsolve.aa_start
    end
    cfg.model_id = model isa B200Model ? model.id : Int32(0)
    cfg.order_discontinuity_t0 = prob.order_discontinuity_t0   # src/solve.jl:107
    cfg.neutral = prob.neutral
    if model isa B200Model
        # DDEFunction(f; mass_matrix) is part of the compiled model (test/interface/mass_matrix.jl:46-58); prob.neutral
        # already follows from it in DDEProblem.  The tag and the problem must agree:
        mdiag = ones(Float64, n)
        mkind = ccall((:b2d_model_mass_matrix, lib), Cint, (Int32, Ptr{Float64}, Int32), model.id, mdiag, n)
        mkind < 0 && check(mkind)
        (prob.f.mass_matrix === I ? mkind == 0 : prob.f.mass_matrix == Diagonal(mdiag)) ||
            error("EnsembleB200: mass matrix of the DDEProblem differs from the one compiled into B200Model($(model.id))")
    else
        prob.f.mass_matrix === I || error("EnsembleB200: a B200UserModel carries its mass matrix in its source (HAS_MASS / mass(i))")
    end
    cfg.n_state, cfg.n_param = n, np
    cfg.n_const_lags = prob.constant_lags === nothing ? 0 : length(prob.constant_lags)
    cfg.n_dep_lags = prob.dependent_lags === nothing ? 0 : length(prob.dependent_lags)
    abstol isa Number && (cfg.abstol = abstol)                 # src/utils.jl:91-133 defaults otherwise
    reltol isa Number && (cfg.reltol = reltol)
    cfg.dt0 = dt
    dtmin === nothing || (cfg.dtmin = dtmin)
    dtmax === nothing || (cfg.dtmax = dtmax)
    cfg.maxiters = maxiters
    callback === nothing || apply_callback!(cfg, callback::B200Callback)        # src/solve.jl:82

    # the saveat points the reference would push into its heap (src/solve.jl:262): interior points only
    grid = saveat isa Number ? collect((t0 + saveat):saveat:tf) : collect(Float64, saveat)
    grid = filter(s -> t0 < s < tf || tf < s < t0, grid)
    n_out = save_start + length(grid) + save_end
    lags = prob.constant_lags === nothing ? Float64[] : collect(Float64, prob.constant_lags)
    idxs = save_idxs === nothing ? Int32[] : Int32.(collect(save_idxs) .- 1)   # save_idxs, 0-based (src/solve.jl:47,217-218)
    n_saved = isempty(idxs) ? n : length(idxs)
    abstol_v = abstol isa AbstractArray ? collect(Float64, abstol) : fill(Float64(cfg.abstol), n)
    reltol_v = reltol isa AbstractArray ? collect(Float64, reltol) : fill(Float64(cfg.reltol), n)
    vector_tol = abstol isa AbstractArray || reltol isa AbstractArray                # src/utils.jl:91-133
    ts = collect(Float64, tstops)                                                     # src/solve.jl:45
    callback === nothing || append!(ts, callback_tstops(callback, t0, tf))
    dd_t = Float64[d isa Discontinuity ? d.t : d for d in d_discontinuities]          # src/solve.jl:46
    dd_o = Int32[d isa Discontinuity ? d.order : 0 for d in d_discontinuities]

    out_u = pinned_array(n_saved, n_out, trajectories)   # page-locked: the GPUs copy straight into it (2x the pageable rate)
    out_t = Vector{Float64}(undef, n_out)
    retcode = Vector{Int32}(undef, trajectories)
    stats = Matrix{Int64}(undef, 8, trajectories)

    handles = map(ensemblealg.devices) do dev
        cfg.device = dev
        h = handle_for(cfg, model)
        # per-solve options of the handle: array-valued tolerances, save_idxs, tstops, d_discontinuities (n = 0 clears)
        check(ccall((:b2d_set_tolerances, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32), h, abstol_v, reltol_v, vector_tol ? n : 0))
        check(ccall((:b2d_set_save_idxs, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int32), h, idxs, length(idxs)))
        check(ccall((:b2d_set_tstops, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32), h, ts, length(ts)))
        check(ccall((:b2d_set_d_discontinuities, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Int32), h, dd_t, dd_o, length(dd_t)))
        # a prob_func that remakes constant_lags: one row of lags per trajectory (n = 0 clears what a cached handle still holds)
        check(ccall((:b2d_set_trajectory_lags, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h, traj_lags, lags_vary ? trajectories : 0))
        h
    end

    elapsed = @elapsed GC.@preserve u0 p lags grid out_u out_t retcode stats handles begin
        if length(handles) == 1
            check(ccall((:b2d_solve, lib), Cint,
                (Ptr{Cvoid}, Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32,
                    Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
                handles[1], trajectories, t0, tf, u0, p, lags, grid, length(grid), save_start, save_end, out_u, out_t, retcode, stats))
        else
            check(ccall((:b2d_solve_multi, lib), Cint,
                (Ptr{Ptr{Cvoid}}, Int32, Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32,
                    Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
                handles, length(handles), trajectories, t0, tf, u0, p, lags, grid, length(grid), save_start, save_end, out_u, out_t,
                retcode, stats))
        end
    end

    sols = LazyTrajectories(prob, alg.alg, ensembleprob.output_func, out_t, out_u, retcode, stats)
    return EnsembleSolution(sols, elapsed, all(==(1), retcode))
end

export EnsembleB200, B200Model, B200UserModel, B200Callback, release_handles

end # module
