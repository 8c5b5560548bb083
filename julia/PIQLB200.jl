# PIQLB200.jl -- Julia side of the drop-in boundary (UNEXECUTED here: this image has no julia).
#
# `using PIQL, PIQLB200` keeps PIQL's public API and problem interface and re-routes the hot
# path (generate_ν, generate_logz, logz_updateV, generate_optimality_gap, plus mc_logz) to
# libpiql_b200.so through `ccall`.  Every call mirrors one entry point of include/piql_b200.h;
# the Python binding in piql.jl_b200/_capi.py exercises exactly the same ABI in the tests.
module PIQLB200

using PIQL
import PIQL: Growable, ControlProblem
# extended below with methods on ControlProblem, so that the reference's own call sites -- generate_ν(ctrl),
# generate_logz(ctrl, V; β), logz_updateV(ctrl, V; β), generate_optimality_gap(ctrl, V; ν) in scripts/experiments.jl:10-39,
# scripts/distributed_experiments.jl:5-33 and test/runtests.jl:10-20 -- run on the GPU unchanged after `using PIQL, PIQLB200`
import PIQL: generate_ν, generate_logz, logz_updateV, generate_z, z_updateV, generate_optimality_gap

const LIB = get(ENV, "PIQL_B200_LIB", joinpath(@__DIR__, "..", "piql.jl_b200", "lib", "libpiql_b200.so"))

struct GridParams          # piql_grid_params
    randomness::Cdouble
    step_cost::Cdouble
    reward::Cdouble
    gamma::Cdouble
end
struct SolveOpts           # piql_solve_opts
    tol::Cdouble
    max_sweeps::Int64
    precision::Int32       # 0 = Float64, 1 = Float32, 2 = Float32 with MUFU ex2/lg2 (tiled kernel)
    check_every::Int32
end
struct SolveStats          # piql_solve_stats (immutable and isbits: a Vector{SolveStats} is the C array the batched
    sweeps::Int64          # entry points fill, a Ref{SolveStats} the single struct)
    residual::Cdouble
    device_seconds::Cdouble
    backups::Int64
    converged::Int32
    kernel::Int32
end
SolveStats() = SolveStats(0, 0.0, 0.0, 0, 0, 0)
struct McOpts              # piql_mc_opts
    npaths::Int64
    path0::Int64
    seed::UInt64
    estimator::Int32
    start_mode::Int32
    horizon::Int64
    starts::Ptr{Int32}
    nstarts::Int64
    injected::Ptr{Cfloat}
    injected_steps::Int64
end
struct McStats             # piql_mc_stats
    transitions::UInt64
    paths::UInt64
    device_seconds::Cdouble
    rollout_seconds::Cdouble
end
McStats() = McStats(0, 0, 0.0, 0.0)

mutable struct Context
    ptr::Ptr{Cvoid}
    walls::Array{Bool,2}
    key::Any        # what is loaded on the device: the tabulated problem's identity (context_for)
    states::Any     # tabulated (non-grid) problems: the discovered states, in table order
    function Context(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:piql_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), ref, device)
        rc == 0 || error("piql_create failed with status $rc (no CPU fallback)")
        ctx = new(ref[], Array{Bool}(undef, 0, 0), nothing, nothing)
        finalizer(c -> ccall((:piql_destroy, LIB), Int32, (Ptr{Cvoid},), c.ptr), ctx)
    end
end

function check(ctx::Context, rc::Int32; allow = ())
    (rc == 0 || rc in allow) && return rc
    msg = unsafe_string(ccall((:piql_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.ptr))
    error("piql_b200 error $rc: $msg")
end

# ---- tabulation: ControlProblem closures -> the flat arguments of piql_grid_load -------------------
# gridworld_ctrl captures `walls`, `gw.goal`, `randomness`, `step_cost`, `reward` in its closures
# (src/Problems/GridWorld.jl:96-121); the wall array and the goal are read from the captured
# Gridworld, the rewards by probing ctrl.reward_function so that modify(ctrl, reward_function = ...)
# (test/runtests.jl:7-8, scripts/experiments.jl:26) is honoured.
function tabulate(ctrl::ControlProblem, gw; randomness)
    goal = gw.goal
    other = first(ii for ii in CartesianIndices(gw.walls) if !gw.walls[ii] && ii != goal)
    reward = ctrl.reward_function(other, 0, goal)
    step_cost = -ctrl.reward_function(other, 0, other)
    return (walls = gw.walls, goal = goal, params = GridParams(randomness, step_cost, reward, ctrl.γ))
end

function load!(ctx::Context, ctrl::ControlProblem, gw; randomness)
    t = tabulate(ctrl, gw; randomness)
    walls8 = UInt8.(t.walls)                       # column-major, border included
    H, W = size(walls8)
    GC.@preserve walls8 check(ctx, ccall((:piql_grid_load, LIB), Int32,
        (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Int64, Ref{GridParams}),
        ctx.ptr, walls8, H, W, t.goal[1] - 1, t.goal[2] - 1, t.params))   # goal is 0-based in the ABI
    ctx.walls = t.walls
    return ctx
end

dense(ctx::Context, V) = [ctx.walls[ii] ? NaN : V(ii) for ii in CartesianIndices(ctx.walls)]
function growable(ctx::Context, A::Matrix{Float64}; default = x -> 0.0)
    Growable(Dict(ii => A[ii] for ii in CartesianIndices(ctx.walls) if !ctx.walls[ii]), default)
end
opts(; tol = 1e-12, max_sweeps = 10^7, precision = 0, check_every = 0) = SolveOpts(tol, max_sweeps, precision, check_every)

# ---- PIQL's solver entry points (src/value_iteration.jl) -----------------------------------------------
function generate_ν(ctx::Context; reward_scale = 1.0, kw...)
    out = Matrix{Float64}(undef, size(ctx.walls)); st = Ref(SolveStats())
    GC.@preserve out check(ctx, ccall((:piql_solve_nu, LIB), Int32,
        (Ptr{Cvoid}, Cdouble, Ref{SolveOpts}, Ptr{Cdouble}, Ref{SolveStats}), ctx.ptr, reward_scale, opts(; kw...), out, st))
    growable(ctx, out)
end
function generate_logz(ctx::Context, V; β = 1.0, kw...)
    Vd = dense(ctx, V); out = similar(Vd); st = Ref(SolveStats())
    GC.@preserve Vd out check(ctx, ccall((:piql_solve_logz, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ref{SolveOpts}, Ptr{Cdouble}, Ref{SolveStats}),
        ctx.ptr, Vd, C_NULL, β, opts(; kw...), out, st))
    growable(ctx, out)
end
function logz_updateV(ctx::Context, V; β = 1.0, kw...)
    Vd = dense(ctx, V); out = similar(Vd); st = Ref(SolveStats())
    GC.@preserve Vd out check(ctx, ccall((:piql_update_v, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ref{SolveOpts}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{SolveStats}),
        ctx.ptr, Vd, β, opts(; kw...), out, C_NULL, st))
    growable(ctx, out; default = V.default)
end
# the linear-space twins (src/value_iteration.jl:42-59); the reference's stop rule for these is 1e-9
function generate_z(ctx::Context, V; β = 1.0, tol = 1e-9, kw...)
    Vd = dense(ctx, V); out = similar(Vd); st = Ref(SolveStats())
    GC.@preserve Vd out check(ctx, ccall((:piql_solve_z, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ref{SolveOpts}, Ptr{Cdouble}, Ref{SolveStats}),
        ctx.ptr, Vd, C_NULL, β, opts(; tol, kw...), out, st))
    growable(ctx, out; default = x -> 1.0)
end
function z_updateV(ctx::Context, V; β = 1.0, tol = 1e-9, kw...)
    Vd = dense(ctx, V); out = similar(Vd); st = Ref(SolveStats())
    GC.@preserve Vd out check(ctx, ccall((:piql_z_update_v, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ref{SolveOpts}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{SolveStats}),
        ctx.ptr, Vd, β, opts(; tol, kw...), out, C_NULL, st))
    growable(ctx, out; default = V.default)
end
function generate_optimality_gap(ctx::Context, V; ν, kw...)
    Vd = dense(ctx, V); nd = dense(ctx, ν); out = similar(Vd); st = Ref(SolveStats())
    GC.@preserve Vd nd out check(ctx, ccall((:piql_solve_gap, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{SolveOpts}, Ptr{Cdouble}, Ref{SolveStats}),
        ctx.ptr, Vd, C_NULL, nd, opts(; kw...), out, st))
    growable(ctx, out)
end

# ---- the reference's own signatures: methods on ControlProblem (src/value_iteration.jl:42,54,77,107,119,144) ----------
# A ControlProblem is a struct of closures.  gridworld_ctrl's closures capture everything piql_grid_load needs
# (src/Problems/GridWorld.jl:96-121: `pa` captures walls and randomness, `s -> s == gw.goal` captures gw), and Julia exposes
# captured variables as fields of the closure object -- the same fields JLD2 serialised into data/*.jld2.  So the problem is
# recognised and tabulated from `ctrl` alone, a context per device is kept in a registry, and the grid is (re)loaded only
# when the tabulated problem changes (content hash of the wall array, goal, the four parameters): the mirror of
# solvers.engine_for on the Python side.  A reward_function replaced through modify() (scripts/experiments.jl:26,
# test/runtests.jl:7-8) keeps the captured gridworld and is probed by tabulate.  Any other ControlProblem (the bandit)
# goes through the sparse-table entry points.
const CONTEXTS = Dict{Int,Context}()
const DEVICE = Ref(0)              # PIQLB200.DEVICE[] = k selects the GPU the ControlProblem methods use

function gridworld_of(ctrl::ControlProblem)
    tc, pa = ctrl.terminal_condition, ctrl.propagator_average
    (hasfield(typeof(tc), :gw) && hasfield(typeof(pa), :walls) && hasfield(typeof(pa), :randomness)) || return nothing
    gw = getfield(tc, :gw)
    getfield(pa, :walls) === gw.walls || return nothing
    return (gw, Float64(getfield(pa, :randomness)))
end

function context_for(ctrl::ControlProblem)
    ctx = get!(() -> Context(DEVICE[]), CONTEXTS, DEVICE[])
    g = gridworld_of(ctrl)
    if g === nothing                  # not the implicit gridworld: tabulate through the closures (bandit, user problems)
        key = (:table, objectid(ctrl))
        if ctx.key != key
            ctx.states = load_table!(ctx, ctrl)
            ctx.key = key
        end
        return ctx
    end
    gw, randomness = g
    t = tabulate(ctrl, gw; randomness)
    key = (:grid, hash(gw.walls), size(gw.walls), Tuple(t.goal), t.params)
    if ctx.key != key
        load!(ctx, ctrl, gw; randomness)
        ctx.key = key; ctx.states = nothing
    end
    return ctx
end
on_table(ctx::Context) = ctx.states !== nothing

function generate_ν(ctrl::ControlProblem; kw...)
    ctx = context_for(ctrl)
    on_table(ctx) ? table_solve(ctx, ctx.states, 0; kw...) : generate_ν(ctx; kw...)
end
function generate_logz(ctrl::ControlProblem, V; β = 1.0, kw...)
    ctx = context_for(ctrl)
    on_table(ctx) ? table_solve(ctx, ctx.states, 1; V, β, kw...) : generate_logz(ctx, V; β, kw...)
end
function logz_updateV(ctrl::ControlProblem, V; β = 1.0, kw...)
    ctx = context_for(ctrl)
    on_table(ctx) ? table_logz_updateV(ctx, ctx.states, V; β, kw...) : logz_updateV(ctx, V; β, kw...)
end
function generate_z(ctrl::ControlProblem, V; β = 1.0, kw...)
    ctx = context_for(ctrl)
    on_table(ctx) && error("generate_z: the linear-space cross-check exists for gridworld problems only")
    generate_z(ctx, V; β, kw...)
end
function z_updateV(ctrl::ControlProblem, V; β = 1.0, kw...)
    ctx = context_for(ctrl)
    on_table(ctx) && error("z_updateV: the linear-space cross-check exists for gridworld problems only")
    z_updateV(ctx, V; β, kw...)
end
function generate_optimality_gap(ctrl::ControlProblem, V; ν = generate_ν(ctrl), kw...)
    ctx = context_for(ctrl)
    on_table(ctx) ? table_solve(ctx, ctx.states, 2; V, ν, kw...) : generate_optimality_gap(ctx, V; ν, kw...)
end
# new entry points of the hot path with the same calling convention
mc_logz(ctrl::ControlProblem, V; kw...) = mc_logz(context_for(ctrl), V; kw...)
mc_logz_updateV(ctrl::ControlProblem, V; kw...) = mc_logz_updateV(context_for(ctrl), V; kw...)
path_estimates(ctrl::ControlProblem, V; kw...) = path_estimates(context_for(ctrl), V; kw...)

# run_experiment(gw; ...) exactly as scripts/experiments.jl:20-39 spells it, batched on the device
run_experiment(gw; kw...) = run_experiment(get!(() -> Context(DEVICE[]), CONTEXTS, DEVICE[]), gw; kw...)

# ---- Monte-Carlo estimate of generate_logz's fixed point (replaces src/path_functions.jl:12-41) ----
function mc_logz(ctx::Context, V; β = 1.0, npaths = 2^20, seed = 0, estimator = 0, horizon = 2^20,
                 starts::Vector{CartesianIndex{2}} = CartesianIndex{2}[])
    H = size(ctx.walls, 1)
    cells = Int32[(s[1] - 1) + H * (s[2] - 1) for s in starts]
    Vd = dense(ctx, V)
    logz_sa = Array{Float64}(undef, 4, size(ctx.walls)...)      # Q-like layout: action fastest
    count = Array{UInt64}(undef, 4, size(ctx.walls)...)
    logz_s = similar(Vd); st = Ref(McStats())
    GC.@preserve cells Vd logz_sa count logz_s begin
        o = McOpts(npaths, 0, seed, estimator, isempty(cells) ? 0 : 1, horizon,
                   isempty(cells) ? C_NULL : pointer(cells), length(cells), C_NULL, 0)
        check(ctx, ccall((:piql_mc_logz, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ref{McOpts}, Ptr{Cdouble}, Ptr{UInt64}, Ptr{Cdouble}, Ref{McStats}),
            ctx.ptr, Vd, C_NULL, β, o, logz_sa, count, logz_s, st))
    end
    (logz = growable(ctx, logz_s), logz_sa = logz_sa, count = count, transitions = st[].transitions)
end

# ---- fitness_reward_estimate / direct_reward_estimate (src/path_functions.jl:34-50) for npaths paths at once --------
# returns (fitness, direct): nsteps x npaths Float32 matrices of running totals (column p = path p), Q = the V-determined
# Q-function; s0 a CartesianIndex, a vector of them, or nothing for every free cell
function path_estimates(ctx::Context, V; s0 = nothing, npaths = 2^16, nsteps = 20, seed = 0)
    Vd = dense(ctx, V); H = size(ctx.walls, 1)
    cells = s0 === nothing ? Int32[] : Int32[(c[1] - 1) + H * (c[2] - 1) for c in (s0 isa CartesianIndex ? [s0] : s0)]
    fit = Matrix{Float32}(undef, nsteps, npaths); dir = similar(fit)
    GC.@preserve cells Vd fit dir begin
        o = McOpts(npaths, 0, seed, 0, isempty(cells) ? 0 : 1, nsteps, isempty(cells) ? C_NULL : pointer(cells), length(cells), C_NULL, 0)
        check(ctx, ccall((:piql_path_estimates, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{McOpts}, Int32, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Int32}, Ptr{Cfloat}, Ptr{Cfloat}),
            ctx.ptr, Vd, C_NULL, o, nsteps, fit, dir, C_NULL, C_NULL, C_NULL))
    end
    (fitness = fit, direct = dir)
end

# ---- the PIQL iteration itself: V <- V + log Z^/beta with log Z^ from rollouts, niter times on the device ----------------
function mc_logz_updateV(ctx::Context, V; β = 1.0, npaths = 2^22, niter = 1, seed = 0, horizon = 2^20)
    Vd = dense(ctx, V); out = similar(Vd); st = Ref(McStats())
    o = McOpts(npaths, 0, seed, 0, 0, horizon, C_NULL, 0, C_NULL, 0)
    GC.@preserve Vd out check(ctx, ccall((:piql_mc_update_v, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ref{McOpts}, Int32, Ptr{Cdouble}, Ref{McStats}), ctx.ptr, Vd, β, o, niter, out, st))
    growable(ctx, out; default = V.default)
end

# ---- the experiment protocol in two ccalls (scripts/experiments.jl:20-39; replaces the pmap over beta and over
# ---- iterates of scripts/distributed_experiments.jl:26,28) ---------------------------------------------------
function learning_traj_batch(ctx::Context, V0, βrange; n = 5, kw...)
    Vd = dense(ctx, V0); βs = collect(Float64, βrange); H, W = size(ctx.walls)
    traj = Array{Float64}(undef, H, W, n + 1, length(βs))          # column-major: [cell, k, beta] as the ABI lays it out
    st = Vector{SolveStats}(undef, n * length(βs))
    GC.@preserve Vd βs traj st check(ctx, ccall((:piql_learning_traj, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Int32, Ref{SolveOpts}, Ptr{Cdouble}, Ptr{SolveStats}),
        ctx.ptr, Vd, βs, length(βs), n, opts(; kw...), traj, st))
    Dict(β => [growable(ctx, traj[:, :, k, b]; default = V0.default) for k in 1:n+1] for (b, β) in enumerate(βs))
end
function optimality_gap_batch(ctx::Context, Vs::Vector, ν; kw...)
    H, W = size(ctx.walls); m = length(Vs)
    Vd = Array{Float64}(undef, H, W, m)
    for (i, V) in enumerate(Vs); Vd[:, :, i] = dense(ctx, V); end
    nd = dense(ctx, ν); out = similar(Vd); st = Vector{SolveStats}(undef, m)
    GC.@preserve Vd nd out st check(ctx, ccall((:piql_solve_gap_batch, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int32, Ptr{Cdouble}, Ref{SolveOpts}, Ptr{Cdouble}, Ptr{SolveStats}),
        ctx.ptr, Vd, m, nd, opts(; kw...), out, st))
    [growable(ctx, out[:, :, i]) for i in 1:m]
end

# ---- run_experiment[_distributed] (scripts/experiments.jl:20-39, scripts/distributed_experiments.jl:15-33) ------------
# Returns the same NamedTuple of PIQL.Growable's the reference builds, so that `@save "data/<name>.jld2" out`
# (scripts/distributed_experiments.jl:49-59) writes files of the layout data/*.jld2 have, and scripts/plotting.jl
# (`out.traj`, `out.gap`, `out.s0`, `out.ν`, `out.randomness`; :46-89) reads GPU-produced results unchanged.
function run_experiment(ctx::Context, gw; randomness = 0.01, γ = 0.99, starting_temp = 0.1, βrange = exp.(range(-6, 0, 8)), n = 5, kw...)
    gwctrl = gridworld_ctrl(gw; randomness, γ)
    load!(ctx, gwctrl, gw; randomness)
    ctx.key = nothing; ctx.states = nothing      # loaded outside the registry: the next context_for reloads
    ν = generate_ν(ctx; kw...)
    (_, furthest) = findmin(ν.dict)
    V_lo = generate_ν(ctx; reward_scale = starting_temp, kw...)        # modify(gwctrl, reward_function = starting_temp * ...)
    experiment = learning_traj_batch(ctx, V_lo, βrange; n, kw...)      # Dict(β => [V_lo, V_1, ..., V_n])
    βs = collect(keys(experiment))
    flat = reduce(vcat, [experiment[β] for β in βs])
    gaps = optimality_gap_batch(ctx, flat, ν; kw...)
    optimality_gap = Dict(β => gaps[(b-1)*(n+1)+1:b*(n+1)] for (b, β) in enumerate(βs))
    return (ν = ν, traj = experiment, gap = optimality_gap, randomness = randomness, γ = γ, ctrl = gwctrl, s0 = furthest)
end
const run_experiment_distributed = run_experiment

# ---- any other ControlProblem (e.g. binary_bandit_problem, src/Problems/BinaryBandit.jl): tabulate through the closures ----
struct TableProblem
    nstates::Int64; nactions::Int32; row_ptr::Ptr{Int64}; succ::Ptr{Int32}; prob::Ptr{Cdouble}; reward::Ptr{Cdouble}
    log_prior::Ptr{Cdouble}; terminal::Ptr{UInt8}; gamma::Cdouble
end
# successors of (s, a) from propagator_average alone: one recording probe, then one indicator probe per successor
function successors(ctrl, s, a)
    seen = Any[]; ctrl.propagator_average(s, a, s1 -> (s1 in seen || push!(seen, s1); 0.0))
    [(s1, ctrl.propagator_average(s, a, x -> x == s1 ? 1.0 : 0.0)) for s1 in seen]
end
function load_table!(ctx::Context, ctrl::ControlProblem; roots = [ctrl.initial_state()])
    states = collect(roots); index = Dict(s => i - 1 for (i, s) in enumerate(states)); acts = collect(ctrl.action_space)
    ptr = Int64[0]; succ = Int32[]; prob = Float64[]; rew = Float64[]; lp = Float64[]; term = UInt8[]
    i = 1
    while i <= length(states)
        s = states[i]; t = ctrl.terminal_condition(s); push!(term, t)
        for a in acts
            if !t
                for (s1, p) in successors(ctrl, s, a)
                    haskey(index, s1) || (push!(states, s1); index[s1] = length(states) - 1)
                    push!(succ, index[s1]); push!(prob, p); push!(rew, ctrl.reward_function(s, a, s1))
                end
            end
            push!(ptr, length(succ)); push!(lp, log(ctrl.action_prior(s, a)))
        end
        i += 1
    end
    GC.@preserve ptr succ prob rew lp term check(ctx, ccall((:piql_table_load, LIB), Int32, (Ptr{Cvoid}, Ref{TableProblem}), ctx.ptr,
        TableProblem(length(states), length(acts), pointer(ptr), pointer(succ), pointer(prob), pointer(rew), pointer(lp), pointer(term), ctrl.γ)))
    states
end
function table_solve(ctx::Context, states, kind; V = nothing, ν = nothing, β = 1.0, reward_scale = 1.0, kw...)
    Vd = V === nothing ? C_NULL : Float64[V(s) for s in states]; nd = ν === nothing ? C_NULL : Float64[ν(s) for s in states]
    out = Vector{Float64}(undef, length(states)); st = Ref(SolveStats())
    GC.@preserve Vd nd out check(ctx, ccall((:piql_table_solve, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ref{SolveOpts}, Ptr{Cdouble}, Ref{SolveStats}),
        ctx.ptr, kind, Vd, nd, β, reward_scale, opts(; kw...), out, st))
    Growable(Dict(s => out[i] for (i, s) in enumerate(states)), x -> 0.0)
end

# a tabulated Q for the table solves / sampler that follow (any Q callable of src/PIQL.jl:72-89); nothing: the V-determined Q
function table_set_q!(ctx::Context, states, actions, Q)
    if Q === nothing
        return check(ctx, ccall((:piql_table_set_q, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, C_NULL))
    end
    Qd = Float64[Q(s, a) for a in actions, s in states]          # action fastest, as the ABI lays it out
    GC.@preserve Qd check(ctx, ccall((:piql_table_set_q, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, Qd))
end

# logz_updateV / Monte-Carlo log Z on the tabulated problem (states as returned by load_table!)
function table_logz_updateV(ctx::Context, states, V; β = 1.0, kw...)
    Vd = Float64[V(s) for s in states]; out = similar(Vd); st = Ref(SolveStats())
    GC.@preserve Vd out check(ctx, ccall((:piql_table_update_v, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ref{SolveOpts}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{SolveStats}),
        ctx.ptr, Vd, β, opts(; kw...), out, C_NULL, st))
    Growable(Dict(s => out[i] for (i, s) in enumerate(states)), V.default)
end
function table_mc_logz(ctx::Context, states, nactions, V; β = 1.0, npaths = 2^20, seed = 0, horizon = 2^20, roots = Int[])
    Vd = Float64[V(s) for s in states]; ns = length(states); cells = Int32.(roots)           # roots: 0-based state indices
    logz_sa = Matrix{Float64}(undef, nactions, ns); count = Matrix{UInt64}(undef, nactions, ns); logz_s = Vector{Float64}(undef, ns)
    st = Ref(McStats())
    GC.@preserve cells Vd logz_sa count logz_s begin
        o = McOpts(npaths, 0, seed, 0, isempty(cells) ? 0 : 1, horizon, isempty(cells) ? C_NULL : pointer(cells), length(cells), C_NULL, 0)
        check(ctx, ccall((:piql_table_mc_logz, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ref{McOpts}, Ptr{Cdouble}, Ptr{UInt64}, Ptr{Cdouble}, Ref{McStats}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}),
            ctx.ptr, Vd, β, o, logz_sa, count, logz_s, st, C_NULL, C_NULL, C_NULL))
    end
    (logz = Growable(Dict(s => logz_s[i] for (i, s) in enumerate(states)), x -> 0.0), logz_sa = logz_sa, count = count,
     transitions = st[].transitions)
end

# ---- sharded rollouts: one worker process per GPU rolls out its block of path ids [path0, path0 + npaths) and keeps the
# ---- statistics on the device until they are exchanged (next section) and finalised -----------------------------------
function policy_table!(ctx::Context, V; β = 1.0)
    Vd = dense(ctx, V)
    GC.@preserve Vd check(ctx, ccall((:piql_policy_table, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cfloat}),
        ctx.ptr, Vd, C_NULL, β, C_NULL))
end
all_free(npaths, path0, seed, horizon) = McOpts(npaths, path0, seed, 0, 0, horizon, C_NULL, 0, C_NULL, 0)
mc_reset!(ctx::Context) = check(ctx, ccall((:piql_mc_reset_dev, LIB), Int32, (Ptr{Cvoid}, Ref{McOpts}), ctx.ptr, all_free(0, 0, 0, 1)))
mc_rollout!(ctx::Context; npaths, path0 = 0, seed = 0, horizon = 2^20) =
    check(ctx, ccall((:piql_mc_rollout_dev, LIB), Int32, (Ptr{Cvoid}, Ref{McOpts}), ctx.ptr, all_free(npaths, path0, seed, horizon)))
function mc_finalize(ctx::Context)
    logz_sa = Array{Float64}(undef, 4, size(ctx.walls)...); count = Array{UInt64}(undef, 4, size(ctx.walls)...)
    logz_s = Matrix{Float64}(undef, size(ctx.walls)); st = Ref(McStats())
    GC.@preserve logz_sa count logz_s check(ctx, ccall((:piql_mc_finalize, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{UInt64}, Ptr{Cdouble}, Ref{McStats}), ctx.ptr, logz_sa, count, logz_s, st))
    (logz = growable(ctx, logz_s), logz_sa = logz_sa, count = count, transitions = st[].transitions)
end

# ---- pipelined iterations: upload / table of iteration k + 1 and download of iteration k overlap the rollouts ----------
# V must stay alive and unchanged until the matching mc_finalize_wait returns (keep a reference; pin it for a truly
# asynchronous copy).  Outputs: Float64 / UInt64 (f32 = false) or Float32 / UInt32 arrays of the shapes of mc_finalize.
function policy_table_async!(ctx::Context, Vd::Matrix{Float64}; β = 1.0)
    check(ctx, ccall((:piql_policy_table_async, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble), ctx.ptr, Vd, C_NULL, β))
end
function mc_finalize_async!(ctx::Context, logz_sa::Array, count::Array, logz_s::Array)
    f32 = eltype(logz_sa) == Float32
    check(ctx, ccall((:piql_mc_finalize_async, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32),
        ctx.ptr, logz_sa, count, logz_s, f32 ? 1 : 0))
end
function mc_finalize_wait(ctx::Context)
    st = Ref(McStats())
    check(ctx, ccall((:piql_mc_finalize_wait, LIB), Int32, (Ptr{Cvoid}, Ref{McStats}), ctx.ptr, st))
    st[]
end

# ---- multi-GPU exchange over NVLink peer memory: one worker process per GPU (replaces addprocs/pmap/@spawn,
# ---- scripts/distributed_experiments.jl:3,26,28,43-46).  The handle blobs travel over Distributed.jl itself.
struct PeerHandles; lse_max::NTuple{64,UInt8}; lse_sum::NTuple{64,UInt8}; count::NTuple{64,UInt8}; flags::NTuple{64,UInt8}; n::Int64; end
function peer_export(ctx::Context)
    h = Ref{PeerHandles}()
    check(ctx, ccall((:piql_peer_export, LIB), Int32, (Ptr{Cvoid}, Ref{PeerHandles}), ctx.ptr, h)); h[]
end
peer_import(ctx::Context, rank, world, blobs::Vector{PeerHandles}) =
    check(ctx, ccall((:piql_peer_import, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{PeerHandles}), ctx.ptr, rank, world, blobs))
mc_allreduce_peer(ctx::Context) = check(ctx, ccall((:piql_mc_allreduce_peer, LIB), Int32, (Ptr{Cvoid},), ctx.ptr))

end # module
