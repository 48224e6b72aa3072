# B200PDE.jl -- reference-side binding of libpdeb200.so (include/pdeb200.h).
#
# Drop-in for the hot path of DistributedConvRL-PDE-Control: the four closures a setup file hands to
# `PDEenv(...)` (src/PDEenv.jl:31-35, called in order at :195-241) are replaced by thin `ccall`
# wrappers; `PDEenv`, `PDEagent`, `PDEhook`, `run(agent, env, ...)` stay unchanged.
#
# NOTE: Julia is not installed in the build image, so this file has been syntax-reviewed only; the same
# entry points are exercised through Python ctypes in tests/ (the C ABI is language neutral).
# julia/B200PDEAgent.jl (included after src/PDEagent.jl) adds the device trajectory with the reference's `update!`
# overload set and the device-backed policy functor.
module B200PDE

const LIB = get(ENV, "PDEB200_LIB", joinpath(@__DIR__, "..", "distributedconvrl-pde-control_b200", "libpdeb200.so"))

const KS, KSEG1D, NS2D, KSEG2D = Int32(0), Int32(1), Int32(2), Int32(3)
const F32, F64 = Int32(0), Int32(1)
const ARR_Y, ARR_P, ARR_STATE, ARR_ACTION, ARR_DELTA_ACTION, ARR_REWARD, ARR_DONE, ARR_TIME, ARR_STEPS, ARR_Y0, ARR_GRADS,
      ARR_LOSSES, ARR_SENSORS, ARR_ACTION_IN, ARR_STATS, ARR_NSUB = Int32.(0:15)
const NET_BEHAVIOR_ACTOR, NET_BEHAVIOR_CRITIC, NET_TARGET_ACTOR, NET_TARGET_CRITIC = Int32.(0:3)
const ACT_IDENTITY, ACT_RELU, ACT_TANH = Int32.(0:2)

# struct pdeb200_config (field order and types must match include/pdeb200.h)
Base.@kwdef mutable struct Config
    struct_size::Int32 = 0
    problem::Int32 = KS
    dtype::Int32 = F64
    nx::Int32 = 0; ny::Int32 = 1
    n_envs::Int32 = 1
    n_sensors::Int32 = 0; n_actuators::Int32 = 0
    window_size::Int32 = 1; temporal_steps::Int32 = 1; memory_size::Int32 = 0
    oversampling::Int32 = 1
    check_max_value::Int32 = 1
    mono::Int32 = 0
    sensors_per_axis::Int32 = 0
    ifpad::Int32 = 1
    Lx::Float64 = 0; Ly::Float64 = 1
    dt::Float64 = 0; te::Float64 = 0; t0::Float64 = 0
    mu::Float64 = 0; nu::Float64 = 0
    agent_power::Float64 = 0; max_value::Float64 = 0
    obs_scale::Float64 = 0
    reward_gain::Float64 = 0; reward_pow::Float64 = 0; reward_div::Float64 = 0; reward_offset::Float64 = 0
    action_punish::Float64 = 0; delta_action_punish::Float64 = 0
    rtol::Float64 = 1e-8; atol::Float64 = 1e-8       # adaptive-step parity mode (KellerSegelSetup.jl:234-239; FluidSetup.jl:178-186)
    adaptive::Int32 = 0; reserved0::Int32 = 0
end

struct Ctx
    ptr::Ptr{Cvoid}
end

function check(rc::Int32, ctx::Ptr{Cvoid} = C_NULL)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:pdeb200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    error("pdeb200 error $rc: $msg")
end

function default_config(problem::Int32)
    cfg = Ref(Config())
    check(ccall((:pdeb200_default_config, LIB), Int32, (Int32, Ref{Config}), problem, cfg))
    cfg[]
end

function create(cfg::Config; device::Integer = 0)
    cfg.struct_size = Int32(sizeof(Config))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pdeb200_create, LIB), Int32, (Ref{Config}, Int32, Ref{Ptr{Cvoid}}), Ref(cfg), Int32(device), out))
    Ctx(out[])
end

destroy(c::Ctx) = ccall((:pdeb200_destroy, LIB), Int32, (Ptr{Cvoid},), c.ptr)

# gaussians / gaussians_actuators are Vector{Vector{Float64}} in the setup files: hcat them so that the
# memory is row-major [n][nx] as the C ABI expects.  a2s is 1-based in Julia, 0-based in C.
function set_bases(c::Ctx, gaussians, gaussians_actuators, actuators_to_sensors; drop_tol = 1e-17)
    sb = reduce(hcat, gaussians); ab = reduce(hcat, gaussians_actuators)      # (nx, n): column-major == [n][nx]
    a2s = Int32.(actuators_to_sensors .- 1)
    check(ccall((:pdeb200_set_bases, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Float64),
                c.ptr, sb, ab, a2s, drop_tol), c.ptr)
end

set_y0(c::Ctx, y0::Array{Float64}; broadcast = true) =
    check(ccall((:pdeb200_set_y0, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32), c.ptr, y0, Int32(broadcast)), c.ptr)

reset!(c::Ctx) = check(ccall((:pdeb200_reset, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), c.ptr, C_NULL), c.ptr)

# env(action): one call = prepare_action + do_step + reward_function + featurize + clock for all envs
function step!(c::Ctx, action::Matrix{Float64}, y::Array{Float64}, reward::Vector{Float64}, state::Matrix{Float64},
               done::Vector{UInt8})
    check(ccall((:pdeb200_step_host, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
                c.ptr, action, y, reward, state, done), c.ptr)
end

get!(c::Ctx, which::Int32, dst::Array) =
    check(ccall((:pdeb200_get, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t), c.ptr, which, dst, sizeof(dst)), c.ptr)

# ---- closures for PDEenv(...) -------------------------------------------------------------------
# In scripts/KS/setup/KSSetup.jl, `initialize_setup()` (:249-262) becomes
#
#   ctx = B200PDE.create(cfg); B200PDE.set_bases(ctx, gaussians, gaussians_actuators, actuators_to_sensors)
#   B200PDE.set_y0(ctx, y0_1D_standard); B200PDE.reset!(ctx)
#   closures = B200PDE.make_closures(ctx, nx, n_act * n_envs, ns)
#   env = PDEenv(do_step = closures.do_step, reward_function = closures.reward_function,
#                featurize = closures.featurize, prepare_action = closures.prepare_action, ...)
#
# PDEenv calls prepare_action -> do_step -> reward_function -> featurize (src/PDEenv.jl:199,217,220,222).
# The fused step runs inside `do_step`; the other three return what that launch already produced.
function make_closures(c::Ctx, nx::Int, ncols::Int, ns::Int, n_envs::Int = 1)
    y = zeros(nx, n_envs); reward = zeros(ncols); state = zeros(ns, ncols); done = zeros(UInt8, n_envs)
    p = zeros(nx, n_envs)
    prepare_action = function (action0 = nothing, t0 = nothing; env = nothing)
        isnothing(env) || B200PDE.get!(c, ARR_P, p)     # after the step; before it p is the reset value
        p
    end
    do_step = function (env)
        step!(c, Matrix{Float64}(env.action), y, reward, state, done)
        B200PDE.get!(c, ARR_P, p); env.p = copy(p)
        copy(y)
    end
    reward_function = env -> copy(reward)
    featurize = function (y0 = nothing, t0 = nothing; env = nothing)
        isnothing(env) && B200PDE.get!(c, ARR_STATE, state)   # constructor / reset!: state of y0
        copy(state)
    end
    (; do_step, reward_function, featurize, prepare_action)
end

# ---- CustomNeuralNetworkApproximator on the device (src/custom_nna.jl) ----------------------------
# Flux models stay the host-side source of truth (save()/load() keep working, KSSetup.jl:378-402):
# push their parameters with net_set!, pull them back with net_get! before saving.
function net_set!(c::Ctx, net::Int32, sizes::Vector{Int32}, acts::Vector{Int32}, flat::Vector{Float32})
    check(ccall((:pdeb200_net_set, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}),
                c.ptr, net, Int32(length(acts)), sizes, acts, flat), c.ptr)
end
net_get!(c::Ctx, net::Int32, flat::Vector{Float32}) =
    check(ccall((:pdeb200_net_get, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Csize_t), c.ptr, net, flat, length(flat)), c.ptr)

# Flux.Chain of Dense -> flat vector in the C ABI's order (per layer: W column-major, then b)
flatten(chain) = reduce(vcat, [vcat(vec(l.weight), l.bias) for l in chain.layers])

policy_act!(c::Ctx; act_noise = 0.0, act_limit = 1.0, noise = nothing) =
    check(ccall((:pdeb200_policy_act, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64),
                c.ptr, isnothing(noise) ? C_NULL : noise, act_noise, act_limit), c.ptr)

# Exploration noise one step ahead (pdeb200_noise_prefetch): its upload runs on a copy stream under the current step's kernels.
# `noise` must stay referenced until the consuming act_step! has returned (the copy reads it asynchronously).
noise_prefetch!(c::Ctx, noise::Matrix{Float64}) =
    check(ccall((:pdeb200_noise_prefetch, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), c.ptr, noise), c.ptr)

# `action = policy(env); env(action)` as ONE call with one synchronisation (pdeb200_act_step_host); packed receives
# [reward | done | state] (result_layout), or only [reward | done] after result_select!(c, false) -- enough for a host whose
# policy and trajectory are on the device.  noise = nothing with act_noise > 0 consumes the oldest prefetched noise.
result_select!(c::Ctx, with_state::Bool) =
    check(ccall((:pdeb200_result_select, LIB), Int32, (Ptr{Cvoid}, Int32), c.ptr, Int32(with_state)), c.ptr)
act_step!(c::Ctx, packed::Vector{UInt8}; act_noise = 0.0, act_limit = 1.0, noise = nothing) =
    check(ccall((:pdeb200_act_step_host, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}),
                c.ptr, isnothing(noise) ? C_NULL : noise, act_noise, act_limit, C_NULL, C_NULL, packed, C_NULL, C_NULL, C_NULL), c.ptr)

ddpg_update!(c::Ctx; γ = 0.99, p = 0.995, lr_actor = 5e-4, lr_critic = 1e-3, literal_q1 = true) =
    check(ccall((:pdeb200_ddpg_update, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Float64, Float64, Int32),
                c.ptr, γ, p, lr_actor, lr_critic, Int32(literal_q1)), c.ptr)

# (app::CustomNeuralNetworkApproximator)(x), src/custom_nna.jl:13: x is (rows, n_cols) Float32; dense layers run on
# the tensor cores (tcgen05, 3xTF32), thin ones on CUDA cores.  path: 0 auto, 1 CUDA cores, 2 tensor cores.
function net_forward(c::Ctx, net::Int32, x::Matrix{Float32}, n_out::Integer; path = 0)
    y = Matrix{Float32}(undef, n_out, size(x, 2))
    used = Ref{Int32}(0)
    check(ccall((:pdeb200_net_forward, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float32}, Ptr{Float32}, Int32, Ref{Int32}),
                c.ptr, net, Int32(size(x, 2)), x, y, Int32(path), used), c.ptr)
    y
end

# A device-backed approximator with the method set of src/custom_nna.jl:7-27.  `model` (a Flux Chain) stays the
# host-side source of truth for save()/load(); `sync!` pulls the trained parameters back into it.
mutable struct DeviceApproximator
    ctx::Ctx
    net::Int32
    model
    n_out::Int
end
(app::DeviceApproximator)(x) = net_forward(app.ctx, app.net, Matrix{Float32}(x), app.n_out)
function upload!(app::DeviceApproximator, sizes::Vector{Int32}, acts::Vector{Int32})
    net_set!(app.ctx, app.net, sizes, acts, Vector{Float32}(flatten(app.model)))
end

# ---- device-resident CircularArraySARTTrajectory (src/PDEagent.jl:237-340) ----------------------------------
# The trajectory `update!` overloads of PDEagent.jl push one transition per actuator column; with the batch
# folded into the column axis these become four calls on device buffers.
traj_create!(c::Ctx, capacity::Integer) =
    check(ccall((:pdeb200_traj_create, LIB), Int32, (Ptr{Cvoid}, Int64), c.ptr, Int64(capacity)), c.ptr)
traj_pre_episode!(c::Ctx) = check(ccall((:pdeb200_traj_pop_tail, LIB), Int32, (Ptr{Cvoid},), c.ptr), c.ptr)       # :237-252
traj_pre_act!(c::Ctx) = check(ccall((:pdeb200_traj_push_pre, LIB), Int32, (Ptr{Cvoid},), c.ptr), c.ptr)           # :254-274
traj_post_act!(c::Ctx) = check(ccall((:pdeb200_traj_push_post, LIB), Int32, (Ptr{Cvoid},), c.ptr), c.ptr)         # :276-289
traj_post_episode!(c::Ctx) = check(ccall((:pdeb200_traj_episode_end, LIB), Int32, (Ptr{Cvoid},), c.ptr), c.ptr)   # :291-314
function traj_length(c::Ctx)
    n = Ref{Int64}(0)
    check(ccall((:pdeb200_traj_length, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), c.ptr, n), c.ptr)
    n[]
end
# pde_sample (:317-340): inds ~ U{1 .. length - number_actuators}; s' = state[inds .+ number_actuators]
sample!(c::Ctx, batch::Integer; seed = 0, offset = 0) =
    check(ccall((:pdeb200_sample, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, UInt64, UInt64),
                c.ptr, Int32(batch), C_NULL, UInt64(seed), UInt64(offset)), c.ptr)

# update!(policy, traj, env, ::PreActStage) (src/PDEagent.jl:342-361): update_loops x { pde_sample ; update! } as ONE call
# (one CUDA graph on the device; a collective after comm_init!)
train_updates!(c::Ctx, n_updates::Integer, batch::Integer; γ = 0.99, p = 0.995, lr_actor = 5e-4, lr_critic = 1e-3,
               literal_q1 = true, seed = 0) =
    check(ccall((:pdeb200_train_updates, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Float64, Float64, Float64, Float64, Int32, UInt64),
                c.ptr, Int32(n_updates), Int32(batch), γ, p, lr_actor, lr_critic, Int32(literal_q1), UInt64(seed)), c.ptr)

# pde_fetch!'s result for the staged batch (src/PDEagent.jl:322-340): (s, a, r, t, s′, inds) in the reference's shapes
function get_batch(c::Ctx, batch::Integer, ns::Integer, na::Integer)
    s = Matrix{Float32}(undef, ns, batch); a = Matrix{Float32}(undef, na, batch); s2 = Matrix{Float32}(undef, ns, batch)
    r = Vector{Float32}(undef, batch); t = Vector{UInt8}(undef, batch); inds = Vector{Int64}(undef, batch)
    check(ccall((:pdeb200_get_batch, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}, Ptr{Float32}, Ptr{Int64}),
                c.ptr, s, a, r, t, s2, inds), c.ptr)
    (state = s, action = a, reward = r, terminal = t .!= 0, next_state = s2), inds .+ 1
end

# ---- batched termination (src/PDEenv.jl:226-240 per environment) ---------------------------------------------
# returns (n_done, n_time_limit, n_diverged_and_reset); sync = false enqueues without reading the counters back
function reset_diverged!(c::Ctx; sync = true)
    counts = zeros(Int32, 3)
    check(ccall((:pdeb200_reset_diverged, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), c.ptr, sync ? pointer(counts) : C_NULL), c.ptr)
    Tuple(counts)
end

# one environment's slice (PDEhook's tracked environment, src/PDEhook.jl:54-62)
get_env!(c::Ctx, which::Int32, env_index::Integer, dst::Array) =
    check(ccall((:pdeb200_get_env, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Csize_t),
                c.ptr, which, Int32(env_index - 1), dst, sizeof(dst)), c.ptr)

# ---- checkpoint state: save() / load() (scripts/KS/setup/KSSetup.jl:378-402) -------------------------------------
# Flux.loadparams! semantics: weights only, optimiser state kept (src/custom_nna.jl:26-27)
net_set_params!(c::Ctx, net::Int32, flat::Vector{Float32}) =
    check(ccall((:pdeb200_net_set_params, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Csize_t), c.ptr, net, flat, length(flat)), c.ptr)
# Flux ADAM state (m, v, βp) of one network in the flat parameter layout: sync into / from `optimizer.state` around save()/load()
function opt_get(c::Ctx, net::Int32, n_params::Integer)
    m = Vector{Float32}(undef, n_params); v = Vector{Float32}(undef, n_params); βp = Vector{Float64}(undef, 2)
    check(ccall((:pdeb200_opt_get, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Float64}, Csize_t),
                c.ptr, net, m, v, βp, n_params), c.ptr)
    m, v, βp
end
opt_set!(c::Ctx, net::Int32, m::Vector{Float32}, v::Vector{Float32}, βp::Vector{Float64}) =
    check(ccall((:pdeb200_opt_set, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Float64}, Csize_t),
                c.ptr, net, m, v, βp, length(m)), c.ptr)
# replay rings <-> RLCore's CircularArrayBuffers: logical order + (first - 1) raw positions, as agent.jld2 stores them
function traj_info(c::Ctx)
    v = [Ref{Int64}(0) for _ in 1:5]
    check(ccall((:pdeb200_traj_info, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}),
                c.ptr, v[1], v[2], v[3], v[4], v[5]), c.ptr)
    (capacity = v[1][], n_sa = v[2][], n_rt = v[3][], first_sa = v[4][], first_rt = v[5][])
end
function traj_get(c::Ctx, ns::Integer, na::Integer)
    i = traj_info(c)
    s = Matrix{Float32}(undef, ns, i.n_sa); a = Matrix{Float32}(undef, na, i.n_sa)
    r = Vector{Float32}(undef, i.n_rt); t = Vector{UInt8}(undef, i.n_rt)
    check(ccall((:pdeb200_traj_get, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}), c.ptr, s, a, r, t), c.ptr)
    s, a, r, t .!= 0
end
traj_set!(c::Ctx, s::Matrix{Float32}, a::Matrix{Float32}, r::Vector{Float32}, t::Vector{UInt8}; first_sa = 0, first_rt = 0) =
    check(ccall((:pdeb200_traj_set, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}),
                c.ptr, size(s, 2), length(r), Int64(first_sa), Int64(first_rt), s, a, r, t), c.ptr)
function rng_get(c::Ctx)
    v = Ref{UInt64}(0)
    check(ccall((:pdeb200_rng_get, LIB), Int32, (Ptr{Cvoid}, Ref{UInt64}), c.ptr, v), c.ptr)
    v[]
end
rng_set!(c::Ctx, offset::Integer) = check(ccall((:pdeb200_rng_set, LIB), Int32, (Ptr{Cvoid}, UInt64), c.ptr, UInt64(offset)), c.ptr)

# ---- multi-GPU: one Julia process (or task with its own device) per GPU (SURVEY.md 8b / 8e) -----------------------
# rank 0:  id = comm_unique_id()  -> ship the 128 bytes to the other ranks (MPI.Bcast!, a file, Distributed.remotecall ...)
# every rank: comm_init!(ctx, id, rank, nranks)   (0-based rank).  Afterwards sample! / ddpg_update! / train_updates! are
# collectives and the gradient exchange runs inside the library's kernels over NVLink peer memory.
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:pdeb200_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    id
end
comm_init!(c::Ctx, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    check(ccall((:pdeb200_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), c.ptr, id, Int32(rank), Int32(nranks)), c.ptr)
comm_destroy!(c::Ctx) = check(ccall((:pdeb200_comm_destroy, LIB), Int32, (Ptr{Cvoid},), c.ptr), c.ptr)
function comm_info(c::Ctx)
    r = Ref{Int32}(0); n = Ref{Int32}(1); t = Ref{Int32}(0)
    check(ccall((:pdeb200_comm_info, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}), c.ptr, r, n, t), c.ptr)
    (rank = r[], nranks = n[], transport = (:none, :nccl, :peer)[t[] + 1])
end
# in-place sum over the ranks of up to 64 host Float64 (episode returns / done counts for PDEhook)
comm_allreduce!(c::Ctx, v::Vector{Float64}) =
    check(ccall((:pdeb200_comm_allreduce_f64, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32), c.ptr, v, Int32(length(v))), c.ptr)

# 0 = auto (fused shared-memory kernels / layer-wise GEMM path for wide networks), 1-3 force the layer-wise path
ddpg_set_path!(c::Ctx, path::Integer) =
    check(ccall((:pdeb200_ddpg_set_path, LIB), Int32, (Ptr{Cvoid}, Int32), c.ptr, Int32(path)), c.ptr)

# fused evaluation roll-out (src/plotting.jl:55-73): n_steps x { actor forward -> env step } on the device
function rollout!(c::Ctx, n_steps::Integer, n_envs::Integer; act_limit = 1.0)
    rs = zeros(Float64, n_envs)
    check(ccall((:pdeb200_rollout, LIB), Int32, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}), c.ptr, Int32(n_steps), act_limit, rs), c.ptr)
    rs
end

end # module
