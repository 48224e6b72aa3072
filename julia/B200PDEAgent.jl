# B200PDEAgent.jl -- the reference's agent-side method set on top of libpdeb200.so.
#
# include() this file AFTER src/PDEagent.jl, src/custom_nna.jl and julia/B200PDE.jl (the reference is not a module: its
# names -- CustomDDPGPolicy, CustomNeuralNetworkApproximator, RLBase, the stage types -- live in Main).  It adds
#
#   DeviceTrajectory            a CircularArraySARTTrajectory kept on the GPU, with exactly the `update!` overloads of
#                               src/PDEagent.jl:237-314 (PreEpisode pop, PreAct push of (s, a) per column, PostAct push of
#                               (r, terminal) per column, PostEpisode dummy push) and `length`;
#   update!(policy, traj::DeviceTrajectory, env, ::PreActStage)
#                               the update trigger of src/PDEagent.jl:342-361: update_loops x { pde_sample ; update! } as ONE
#                               library call (B200PDE.train_updates!, a CUDA graph on the device);
#   DevicePolicyForward         `(policy)(env; learning)` of src/PDEagent.jl:175-209 with the actor evaluated on the device:
#                               the action stays in the context's staging buffer, the returned host matrix is what hooks /
#                               start policies see;
#   sync_to_flux! / sync_from_flux!
#                               move weights and ADAM state between the device and the policy's Flux objects around
#                               save() / load() (scripts/KS/setup/KSSetup.jl:378-402).
#
# NOTE: Julia is not installed in the build image: this file is written against src/PDEagent.jl's signatures and has been
# syntax-reviewed only.  The same call sequences are executed by distributedconvrl-pde-control_b200/agent.py (run_episode,
# CustomDDPGPolicy.maybe_update) and tested in tests/test_agent_gpu.py.

using .B200PDE

struct DeviceTrajectory <: AbstractTrajectory
    ctx::B200PDE.Ctx
    capacity::Int
end

function DeviceTrajectory(ctx::B200PDE.Ctx; capacity::Int)
    B200PDE.traj_create!(ctx, capacity)
    DeviceTrajectory(ctx, capacity)
end

Base.length(t::DeviceTrajectory) = Int(B200PDE.traj_length(t.ctx))

# ---- trajectory update! overloads: src/PDEagent.jl:237-314 ------------------------------------------------------------
# PreEpisode: pop the n_columns dummy (state, action) pairs the previous PostEpisode pushed (:237-252)
RLBase.update!(t::DeviceTrajectory, ::CustomDDPGPolicy, ::AbstractEnv, ::PreEpisodeStage) = B200PDE.traj_pre_episode!(t.ctx)
# PreAct: push (state[:, i], action[:, i]) for every column; the action is the one `policy(env)` just staged (:254-274)
RLBase.update!(t::DeviceTrajectory, ::CustomDDPGPolicy, ::AbstractEnv, ::PreActStage, action) = B200PDE.traj_pre_act!(t.ctx)
# PostAct: push reward[i] and is_terminated(env) for every column (:276-289); with B environments per context the
# terminal flag is each environment's own `done`
RLBase.update!(t::DeviceTrajectory, ::CustomDDPGPolicy, ::AbstractEnv, ::PostActStage) = B200PDE.traj_post_act!(t.ctx)
# PostEpisode: final state + zero action per column (:291-314)
RLBase.update!(t::DeviceTrajectory, ::CustomDDPGPolicy, ::AbstractEnv, ::PostEpisodeStage) = B200PDE.traj_post_episode!(t.ctx)

# ---- update trigger: src/PDEagent.jl:342-361 ------------------------------------------------------------------------------
function RLBase.update!(policy::CustomDDPGPolicy, traj::DeviceTrajectory, ::AbstractEnv, ::PreActStage)
    number_actuators = length(size(policy.action_space)) == 2 ? size(policy.action_space)[2] : 1
    length(traj) > policy.update_after * number_actuators || return
    policy.update_step % policy.update_freq == 0 || return
    B200PDE.train_updates!(traj.ctx, policy.update_loops, policy.batch_size; γ = policy.y, p = policy.p,
                           lr_actor = policy.behavior_actor.optimizer.eta, lr_critic = policy.behavior_critic.optimizer.eta,
                           literal_q1 = true, seed = B200PDE_SEED[])
    nothing
end
const B200PDE_SEED = Ref{UInt64}(0)          # Philox stream of the device sampler (policy.rng is a host StableRNG)

function RLBase.update!(policy::CustomDDPGPolicy, ::DeviceTrajectory, ::AbstractEnv, stage::Union{PostEpisodeStage,PostExperimentStage})
    stage == policy.reset_stage && (policy.update_step = 0)             # src/PDEagent.jl:211-235
end

# ---- policy forward: src/PDEagent.jl:175-209 ------------------------------------------------------------------------------
# Wrap the reference policy so that `agent(env)` evaluates the behavior actor on the device.  The action is left in the
# context's staging buffer (pdeb200_step_device / traj_pre_act! read it there); the returned matrix is a host copy.
struct DevicePolicyForward{P<:CustomDDPGPolicy} <: AbstractPolicy
    policy::P
    ctx::B200PDE.Ctx
end

function (dp::DevicePolicyForward)(env; learning = true, test = false)
    policy = dp.policy
    learning && (policy.update_step += 1)
    na, ncols = size(policy.action_space)
    actions = Matrix{Float64}(undef, na, ncols)
    if policy.update_step <= policy.start_steps
        actions .= policy.start_policy(env)
        check_set = ccall((:pdeb200_set, B200PDE.LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t),
                          dp.ctx.ptr, B200PDE.ARR_ACTION_IN, actions, sizeof(actions))
        B200PDE.check(check_set, dp.ctx.ptr)
    else
        noise = learning ? randn(policy.rng, na - policy.memory_size, ncols) : nothing     # host StableRNG, as in the reference
        B200PDE.policy_act!(dp.ctx; act_noise = learning ? policy.act_noise : 0.0, act_limit = policy.act_limit, noise = noise)
        B200PDE.get!(dp.ctx, B200PDE.ARR_ACTION_IN, actions)
    end
    actions
end
(dp::DevicePolicyForward)(stage::AbstractStage, env::AbstractEnv) = nothing
RLBase.update!(dp::DevicePolicyForward, args...) = RLBase.update!(dp.policy, args...)

# ---- save() / load(): scripts/KS/setup/KSSetup.jl:378-402 ----------------------------------------------------------------------
const _NETS = ((:behavior_actor, B200PDE.NET_BEHAVIOR_ACTOR), (:behavior_critic, B200PDE.NET_BEHAVIOR_CRITIC),
               (:target_actor, B200PDE.NET_TARGET_ACTOR), (:target_critic, B200PDE.NET_TARGET_CRITIC))

_acts(chain) = Int32[l.σ === identity ? B200PDE.ACT_IDENTITY : (l.σ === tanh ? B200PDE.ACT_TANH : B200PDE.ACT_RELU) for l in chain.layers]
_sizes(chain) = Int32[size(chain.layers[1].weight, 2); [size(l.weight, 1) for l in chain.layers]]

# after create_agent / load(): Flux objects -> device (weights; ADAM moments when the optimiser already has state)
function sync_from_flux!(ctx::B200PDE.Ctx, policy::CustomDDPGPolicy)
    for (field, id) in _NETS
        app = getfield(policy, field)
        chain = app.model
        B200PDE.net_set!(ctx, id, _sizes(chain), _acts(chain), Vector{Float32}(B200PDE.flatten(chain)))
        st = app.optimizer.state
        isempty(st) && continue
        m = Float32[]; v = Float32[]; βp = [0.9, 0.999]
        for l in chain.layers, p in (l.weight, l.bias)
            mt, vt, b = st[p]
            append!(m, vec(mt)); append!(v, vec(vt)); βp = collect(Float64, b)
        end
        B200PDE.opt_set!(ctx, id, m, v, βp)
    end
end

# before save(): device -> Flux objects, so that FileIO.save writes what a CPU run would have written
function sync_to_flux!(ctx::B200PDE.Ctx, policy::CustomDDPGPolicy)
    for (field, id) in _NETS
        app = getfield(policy, field)
        chain = app.model
        n = sum(length(l.weight) + length(l.bias) for l in chain.layers)
        flat = Vector{Float32}(undef, n)
        B200PDE.net_get!(ctx, id, flat)
        m, v, βp = B200PDE.opt_get(ctx, id, n)
        o = 0
        for l in chain.layers, p in (l.weight, l.bias)
            k = length(p)
            p .= reshape(flat[o+1:o+k], size(p))
            app.optimizer.state[p] = (reshape(m[o+1:o+k], size(p)), reshape(v[o+1:o+k], size(p)), copy(βp))
            o += k
        end
    end
end
