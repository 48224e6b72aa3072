/*
 * pdeb200.h -- C ABI of the B200-native batched PDE-control hot path.
 *
 * Drop-in boundary for janstenner/DistributedConvRL-PDE-Control: each entry
 * point replaces one closure / method of the reference's plugin interface and is
 * what a Julia `ccall` (or, in this repository, Python `ctypes`) binds.  Plain C,
 * no exceptions across the boundary, no torch types.  Reference citations are
 * relative to the reference repository root.
 *
 *   reference interface                                   replaced by
 *   ----------------------------------------------------  -----------------------------
 *   PDEenv(...) constructor        src/PDEenv.jl:64-170    pdeb200_create + set_bases + set_y0 + reset
 *   RLBase.reset!(env)             src/PDEenv.jl:183-193   pdeb200_reset
 *   (env::PDEenv)(action)          src/PDEenv.jl:195-241   pdeb200_step / pdeb200_step_device
 *     prepare_action               scripts/KS/setup/KSSetup.jl:231-245 (KSeg :318-332, Fluid :247-261)
 *     do_step                      KSSetup.jl:130-160; KellerSegelSetup.jl:213-239; fluid_rk4.jl:122-229 + FluidSetup.jl:163-172
 *     reward_function              KSSetup.jl:162-184; KellerSegelSetup.jl:241-263; FluidSetup.jl:188-202
 *     featurize                    KSSetup.jl:190-229; KellerSegelSetup.jl:265-316; FluidSetup.jl:204-245
 *   env.y / env.p / env.state / env.reward / env.done ...  pdeb200_get / pdeb200_device_ptr
 *   CustomNeuralNetworkApproximator(model, optimizer)
 *                                  src/custom_nna.jl:7-27  pdeb200_net_set / pdeb200_net_get
 *   (app::CustomNeuralNetworkApproximator)(x)   custom_nna.jl:13  pdeb200_net_forward
 *   (policy::CustomDDPGPolicy)(env) src/PDEagent.jl:175-209 pdeb200_policy_act
 *   policy loop in plot_heat        src/plotting.jl:55-73   pdeb200_rollout (fused actor + env step, K steps / launch)
 *   trajectory update! overloads   src/PDEagent.jl:237-314 pdeb200_traj_push_pre / _post / _episode_end / _pop_tail
 *   pde_sample / pde_fetch!        src/PDEagent.jl:317-340 pdeb200_sample
 *   update!(policy, batch)         src/PDEagent.jl:363-418 pdeb200_ddpg_update (gradient exchange over NVLink peer memory inside)
 *   update!(policy, traj, env, ::PreActStage)  src/PDEagent.jl:342-361  pdeb200_train_updates (update_loops x {sample, update}, one CUDA graph)
 *   Flux.Optimise.ADAM state / save() / load()  scripts/KS/setup/KSSetup.jl:378-402  pdeb200_opt_get/set, pdeb200_traj_get/set
 *   (no counterpart: the reference is single-process)       pdeb200_comm_unique_id / pdeb200_comm_init
 *
 * Conventions
 *   - every call returns int32 status: 0 = OK, negative = error; the message is
 *     available from pdeb200_last_error(ctx) (or pdeb200_last_error(NULL) for
 *     create failures);
 *   - host pointers are borrowed for the duration of the call only;
 *   - calls are synchronous for the caller unless the name ends in _async /
 *     _device (those enqueue on the context's stream and return);
 *   - a context is not thread-safe; one Julia task / Python thread drives it;
 *   - array layouts are exactly the memory of the reference's column-major Julia
 *     arrays with the environment batch folded into the actuator (column) axis:
 *       y       Julia (nx, B)  [KS]  (2, nx, B) [KSeg]  (2, nx, ny, B) [KSeg 2-D]  complex (ny, nx, B) [NS]   -> env-major contiguous
 *       state   Julia (ns, n_act*B)      -> [B][n_act][ns]
 *       action  Julia (1+mem, n_act*B)   -> [B][n_act][1+mem]
 *       reward  Julia (n_act*B,)         -> [B][n_act]
 *     Element type of y/p/state/action/reward is the context dtype (f32 or f64);
 *     networks and the replay buffer are always f32 (src/PDEagent.jl:112-117).
 */
#ifndef PDEB200_H
#define PDEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEB200_ABI_VERSION 2

typedef struct pdeb200_ctx pdeb200_ctx;

enum { PDEB200_OK = 0, PDEB200_EINVAL = -1, PDEB200_ECUDA = -2, PDEB200_EUNSUPPORTED = -3, PDEB200_ESTATE = -4, PDEB200_ECOMM = -5 };
enum { PDEB200_F32 = 0, PDEB200_F64 = 1 };
enum { PDEB200_KS = 0, PDEB200_KSEG1D = 1, PDEB200_NS2D = 2, PDEB200_KSEG2D = 3 };
enum { PDEB200_CHECK_NONE = 0, PDEB200_CHECK_Y = 1, PDEB200_CHECK_REWARD = 2 };   /* src/PDEenv.jl:226-240 */
enum { PDEB200_ACT_IDENTITY = 0, PDEB200_ACT_RELU = 1, PDEB200_ACT_TANH = 2 };
enum { PDEB200_NET_BEHAVIOR_ACTOR = 0, PDEB200_NET_BEHAVIOR_CRITIC = 1, PDEB200_NET_TARGET_ACTOR = 2, PDEB200_NET_TARGET_CRITIC = 3 };

/* arrays addressable through pdeb200_get / pdeb200_set / pdeb200_device_ptr */
enum {
    PDEB200_ARR_Y = 0,            /* env.y                                   dtype   */
    PDEB200_ARR_P = 1,            /* env.p (actuation field)                 dtype   */
    PDEB200_ARR_STATE = 2,        /* env.state  [B][n_act][ns]               dtype   */
    PDEB200_ARR_ACTION = 3,       /* env.action [B][n_act][1+mem]            dtype   */
    PDEB200_ARR_DELTA_ACTION = 4, /* env.delta_action                        dtype   */
    PDEB200_ARR_REWARD = 5,       /* env.reward [B][n_rew]                   dtype   */
    PDEB200_ARR_DONE = 6,         /* env.done   [B]                          uint8   */
    PDEB200_ARR_TIME = 7,         /* env.time   [B]                          float64 */
    PDEB200_ARR_STEPS = 8,        /* env.steps  [B]                          int32   */
    PDEB200_ARR_Y0 = 9,           /* env.y0                                  dtype   */
    PDEB200_ARR_GRADS = 10,       /* flat [critic | actor] gradient (+ tail) float32 */
    PDEB200_ARR_LOSSES = 11,      /* {critic_loss, actor_loss}               float32 */
    PDEB200_ARR_SENSORS = 12,     /* raw sensor dots [B][fields][n_sensors]  dtype   */
    PDEB200_ARR_ACTION_IN = 13,   /* staged action of the last policy_act    dtype   */
    PDEB200_ARR_NSUB = 15,        /* adaptive mode: {accepted, rejected} substeps of the last env step  int32 [B][2] */
    PDEB200_ARR_STATS = 14        /* batch sums over ALL ranks {sum r, sum r^2, n, sum c, sum c^2, sum q(actor)}  float64[8] (quirk Q1 r-bar) */
};

typedef struct pdeb200_config {
    int32_t struct_size;          /* = sizeof(pdeb200_config), ABI check                       */
    int32_t problem;              /* PDEB200_KS / KSEG1D / NS2D / KSEG2D                       */
    int32_t dtype;                /* PDEB200_F32 / F64: arithmetic type of the PDE path        */
    int32_t nx, ny;               /* grid; ny = 1 for 1-D problems                             */
    int32_t n_envs;               /* B, independent environments on this GPU                   */
    int32_t n_sensors;            /* length(sensor_positions)                                  */
    int32_t n_actuators;          /* length(actuator_positions)                                */
    int32_t window_size;          /* KSSetup.jl:46 ; 2-D problems use window_size^2 rows       */
    int32_t temporal_steps;       /* KSSetup.jl:43                                             */
    int32_t memory_size;          /* KSSetup.jl:39                                             */
    int32_t oversampling;         /* substeps per env step (KS: CNAB2, KSeg/NS: RK4)           */
    int32_t check_max_value;      /* PDEB200_CHECK_*                                           */
    int32_t mono;                 /* 1 = global-agent variant (KSglobalSetup.jl): one column   */
    int32_t sensors_per_axis;     /* NS2D, KSEG2D: sensor lattice side (FluidSetup.jl:61)      */
    int32_t ifpad;                /* NS2D: 3/2-rule de-aliasing (FluidSetup.jl:101)            */
    double Lx, Ly;                /* domain                                                    */
    double dt, te, t0;            /* env step, episode end, start                              */
    double mu;                    /* KS inhomogeneous forcing amplitude (KSSetup.jl:155)       */
    double nu;                    /* NS viscosity (FluidSetup.jl:28)                           */
    double agent_power;           /* KSSetup.jl:51                                             */
    double max_value;             /* divergence guard (PDEenv.jl:226-237)                      */
    double obs_scale;             /* sensor value = <y,g> * obs_scale  (1/30, 1/4, 1/70)       */
    double reward_gain;           /* r = -|gain*(<y,g> - offset*sum(g))|^pow / div             */
    double reward_pow;            /*     - action_punish a^2 - delta_action_punish da^2        */
    double reward_div;
    double reward_offset;
    double action_punish;
    double delta_action_punish;
    /* Adaptive-step parity mode.  The reference's ACTIVE Keller-Segel stepper is OrdinaryDiffEq's
     * adaptive RK4() at reltol = abstol = 1e-8 (KellerSegelSetup.jl:234-239): adaptive = 1 integrates each environment
     * with its own error-controlled step sequence (classical RK4 + step doubling, RMS error norm like OrdinaryDiffEq's
     * default) instead of `oversampling` fixed substeps.  For PDEB200_NS2D the same controller takes the role of the wired-in
     * `do_step2` (FluidSetup.jl:178-186, reltol = abstol = 1e0 as shipped): the environments advance through attempts
     * (3 RK4 steps each) in lock step, each with its own step size and accept / reject decision.  `oversampling` gives
     * the first trial step dt / oversampling; PDEB200_ARR_NSUB reports {accepted, rejected} per environment.
     * KSEG1D and NS2D. */
    double rtol, atol;
    int32_t adaptive;
    int32_t reserved0;
} pdeb200_config;

/* ---- lifecycle -------------------------------------------------------------------------- */
int32_t pdeb200_abi_version(void);
/* Fills the problem-specific reward/obs constants of the reference setup file for `problem`. */
int32_t pdeb200_default_config(int32_t problem, pdeb200_config* cfg);
int32_t pdeb200_create(const pdeb200_config* cfg, int32_t device, pdeb200_ctx** out);
int32_t pdeb200_destroy(pdeb200_ctx* ctx);
const char* pdeb200_last_error(const pdeb200_ctx* ctx);
/* Use an externally owned cudaStream_t (e.g. torch's current stream); NULL = own stream. */
int32_t pdeb200_set_stream(pdeb200_ctx* ctx, void* cuda_stream);
int32_t pdeb200_synchronize(pdeb200_ctx* ctx);

/* ---- constants -------------------------------------------------------------------------- */
/* gaussians / gaussians_actuators of the setup files, as dense float64 row-major
 * [n_sensors][npts] and [n_actuators][npts] (npts = nx*ny; NS: Julia's (nx,ny) column-major
 * flattening), a2s = actuators_to_sensors, 0-based.  Entries with |w| <= drop_tol*max|w|
 * are dropped when the banded/sparse device tables are built (0 keeps every non-zero). */
int32_t pdeb200_set_bases(pdeb200_ctx* ctx, const double* sensor_basis, const double* actuator_basis,
                          const int32_t* a2s, double drop_tol);
/* y0 (env.y0): float64 host array, either one environment (broadcast = 1) or [B][...] */
int32_t pdeb200_set_y0(pdeb200_ctx* ctx, const double* y0, int32_t broadcast);

/* ---- environment ------------------------------------------------------------------------ */
/* RLBase.reset!(env) for the masked environments (mask == NULL: all). */
int32_t pdeb200_reset(pdeb200_ctx* ctx, const uint8_t* mask);
/* Batched termination.  The reference ends the episode of THE environment that diverged (src/PDEenv.jl:226-240); with B
 * environments per context the ones whose done flag is set although their clock has not reached te are reset in place
 * (reset! semantics) while the others keep stepping -- their terminal flag has already been pushed to the replay ring, so
 * the TD target never bootstraps across the reset.  counts: optional HOST int32[3] = {done, time limit reached, diverged
 * and reset}; NULL: enqueue and return without synchronising. */
int32_t pdeb200_reset_diverged(pdeb200_ctx* ctx, int32_t* counts);
/* env(action): actions is a HOST array [B][n_act][1+mem] of the context dtype.
 * Copies it to the device, runs the fused step, leaves results on the device. */
int32_t pdeb200_step(pdeb200_ctx* ctx, const void* actions_host);
/* same with a DEVICE action buffer; NULL = the context's own action buffer (as
 * written by pdeb200_policy_act).  Enqueues and returns. */
int32_t pdeb200_step_device(pdeb200_ctx* ctx, const void* actions_dev);
/* Convenience for the drop-in closures: step + copy back y/reward/state/done in one call.
 * Any output pointer may be NULL. */
int32_t pdeb200_step_host(pdeb200_ctx* ctx, const void* actions_host, void* y_out, void* reward_out,
                          void* state_out, uint8_t* done_out);
/* `action = policy(env); env(action)` with HOST buffers as ONE call with ONE synchronisation (the drop-in closures' two calls
 * back to back): actor forward (+ optional host noise, as pdeb200_policy_act) -> the action is copied to action_out (HOST) and
 * taken back from that buffer as the step's input (stream-ordered D2H then H2D; action_out = NULL: the action stays on the
 * device, for hosts whose trajectory and hooks live on the device too) -> env step -> results to the host.
 * result_packed: optional HOST buffer of pdeb200_result_layout's total_bytes receiving [reward | done | state] in ONE copy
 * (they share one device allocation; pdeb200_result_select(ctx, 0) shortens the copy to the [reward | done] prefix);
 * reward_out / state_out / done_out: optional separate destinations instead. */
int32_t pdeb200_act_step_host(pdeb200_ctx* ctx, const double* noise_host, double act_noise, double act_limit, void* action_out,
                              void* y_out, void* result_packed, void* reward_out, void* state_out, uint8_t* done_out);
/* Exploration noise never depends on the environment, so a host can hand over the NEXT step's noise before it makes the
 * (synchronous) call for the current step -- the reference draws it inside the policy call, src/PDEagent.jl:201.  This enqueues
 * the upload on a copy stream of the context's own and returns; pdeb200_act_step_host calls with noise_host = NULL and
 * act_noise > 0 consume the prefetches in order.  At most two may be outstanding (PDEB200_ESTATE otherwise), which is what the
 * pipelined loop needs:   prefetch(noise[0]);  for i = 0, 1, ...: { prefetch(noise[i+1]); act_step_host(NULL, ...) }   -- the
 * upload of step i+1's noise then runs under step i's kernels instead of in front of its own. */
int32_t pdeb200_noise_prefetch(pdeb200_ctx* ctx, const double* noise_host);
int32_t pdeb200_result_layout(const pdeb200_ctx* ctx, size_t* reward_off, size_t* state_off, size_t* done_off, size_t* total_bytes);
/* with_state = 0: result_packed receives only [reward | done] (total_bytes of pdeb200_result_layout shrinks accordingly).  For
 * hosts whose policy and trajectory are on the device (DevicePolicyForward / DeviceTrajectory of the Julia shim): nothing on the
 * host reads env.state then -- the hook reads env.reward (src/PDEhook.jl:66-76), the run loop env.done (PDEenv.jl:226-240) --
 * and pdeb200_get(PDEB200_ARR_STATE) still fetches it on demand.  Default: with_state = 1. */
int32_t pdeb200_result_select(pdeb200_ctx* ctx, int32_t with_state);
int32_t pdeb200_get(pdeb200_ctx* ctx, int32_t which, void* host_dst, size_t bytes);
/* One environment's slice of a per-environment array (PDEhook's tracked environment: src/PDEhook.jl:54-62). */
int32_t pdeb200_get_env(pdeb200_ctx* ctx, int32_t which, int32_t env_index, void* host_dst, size_t bytes);
int32_t pdeb200_set(pdeb200_ctx* ctx, int32_t which, const void* host_src, size_t bytes);
int32_t pdeb200_device_ptr(pdeb200_ctx* ctx, int32_t which, void** ptr, size_t* bytes);
int32_t pdeb200_obs_rows(const pdeb200_ctx* ctx);      /* ns = size(state_space)[1] */
int32_t pdeb200_obs_cols(const pdeb200_ctx* ctx);      /* columns per environment   */

/* ---- networks (always float32, Flux Dense layout: W is (out, in) column-major) ---------- */
/* n_layers Dense layers; sizes[0..n_layers] = in, hidden..., out; activations[n_layers];
 * params = concatenation over layers of (W column-major (out,in), b).  */
int32_t pdeb200_net_set(pdeb200_ctx* ctx, int32_t net, int32_t n_layers, const int32_t* sizes,
                        const int32_t* activations, const float* params);
int32_t pdeb200_net_get(pdeb200_ctx* ctx, int32_t net, float* params, size_t n_params);
int32_t pdeb200_net_num_params(const pdeb200_ctx* ctx, int32_t net);

/* The approximator call  y = model(x)  on a (rows, n_cols) matrix (src/custom_nna.jl:13): x_host is float32
 * [n_cols][in] (Julia column-major (in, n_cols)), y_host float32 [n_cols][out].  Dense layers (in, out >= 32) run
 * on the tensor cores (tcgen05, 3xTF32, fp32-accurate); thin layers on CUDA cores.  path: 0 = auto, 1 = CUDA cores
 * only, 2 = tensor cores wherever the layout allows.  n_tensor_layers (optional) receives how many layers used them. */
int32_t pdeb200_net_forward(pdeb200_ctx* ctx, int32_t net, int32_t n_cols, const float* x_host, float* y_host, int32_t path,
                            int32_t* n_tensor_layers);
/* same on DEVICE buffers with leading dimensions (floats); ldx must be a multiple of 4 for the tensor-core path and
 * the padding columns [in, ldx) must be zero. */
int32_t pdeb200_net_forward_device(pdeb200_ctx* ctx, int32_t net, int32_t n_cols, const float* x_dev, int64_t ldx, float* y_dev,
                                   int64_t ldy, int32_t path, int32_t* n_tensor_layers);

/* ---- policy ----------------------------------------------------------------------------- */
/* actions = clamp(behavior_actor(state) + noise*act_noise, +-act_limit)   (PDEagent.jl:189-204)
 * noise_host: NULL (learning = false) or HOST float64 [B][n_cols][na - mem] standard normals
 * supplied by the caller (the reference draws them from a StableRNG, PDEagent.jl:201).
 * Result stays in the context's action buffer (ARR_ACTION is updated by the next step). */
int32_t pdeb200_policy_act(pdeb200_ctx* ctx, const double* noise_host, double act_noise, double act_limit);
/* Same, with noise generated on the device (Philox, seed/offset) -- the batched-training path. */
int32_t pdeb200_policy_act_rng(pdeb200_ctx* ctx, uint64_t seed, uint64_t offset, double act_noise, double act_limit);
/* Fused roll-out: n_steps x { actor forward (no noise) -> env step } in ONE launch with the PDE
 * state resident on chip (the evaluation loop of src/plotting.jl:55-73, batched).
 * reward_sum_out: optional HOST float64 [B] = sum over steps of mean_i reward. */
int32_t pdeb200_rollout(pdeb200_ctx* ctx, int32_t n_steps, double act_limit, double* reward_sum_out);

/* ---- replay buffer (device resident CircularArraySARTTrajectory) ------------------------ */
int32_t pdeb200_traj_create(pdeb200_ctx* ctx, int64_t capacity);
int32_t pdeb200_traj_length(const pdeb200_ctx* ctx, int64_t* length);
int32_t pdeb200_traj_push_pre(pdeb200_ctx* ctx);          /* PreActStage  PDEagent.jl:254-274 */
int32_t pdeb200_traj_push_post(pdeb200_ctx* ctx);         /* PostActStage PDEagent.jl:276-289 */
int32_t pdeb200_traj_episode_end(pdeb200_ctx* ctx);       /* PostEpisode  PDEagent.jl:291-314 */
int32_t pdeb200_traj_pop_tail(pdeb200_ctx* ctx);          /* PreEpisode   PDEagent.jl:237-252 */
/* batch of `batch` transitions; inds_host (0-based, < length - n_cols_total) or NULL to draw
 * uniformly on the device from (seed, offset). */
int32_t pdeb200_sample(pdeb200_ctx* ctx, int32_t batch, const int64_t* inds_host, uint64_t seed, uint64_t offset);
/* Explicit batch (host float32): s [batch][ns], a [batch][na], r [batch], t [batch] (uint8), s' [batch][ns]. */
int32_t pdeb200_set_batch(pdeb200_ctx* ctx, int32_t batch, const float* s, const float* a, const float* r,
                          const uint8_t* t, const float* snext);

/* ---- DDPG update (PDEagent.jl:363-418) -------------------------------------------------- */
/* The whole update on the batch staged by pdeb200_sample / pdeb200_set_batch: targets, critic gradient, ADAM on the
 * critic, actor gradient through the UPDATED critic, ADAM on the actor, Polyak on both targets, losses.
 * literal_q1 = 1 reproduces the reference's (1,B) x (B,) broadcast in the critic loss; 0 = per-sample
 * TD target.  Shipped network sizes: two launches; the last CTA of each gradient kernel reduces the per-CTA partials in a
 * fixed order and applies the optimiser.  After pdeb200_comm_init this call is a COLLECTIVE: that same CTA exchanges
 * the reduced gradient with every peer GPU over NVLink peer memory (fixed rank-order sum, weights stay bit-identical on
 * all ranks); gradients and r-bar are means over the GLOBAL batch (sum of the ranks' local batches). */
int32_t pdeb200_ddpg_update(pdeb200_ctx* ctx, double gamma, double polyak, double lr_actor, double lr_critic,
                            int32_t literal_q1);
/* update!(policy, traj, env, ::PreActStage) (PDEagent.jl:342-361): n_updates x { pde_sample(batch) ; update!(batch) } with
 * device-drawn indices (Philox stream `seed`, counter kept on the device), enqueued as ONE CUDA graph on the context's
 * stream (captured on first use, replayed afterwards).  Collective after pdeb200_comm_init.  Enqueues and returns. */
int32_t pdeb200_train_updates(pdeb200_ctx* ctx, int32_t n_updates, int32_t batch, double gamma, double polyak, double lr_actor,
                              double lr_critic, int32_t literal_q1, uint64_t seed);
/* Which kernels run the update: 0 = auto (fused shared-memory kernels when the four networks fit, otherwise the
 * layer-wise GEMM path: tcgen05 tensor cores for dense layers, CUDA cores for thin ones); 1 = layer-wise, CUDA cores
 * only; 2 = layer-wise, tensor cores wherever the layout allows; 3 = layer-wise, automatic per-layer choice. */
int32_t pdeb200_ddpg_set_path(pdeb200_ctx* ctx, int32_t path);
/* The four phases separately, for hosts that run their OWN gradient exchange between them (never needed with
 * pdeb200_comm_init): critic gradient into ARR_GRADS[0 .. n_critic) scaled by 1/global_batch, ADAM on the critic from
 * ARR_GRADS, actor gradient through the updated critic into ARR_GRADS[n_critic ..), ADAM on the actor + Polyak. */
int32_t pdeb200_ddpg_critic_grads(pdeb200_ctx* ctx, double gamma, int32_t literal_q1, int64_t global_batch);
int32_t pdeb200_ddpg_critic_apply(pdeb200_ctx* ctx, double lr);
int32_t pdeb200_ddpg_actor_grads(pdeb200_ctx* ctx, int64_t global_batch);
int32_t pdeb200_ddpg_actor_apply(pdeb200_ctx* ctx, double lr, double polyak);

/* ---- checkpoint state (save() / load(), scripts/KS/setup/KSSetup.jl:378-402) ----------------------------------- */
/* Overwrite the weights only, keeping the optimiser state (Flux.loadparams!, src/custom_nna.jl:26-27);
 * pdeb200_net_set (re)creates the network and resets ADAM like constructing a fresh Flux.ADAM. */
int32_t pdeb200_net_set_params(pdeb200_ctx* ctx, int32_t net, const float* params, size_t n_params);
/* Flux ADAM state of one network: first / second moments (float32, parameter layout) and the running beta powers
 * `(beta1^t, beta2^t)` (the `(2,)` Float64 arrays of agent.jld2).  Any output pointer may be NULL. */
int32_t pdeb200_opt_get(pdeb200_ctx* ctx, int32_t net, float* m, float* v, double* beta_p2, size_t n_params);
int32_t pdeb200_opt_set(pdeb200_ctx* ctx, int32_t net, const float* m, const float* v, const double* beta_p2, size_t n_params);
/* The replay rings in LOGICAL order (index 0 = oldest column): state [n_sa][ns], action [n_sa][na] (n_sa state/action
 * columns), reward [n_rt], terminal [n_rt].  traj_info returns the counts and the raw (0-based) ring positions of the
 * oldest column of the state/action and reward/terminal rings -- RLCore's CircularArrayBuffer `nframes` and `first - 1`
 * as saved in agent.jld2; get/set copy whole rings (set: n_sa <= capacity+1, n_rt <= capacity; first_* place the oldest
 * column so that a restored ring keeps writing where the saved one would).  Any output pointer may be NULL. */
int32_t pdeb200_traj_info(const pdeb200_ctx* ctx, int64_t* capacity, int64_t* n_sa, int64_t* n_rt, int64_t* first_sa,
                          int64_t* first_rt);
int32_t pdeb200_traj_get(pdeb200_ctx* ctx, float* state, float* action, float* reward, uint8_t* terminal);
int32_t pdeb200_traj_set(pdeb200_ctx* ctx, int64_t n_sa, int64_t n_rt, int64_t first_sa, int64_t first_rt, const float* state,
                         const float* action, const float* reward, const uint8_t* terminal);
/* Sampler counter of pdeb200_train_updates (device resident; part of a resumable checkpoint). */
int32_t pdeb200_rng_get(pdeb200_ctx* ctx, uint64_t* offset);
int32_t pdeb200_rng_set(pdeb200_ctx* ctx, uint64_t offset);
/* The batch staged by the last pdeb200_sample / pdeb200_set_batch (host float32 / uint8 / int64 outputs, any may be NULL):
 * what pde_fetch! returns (PDEagent.jl:322-340). */
int32_t pdeb200_get_batch(pdeb200_ctx* ctx, float* s, float* a, float* r, uint8_t* t, float* snext, int64_t* inds);

/* ---- multi-GPU (the reference itself is single-process) ----------------------------------- */
/* One process (or host thread) per GPU, environments sharded by batch index, ONE exchange per DDPG phase.
 * Rank 0 calls pdeb200_comm_unique_id and distributes the 128 bytes by any host-side means (MPI, a file, a socket,
 * torch.distributed); every rank then calls pdeb200_comm_init(ctx, id, rank, nranks) -- a collective that (1) joins an
 * NCCL communicator (bootstrap + fallback transport) and (2) maps every peer's exchange buffer through CUDA IPC so that
 * the gradient kernels exchange over NVLink / NVSwitch peer memory themselves (transport PDEB200_COMM_PEER).  If peer
 * mapping is impossible the update falls back to ncclAllReduce between its phases (PDEB200_COMM_NCCL);
 * PDEB200_COMM_TRANSPORT=nccl|peer forces one.  From then on pdeb200_sample, pdeb200_set_batch, pdeb200_ddpg_update and
 * pdeb200_train_updates must be called by all ranks in the same order. */
#define PDEB200_UNIQUE_ID_BYTES 128
enum { PDEB200_COMM_NONE = 0, PDEB200_COMM_NCCL = 1, PDEB200_COMM_PEER = 2 };
int32_t pdeb200_comm_unique_id(uint8_t* id128);
int32_t pdeb200_comm_init(pdeb200_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t nranks);
int32_t pdeb200_comm_destroy(pdeb200_ctx* ctx);
int32_t pdeb200_comm_info(const pdeb200_ctx* ctx, int32_t* rank, int32_t* nranks, int32_t* transport);
/* In-place sum over all ranks of n <= 64 host doubles (episode returns, done counts for PDEhook; NCCL on the stream). */
int32_t pdeb200_comm_allreduce_f64(pdeb200_ctx* ctx, double* host_inout, int32_t n);

/* ---- introspection ---------------------------------------------------------------------- */
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t pdeb200_launch_count(const pdeb200_ctx* ctx);
/* Time the n most recent env-step kernels took on the device, via CUDA events on the
 * context's stream (ms); used for the roofline line.  */
int32_t pdeb200_last_step_ms(pdeb200_ctx* ctx, float* ms);
int32_t pdeb200_enable_step_timing(pdeb200_ctx* ctx, int32_t on);
/* Device time of the problem's core kernel(s) -- (y, p) -> (y', sensor dots): the dominant kernel of a step -- in the
 * most recent env step, from CUDA events recorded around it on the context's stream (needs step timing enabled). */
int32_t pdeb200_last_core_ms(pdeb200_ctx* ctx, float* ms);
/* Name of the core kernel the most recent env step launched (KS: which ks_step variant the dispatch chose); "" if none. */
const char* pdeb200_last_core_kernel(const pdeb200_ctx* ctx);
/* The three phases of the most recent single env step {actuation, core, observe} in ms (events on the stream). */
int32_t pdeb200_last_phase_ms(pdeb200_ctx* ctx, float* ms3);
/* Measured CUDA-core FMA peak of this GPU (TFLOP/s, 2 flops per FMA) for dtype PDEB200_F32 / F64: the roofline
 * denominator of the FP-pipe bound kernels (MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers). */
int32_t pdeb200_measure_fma_peak(pdeb200_ctx* ctx, int32_t dtype, double* tflops);
/* With PDEB200_DDPG_TIMELINE=1 in the environment the register-resident DDPG kernels stamp %globaltimer at their phase
 * boundaries; this copies out and resets the records (8 uint64 each: phase 0 critic / 1 actor, t_entry, t_loop_done,
 * t_tail_start, t_reduced, t_exchanged, t_end, unused; ns).  Measurement support for tools/bench_train.py. */
int32_t pdeb200_debug_timeline(pdeb200_ctx* ctx, uint64_t* out, int32_t max_records, int32_t* n);
/* Algorithmic HBM bytes and flops of one env step for this configuration. */
int32_t pdeb200_step_cost(const pdeb200_ctx* ctx, double* hbm_bytes_per_env, double* flops_per_env);

#ifdef __cplusplus
}
#endif
#endif /* PDEB200_H */
