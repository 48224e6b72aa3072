// Problem-independent kernels on either side of the PDE core kernels.
//
//   actuate_kernel : [fused actor forward]      src/PDEagent.jl:189,204
//                    delta_action / action      src/PDEenv.jl:196-197
//                    prepare_action             KSSetup.jl:231-245, KellerSegelSetup.jl:318-332,
//                                               FluidSetup.jl:247-259 (the physical-space sum)
//   observe_kernel : reward_function, featurize, clock/done   (see obs_reward.cuh, PDEenv.jl:220-240)
//
// They run at high occupancy with one thread per column / grid point (actuation: persistent CTAs over groups of
// environments; observation: one warp per environment); the register- and shared-memory-heavy core kernels
// (ks_step.cuh, ...) only see  (y, p) -> (y', sensor dots, max|y|).  actuate_conv_kernel is the shape-specialised
// actuation of the conv agent, actuate_kernel the runtime-shaped one; both give bit-identical results.
#pragma once
#include "common.cuh"
#include "obs_reward.cuh"

namespace pdeb200 {

template <typename T>
struct ActuateArgs {
    int n_envs, envs_per_cta;
    int n_act, a_rows, obs_rows, mono, npts, use_actor;
    int actor_wmax, actor_np;      // widest layer / parameter count of the actor
    int stage_table;               // 1: the gather table fits the CTA's shared memory
    const int* act_idx;            // ELL over grid points: [nnz][npts] actuator index
    const T* act_w;                //                        [nnz][npts] weight
    int act_nnz;
    T power, act_limit;
    NetDev actor;
    const T* actions_in;           // [B][n_act][a_rows] (ignored when use_actor)
    const T* state;                // [B][n_cols][obs_rows]
    T* action; T* delta_action;    // [B][n_act][a_rows]
    T* p;                          // [B][npts] physical actuation field
    // fused actor of actuate_conv_kernel only: exploration noise [B][n_act] added before the clamp (PDEagent.jl:201-202),
    // and the policy's output array (ARR_ACTION_IN) kept up to date
    const T* noise; T act_noise; T* action_in_out;
};

// Shared memory: s_a [E][n_act] T | staged table (w, idx) | actor params | activations [2][wmax][blockDim] f32
template <typename T>
__host__ __device__ inline size_t actuate_smem_bytes(int E, int n_act, int npts, int nnz, int stage, int use_actor,
                                                     int actor_np, int wmax, int block) {
    size_t b = (size_t)E * n_act * sizeof(T);
    b = (b + 15) & ~(size_t)15;
    if (stage) b += (size_t)nnz * npts * (sizeof(T) + sizeof(int));
    b = (b + 15) & ~(size_t)15;
    if (use_actor) b += (size_t)actor_np * sizeof(float) + (size_t)2 * wmax * block * sizeof(float);
    return b;
}

// MLP forward with per-thread activations held in shared memory ([unit][thread], conflict free) and
// weights broadcast from shared memory.  Restates Flux Dense: y = act.(W*x .+ b)  (src/PDEagent.jl:18-30).
// Input in xa[i*BD]; returns the buffer holding the output.
__device__ __forceinline__ float* mlp_forward_smem(const NetDev& net, const float* params, float* xa, float* xb, int BD) {
    float* in = xa;
    float* out = xb;
    for (int l = 0; l < net.n_layers; ++l) {
        const int ni = net.sizes[l], no = net.sizes[l + 1];
        const float* W = params + net.offs[l];
        const float* b = W + ni * no;
        for (int o = 0; o < no; ++o) {
            float acc = 0.f;
            for (int i = 0; i < ni; ++i) acc = fmaf(W[o + no * i], in[i * BD], acc);
            out[o * BD] = act_apply(net.acts[l], acc + b[o]);
        }
        float* tmp = in; in = out; out = tmp;
    }
    return in;
}

template <typename T>
__global__ void __launch_bounds__(256) actuate_kernel(const __grid_constant__ ActuateArgs<T> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int E = A.envs_per_cta, BD = blockDim.x, tid = threadIdx.x;
    T* s_a = reinterpret_cast<T*>(smem_raw);                       // row 0 of the new actions, [E][n_act]
    size_t off = ((size_t)E * A.n_act * sizeof(T) + 15) & ~(size_t)15;
    const size_t ntab = (size_t)A.act_nnz * A.npts;
    T* s_w = reinterpret_cast<T*>(smem_raw + off);
    int* s_idx = reinterpret_cast<int*>(s_w + ntab);
    if (A.stage_table) off += ntab * (sizeof(T) + sizeof(int));
    off = (off + 15) & ~(size_t)15;
    float* s_par = reinterpret_cast<float*>(smem_raw + off);
    float* s_x = s_par + A.actor_np;
    if (A.stage_table)
        for (size_t i = tid; i < ntab; i += BD) { s_w[i] = __ldg(A.act_w + i); s_idx[i] = __ldg(A.act_idx + i); }
    if (A.use_actor)
        for (int i = tid; i < A.actor_np; i += BD) s_par[i] = A.actor.params[i];
    __syncthreads();
    // ---- fast path (conv agent, E * n_act <= blockDim, npts <= blockDim, <= 8 taps per grid point, actor width <= 8):
    // no integer division in the loops, the thread's gather taps and the actor activations live in registers
    int actor_w = 0;
    if (A.use_actor) for (int l = 0; l <= A.actor.n_layers; ++l) actor_w = max(actor_w, A.actor.sizes[l]);
    if (!A.mono && A.stage_table && E * A.n_act <= BD && A.npts <= BD && A.act_nnz <= 8 && actor_w <= 8 && A.a_rows <= 8) {
        const int ce = tid / A.n_act, cj = tid - ce * A.n_act;            // this thread's column slot (fixed for all groups)
        const bool col_on = tid < E * A.n_act;
        int ti[8]; T tw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool on = j < A.act_nnz && tid < A.npts;
            ti[j] = on ? s_idx[j * A.npts + tid] : 0;
            tw[j] = on ? s_w[j * A.npts + tid] : T(0);
        }
        // the column inputs (previous action, observation or given action) of the NEXT group are requested before this
        // group's gather phase: their cold-L2 latency (~2 us after a flush) then hides behind it instead of heading every group
        T prev[8], in8[8];
        auto request = [&](int env0n) {
            const int envn = env0n + ce;
            const bool onn = col_on && env0n < A.n_envs && envn < A.n_envs;
            const size_t coln = (size_t)(onn ? envn : 0) * A.n_act + (onn ? cj : 0);
            const T* acoln = A.action + coln * A.a_rows;
            const T* srcn = A.use_actor ? A.state + coln * A.obs_rows : A.actions_in + coln * A.a_rows;
            const int nin = A.use_actor ? A.obs_rows : A.a_rows;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                prev[r] = (onn && r < A.a_rows) ? acoln[r] : T(0);
                in8[r] = (onn && r < nin) ? srcn[r] : T(0);
            }
        };
        request(blockIdx.x * E);
        for (int env0 = blockIdx.x * E; env0 < A.n_envs; env0 += gridDim.x * E) {
            const int env = env0 + ce;
            T first = T(0);
            T v[8];
            const bool act_on = col_on && env < A.n_envs;
            if (act_on) {
                if (A.use_actor) {
                    // tiny runtime-shaped MLP: activations in shared memory ([unit][thread], conflict free) with plain
                    // runtime loops -- for a 1-6-1 actor that is ~130 instructions per column; a fully unrolled,
                    // predicated 8 x 8 register version issues ~500
                    float* xa = s_x + tid;
                    float* xh = s_x + (size_t)A.actor_wmax * BD + tid;
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        if (r < A.obs_rows) xa[r * BD] = (float)in8[r];
                    const float* out = mlp_forward_smem(A.actor, s_par, xa, xh, BD);
#pragma unroll
                    for (int r = 0; r < 8; ++r) v[r] = r < A.a_rows ? clamp_t<T>((T)out[r * BD], A.act_limit) : T(0);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) v[r] = in8[r];
                }
                const size_t col = (size_t)env * A.n_act + cj;
                T* acol = A.action + col * A.a_rows;
                T* dcol = A.delta_action + col * A.a_rows;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r < A.a_rows) { dcol[r] = v[r] - prev[r]; acol[r] = v[r]; }
                first = v[0];
            }
            request(env0 + gridDim.x * E);
            if (col_on) s_a[tid] = first;
            __syncthreads();
            if (tid < A.npts) {
                for (int e = 0; e < E; ++e) {
                    if (env0 + e >= A.n_envs) break;
                    const T* sa = s_a + e * A.n_act;
                    T acc = T(0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < A.act_nnz) acc += (A.power * sa[ti[j]]) * tw[j];              // ascending actuator index
                    A.p[(size_t)(env0 + e) * A.npts + tid] = acc;
                }
            }
            __syncthreads();
        }
        return;
    }
    // persistent over groups of E environments: the tables above are staged once per CTA
    for (int env0 = blockIdx.x * E; env0 < A.n_envs; env0 += gridDim.x * E) {
    if (A.use_actor && A.mono) {
        // global agent: one column per env with n_act outputs
        for (int e = tid; e < E; e += BD) {
            const int env = env0 + e;
            if (env >= A.n_envs) continue;
            float* xa = s_x + tid;
            float* xb = s_x + (size_t)A.actor_wmax * BD + tid;
            const T* scol = A.state + (size_t)env * A.obs_rows;
            for (int r = 0; r < A.obs_rows; ++r) xa[r * BD] = (float)scol[r];
            const float* out = mlp_forward_smem(A.actor, s_par, xa, xb, BD);
            const size_t abase = (size_t)env * A.n_act;
            for (int j = 0; j < A.n_act; ++j) {
                const T v = clamp_t<T>((T)out[j * BD], A.act_limit);
                A.delta_action[abase + j] = v - A.action[abase + j];
                A.action[abase + j] = v;
                s_a[e * A.n_act + j] = v;
            }
        }
    } else {
        for (int q = tid; q < E * A.n_act; q += BD) {
            const int e = q / A.n_act, j = q % A.n_act, env = env0 + e;
            if (env >= A.n_envs) { s_a[q] = T(0); continue; }
            const size_t col = (size_t)env * A.n_act + j;
            T* acol = A.action + col * A.a_rows;
            T* dcol = A.delta_action + col * A.a_rows;
            if (A.use_actor) {
                float* xa = s_x + tid;
                float* xb = s_x + (size_t)A.actor_wmax * BD + tid;
                const T* scol = A.state + col * A.obs_rows;
                for (int r = 0; r < A.obs_rows; ++r) xa[r * BD] = (float)scol[r];
                const float* out = mlp_forward_smem(A.actor, s_par, xa, xb, BD);
                for (int r = 0; r < A.a_rows; ++r) {
                    const T v = clamp_t<T>((T)out[r * BD], A.act_limit);
                    dcol[r] = v - acol[r]; acol[r] = v;
                    if (r == 0) s_a[q] = v;
                }
            } else {
                const T* icol = A.actions_in + col * A.a_rows;
                for (int r = 0; r < A.a_rows; ++r) {
                    const T v = icol[r];
                    dcol[r] = v - acol[r]; acol[r] = v;
                    if (r == 0) s_a[q] = v;
                }
            }
        }
    }
    __syncthreads();
    // p[n] = sum_i (power * a_i) * g_i[n], ascending actuator index like the reference loop
    const int* idxp = A.stage_table ? s_idx : A.act_idx;
    const T* wp = A.stage_table ? s_w : A.act_w;
    for (int q = tid; q < E * A.npts; q += BD) {
        const int e = q / A.npts, n = q % A.npts, env = env0 + e;
        if (env >= A.n_envs) continue;
        T acc = T(0);
        for (int j = 0; j < A.act_nnz; ++j) {
            const int i = idxp[(size_t)j * A.npts + n];
            acc += (A.power * s_a[e * A.n_act + i]) * wp[(size_t)j * A.npts + n];
        }
        A.p[(size_t)env * A.npts + n] = acc;
    }
    __syncthreads();                     // s_a / activations are reused by the next group
    }
}

// ---- shape-specialised actuation for the conv agent (one action row per actuator) ------------------------------
// Same arithmetic as the fast path of actuate_kernel, with the loop shapes known at compile time: NNZ taps per grid
// point held in registers (fetched straight from the global table, no staging pass), the actor either absent
// (NIN = 0: actions given), a NIN-NH-1 network evaluated in registers with its parameters hoisted out of the group
// loop, or of runtime shape through shared memory (NIN = -1).  The actions of a group go through a double-buffered
// shared line already multiplied by agent_power, so a group costs one barrier and the gather is one shared load and
// one FMA per tap.  ncu on the generic kernel (profiles/r1_ks_step_f64.md, "actuate"): 1066 warp instructions per
// warp and group, 29 % of them in the predicated 8-tap gather and 28 % in the runtime-shaped MLP loops.
template <int NIN, int NH> struct ActorRegs {
    static constexpr int NP = NIN > 0 ? NIN * NH + NH + NH + 1 : 1;
};

template <typename T, int NNZ, int NIN, int NH>
__global__ void __launch_bounds__(256, 3) actuate_conv_kernel(const __grid_constant__ ActuateArgs<T> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int E = A.envs_per_cta, tid = threadIdx.x;
    constexpr int BD = 256;
    const int ncol = E * A.n_act;                                   // <= BD
    T* s_a = reinterpret_cast<T*>(smem_raw);                        // [2][ncol]  power * action
    float* s_par = reinterpret_cast<float*>(smem_raw + (((size_t)2 * ncol * sizeof(T) + 15) & ~(size_t)15));
    float* s_x = s_par + ((A.actor_np + 3) & ~3);

    const bool pt_on = tid < A.npts;
    int tb[NNZ]; T tw[NNZ];                                         // byte offset of the tap's actuator in s_a, weight
#pragma unroll
    for (int j = 0; j < NNZ; ++j) {
        const bool on = pt_on && j < A.act_nnz;
        tb[j] = on ? __ldg(A.act_idx + j * A.npts + tid) * (int)sizeof(T) : 0;
        tw[j] = on ? __ldg(A.act_w + j * A.npts + tid) : T(0);
    }
#pragma unroll
    for (int j = 1; j < NNZ; ++j) if (j >= A.act_nnz) tb[j] = tb[0];  // padding taps: weight 0 on a tap the point already has

    float w[ActorRegs<NIN, NH>::NP];
    if (NIN != 0) {
        for (int i = tid; i < A.actor_np; i += BD) s_par[i] = A.actor.params[i];
        __syncthreads();
        if (NIN > 0) {
            // W0 (NH x NIN, column-major) | b0 | W1 (1 x NH) | b1 are contiguous in the flat parameter vector
#pragma unroll
            for (int i = 0; i < ActorRegs<NIN, NH>::NP; ++i) w[i] = s_par[i];
        }
    }
    const int act0 = A.actor.acts[0], act1 = A.actor.acts[1];

    const int ce = tid / A.n_act, cj = tid - ce * A.n_act;          // this thread's column slot (fixed for all groups)
    const bool col_on = tid < ncol;
    constexpr int NX = NIN > 0 ? NIN : (NIN < 0 ? 8 : 1);
    T prev, xin[NX], nz = T(0);
    auto request = [&](int env0n) {
        const int envn = env0n + ce;
        const bool onn = col_on && envn < A.n_envs;
        const size_t coln = onn ? (size_t)envn * A.n_act + cj : 0;
        prev = onn ? A.action[coln] : T(0);
        if (NIN != 0 && A.noise) nz = onn ? A.noise[coln] : T(0);
        if (NIN == 0) xin[0] = onn ? A.actions_in[coln] : T(0);
        else {
            const T* srcn = A.state + coln * A.obs_rows;
#pragma unroll
            for (int r = 0; r < NX; ++r) xin[r] = (onn && (NIN > 0 || r < A.obs_rows)) ? srcn[r] : T(0);
        }
    };
    request(blockIdx.x * E);
    int buf = 0;
    for (int env0 = blockIdx.x * E; env0 < A.n_envs; env0 += gridDim.x * E, buf ^= 1) {
        const int env = env0 + ce;
        T* sa = s_a + buf * ncol;
        if (col_on) {
            T v = T(0);
            if (env < A.n_envs) {
                if (NIN == 0) v = xin[0];
                else {
                    float o;
                    if (NIN > 0) {
                        float xf[NIN > 0 ? NIN : 1];
#pragma unroll
                        for (int i = 0; i < NIN; ++i) xf[i] = (float)xin[i];
                        o = actor_two_layer<(NIN > 0 ? NIN : 1), (NH > 0 ? NH : 1)>(w, xf, act0, act1);
                    } else {
                        float* xa = s_x + tid;
                        float* xh = s_x + (size_t)A.actor_wmax * BD + tid;
#pragma unroll
                        for (int r = 0; r < NX; ++r)
                            if (r < A.obs_rows) xa[r * BD] = (float)xin[r];
                        o = mlp_forward_smem(A.actor, s_par, xa, xh, BD)[0];
                    }
                    T vv = (T)o;
                    if (A.noise) vv += nz * A.act_noise;
                    v = clamp_t<T>(vv, A.act_limit);
                }
                const size_t col = (size_t)env * A.n_act + cj;
                if (NIN != 0 && A.action_in_out) A.action_in_out[col] = v;
                A.delta_action[col] = v - prev;
                A.action[col] = v;
            }
            sa[tid] = A.power * v;
        }
        request(env0 + gridDim.x * E);                  // next group's inputs travel during this group's gather
        __syncthreads();
        if (pt_on) {
            const unsigned char* sab = reinterpret_cast<const unsigned char*>(sa);
            T* pout = A.p + (size_t)env0 * A.npts + tid;
            const int ne = min(E, A.n_envs - env0);
            for (int e = 0; e < ne; ++e) {
                T a[NNZ];
#pragma unroll
                for (int j = 0; j < NNZ; ++j) a[j] = *reinterpret_cast<const T*>(sab + tb[j]);
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < NNZ; ++j) acc = fma(a[j], tw[j], acc);          // ascending actuator index
                *pout = acc;
                sab += A.n_act * sizeof(T); pout += A.npts;
            }
        }
    }
}

template <typename T>
struct ObserveArgs {
    ObsRewardParams<T> P;
    int n_envs;
    int fresh;                     // 1: reset!/constructor semantics (no clock advance, zero reward)
    const uint8_t* mask;           // optional per-env mask (reset)
    const T* sensors;              // [B][fields][n_sensors] raw dots
    const T* vmax;                 // [B] max |y| (CHECK_Y)
    T* state; T* action; T* delta_action; T* action_in; T* reward;
    uint8_t* done; double* time; int* steps; double* reward_sum;
    const int* list; const int* list_n;   // optional: only these environments (see observe_kernel)
};

// One WARP per environment: lanes stride over the actuator columns, the sensor dots are read straight from global
// memory (512 B per environment, L1-resident after the first touch), the per-environment (sum, max) of the column
// rewards is a shuffle reduction and lane 0 carries the clock / done flag, whose inputs it requested up front.  No
// shared memory and no CTA barrier: the earlier one-CTA-per-environment version serialised three DRAM round trips
// (sensors -> barrier -> columns -> barrier -> scalars) per CTA.
// PLAIN = the 1-D conv agent without temporal stacking or action memory (one field, window <= n_sensors): the column
// loop then has no runtime-shaped inner loops and no integer division (every shipped KS script).
template <typename T, bool PLAIN>
__device__ __forceinline__ void observe_env(const ObserveArgs<T>& A, const int env, const int lane) {
    const ObsRewardParams<T>& P = A.P;
    if (A.mask && !A.mask[env]) return;
    const bool fresh = A.fresh != 0;
    double tm0 = 0.0; T vm = T(0); int st = 0; double rs0 = 0.0;
    if (lane == 0 && !fresh) {
        tm0 = A.time[env]; st = A.steps[env];
        if (P.check_max == 1) vm = A.vmax[env];
        if (A.reward_sum) rs0 = A.reward_sum[env];
    }
    const T* sens = A.sensors + (size_t)env * P.fields * P.n_sensors;
    auto sv = [&](int f, int i) { return sens[f * P.n_sensors + i]; };
    T racc = T(0), rmax = T(0);
    for (int j = lane; j < P.n_act; j += 32) {
        const size_t col = (size_t)env * P.n_act + j;
        T* acol = A.action + col * P.a_rows;
        T* dcol = A.delta_action + col * P.a_rows;
        if (fresh)
            for (int r = 0; r < P.a_rows; ++r) { acol[r] = T(0); dcol[r] = T(0); A.action_in[col * P.a_rows + r] = T(0); }
        const T a0 = fresh ? T(0) : acol[0], d0 = fresh ? T(0) : dcol[0];
        T rj;
        if (PLAIN) {
            const int m = P.a2s[j], ns = P.n_sensors, h = P.window / 2;
            const T sm = sens[m];
            T* scol = A.state + col * P.window;                       // obs_rows == window
            for (int r = 0; r < P.window; ++r) {
                int i = m - (r - h);                                  // circshift(v, i)[j] == v[j - i]
                i = i < 0 ? i + ns : (i >= ns ? i - ns : i);
                scol[r] = (r == h ? sm : sens[i]) * P.obs_scale;
            }
            const T raw = sm - P.r_offset * P.sens_sum[m];
            T s;
            if (P.r_pow == T(2)) { const T g = P.r_gain * raw; s = g * g / P.r_div; }
            else s = pow_t<T>(fabs(P.r_gain * raw), P.r_pow) / P.r_div;
            rj = -fabs(s) - P.a_pun * a0 * a0 - P.da_pun * d0 * d0;
            A.reward[col] = fresh ? T(0) : rj;
        } else if (!P.mono) {
            rj = assemble_column<T>(P, sv, j, a0, d0, acol, A.state + col * P.obs_rows, fresh);
            A.reward[col] = fresh ? T(0) : rj;
        } else {
            const int m = P.a2s[j];
            const T raw = sv(0, m) - P.r_offset * P.sens_sum[m];
            const T s = pow_t<T>(fabs(P.r_gain * raw), P.r_pow) / P.r_div;
            rj = -fabs(s) - P.a_pun * a0 * a0 - P.da_pun * d0 * d0;
        }
        racc += rj; rmax = fmax(rmax, fabs(rj));
    }
    if (!PLAIN && P.mono) {
        // state = reshape(sensors, (n_sensors, 1)) [+ temporal stacking], KSglobalSetup.jl:222-238
        T* scol = A.state + (size_t)env * P.obs_rows;
        if (P.temporal > 1 && !fresh) {
            if (lane == 0)
                for (int r = P.obs_rows - P.memory - 1; r >= P.n_sensors; --r) scol[r] = scol[r - P.n_sensors];
            __syncwarp();
        }
        for (int i = lane; i < P.n_sensors; i += 32) {
            const T v = sens[i] * P.obs_scale;
            scol[i] = v;
            if (fresh) for (int k = 1; k < P.temporal; ++k) scol[k * P.n_sensors + i] = v;
        }
        for (int k = lane; k < P.memory; k += 32) scol[P.obs_rows - P.memory + k] = T(0);
    }
    // (sum, max) over the columns; lanes hold partial sums of columns lane, lane + 32, ...
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        racc += __shfl_xor_sync(0xffffffffu, racc, o);
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    }
    if (lane == 0) {
        if (fresh) {
            if (P.mono) A.reward[env] = T(0);
            A.done[env] = 0; A.time[env] = 0.0; A.steps[env] = 0;
        } else {
            const T rmean = racc / T(P.n_act);
            T mx = rmax;
            if (P.mono) { A.reward[env] = rmean; mx = fabs(rmean); }
            const double tm = tm0 + P.dt;                         // env.time += env.dt (Float64, quirk Q7)
            bool dn = tm >= P.te;
            if (P.check_max == 1) dn = dn || (vm > P.max_value);
            else if (P.check_max == 2) dn = dn || (mx > P.max_value);
            A.done[env] = dn ? 1 : 0;
            A.time[env] = tm;
            A.steps[env] = st + 1;
            if (A.reward_sum) A.reward_sum[env] = rs0 + (double)rmean;
        }
    }
}

// one warp per environment; A.list != nullptr: the warps of a small grid walk the environments list[0 .. *list_n)
// (pdeb200_reset_diverged: usually none -- a full-size grid of early-exiting CTAs costs ~8 us, this one ~2)
template <typename T, bool PLAIN>
__global__ void __launch_bounds__(128) observe_kernel(const __grid_constant__ ObserveArgs<T> A) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (A.list) {
        const int n = *A.list_n, nw = gridDim.x * (blockDim.x >> 5);
        for (int k = w; k < n; k += nw) observe_env<T, PLAIN>(A, A.list[k], lane);
        return;
    }
    if (w < A.n_envs) observe_env<T, PLAIN>(A, w, lane);
}

// Sensor dots straight from physical fields in global memory (reset! / set_state path; the core
// kernels produce them from on-chip state).  One CTA per environment.
template <typename T>
__global__ void __launch_bounds__(128) sensors_phys_kernel(int fields, int npts, int n_sensors, EllTable<T> sens,
                                                           const uint8_t* __restrict__ mask, const T* __restrict__ y,
                                                           int interleaved, T* sensors_out, T* vmax_out,
                                                           const int* __restrict__ list = nullptr,
                                                           const int* __restrict__ list_n = nullptr) {
    // list != nullptr: a small grid walks the environments list[0 .. *list_n) (pdeb200_reset_diverged: usually none)
    const int n_items = list ? *list_n : (int)gridDim.x;
    for (int k = blockIdx.x; k < n_items; k += gridDim.x) {
        const int env = list ? list[k] : k;
        if (mask && !mask[env]) continue;
        const T* ye = y + (size_t)env * fields * npts;
        for (int q = threadIdx.x; q < fields * n_sensors; q += blockDim.x) {
            const int f = q / n_sensors, i = q % n_sensors;
            T acc = T(0);
            for (int j = 0; j < sens.nnz_max; ++j) {
                const int n = sens.idx[j * n_sensors + i];
                acc += (interleaved ? ye[n * fields + f] : ye[f * npts + n]) * sens.w[j * n_sensors + i];
            }
            sensors_out[(size_t)env * fields * n_sensors + q] = acc;
        }
        if (vmax_out && threadIdx.x == 0) vmax_out[env] = T(0);
    }
}

}  // namespace pdeb200
