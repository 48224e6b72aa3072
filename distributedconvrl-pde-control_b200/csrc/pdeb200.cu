// libpdeb200.so -- C ABI (include/pdeb200.h): context, constants, environment entry points.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "glue.cuh"

namespace pdeb200 {

thread_local std::string g_last_error;

namespace {

constexpr int kListGrid = 296;      // CTAs of the list-driven reset kernels (two per SM)

// reset!: y = y0, p = prepare_action(action0) = 0 for the masked environments (src/PDEenv.jl:183-193)
template <typename T>
__global__ void reset_copy_kernel(int y_elems, int p_elems, const uint8_t* __restrict__ mask, const T* __restrict__ y0,
                                  T* y, T* p, const int* __restrict__ list, const int* __restrict__ list_n) {
    // list != nullptr: a small grid walks the environments list[0 .. *list_n)
    const int n_items = list ? *list_n : (int)gridDim.x;
    for (int k = blockIdx.x; k < n_items; k += gridDim.x) {
        const int env = list ? list[k] : k;
        if (mask && !mask[env]) continue;
        for (int i = threadIdx.x; i < y_elems; i += blockDim.x) y[(size_t)env * y_elems + i] = y0[(size_t)env * y_elems + i];
        for (int i = threadIdx.x; i < p_elems; i += blockDim.x) p[(size_t)env * p_elems + i] = T(0);
    }
}

// Batched termination: mask[b] = done[b] although the clock of b has not reached te, i.e. b diverged (PDEenv.jl:226-237);
// counts = {done, time limit, diverged}
// list[0 .. counts[2]) = the diverged environments (any order): the reset kernels that follow walk it with small grids
__global__ void diverged_mask_kernel(int n_envs, const uint8_t* __restrict__ done, const double* __restrict__ time, double te,
                                     uint8_t* mask, int* counts, int* list) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_envs) return;
    const bool dn = done[b] != 0, tl = time[b] >= te;
    mask[b] = dn && !tl;
    if (dn) atomicAdd(counts + 0, 1);
    if (dn && tl) atomicAdd(counts + 1, 1);
    if (dn && !tl) list[atomicAdd(counts + 2, 1)] = b;
}

// ------------------------------------------------------------------------------------------------
// Policy forward: actions = clamp(actor(state) [+ noise * act_noise], +-act_limit)
// src/PDEagent.jl:189, 201-204.  One thread per column.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
    const uint64_t p = (uint64_t)a * b; *hi = (uint32_t)(p >> 32); return (uint32_t)p;
}
// Philox4x32-10 counter RNG (public algorithm, Salmon et al. 2011) -> two standard normals.
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint64_t ctr, float* n0, float* n1) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, h1;
        const uint32_t l0 = mulhilo(0xD2511F53u, c0, &h0), l1 = mulhilo(0xCD9E8D57u, c2, &h1);
        const uint32_t n0_ = h1 ^ c1 ^ k0, n2_ = h0 ^ c3 ^ k1;
        c0 = n0_; c1 = l1; c2 = n2_; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const float u0 = ((float)c0 + 0.5f) * 2.3283064365386963e-10f;
    const float u1 = ((float)c1 + 0.5f) * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.f * logf(u0));
    float s, co;
    sincosf(6.283185307179586f * u1, &s, &co);
    *n0 = rad * co; *n1 = rad * s;
}

// Per-column actor forward.  Activations live in shared memory as [unit][thread] (conflict free) and the layers are
// plain runtime loops: for the tiny runtime-shaped networks of this path (1-6-1 ... 12-20-1) that is ~4x fewer issued
// instructions than a register version that has to be unrolled (and predicated) to a compile-time maximum width.
// NIN > 0: NIN-NH-1 actor in registers (one output row: the conv agent of every shipped KS script); NIN = -1: runtime shape
template <typename T, int NIN, int NH>
__global__ void __launch_bounds__(128) policy_kernel(NetDev net, int n_params, int wmax, int n_columns, int obs_rows, int a_rows,
                                                     int memory, const T* __restrict__ state, T* __restrict__ action_in,
                                                     const T* __restrict__ noise, int use_rng, uint64_t seed, uint64_t offset,
                                                     T act_noise, T act_limit) {
    extern __shared__ float s_dyn[];
    float* s_par = s_dyn;
    float* s_x = s_dyn + n_params;                                   // [2][wmax][blockDim]
    const int BD = blockDim.x;
    for (int i = threadIdx.x; i < n_params; i += BD) s_par[i] = net.params[i];
    __syncthreads();
    const int col = blockIdx.x * BD + threadIdx.x;
    if (col >= n_columns) return;
    float one[1];
    const float* out = one;
    int ostride = 1;
    if (NIN > 0) {
        constexpr int NP = NIN > 0 ? NIN * NH + 2 * NH + 1 : 1;
        float w[NP], x[NIN > 0 ? NIN : 1];
#pragma unroll
        for (int i = 0; i < NP; ++i) w[i] = s_par[i];
#pragma unroll
        for (int r = 0; r < NIN; ++r) x[r] = (float)state[(size_t)col * (NIN > 0 ? NIN : 1) + r];
        one[0] = actor_two_layer<(NIN > 0 ? NIN : 1), (NH > 0 ? NH : 1)>(w, x, net.acts[0], net.acts[1]);
    } else {
        float* xa = s_x + threadIdx.x;
        float* xh = s_x + (size_t)wmax * BD + threadIdx.x;
        for (int r = 0; r < obs_rows; ++r) xa[r * BD] = (float)state[(size_t)col * obs_rows + r];
        out = mlp_forward_smem(net, s_par, xa, xh, BD);
        ostride = BD;
    }
    const int n_out = net.sizes[net.n_layers];            // == a_rows (conv agent) or n_act (mono)
    const int noisy = n_out - memory;
    for (int r = 0; r < n_out; ++r) {
        T v = (T)out[r * ostride];
        if (r < noisy) {
            if (noise) v += noise[(size_t)col * noisy + r] * act_noise;
            else if (use_rng) {
                float g0, g1;
                philox_normal2(seed, offset + (uint64_t)col * noisy + r, &g0, &g1);
                v += (T)g0 * act_noise;
            }
        }
        action_in[(size_t)col * n_out + r] = clamp_t<T>(v, act_limit);
    }
}

template <typename T>
int32_t build_ell(pdeb200_ctx* c, const std::vector<std::vector<std::pair<int, double>>>& rows, EllHost* out) {
    int nnz = 1;
    for (auto& r : rows) nnz = std::max(nnz, (int)r.size());
    const int n = (int)rows.size();
    std::vector<int> idx((size_t)nnz * n, 0);
    std::vector<T> w((size_t)nnz * n, T(0));
    for (int i = 0; i < n; ++i) {
        const int pad = rows[i].empty() ? 0 : rows[i][0].first;
        for (int j = 0; j < nnz; ++j) {
            if (j < (int)rows[i].size()) { idx[(size_t)j * n + i] = rows[i][j].first; w[(size_t)j * n + i] = (T)rows[i][j].second; }
            else idx[(size_t)j * n + i] = pad;
        }
    }
    if (out->d_idx) cudaFree(out->d_idx);
    if (out->d_w) cudaFree(out->d_w);
    PDEB_CUDA(c, cudaMalloc(&out->d_idx, idx.size() * sizeof(int)));
    PDEB_CUDA(c, cudaMalloc(&out->d_w, w.size() * sizeof(T)));
    PDEB_CUDA(c, cudaMemcpy(out->d_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    PDEB_CUDA(c, cudaMemcpy(out->d_w, w.data(), w.size() * sizeof(T), cudaMemcpyHostToDevice));
    out->nnz_max = nnz; out->n_rows = n;
    return PDEB200_OK;
}

template <typename T>
int32_t set_bases_t(pdeb200_ctx* c, const double* sb, const double* ab, const int32_t* a2s, double tol) {
    const pdeb200_config& g = c->cfg;
    const int np = c->npts;
    std::vector<std::vector<std::pair<int, double>>> srows(g.n_sensors), arows(np);
    std::vector<T> ssum(g.n_sensors);
    for (int i = 0; i < g.n_sensors; ++i) {
        double mx = 0, sum = 0;
        for (int n = 0; n < np; ++n) { mx = std::max(mx, std::fabs(sb[(size_t)i * np + n])); sum += sb[(size_t)i * np + n]; }
        ssum[i] = (T)sum;
        for (int n = 0; n < np; ++n) {
            const double w = sb[(size_t)i * np + n];
            if (w != 0.0 && std::fabs(w) > tol * mx) srows[i].push_back({n, w});
        }
    }
    for (int i = 0; i < g.n_actuators; ++i) {
        double mx = 0;
        for (int n = 0; n < np; ++n) mx = std::max(mx, std::fabs(ab[(size_t)i * np + n]));
        for (int n = 0; n < np; ++n) {
            const double w = ab[(size_t)i * np + n];
            if (w != 0.0 && std::fabs(w) > tol * mx) arows[n].push_back({i, w});   // ascending actuator index
        }
    }
    int32_t rc;
    if ((rc = build_ell<T>(c, srows, &c->sens))) return rc;
    if ((rc = build_ell<T>(c, arows, &c->actT))) return rc;
    std::vector<int> a(g.n_actuators);
    for (int i = 0; i < g.n_actuators; ++i) {
        if (a2s[i] < 0 || a2s[i] >= g.n_sensors) return fail(c, PDEB200_EINVAL, "set_bases: a2s out of range (0-based)");
        a[i] = a2s[i];
    }
    PDEB_CUDA(c, cudaMemcpy(c->d_a2s, a.data(), a.size() * sizeof(int), cudaMemcpyHostToDevice));
    PDEB_CUDA(c, cudaMemcpy(c->d_sens_sum, ssum.data(), ssum.size() * sizeof(T), cudaMemcpyHostToDevice));
    c->bases_set = true;
    c->ks_layout_dirty = true;
    if (c->cfg.problem == PDEB200_KS) return ks_bases_changed(c, sb);
    return PDEB200_OK;
}

template <typename T>
int32_t observe_t(pdeb200_ctx* c, int fresh, const uint8_t* d_mask, double* d_rsum, const int* list = nullptr,
                  const int* list_n = nullptr) {
    ObserveArgs<T> O;
    O.list = list; O.list_n = list_n;
    O.P = make_obs_params<T>(c);
    O.n_envs = c->cfg.n_envs; O.fresh = fresh; O.mask = d_mask; O.sensors = (const T*)c->sensors; O.vmax = (const T*)c->vmax;
    O.state = (T*)c->state; O.action = (T*)c->action; O.delta_action = (T*)c->delta_action; O.action_in = (T*)c->action_in;
    O.reward = (T*)c->reward; O.done = c->done; O.time = c->time; O.steps = c->steps; O.reward_sum = d_rsum;
    const int tpb = 128;                                          // one warp per environment
    const bool plain = !O.P.mono && O.P.spa == 0 && O.P.temporal == 1 && O.P.memory == 0 && O.P.fields == 1 &&
                       O.P.a_rows == 1 && O.P.window <= O.P.n_sensors && O.P.obs_rows == O.P.window;
    auto kern = plain ? observe_kernel<T, true> : observe_kernel<T, false>;
    const int full = (c->cfg.n_envs + tpb / 32 - 1) / (tpb / 32);
    kern<<<list ? std::min(full, kListGrid) : full, tpb, 0, c->stream>>>(O);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

template <typename T> using ActuateFn = void (*)(ActuateArgs<T>);

// mode: 0 actions given, 1 / 3: 1-6-1 / 3-6-1 actor in registers, -1: actor of runtime shape through shared memory
template <typename T, int NNZ>
ActuateFn<T> conv_kernel_for(int mode) {
    switch (mode) {
        case 0: return actuate_conv_kernel<T, NNZ, 0, 0>;
        case 1: return actuate_conv_kernel<T, NNZ, 1, 6>;
        case 3: return actuate_conv_kernel<T, NNZ, 3, 6>;
        default: return actuate_conv_kernel<T, NNZ, -1, 0>;
    }
}
template <typename T>
ActuateFn<T> conv_kernel_for(int nnz, int mode) {
    if (nnz <= 2) return conv_kernel_for<T, 2>(mode);
    if (nnz <= 4) return conv_kernel_for<T, 4>(mode);
    if (nnz <= 6) return conv_kernel_for<T, 6>(mode);
    return conv_kernel_for<T, 8>(mode);
}

template <typename T>
int32_t actuate_t(pdeb200_ctx* c, const void* actions_dev, int use_actor, double act_limit, const void* d_noise = nullptr,
                  double act_noise = 0.0, bool want_policy_out = false) {
    ActuateArgs<T> A;
    A.noise = nullptr; A.act_noise = T(0); A.action_in_out = nullptr;
    const HostNet& net = c->nets[PDEB200_NET_BEHAVIOR_ACTOR];
    A.n_envs = c->cfg.n_envs;
    A.n_act = c->cfg.n_actuators; A.a_rows = c->a_rows; A.obs_rows = c->obs_rows; A.mono = c->cfg.mono;
    A.npts = c->npts; A.use_actor = use_actor;
    A.actor_np = use_actor ? net.n_params : 0;
    A.actor_wmax = 1;
    if (use_actor) for (int l = 0; l <= net.n_layers; ++l) A.actor_wmax = std::max(A.actor_wmax, net.sizes[l]);
    A.act_idx = c->actT.d_idx; A.act_w = (const T*)c->actT.d_w; A.act_nnz = c->actT.nnz_max;
    A.power = (T)c->cfg.agent_power; A.act_limit = (T)act_limit;
    A.actor = net.dev();
    A.actions_in = (const T*)actions_dev; A.state = (const T*)c->state;
    A.action = (T*)c->action; A.delta_action = (T*)c->delta_action;
    // NS keeps env.p spectral; its physical-space sum goes to a scratch field owned by the NS back-end
    A.p = (T*)(c->cfg.problem == PDEB200_NS2D ? c->prob_p_phys : c->p);
    const int tpb = 256;
    // several environments per CTA so that the column phase fills the block and the table is staged once
    int E = std::max(1, std::min(8, tpb / std::max(1, c->cfg.n_actuators)));
    E = std::min(E, std::max(1, 4096 / std::max(1, c->npts)));
    A.envs_per_cta = E;
    const size_t tab = (size_t)A.act_nnz * A.npts * (sizeof(T) + sizeof(int));
    A.stage_table = tab <= 64 * 1024;
    // the conv agent with one action row per actuator (every shipped KS / Keller-Segel 1-D / NS script) runs the
    // shape-specialised kernel; everything else (global agent, action memory rows, > 8 taps, > 256 points) the generic one
    ActuateFn<T> kern = actuate_kernel<T>;
    size_t smem = actuate_smem_bytes<T>(E, A.n_act, A.npts, A.act_nnz, A.stage_table, use_actor, A.actor_np,
                                        A.actor_wmax, tpb);
    static const bool generic_only = [] { const char* e = getenv("PDEB200_ACTUATE_GENERIC"); return e && atoi(e) != 0; }();
    if (!generic_only && !A.mono && A.a_rows == 1 && E * A.n_act <= tpb && A.npts <= tpb && A.act_nnz <= 8 &&
        (!use_actor || (A.obs_rows <= 8 && net.sizes[net.n_layers] == 1))) {
        const bool two = use_actor && net.n_layers == 2 && net.offs[0] == 0 &&
                         net.offs[1] == net.sizes[0] * net.sizes[1] + net.sizes[1] && net.sizes[1] == 6;
        const int mode = !use_actor ? 0 : (two && net.sizes[0] == 1) ? 1 : (two && net.sizes[0] == 3) ? 3 : -1;
        kern = conv_kernel_for<T>(A.act_nnz, mode);
        smem = (((size_t)2 * E * A.n_act * sizeof(T) + 15) & ~(size_t)15) + (size_t)((A.actor_np + 3) & ~3) * sizeof(float) +
               (mode == -1 ? (size_t)2 * A.actor_wmax * tpb * sizeof(float) : 0);
        if (use_actor && want_policy_out) {
            A.noise = (const T*)d_noise; A.act_noise = (T)act_noise; A.action_in_out = (T*)c->action_in;
        }
    } else if (want_policy_out) {
        return 1;                        // the policy-with-noise fusion exists in the shape-specialised kernel only: caller falls back
    }
    if (smem > 200 * 1024) return fail(c, PDEB200_EUNSUPPORTED, "actuate: actor too large for shared memory");
    PDEB_CUDA(c, ensure_dyn_smem(kern, smem, c->device));
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
    const int groups = (c->cfg.n_envs + E - 1) / E;
    // CTAs that are really co-resident (registers AND shared memory): a grid one CTA larger than a whole wave costs a
    // full extra CTA lifetime
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tpb, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    // whole waves of persistent CTAs: every CTA gets the same number of groups when the batch allows it
    const int resident = n_sm * per_sm;
    const int rounds = (groups + resident - 1) / resident;
    const int grid = std::min(groups, (groups + rounds - 1) / rounds);
    kern<<<grid, tpb, smem, c->stream>>>(A);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

template <typename T>
int32_t reset_t(pdeb200_ctx* c, const uint8_t* d_mask, const int* list = nullptr, const int* list_n = nullptr) {
    if (c->cfg.problem == PDEB200_NS2D) list = list_n = nullptr;       // the NS sensor path (2-D transforms) is mask driven
    const int lg = list ? std::min(c->cfg.n_envs, kListGrid) : c->cfg.n_envs;
    reset_copy_kernel<T><<<lg, 256, 0, c->stream>>>(c->y_elems, c->p_elems, d_mask, (const T*)c->y0,
                                                   (T*)c->y, (T*)c->p, list, list_n);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    if (c->cfg.problem == PDEB200_NS2D) {
        int32_t rc = ns_sensors(c, d_mask);
        if (rc) return rc;
    } else {
        sensors_phys_kernel<T><<<lg, 128, 0, c->stream>>>(
            c->fields, c->npts, c->cfg.n_sensors,
            EllTable<T>{c->sens.d_idx, (const T*)c->sens.d_w, c->sens.nnz_max, c->sens.n_rows}, d_mask, (const T*)c->y,
            (c->cfg.problem == PDEB200_KSEG1D || c->cfg.problem == PDEB200_KSEG2D) ? 1 : 0, (T*)c->sensors, (T*)c->vmax,
            list, list_n);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 1;
    }
    return observe_t<T>(c, 1, d_mask, nullptr, list, list_n);
}

int32_t core_step(pdeb200_ctx* c) {
    switch (c->cfg.problem) {
        case PDEB200_KS: return ks_core(c);
        case PDEB200_KSEG1D: return kseg_core(c);
        case PDEB200_NS2D: return ns_core(c);
        case PDEB200_KSEG2D: return kseg2d_core(c);
    }
    return fail(c, PDEB200_EINVAL, "bad problem");
}

// env(action) for all environments: actuate -> core -> observe, `n_steps` times (n_steps > 1 only with
// the fused actor).  Everything is enqueued on the context's stream; no host synchronisation.
int32_t do_step(pdeb200_ctx* c, const void* actions_dev, int n_steps, int use_actor, double act_limit, double* d_rsum,
                const void* d_noise = nullptr, double act_noise = 0.0, bool policy_fused = false) {
    if (!c->bases_set) return fail(c, PDEB200_ESTATE, "step: call pdeb200_set_bases first");
    if (!c->y0_set) return fail(c, PDEB200_ESTATE, "step: call pdeb200_set_y0 + pdeb200_reset first");
    if (use_actor) {
        const HostNet& a = c->nets[PDEB200_NET_BEHAVIOR_ACTOR];
        if (!a.n_layers) return fail(c, PDEB200_ESTATE, "rollout: behavior actor not set");
        const int expect_out = c->cfg.mono ? c->cfg.n_actuators * c->a_rows : c->a_rows;
        if (a.sizes[0] != c->obs_rows || a.sizes[a.n_layers] != expect_out)
            return fail(c, PDEB200_EINVAL, "rollout: actor in/out sizes do not match the env's state/action spaces");
        for (int l = 0; l <= a.n_layers; ++l)
            if (a.sizes[l] > kFusedActorMaxWidth) return fail(c, PDEB200_EUNSUPPORTED, "fused actor: layer wider than 64");
    }
    const bool f64 = c->cfg.dtype == PDEB200_F64;
    if (c->timing) cudaEventRecord(c->ev0, c->stream);
    for (int s = 0; s < n_steps; ++s) {
        int32_t rc = f64 ? actuate_t<double>(c, actions_dev, use_actor, act_limit, d_noise, act_noise, policy_fused)
                         : actuate_t<float>(c, actions_dev, use_actor, act_limit, d_noise, act_noise, policy_fused);
        if (rc) return rc;               // 1: policy fusion not available for this shape (nothing was launched)
        const bool tc = c->timing && s == n_steps - 1;            // CUDA events around the core (dominant) kernel
        if (tc) cudaEventRecord(c->evc0, c->stream);
        if ((rc = core_step(c))) return rc;
        if (tc) cudaEventRecord(c->evc1, c->stream);
        rc = f64 ? observe_t<double>(c, 0, nullptr, d_rsum) : observe_t<float>(c, 0, nullptr, d_rsum);
        if (rc) return rc;
    }
    if (c->timing) { cudaEventRecord(c->ev1, c->stream); c->timed = true; }
    return PDEB200_OK;
}

// FMA micro-benchmark: measured CUDA-core peak for the roofline denominators (MEASURED_PEAKS.json has HBM and
// bf16 tensor numbers only).  8 independent chains per thread, 2 flops per FMA.
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, T seed, T* out) {
    T a0 = seed, a1 = seed + T(1), a2 = seed + T(2), a3 = seed + T(3), a4 = seed + T(4), a5 = seed + T(5), a6 = seed + T(6),
      a7 = seed + T(7);
    const T m = T(0.999999), c = T(1e-6) * (T)threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    }
    const T r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == T(-1)) out[0] = r;                       // never true: keeps the chains alive
}

struct ArrInfo { void* ptr; size_t bytes; };
ArrInfo arr_info(pdeb200_ctx* c, int which) {
    const size_t B = c->cfg.n_envs, e = c->esz;
    switch (which) {
        case PDEB200_ARR_Y: return {c->y, B * c->y_elems * e};
        case PDEB200_ARR_Y0: return {c->y0, B * c->y_elems * e};
        case PDEB200_ARR_P: return {c->p, B * c->p_elems * e};
        case PDEB200_ARR_STATE: return {c->state, B * c->n_cols * c->obs_rows * e};
        case PDEB200_ARR_ACTION: return {c->action, B * c->cfg.n_actuators * c->a_rows * e};
        case PDEB200_ARR_ACTION_IN: return {c->action_in, B * c->cfg.n_actuators * c->a_rows * e};
        case PDEB200_ARR_DELTA_ACTION: return {c->delta_action, B * c->cfg.n_actuators * c->a_rows * e};
        case PDEB200_ARR_REWARD: return {c->reward, B * c->n_rew * e};
        case PDEB200_ARR_DONE: return {c->done, B};
        case PDEB200_ARR_TIME: return {c->time, B * sizeof(double)};
        case PDEB200_ARR_STEPS: return {c->steps, B * sizeof(int)};
        case PDEB200_ARR_GRADS: return {c->d_grads, (size_t)c->n_grads * sizeof(float)};
        case PDEB200_ARR_LOSSES: return {c->d_losses, 2 * sizeof(float)};
        case PDEB200_ARR_STATS: return {agent_stats(c), 8 * sizeof(double)};
        case PDEB200_ARR_SENSORS: return {c->sensors, B * c->fields * c->cfg.n_sensors * e};
        case PDEB200_ARR_NSUB: return {c->d_nsub, B * 2 * sizeof(int)};
    }
    return {nullptr, 0};
}

}  // namespace
}  // namespace pdeb200

using namespace pdeb200;

extern "C" {

int32_t pdeb200_abi_version(void) { return PDEB200_ABI_VERSION; }

int32_t pdeb200_default_config(int32_t problem, pdeb200_config* cfg) {
    if (!cfg) return fail(nullptr, PDEB200_EINVAL, "default_config: null cfg");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (int32_t)sizeof(pdeb200_config);
    cfg->problem = problem; cfg->dtype = PDEB200_F64; cfg->ny = 1; cfg->n_envs = 1;
    cfg->temporal_steps = 1; cfg->memory_size = 0; cfg->mono = 0; cfg->ifpad = 1;
    cfg->adaptive = 0; cfg->rtol = 1e-8; cfg->atol = 1e-8;
    cfg->Ly = 1.0; cfg->t0 = 0.0;
    switch (problem) {
        case PDEB200_KS:          // scripts/KS/setup/KSSetup.jl:20-51, 162-184, 201
            cfg->nx = 240; cfg->Lx = 200.0; cfg->dt = 0.1; cfg->te = 5.0; cfg->oversampling = 30;
            cfg->window_size = 1; cfg->check_max_value = PDEB200_CHECK_Y; cfg->max_value = 30.0;
            cfg->agent_power = 7.5; cfg->obs_scale = 1.0 / 30.0; cfg->reward_gain = 6.0; cfg->reward_pow = 1.3;
            cfg->reward_div = 90.0; cfg->reward_offset = 0.0; cfg->action_punish = 0.002; cfg->delta_action_punish = 0.002;
            break;
        case PDEB200_KSEG1D:      // scripts/Keller-Segel/setup/KellerSegelSetup.jl:26-57, 241-263, 276
            cfg->nx = 100; cfg->Lx = 10.0; cfg->dt = 0.006; cfg->te = 8.0; cfg->oversampling = 40;
            cfg->window_size = 3; cfg->temporal_steps = 2; cfg->check_max_value = PDEB200_CHECK_Y; cfg->max_value = 20.0;
            cfg->agent_power = 10.0; cfg->obs_scale = 0.25; cfg->reward_gain = 1.0; cfg->reward_pow = 2.0;
            cfg->reward_div = 800.0; cfg->reward_offset = 1.0; cfg->action_punish = 0.0; cfg->delta_action_punish = 0.0;
            break;
        case PDEB200_NS2D:        // scripts/Fluid/setup/FluidSetup.jl:28-77, 188-217
            cfg->nx = 128; cfg->ny = 128; cfg->Lx = 1.0; cfg->Ly = 1.0; cfg->dt = 0.02; cfg->te = 6.0;
            cfg->oversampling = 40; cfg->nu = 0.00005; cfg->window_size = 3; cfg->sensors_per_axis = 16;
            cfg->check_max_value = PDEB200_CHECK_REWARD; cfg->max_value = 3.0; cfg->agent_power = 70.0;
            cfg->obs_scale = 1.0 / 70.0; cfg->reward_gain = 1.0; cfg->reward_pow = 1.1; cfg->reward_div = 320.0;
            cfg->action_punish = 0.002; cfg->delta_action_punish = 0.002;
            break;
        case PDEB200_KSEG2D:      // 2-D generalisation of the Keller-Segel setup (BASELINE config 3)
            cfg->nx = 128; cfg->ny = 128; cfg->Lx = 12.8; cfg->Ly = 12.8; cfg->dt = 0.006; cfg->te = 8.0; cfg->oversampling = 40;
            cfg->window_size = 3; cfg->temporal_steps = 2; cfg->sensors_per_axis = 16;
            cfg->check_max_value = PDEB200_CHECK_Y; cfg->max_value = 20.0;
            cfg->agent_power = 10.0; cfg->obs_scale = 1.0 / 20.0; cfg->reward_gain = 1.0; cfg->reward_pow = 2.0;
            cfg->reward_div = 20000.0; cfg->reward_offset = 1.0; cfg->action_punish = 0.0; cfg->delta_action_punish = 0.0;
            break;
        default: return fail(nullptr, PDEB200_EINVAL, "default_config: unknown problem");
    }
    return PDEB200_OK;
}

int32_t pdeb200_create(const pdeb200_config* cfg, int32_t device, pdeb200_ctx** out) {
    if (!cfg || !out) return fail(nullptr, PDEB200_EINVAL, "create: null argument");
    if (cfg->struct_size != (int32_t)sizeof(pdeb200_config))
        return fail(nullptr, PDEB200_EINVAL, "create: pdeb200_config.struct_size mismatch (ABI)");
    if (cfg->n_envs < 1 || cfg->nx < 1 || cfg->n_sensors < 1 || cfg->n_actuators < 1 || cfg->window_size < 1 ||
        cfg->window_size % 2 == 0 || cfg->temporal_steps < 1 || cfg->memory_size < 0)
        return fail(nullptr, PDEB200_EINVAL, "create: bad sizes (n_envs/nx/n_sensors/n_actuators/window/temporal/memory)");
    if (cfg->dtype != PDEB200_F32 && cfg->dtype != PDEB200_F64) return fail(nullptr, PDEB200_EINVAL, "create: bad dtype");
    if (cfg->adaptive && cfg->problem != PDEB200_KSEG1D && cfg->problem != PDEB200_NS2D)
        return fail(nullptr, PDEB200_EUNSUPPORTED, "create: adaptive = 1 exists for PDEB200_KSEG1D and PDEB200_NS2D only");
    if (cfg->adaptive && !(cfg->rtol > 0.0 && cfg->atol > 0.0)) return fail(nullptr, PDEB200_EINVAL, "create: adaptive needs rtol, atol > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, PDEB200_ECUDA, "create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, PDEB200_EINVAL, "create: bad device index");
    pdeb200_ctx* c = new pdeb200_ctx();
    c->cfg = *cfg; c->device = device;
    auto bail = [&](int32_t rc) { std::string m = c->err; pdeb200_destroy(c); g_last_error = m; return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(fail(c, PDEB200_ECUDA, "cudaSetDevice failed"));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) return bail(fail(c, PDEB200_EUNSUPPORTED, "pdeb200 kernels are built for sm_100a (B200) only"));
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(c, PDEB200_ECUDA, "cudaStreamCreate failed"));
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1); cudaEventCreate(&c->evc0); cudaEventCreate(&c->evc1);
    c->esz = cfg->dtype == PDEB200_F64 ? 8 : 4;
    c->a_rows = 1 + cfg->memory_size;
    switch (cfg->problem) {
        case PDEB200_KS: c->npts = cfg->nx; c->fields = 1; c->y_elems = c->npts; c->p_elems = c->npts; c->wrows = cfg->window_size; break;
        case PDEB200_KSEG1D: c->npts = cfg->nx; c->fields = 2; c->y_elems = 2 * c->npts; c->p_elems = c->npts; c->wrows = cfg->window_size; break;
        case PDEB200_NS2D: c->npts = cfg->nx * cfg->ny; c->fields = 1; c->y_elems = 2 * c->npts; c->p_elems = 2 * c->npts;
            c->wrows = cfg->window_size * cfg->window_size; break;
        case PDEB200_KSEG2D: c->npts = cfg->nx * cfg->ny; c->fields = 2; c->y_elems = 2 * c->npts; c->p_elems = c->npts;
            c->wrows = cfg->window_size * cfg->window_size; break;
        default: return bail(fail(c, PDEB200_EINVAL, "create: unknown problem"));
    }
    if (cfg->mono) {
        if (cfg->problem != PDEB200_KS) return bail(fail(c, PDEB200_EUNSUPPORTED, "mono (global agent) exists for KS only"));
        if (cfg->memory_size > 0)      // KSglobalSetup.jl:240-246 appends env.action rows; no shipped script uses it
            return bail(fail(c, PDEB200_EUNSUPPORTED, "mono (global agent) with memory_size > 0 is not implemented"));
        c->n_cols = 1; c->n_rew = 1; c->obs_rows = cfg->n_sensors * cfg->temporal_steps + cfg->memory_size;
    } else {
        c->n_cols = cfg->n_actuators; c->n_rew = cfg->n_actuators;
        c->obs_rows = c->wrows * c->fields * cfg->temporal_steps + cfg->memory_size;
    }
    const size_t B = cfg->n_envs, e = c->esz;
    auto alloc = [&](void** p, size_t bytes) {
        if (cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) return false;
        return cudaMemset(*p, 0, bytes ? bytes : 16) == cudaSuccess;
    };
    const size_t nact = B * cfg->n_actuators * c->a_rows;
    // what a host reads back after every env step lives in ONE block [reward | done | state]: a single D2H copy of the
    // whole block, or of its [reward | done] prefix when the policy and the trajectory live on the device and no host
    // code consumes the observation (pdeb200_result_select)
    auto up256 = [](size_t b) { return (b + 255) / 256 * 256; };
    c->res_reward_off = 0;
    c->res_done_off = up256(B * c->n_rew * e);
    c->res_state_off = c->res_done_off + up256(B);
    c->res_bytes = c->res_state_off + up256(B * c->n_cols * c->obs_rows * e);
    c->res_copy_bytes = c->res_bytes;
    bool ok = alloc(&c->y, B * c->y_elems * e) && alloc(&c->y0, B * c->y_elems * e) && alloc(&c->p, B * c->p_elems * e) &&
              alloc(&c->result_block, c->res_bytes) && alloc(&c->action, nact * e) &&
              alloc(&c->action_in, nact * e) && alloc(&c->delta_action, nact * e) &&
              alloc(&c->sensors, B * c->fields * cfg->n_sensors * e) &&
              alloc((void**)&c->time, B * 8) && alloc((void**)&c->steps, B * 4) && alloc((void**)&c->d_mask, B) && alloc((void**)&c->d_list, B * 4) &&
              alloc((void**)&c->d_rsum, B * 8) && alloc(&c->d_noise, nact * e) && alloc(&c->vmax, B * e) &&
              alloc((void**)&c->d_a2s, cfg->n_actuators * 4) && alloc(&c->d_sens_sum, cfg->n_sensors * e) &&
              alloc((void**)&c->d_losses, 8) && alloc((void**)&c->d_counts, 16) && alloc((void**)&c->d_nsub, B * 8) &&
              alloc(&c->d_hlast, B * 8);
    if (!ok) return bail(fail(c, PDEB200_ECUDA, std::string("cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError())));
    c->reward = (char*)c->result_block + c->res_reward_off;
    c->state = (char*)c->result_block + c->res_state_off;
    c->done = (uint8_t*)((char*)c->result_block + c->res_done_off);
    int32_t rc = PDEB200_OK;
    switch (cfg->problem) {
        case PDEB200_KS: rc = ks_setup(c); break;
        case PDEB200_KSEG1D: rc = kseg_setup(c); break;
        case PDEB200_NS2D: rc = ns_setup(c); break;
        case PDEB200_KSEG2D: rc = kseg2d_setup(c); break;
    }
    if (rc) return bail(rc);
    *out = c;
    return PDEB200_OK;
}

int32_t pdeb200_destroy(pdeb200_ctx* c) {
    if (!c) return PDEB200_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    ks_free(c); kseg_free(c); ns_free(c); agent_free(c); comm_free(c);
    for (void* p : {c->y, c->y0, c->p, c->result_block, c->action, c->action_in, c->delta_action, c->sensors,
                    (void*)c->time, (void*)c->steps, (void*)c->d_mask, (void*)c->d_list, (void*)c->d_rsum, c->d_noise, c->vmax,
                    (void*)c->d_a2s, c->d_sens_sum, (void*)c->sens.d_idx, c->sens.d_w, (void*)c->actT.d_idx, c->actT.d_w,
                    (void*)c->d_grads, (void*)c->d_losses, (void*)c->d_counts, (void*)c->d_nsub, c->d_hlast})
        if (p) cudaFree(p);
    for (auto& n : c->nets)
        for (void* p : {(void*)n.d_params, (void*)n.d_m, (void*)n.d_v, (void*)n.d_betap})
            if (p) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evc0) cudaEventDestroy(c->evc0);
    if (c->evc1) cudaEventDestroy(c->evc1);
    for (int i = 0; i < 2; ++i) {
        if (c->d_noise_q[i]) cudaFree(c->d_noise_q[i]);
        if (c->noise_ready[i]) cudaEventDestroy(c->noise_ready[i]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return PDEB200_OK;
}

const char* pdeb200_last_error(const pdeb200_ctx* c) { return c ? c->err.c_str() : g_last_error.c_str(); }

int32_t pdeb200_set_stream(pdeb200_ctx* c, void* s) {
    if (!c) return PDEB200_EINVAL;
    cudaStreamSynchronize(c->stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (s) { c->stream = (cudaStream_t)s; c->own_stream = false; }
    else { PDEB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    return PDEB200_OK;
}

int32_t pdeb200_synchronize(pdeb200_ctx* c) {
    if (!c) return PDEB200_EINVAL;
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_set_bases(pdeb200_ctx* c, const double* sb, const double* ab, const int32_t* a2s, double tol) {
    if (!c || !sb || !ab || !a2s) return fail(c, PDEB200_EINVAL, "set_bases: null argument");
    cudaSetDevice(c->device);
    return c->cfg.dtype == PDEB200_F64 ? set_bases_t<double>(c, sb, ab, a2s, tol) : set_bases_t<float>(c, sb, ab, a2s, tol);
}

int32_t pdeb200_set_y0(pdeb200_ctx* c, const double* y0, int32_t broadcast) {
    if (!c || !y0) return fail(c, PDEB200_EINVAL, "set_y0: null argument");
    cudaSetDevice(c->device);
    const size_t B = c->cfg.n_envs, ye = c->y_elems;
    std::vector<unsigned char> buf(B * ye * c->esz);
    for (size_t b = 0; b < B; ++b)
        for (size_t i = 0; i < ye; ++i) {
            const double v = y0[(broadcast ? 0 : b * ye) + i];
            if (c->esz == 8) ((double*)buf.data())[b * ye + i] = v;
            else ((float*)buf.data())[b * ye + i] = (float)v;
        }
    PDEB_CUDA(c, cudaMemcpyAsync(c->y0, buf.data(), buf.size(), cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->y0_set = true;
    return PDEB200_OK;
}

int32_t pdeb200_reset(pdeb200_ctx* c, const uint8_t* mask) {
    if (!c) return PDEB200_EINVAL;
    if (!c->bases_set || !c->y0_set) return fail(c, PDEB200_ESTATE, "reset: set_bases and set_y0 first");
    cudaSetDevice(c->device);
    const uint8_t* dm = nullptr;
    if (mask) {
        PDEB_CUDA(c, cudaMemcpyAsync(c->d_mask, mask, c->cfg.n_envs, cudaMemcpyHostToDevice, c->stream));
        dm = c->d_mask;
    }
    int32_t rc = c->cfg.dtype == PDEB200_F64 ? reset_t<double>(c, dm) : reset_t<float>(c, dm);
    if (rc) return rc;
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_reset_diverged(pdeb200_ctx* c, int32_t* counts) {
    if (!c) return PDEB200_EINVAL;
    if (!c->bases_set || !c->y0_set) return fail(c, PDEB200_ESTATE, "reset_diverged: set_bases and set_y0 first");
    cudaSetDevice(c->device);
    PDEB_CUDA(c, cudaMemsetAsync(c->d_counts, 0, 4 * sizeof(int), c->stream));
    diverged_mask_kernel<<<(c->cfg.n_envs + 255) / 256, 256, 0, c->stream>>>(c->cfg.n_envs, c->done, c->time, c->cfg.te, c->d_mask,
                                                                               c->d_counts, c->d_list);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    // the reset kernels walk the list of diverged environments (usually empty) with small grids instead of launching one
    // early-exiting CTA / warp per environment
    int32_t rc = c->cfg.dtype == PDEB200_F64 ? reset_t<double>(c, c->d_mask, c->d_list, c->d_counts + 2)
                                             : reset_t<float>(c, c->d_mask, c->d_list, c->d_counts + 2);
    if (rc) return rc;
    if (counts) {
        PDEB_CUDA(c, cudaMemcpyAsync(counts, c->d_counts, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return PDEB200_OK;
}

int32_t pdeb200_step_device(pdeb200_ctx* c, const void* actions_dev) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    return do_step(c, actions_dev ? actions_dev : c->action_in, 1, 0, 0.0, nullptr);
}

int32_t pdeb200_step(pdeb200_ctx* c, const void* actions_host) {
    if (!c || !actions_host) return fail(c, PDEB200_EINVAL, "step: null argument");
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)c->cfg.n_envs * c->cfg.n_actuators * c->a_rows * c->esz;
    PDEB_CUDA(c, cudaMemcpyAsync(c->action_in, actions_host, bytes, cudaMemcpyHostToDevice, c->stream));
    int32_t rc = do_step(c, c->action_in, 1, 0, 0.0, nullptr);
    if (rc) return rc;
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_step_host(pdeb200_ctx* c, const void* actions_host, void* y_out, void* reward_out, void* state_out,
                          uint8_t* done_out) {
    if (!c || !actions_host) return fail(c, PDEB200_EINVAL, "step_host: null argument");
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)c->cfg.n_envs * c->cfg.n_actuators * c->a_rows * c->esz;
    PDEB_CUDA(c, cudaMemcpyAsync(c->action_in, actions_host, bytes, cudaMemcpyHostToDevice, c->stream));
    int32_t rc = do_step(c, c->action_in, 1, 0, 0.0, nullptr);
    if (rc) return rc;
    auto back = [&](int which, void* dst) -> cudaError_t {
        if (!dst) return cudaSuccess;
        ArrInfo a = arr_info(c, which);
        return cudaMemcpyAsync(dst, a.ptr, a.bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    PDEB_CUDA(c, back(PDEB200_ARR_Y, y_out));
    PDEB_CUDA(c, back(PDEB200_ARR_REWARD, reward_out));
    PDEB_CUDA(c, back(PDEB200_ARR_STATE, state_out));
    PDEB_CUDA(c, back(PDEB200_ARR_DONE, done_out));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

static int32_t policy_launch(pdeb200_ctx* c, const void* d_noise, int use_rng, uint64_t seed, uint64_t offset, double act_noise,
                             double act_limit);

int32_t pdeb200_result_layout(const pdeb200_ctx* c, size_t* reward_off, size_t* state_off, size_t* done_off, size_t* total_bytes) {
    if (!c) return PDEB200_EINVAL;
    if (reward_off) *reward_off = c->res_reward_off;
    if (state_off) *state_off = c->res_state_off;
    if (done_off) *done_off = c->res_done_off;
    if (total_bytes) *total_bytes = c->res_copy_bytes;
    return PDEB200_OK;
}

int32_t pdeb200_result_select(pdeb200_ctx* c, int32_t with_state) {
    if (!c) return PDEB200_EINVAL;
    c->res_copy_bytes = with_state ? c->res_bytes : c->res_state_off;
    return PDEB200_OK;
}

static size_t policy_noise_count(const pdeb200_ctx* c) {
    const int n_out = c->cfg.mono ? c->cfg.n_actuators * c->a_rows : c->a_rows;
    return (size_t)c->cfg.n_envs * c->n_cols * (n_out - (c->cfg.mono ? 0 : c->cfg.memory_size));
}

int32_t pdeb200_noise_prefetch(pdeb200_ctx* c, const double* noise_host) {
    if (!c || !noise_host) return fail(c, PDEB200_EINVAL, "noise_prefetch: null argument");
    if (c->noise_put - c->noise_got >= 2) return fail(c, PDEB200_ESTATE, "noise_prefetch: two prefetches are already waiting for their policy calls");
    cudaSetDevice(c->device);
    const size_t n = policy_noise_count(c);
    if (!c->copy_stream) {
        PDEB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            PDEB_CUDA(c, cudaEventCreateWithFlags(&c->noise_ready[i], cudaEventDisableTiming));
            PDEB_CUDA(c, cudaMalloc(&c->d_noise_q[i], n * c->esz));
        }
    }
    // slot s was last read by the policy kernel of consumption (put - 2); put - got < 2 means that call has been made, and every
    // consuming call synchronises before it returns, so the buffer is free: no ordering against c->stream is needed
    const int s = (int)(c->noise_put & 1);
    if (c->esz == 8) PDEB_CUDA(c, cudaMemcpyAsync(c->d_noise_q[s], noise_host, n * 8, cudaMemcpyHostToDevice, c->copy_stream));
    else {
        std::vector<float> tmp(n);
        for (size_t i = 0; i < n; ++i) tmp[i] = (float)noise_host[i];
        PDEB_CUDA(c, cudaMemcpyAsync(c->d_noise_q[s], tmp.data(), n * 4, cudaMemcpyHostToDevice, c->copy_stream));
        PDEB_CUDA(c, cudaStreamSynchronize(c->copy_stream));          // tmp is pageable and local
    }
    PDEB_CUDA(c, cudaEventRecord(c->noise_ready[s], c->copy_stream));
    c->noise_put += 1;
    return PDEB200_OK;
}

// PDEB200_E2E_TRACE=<file prefix>: per call, stream events at {entry, noise on the device, kernels done, results on the host} and
// host clock at {entry, return}, appended to <prefix>.<ctx>.csv by the calling thread (tools/e2e_timeline.py draws the shards'
// overlap from these).  Off (one getenv per process) in normal use.
namespace {
struct E2eTrace {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    FILE* f = nullptr;
};
cudaEvent_t g_trace_base = nullptr;
std::mutex g_trace_mu;
std::map<pdeb200_ctx*, E2eTrace> g_traces;
const char* trace_prefix() { static const char* p = getenv("PDEB200_E2E_TRACE"); return p; }
E2eTrace* trace_get(pdeb200_ctx* c) {
    std::lock_guard<std::mutex> lock(g_trace_mu);
    if (!g_trace_base) { cudaEventCreate(&g_trace_base); cudaEventRecord(g_trace_base, c->stream); cudaEventSynchronize(g_trace_base); }
    E2eTrace& t = g_traces[c];
    if (!t.f) {
        for (auto& e : t.ev) cudaEventCreate(&e);
        char name[512];
        snprintf(name, sizeof name, "%s.%p.csv", trace_prefix(), (void*)c);
        t.f = fopen(name, "w");
    }
    return &t;
}
double host_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

int32_t pdeb200_act_step_host(pdeb200_ctx* c, const double* noise_host, double act_noise, double act_limit, void* action_out,
                              void* y_out, void* result_packed, void* reward_out, void* state_out, uint8_t* done_out) {
    if (!c) return fail(c, PDEB200_EINVAL, "act_step_host: null argument");
    cudaSetDevice(c->device);
    E2eTrace* tr = trace_prefix() ? trace_get(c) : nullptr;
    const double h0 = tr ? host_now_ms() : 0.0;
    if (tr) cudaEventRecord(tr->ev[0], c->stream);
    const void* dn = nullptr;
    std::vector<float> tmp;
    if (noise_host) {
        const int n_out = c->cfg.mono ? c->cfg.n_actuators * c->a_rows : c->a_rows;
        const size_t n = (size_t)c->cfg.n_envs * c->n_cols * (n_out - (c->cfg.mono ? 0 : c->cfg.memory_size));
        if (c->esz == 8) PDEB_CUDA(c, cudaMemcpyAsync(c->d_noise, noise_host, n * 8, cudaMemcpyHostToDevice, c->stream));
        else {
            tmp.resize(n);
            for (size_t i = 0; i < n; ++i) tmp[i] = (float)noise_host[i];
            PDEB_CUDA(c, cudaMemcpyAsync(c->d_noise, tmp.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
        }
        dn = c->d_noise;
    } else if (c->noise_put > c->noise_got && act_noise > 0.0) {
        // noise uploaded by pdeb200_noise_prefetch while the previous step ran: wait for that copy, use its buffer
        const int s = (int)(c->noise_got & 1);
        PDEB_CUDA(c, cudaStreamWaitEvent(c->stream, c->noise_ready[s], 0));
        c->noise_got += 1;
        dn = c->d_noise_q[s];
    }
    if (tr) cudaEventRecord(tr->ev[1], c->stream);
    // host does not want the action: policy(env) (actor + noise + clamp) runs inside the actuation kernel when the
    // shape-specialised one applies (one launch and one pass over the state fewer per step); otherwise policy kernel first
    int32_t rc = 1;
    if (!action_out) {
        const HostNet& pa = c->nets[PDEB200_NET_BEHAVIOR_ACTOR];
        bool narrow = pa.n_layers > 0 && c->cfg.memory_size == 0;
        for (int l = 0; narrow && l <= pa.n_layers; ++l) narrow = pa.sizes[l] <= kFusedActorMaxWidth;
        if (narrow) rc = do_step(c, nullptr, 1, 1, act_limit, nullptr, dn, act_noise, true);
        if (rc < 0) return rc;
    }
    const bool fused_policy = rc == 0;
    if (!fused_policy && (rc = policy_launch(c, dn, 0, 0, 0, act_noise, act_limit))) return rc;
    // policy(env) hands the action to the host; env(action) takes it from there (stream order: D2H, then H2D of the same buffer)
    // action_out == NULL: the action never leaves the device (device policy + device trajectory: no host consumer)
    const size_t abytes = (size_t)c->cfg.n_envs * c->cfg.n_actuators * c->a_rows * c->esz;
    if (action_out) {
        PDEB_CUDA(c, cudaMemcpyAsync(action_out, c->action_in, abytes, cudaMemcpyDeviceToHost, c->stream));
        PDEB_CUDA(c, cudaMemcpyAsync(c->action_in, action_out, abytes, cudaMemcpyHostToDevice, c->stream));
    }
    if (!fused_policy && (rc = do_step(c, c->action_in, 1, 0, 0.0, nullptr))) return rc;
    if (tr) cudaEventRecord(tr->ev[2], c->stream);
    if (y_out) PDEB_CUDA(c, cudaMemcpyAsync(y_out, c->y, (size_t)c->cfg.n_envs * c->y_elems * c->esz, cudaMemcpyDeviceToHost, c->stream));
    if (result_packed) PDEB_CUDA(c, cudaMemcpyAsync(result_packed, c->result_block, c->res_copy_bytes, cudaMemcpyDeviceToHost, c->stream));
    auto back = [&](int which, void* dst) -> cudaError_t {
        if (!dst) return cudaSuccess;
        ArrInfo a = arr_info(c, which);
        return cudaMemcpyAsync(dst, a.ptr, a.bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    PDEB_CUDA(c, back(PDEB200_ARR_REWARD, reward_out));
    PDEB_CUDA(c, back(PDEB200_ARR_STATE, state_out));
    PDEB_CUDA(c, back(PDEB200_ARR_DONE, done_out));
    if (tr) cudaEventRecord(tr->ev[3], c->stream);
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (tr && tr->f) {
        float t[4];
        for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], g_trace_base, tr->ev[i]);
        fprintf(tr->f, "%.4f,%.4f,%.4f,%.4f,%.4f,%.4f\n", t[0], t[1], t[2], t[3], h0, host_now_ms());
        fflush(tr->f);
    }
    return PDEB200_OK;
}

int32_t pdeb200_get(pdeb200_ctx* c, int32_t which, void* dst, size_t bytes) {
    if (!c || !dst) return fail(c, PDEB200_EINVAL, "get: null argument");
    cudaSetDevice(c->device);
    ArrInfo a = arr_info(c, which);
    if (!a.ptr) return fail(c, PDEB200_EINVAL, "get: unknown or unallocated array");
    if (bytes != a.bytes) return fail(c, PDEB200_EINVAL, "get: size mismatch (expected " + std::to_string(a.bytes) + " bytes)");
    PDEB_CUDA(c, cudaMemcpyAsync(dst, a.ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_get_env(pdeb200_ctx* c, int32_t which, int32_t env_index, void* dst, size_t bytes) {
    if (!c || !dst) return fail(c, PDEB200_EINVAL, "get_env: null argument");
    if (env_index < 0 || env_index >= c->cfg.n_envs) return fail(c, PDEB200_EINVAL, "get_env: environment index out of range");
    if (which == PDEB200_ARR_GRADS || which == PDEB200_ARR_LOSSES || which == PDEB200_ARR_STATS)
        return fail(c, PDEB200_EINVAL, "get_env: not a per-environment array");
    cudaSetDevice(c->device);
    ArrInfo a = arr_info(c, which);
    if (!a.ptr) return fail(c, PDEB200_EINVAL, "get_env: unknown or unallocated array");
    const size_t per = a.bytes / (size_t)c->cfg.n_envs;
    if (bytes != per) return fail(c, PDEB200_EINVAL, "get_env: size mismatch (expected " + std::to_string(per) + " bytes)");
    PDEB_CUDA(c, cudaMemcpyAsync(dst, (const char*)a.ptr + per * (size_t)env_index, per, cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_set(pdeb200_ctx* c, int32_t which, const void* src, size_t bytes) {
    if (!c || !src) return fail(c, PDEB200_EINVAL, "set: null argument");
    cudaSetDevice(c->device);
    ArrInfo a = arr_info(c, which);
    if (!a.ptr) return fail(c, PDEB200_EINVAL, "set: unknown or unallocated array");
    if (bytes != a.bytes) return fail(c, PDEB200_EINVAL, "set: size mismatch (expected " + std::to_string(a.bytes) + " bytes)");
    PDEB_CUDA(c, cudaMemcpyAsync(a.ptr, src, bytes, cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (which == PDEB200_ARR_Y0) c->y0_set = true;
    return PDEB200_OK;
}

int32_t pdeb200_device_ptr(pdeb200_ctx* c, int32_t which, void** ptr, size_t* bytes) {
    if (!c || !ptr) return fail(c, PDEB200_EINVAL, "device_ptr: null argument");
    ArrInfo a = arr_info(c, which);
    if (!a.ptr) return fail(c, PDEB200_EINVAL, "device_ptr: unknown or unallocated array");
    *ptr = a.ptr;
    if (bytes) *bytes = a.bytes;
    return PDEB200_OK;
}

int32_t pdeb200_obs_rows(const pdeb200_ctx* c) { return c ? c->obs_rows : PDEB200_EINVAL; }
int32_t pdeb200_obs_cols(const pdeb200_ctx* c) { return c ? c->n_cols : PDEB200_EINVAL; }

// ---- networks -----------------------------------------------------------------------------------
int32_t pdeb200_net_set(pdeb200_ctx* c, int32_t net, int32_t n_layers, const int32_t* sizes, const int32_t* acts,
                        const float* params) {
    if (!c || net < 0 || net > 3 || !sizes || !acts || !params) return fail(c, PDEB200_EINVAL, "net_set: bad argument");
    if (n_layers < 1 || n_layers > kMaxLayers) return fail(c, PDEB200_EUNSUPPORTED, "net_set: 1..4 Dense layers supported");
    cudaSetDevice(c->device);
    HostNet& n = c->nets[net];
    int total = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (sizes[l] < 1 || sizes[l + 1] < 1) return fail(c, PDEB200_EINVAL, "net_set: layer size < 1");
        n.offs[l] = total;
        total += sizes[l] * sizes[l + 1] + sizes[l + 1];
        n.acts[l] = acts[l];
    }
    for (int l = 0; l <= n_layers; ++l) n.sizes[l] = sizes[l];
    if (total != n.n_params || !n.d_params) {
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
        for (float** p : {&n.d_params, &n.d_m, &n.d_v}) { if (*p) cudaFree(*p); *p = nullptr; }
        PDEB_CUDA(c, cudaMalloc(&n.d_params, total * sizeof(float)));
        PDEB_CUDA(c, cudaMalloc(&n.d_m, total * sizeof(float)));
        PDEB_CUDA(c, cudaMalloc(&n.d_v, total * sizeof(float)));
    }
    if (!n.d_betap) PDEB_CUDA(c, cudaMalloc(&n.d_betap, 2 * sizeof(double)));
    n.n_layers = n_layers; n.n_params = total;
    agent_invalidate_graph(c);                      // a captured update graph holds the old pointers / shapes
    const double bp0[2] = {0.9, 0.999};             // a fresh Flux.ADAM: zero moments, beta powers (beta1, beta2)
    PDEB_CUDA(c, cudaMemcpyAsync(n.d_params, params, total * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaMemsetAsync(n.d_m, 0, total * sizeof(float), c->stream));
    PDEB_CUDA(c, cudaMemsetAsync(n.d_v, 0, total * sizeof(float), c->stream));
    PDEB_CUDA(c, cudaMemcpyAsync(n.d_betap, bp0, sizeof(bp0), cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_net_set_params(pdeb200_ctx* c, int32_t net, const float* params, size_t n_params) {
    if (!c || net < 0 || net > 3 || !params) return fail(c, PDEB200_EINVAL, "net_set_params: bad argument");
    HostNet& n = c->nets[net];
    if (!n.d_params || (size_t)n.n_params != n_params) return fail(c, PDEB200_EINVAL, "net_set_params: network unset or size mismatch");
    cudaSetDevice(c->device);
    PDEB_CUDA(c, cudaMemcpyAsync(n.d_params, params, n_params * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_opt_get(pdeb200_ctx* c, int32_t net, float* m, float* v, double* beta_p2, size_t n_params) {
    if (!c || net < 0 || net > 3) return fail(c, PDEB200_EINVAL, "opt_get: bad argument");
    HostNet& n = c->nets[net];
    if (!n.d_params || (size_t)n.n_params != n_params) return fail(c, PDEB200_EINVAL, "opt_get: network unset or size mismatch");
    cudaSetDevice(c->device);
    if (m) PDEB_CUDA(c, cudaMemcpyAsync(m, n.d_m, n_params * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (v) PDEB_CUDA(c, cudaMemcpyAsync(v, n.d_v, n_params * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (beta_p2) PDEB_CUDA(c, cudaMemcpyAsync(beta_p2, n.d_betap, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_opt_set(pdeb200_ctx* c, int32_t net, const float* m, const float* v, const double* beta_p2, size_t n_params) {
    if (!c || net < 0 || net > 3) return fail(c, PDEB200_EINVAL, "opt_set: bad argument");
    HostNet& n = c->nets[net];
    if (!n.d_params || (size_t)n.n_params != n_params) return fail(c, PDEB200_EINVAL, "opt_set: network unset or size mismatch");
    cudaSetDevice(c->device);
    if (m) PDEB_CUDA(c, cudaMemcpyAsync(n.d_m, m, n_params * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (v) PDEB_CUDA(c, cudaMemcpyAsync(n.d_v, v, n_params * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (beta_p2) PDEB_CUDA(c, cudaMemcpyAsync(n.d_betap, beta_p2, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_net_get(pdeb200_ctx* c, int32_t net, float* params, size_t n_params) {
    if (!c || net < 0 || net > 3 || !params) return fail(c, PDEB200_EINVAL, "net_get: bad argument");
    HostNet& n = c->nets[net];
    if (!n.d_params || (size_t)n.n_params != n_params) return fail(c, PDEB200_EINVAL, "net_get: network unset or size mismatch");
    cudaSetDevice(c->device);
    PDEB_CUDA(c, cudaMemcpyAsync(params, n.d_params, n_params * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_net_num_params(const pdeb200_ctx* c, int32_t net) {
    if (!c || net < 0 || net > 3) return PDEB200_EINVAL;
    return c->nets[net].n_params;
}

// ---- policy -------------------------------------------------------------------------------------
static int32_t policy_launch(pdeb200_ctx* c, const void* d_noise, int use_rng, uint64_t seed, uint64_t offset,
                             double act_noise, double act_limit) {
    const HostNet& a = c->nets[PDEB200_NET_BEHAVIOR_ACTOR];
    if (!a.n_layers) return fail(c, PDEB200_ESTATE, "policy_act: behavior actor not set");
    const int n_out = a.sizes[a.n_layers];
    const int expect_out = c->cfg.mono ? c->cfg.n_actuators * c->a_rows : c->a_rows;
    if (a.sizes[0] != c->obs_rows || n_out != expect_out)
        return fail(c, PDEB200_EINVAL, "policy_act: actor in/out sizes do not match the env's state/action spaces");
    for (int l = 0; l <= a.n_layers; ++l)
        if (a.sizes[l] > kFusedActorMaxWidth) return fail(c, PDEB200_EUNSUPPORTED, "policy_act: layer wider than 64");
    const int ncol = c->cfg.n_envs * c->n_cols;
    const int mem = c->cfg.mono ? 0 : c->cfg.memory_size;
    const int tpb = 128, grid = (ncol + tpb - 1) / tpb;
    int wmax = 1;
    for (int l = 0; l <= a.n_layers; ++l) wmax = std::max(wmax, a.sizes[l]);
    const size_t smem = ((size_t)a.n_params + (size_t)2 * wmax * tpb) * sizeof(float);
    if (smem > 200 * 1024) return fail(c, PDEB200_EUNSUPPORTED, "policy_act: actor too large for shared memory");
    const bool two = a.n_layers == 2 && a.offs[0] == 0 && a.offs[1] == a.sizes[0] * a.sizes[1] + a.sizes[1] && a.sizes[1] == 6 &&
                     a.sizes[2] == 1 && c->obs_rows == a.sizes[0];
    const int mode = (two && a.sizes[0] == 1) ? 1 : (two && a.sizes[0] == 3) ? 3 : -1;
    if (c->cfg.dtype == PDEB200_F64) {
        auto kern = mode == 1 ? policy_kernel<double, 1, 6> : mode == 3 ? policy_kernel<double, 3, 6> : policy_kernel<double, -1, 0>;
        PDEB_CUDA(c, ensure_dyn_smem(kern, smem, c->device));
        kern<<<grid, tpb, smem, c->stream>>>(a.dev(), a.n_params, wmax, ncol, c->obs_rows, c->a_rows, mem, (const double*)c->state,
                                            (double*)c->action_in, (const double*)d_noise, use_rng, seed, offset, act_noise, act_limit);
    } else {
        auto kern = mode == 1 ? policy_kernel<float, 1, 6> : mode == 3 ? policy_kernel<float, 3, 6> : policy_kernel<float, -1, 0>;
        PDEB_CUDA(c, ensure_dyn_smem(kern, smem, c->device));
        kern<<<grid, tpb, smem, c->stream>>>(a.dev(), a.n_params, wmax, ncol, c->obs_rows, c->a_rows, mem, (const float*)c->state,
                                            (float*)c->action_in, (const float*)d_noise, use_rng, seed, offset, (float)act_noise,
                                            (float)act_limit);
    }
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

int32_t pdeb200_policy_act(pdeb200_ctx* c, const double* noise_host, double act_noise, double act_limit) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    const void* dn = nullptr;
    std::vector<float> tmp;
    if (noise_host) {
        const int n_out = c->cfg.mono ? c->cfg.n_actuators * c->a_rows : c->a_rows;
        const int noisy = n_out - (c->cfg.mono ? 0 : c->cfg.memory_size);
        const size_t n = (size_t)c->cfg.n_envs * c->n_cols * noisy;
        if (c->esz == 8) PDEB_CUDA(c, cudaMemcpyAsync(c->d_noise, noise_host, n * 8, cudaMemcpyHostToDevice, c->stream));
        else {
            tmp.resize(n);
            for (size_t i = 0; i < n; ++i) tmp[i] = (float)noise_host[i];
            PDEB_CUDA(c, cudaMemcpyAsync(c->d_noise, tmp.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
        }
        dn = c->d_noise;
    }
    int32_t rc = policy_launch(c, dn, 0, 0, 0, act_noise, act_limit);
    if (rc) return rc;
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_policy_act_rng(pdeb200_ctx* c, uint64_t seed, uint64_t offset, double act_noise, double act_limit) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    return policy_launch(c, nullptr, 1, seed, offset, act_noise, act_limit);
}

int32_t pdeb200_rollout(pdeb200_ctx* c, int32_t n_steps, double act_limit, double* reward_sum_out) {
    if (!c || n_steps < 1) return fail(c, PDEB200_EINVAL, "rollout: bad argument");
    cudaSetDevice(c->device);
    double* rs = nullptr;
    if (reward_sum_out) {
        PDEB_CUDA(c, cudaMemsetAsync(c->d_rsum, 0, (size_t)c->cfg.n_envs * 8, c->stream));
        rs = c->d_rsum;
    }
    int32_t rc = do_step(c, nullptr, n_steps, 1, act_limit, rs);
    if (rc) return rc;
    if (reward_sum_out) {
        PDEB_CUDA(c, cudaMemcpyAsync(reward_sum_out, c->d_rsum, (size_t)c->cfg.n_envs * 8, cudaMemcpyDeviceToHost, c->stream));
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return PDEB200_OK;
}

// ---- introspection ------------------------------------------------------------------------------
int64_t pdeb200_launch_count(const pdeb200_ctx* c) { return c ? c->launches : 0; }

int32_t pdeb200_enable_step_timing(pdeb200_ctx* c, int32_t on) {
    if (!c) return PDEB200_EINVAL;
    c->timing = on != 0; c->timed = false;
    return PDEB200_OK;
}

int32_t pdeb200_last_step_ms(pdeb200_ctx* c, float* ms) {
    if (!c || !ms) return PDEB200_EINVAL;
    if (!c->timed) return fail(c, PDEB200_ESTATE, "last_step_ms: timing not enabled or no step yet");
    PDEB_CUDA(c, cudaEventSynchronize(c->ev1));
    PDEB_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return PDEB200_OK;
}

int32_t pdeb200_measure_fma_peak(pdeb200_ctx* c, int32_t dtype, double* tflops) {
    if (!c || !tflops) return fail(c, PDEB200_EINVAL, "measure_fma_peak: null argument");
    cudaSetDevice(c->device);
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
    const int iters = 4096, grid = n_sm * 8, tpb = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, c->stream);
        if (dtype == PDEB200_F64) fma_peak_kernel<double><<<grid, tpb, 0, c->stream>>>(iters, 1.0, (double*)c->vmax);
        else fma_peak_kernel<float><<<grid, tpb, 0, c->stream>>>(iters, 1.f, (float*)c->vmax);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 64.0 * iters * (double)grid * tpb;
        if (rep > 0 && ms > 0.f) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 4;
    *tflops = best;
    return PDEB200_OK;
}

int32_t pdeb200_last_phase_ms(pdeb200_ctx* c, float* ms3) {
    if (!c || !ms3) return PDEB200_EINVAL;
    if (!c->timed) return fail(c, PDEB200_ESTATE, "last_phase_ms: timing not enabled or no step yet");
    PDEB_CUDA(c, cudaEventSynchronize(c->ev1));
    PDEB_CUDA(c, cudaEventElapsedTime(ms3 + 0, c->ev0, c->evc0));      // actuation (policy + prepare_action)
    PDEB_CUDA(c, cudaEventElapsedTime(ms3 + 1, c->evc0, c->evc1));     // core (do_step + sensor dots)
    PDEB_CUDA(c, cudaEventElapsedTime(ms3 + 2, c->evc1, c->ev1));      // observe (reward, featurize, clock)
    return PDEB200_OK;
}

int32_t pdeb200_last_core_ms(pdeb200_ctx* c, float* ms) {
    if (!c || !ms) return PDEB200_EINVAL;
    if (!c->timed) return fail(c, PDEB200_ESTATE, "last_core_ms: timing not enabled or no step yet");
    PDEB_CUDA(c, cudaEventSynchronize(c->evc1));
    PDEB_CUDA(c, cudaEventElapsedTime(ms, c->evc0, c->evc1));
    return PDEB200_OK;
}

const char* pdeb200_last_core_kernel(const pdeb200_ctx* c) { return c ? c->core_kernel : ""; }

int32_t pdeb200_step_cost(const pdeb200_ctx* c, double* bytes, double* flops) {
    if (!c) return PDEB200_EINVAL;
    switch (c->cfg.problem) {
        case PDEB200_KS: return ks_cost(c, bytes, flops);
        case PDEB200_KSEG1D: return kseg_cost(c, bytes, flops);
        case PDEB200_NS2D: return ns_cost(c, bytes, flops);
        case PDEB200_KSEG2D: return kseg2d_cost(c, bytes, flops);
    }
    return PDEB200_EINVAL;
}

}  // extern "C"
