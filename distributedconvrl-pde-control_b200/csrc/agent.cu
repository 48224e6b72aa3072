// Agent back-end: device-resident replay trajectory, sampler and the DDPG update.
//
// Restates (batched; one replay "transition" = one actuator column, as in the reference):
//   trajectory update! overloads   /root/reference/src/PDEagent.jl:237-314
//       on RLCore 0.8.13's CircularArraySARTTrajectory: state/action rings of capacity+1
//       columns, reward/terminal rings of capacity, each a plain circular buffer
//   pde_sample / pde_fetch!        PDEagent.jl:317-340   (s' = state[inds + n_columns])
//   update!(policy, batch)         PDEagent.jl:363-418   (critic step, actor step through the
//                                                         UPDATED critic, Polyak on both targets)
//   Flux.Optimise.ADAM             (third-party; eta/beta Float64 applied to Float32 arrays)
//
// The networks are tiny (KS: 19 + 561 parameters), so the contraction is far too thin for tensor
// cores: one CTA processes tiles of 32 samples, thread-per-unit forward, thread-per-(unit,input)
// gradient accumulation in shared memory, per-CTA partials reduced in fixed order (deterministic).
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "comm.cuh"
#include "ctx.hpp"

namespace pdeb200 {

namespace {

constexpr int TS = 32;                 // samples per tile

struct Ring { int64_t cap = 0, start = 0, len = 0; };

// Device-resident mirror of the state a captured update graph must not receive by value: the ring positions (written by
// the push kernels), and the sampler's Philox counter (advanced by the sampler itself).
struct AgentDev {
    Ring sa, rt;
    unsigned long long rng_offset;
};

// stats[] layout (float64[8], PDEB200_ARR_STATS): sums over the GLOBAL batch (all ranks)
enum { ST_R = 0, ST_R2 = 1, ST_N = 2, ST_C = 3, ST_C2 = 4, ST_Q = 5 };

struct Agent {
    int ns = 0, na = 0;                // rows per state / action column
    int64_t ncols = 0;                 // columns pushed per env step (= B * n_cols)
    Ring sa, rt;
    float *state = nullptr, *action = nullptr, *reward = nullptr;
    uint8_t* terminal = nullptr;
    // sampled batch
    int batch = 0, batch_cap = 0;
    float *bs = nullptr, *ba = nullptr, *br = nullptr, *bs2 = nullptr;
    uint8_t* bt = nullptr;
    int64_t* inds = nullptr;
    double* stats = nullptr;           // [0] sum r  [1] sum r^2  [2] sum c  [3] sum c^2  [4] sum q(actor)
    float* partials = nullptr; size_t partials_floats = 0;
    int n_blocks = 0;
    double* stat_part = nullptr; int stat_part_cap = 0;   // per-CTA (sum r, sum r^2) of the sampler
    unsigned int* tickets = nullptr;                     // [0] sampler, [1] critic, [2] actor: "last CTA" counters (self-resetting)
    float* arena = nullptr; size_t arena_cap = 0, arena_used = 0;   // activations of the layer-wise (wide network) path
    int force_wide = 0;                // 1: always use the layer-wise path (tests / measurements)
    int wide_path = 0;                 // layer dispatch there: 0 auto, 1 CUDA cores only, 2 tensor cores wherever possible
    unsigned long long* timeline = nullptr;       // PDEB200_DDPG_TIMELINE=1: device timestamps of the update kernels' phases
    AgentDev* dev = nullptr;           // device mirror (see AgentDev)
    bool ring_dirty = true;            // host rings changed without a push kernel (create / pop_tail / set)
    float* xbuf = nullptr; int xbuf_cap = 0;      // reduced gradient (+ 2 loss sums) of one phase: the exchange's send / receive vector
    // captured graph of pdeb200_train_updates
    cudaGraphExec_t graph = nullptr;
    struct GraphKey { int n = 0, batch = 0, literal = 0; double gamma = 0, polyak = 0, lr_a = 0, lr_c = 0; uint64_t seed = 0; } gkey;
};

Agent* ag(pdeb200_ctx* c) { return static_cast<Agent*>(c->agent); }

// ---------------------------------------------------------------------------------------------
// replay kernels
// ---------------------------------------------------------------------------------------------
// `after`: the ring as it stands once this push has landed -- mirrored into device memory for the captured sampler
template <typename T>
__global__ void push_sa_kernel(int64_t n, int ns, int na, int64_t cap, int64_t pos0, const T* __restrict__ state,
                               const T* __restrict__ action, int zero_action, float* rstate, float* raction, AgentDev* dev,
                               Ring after) {
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (col == 0) dev->sa = after;
    if (col >= n) return;
    const int64_t dst = (pos0 + col) % cap;
    for (int r = 0; r < ns; ++r) rstate[dst * ns + r] = (float)state[col * ns + r];
    for (int r = 0; r < na; ++r) raction[dst * na + r] = zero_action ? 0.f : (float)action[col * na + r];
}

template <typename T>
__global__ void push_rt_kernel(int64_t n, int cols_per_env, int64_t cap, int64_t pos0, const T* __restrict__ reward,
                               const uint8_t* __restrict__ done, float* rreward, uint8_t* rterminal, AgentDev* dev, Ring after) {
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (col == 0) dev->rt = after;
    if (col >= n) return;
    const int64_t dst = (pos0 + col) % cap;
    rreward[dst] = (float)reward[col];
    rterminal[dst] = done[col / cols_per_env];
}

__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
    const uint64_t p = (uint64_t)a * b; *hi = (uint32_t)(p >> 32); return (uint32_t)p;
}
__device__ __forceinline__ uint64_t philox_u64(uint64_t seed, uint64_t ctr) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, h1;
        const uint32_t l0 = mulhilo32(0xD2511F53u, c0, &h0), l1 = mulhilo32(0xCD9E8D57u, c2, &h1);
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return ((uint64_t)c0 << 32) | c1;
}

__global__ void ring_sync_kernel(AgentDev* dev, Ring sa, Ring rt) { dev->sa = sa; dev->rt = rt; }
__global__ void rng_set_kernel(AgentDev* dev, unsigned long long off) { dev->rng_offset = off; }

// Completion of the batch statistics by ONE thread: sum over all ranks (peer-memory exchange, comm.cuh), publish
__device__ __forceinline__ void publish_stats(const CommDev& cm, double sum_r, double sum_r2, int n_local, double* stats) {
    double v3[3] = {sum_r, sum_r2, (double)n_local};
    if (cm.nranks > 1) comm_allreduce_stats(cm, v3);
    stats[ST_R] = v3[0]; stats[ST_R2] = v3[1]; stats[ST_N] = v3[2];
    stats[ST_C] = stats[ST_C2] = stats[ST_Q] = 0.0;
}

// sampled batch: inds ~ U{0 .. range-1} (reference: rand(rng, 1:length(t)-number_actuators, batch_size)), gather,
// reward statistics.  dev_state = 1: ring positions and the Philox counter come from the device mirror (captured
// graph: pdeb200_train_updates), and the counter is advanced by n; 0: from the arguments (pdeb200_sample).
__global__ void __launch_bounds__(128)
fetch_kernel(int n, int ns, int na, int64_t ncols, Ring sa, Ring rt, int64_t* inds, int draw, uint64_t seed,
             uint64_t offset, int dev_state, AgentDev* dev, const float* __restrict__ rstate, const float* __restrict__ raction,
             const float* __restrict__ rreward, const uint8_t* __restrict__ rterminal, float* bs, float* ba, float* br,
             uint8_t* bt, float* bs2, double* stat_part, unsigned int* ticket, double* stats, const __grid_constant__ CommDev cm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (dev_state) { sa = dev->sa; rt = dev->rt; offset = dev->rng_offset; }
    const int64_t range = rt.len - ncols;
    double r1 = 0.0, r2 = 0.0;
    if (i < n) {
        int64_t ind;
        if (draw) {          // inds ~ U{0 .. range-1}
            const uint64_t u = philox_u64(seed, offset + i);
            ind = (int64_t)(((unsigned __int128)u * (unsigned __int128)range) >> 64);
            inds[i] = ind;
        } else ind = inds[i];
        const int64_t ps = (sa.start + ind) % sa.cap, ps2 = (sa.start + ind + ncols) % sa.cap, pr = (rt.start + ind) % rt.cap;
        for (int r = 0; r < ns; ++r) { bs[(size_t)i * ns + r] = rstate[ps * ns + r]; bs2[(size_t)i * ns + r] = rstate[ps2 * ns + r]; }
        for (int r = 0; r < na; ++r) ba[(size_t)i * na + r] = raction[ps * na + r];
        const float rv = rreward[pr];
        br[i] = rv;
        bt[i] = rterminal[pr];
        r1 = rv; r2 = (double)rv * rv;
    }
    // sum r, sum r^2 over the local batch in a fixed order: shuffle tree per warp, warps and CTAs ascending; the last CTA
    // to finish (ticket counter) adds the per-CTA sums and runs the cross-rank exchange -- no separate launch
    __shared__ double s_p[2][4];
    __shared__ bool s_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r1 += __shfl_xor_sync(0xffffffffu, r1, o); r2 += __shfl_xor_sync(0xffffffffu, r2, o); }
    if ((threadIdx.x & 31) == 0) { s_p[0][threadIdx.x >> 5] = r1; s_p[1][threadIdx.x >> 5] = r2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_p[0][w]; b += s_p[1][w]; }
        stat_part[2 * blockIdx.x] = a; stat_part[2 * blockIdx.x + 1] = b;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (s_last) {
            __threadfence();
            double sa_ = 0.0, sb_ = 0.0;
            for (unsigned int k = 0; k < gridDim.x; ++k) { sa_ += __ldcg(stat_part + 2 * k); sb_ += __ldcg(stat_part + 2 * k + 1); }
            publish_stats(cm, sa_, sb_, n, stats);
            if (dev_state) dev->rng_offset = offset + (unsigned long long)n;
            *ticket = 0;
        }
    }
}

// sum r, sum r^2 of an explicit batch (single CTA, fixed order), then summed over the ranks like the sampler's
__global__ void reward_stats_kernel(int n, const float* __restrict__ br, double* stats, const __grid_constant__ CommDev cm) {
    __shared__ double s0[256], s1[256];
    double a = 0, b = 0;
    int i = threadIdx.x;
    // same per-thread order as a plain strided loop, loads issued 16 at a time
    for (; i + 15 * (int)blockDim.x < n; i += 16 * blockDim.x) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = br[i + u * blockDim.x];
#pragma unroll
        for (int u = 0; u < 16; ++u) { const double r = v[u]; a += r; b += r * r; }
    }
    for (; i < n; i += blockDim.x) { const double r = br[i]; a += r; b += r * r; }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) publish_stats(cm, s0[0], s1[0], n, stats);
}

// ---------------------------------------------------------------------------------------------
// tile-level MLP pieces (activations in shared memory as [sample][unit])
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(int kind, float out) {
    if (kind == 1) return out > 0.f ? 1.f : 0.f;         // relu
    if (kind == 2) return 1.f - out * out;               // tanh
    return 1.f;
}

__device__ void layer_forward(const float* __restrict__ W, int ni, int no, int act, const float* in, float* out) {
    const float* b = W + (size_t)ni * no;
    if (no * 8 <= (int)blockDim.x) {
        // narrow layer (e.g. the 140 -> 1 critic head): one thread per (sample, unit) instead of one per unit, so the
        // contraction over ni is not serialised into a single thread's 32 accumulators
        for (int q = threadIdx.x; q < TS * no; q += blockDim.x) {
            const int i = q / no, k = q - i * no;
            float acc = 0.f;
            for (int j = 0; j < ni; ++j) acc = fmaf(W[k + (size_t)no * j], in[i * ni + j], acc);
            out[q] = act_apply(act, acc + b[k]);
        }
        __syncthreads();
        return;
    }
    for (int k = threadIdx.x; k < no; k += blockDim.x) {
        float acc[TS];
#pragma unroll
        for (int i = 0; i < TS; ++i) acc[i] = 0.f;
        for (int j = 0; j < ni; ++j) {
            const float w = W[k + (size_t)no * j];
#pragma unroll
            for (int i = 0; i < TS; ++i) acc[i] = fmaf(w, in[i * ni + j], acc[i]);
        }
        const float bk = b[k];
#pragma unroll
        for (int i = 0; i < TS; ++i) out[i * no + k] = act_apply(act, acc[i] + bk);
    }
    __syncthreads();
}

// whole-network forward; acts[l] = pointer to layer-l activations ([TS][sizes[l]])
__device__ void net_forward(const NetDev& net, float* const* acts) {
    for (int l = 0; l < net.n_layers; ++l)
        layer_forward(net.params + net.offs[l], net.sizes[l], net.sizes[l + 1], net.acts[l], acts[l], acts[l + 1]);
}

// Backward through the network.  d_out: dLoss/d(output) [TS][n_out] (overwritten).  acc: shared
// accumulator over the flat parameter vector (or nullptr: no parameter gradients).  d_a/d_b:
// ping-pong scratch [TS][wmax].  Returns the buffer holding dLoss/d(input) [TS][sizes[0]].
__device__ float* net_backward(const NetDev& net, float* const* acts, float* d_out, float* d_other, float* acc,
                               bool want_input_grad) {
    float* d = d_out;
    float* dn = d_other;
    for (int l = net.n_layers - 1; l >= 0; --l) {
        const int ni = net.sizes[l], no = net.sizes[l + 1];
        const float* W = net.params + net.offs[l];
        const float* out = acts[l + 1];
        const float* in = acts[l];
        // delta_pre = delta_out * act'(out)
        if (net.acts[l] != 0) {
            for (int q = threadIdx.x; q < TS * no; q += blockDim.x) d[q] *= act_grad(net.acts[l], out[q]);
            __syncthreads();
        }
        if (acc) {
            const int n_pairs = (ni + 1) * no;             // flat index offs + q: W column-major then bias
            for (int q = threadIdx.x; q < n_pairs; q += blockDim.x) {
                const int k = q % no, j = q / no;
                float g = 0.f;
                if (j < ni) {
#pragma unroll 8
                    for (int i = 0; i < TS; ++i) g = fmaf(d[i * no + k], in[i * ni + j], g);
                } else {
#pragma unroll 8
                    for (int i = 0; i < TS; ++i) g += d[i * no + k];
                }
                acc[net.offs[l] + q] += g;
            }
        }
        if ((l > 0 || want_input_grad) && ni * 8 <= (int)blockDim.x) {
            // few inputs (e.g. the critic's (s, a) layer): one thread per (sample, input)
            for (int q = threadIdx.x; q < TS * ni; q += blockDim.x) {
                const int i = q / ni, j = q - i * ni;
                float sacc = 0.f;
                for (int k = 0; k < no; ++k) sacc = fmaf(W[k + (size_t)no * j], d[i * no + k], sacc);
                dn[q] = sacc;
            }
        } else if (l > 0 || want_input_grad) {
            for (int j = threadIdx.x; j < ni; j += blockDim.x) {
                float s[TS];
#pragma unroll
                for (int i = 0; i < TS; ++i) s[i] = 0.f;
                for (int k = 0; k < no; ++k) {
                    const float w = W[k + (size_t)no * j];
#pragma unroll
                    for (int i = 0; i < TS; ++i) s[i] = fmaf(w, d[i * no + k], s[i]);
                }
#pragma unroll
                for (int i = 0; i < TS; ++i) dn[i * ni + j] = s[i];
            }
        }
        __syncthreads();
        float* t = d; d = dn; dn = t;
    }
    return d;
}

struct DdpgArgs {
    NetDev A, C, At, Ct;
    int batch, ns, na;
    const float *s, *a, *r, *s2;
    const uint8_t* t;
    float gamma; int literal_q1;
    double inv_global_batch;        // > 0: given by the caller (four-phase API); 0: 1 / stats[ST_N] (sampler's global count)
    const double* stats;            // stats[ST_R] = GLOBAL sum of rewards
    float* partials;                // [gridDim.x][n_acc + 2]
    int n_acc;                      // parameters accumulated (critic or actor)
    int wmax;                       // widest activation row over all nets (incl. ns+na)
    // fused tail: the last CTA to finish reduces the partials (fixed order), exchanges the result with the peer GPUs
    // (cm.nranks > 1) and applies ADAM + Polyak
    int fuse;
    unsigned int* ticket;
    float* grads; double* stats_out; int stat0;
    float* xbuf;                    // n_acc + 2 floats: reduced gradient + loss sums, send / receive vector of the exchange
    float *x, *m, *v, *target;
    double* betap;                  // device-resident beta powers of this network's ADAM (read, then advanced)
    double eta, b1, b2, eps; float polyak;
    float* losses; int literal_loss;
    CommDev cm;
};

__device__ __forceinline__ double inv_gb(const DdpgArgs& D) { return D.inv_global_batch > 0.0 ? D.inv_global_batch : 1.0 / D.stats[ST_N]; }

__device__ void fused_tail(const DdpgArgs& D);

__device__ __forceinline__ int carve_acts(const NetDev& net, float* base, float** acts) {
    int off = 0;
    for (int l = 0; l <= net.n_layers; ++l) { acts[l] = base + off; off += TS * net.sizes[l]; }
    return off;
}

// Critic phase: targets from (A_t, C_t), critic loss gradient (PDEagent.jl:385-398).
__global__ void __launch_bounds__(512) ddpg_critic_kernel(const __grid_constant__ DdpgArgs D) {
    extern __shared__ __align__(16) float sm[];
    float* actsC[kMaxLayers + 1];
    int off = carve_acts(D.C, sm, actsC);
    float* scratch0 = sm + off; off += TS * D.wmax;      // target-net ping-pong / deltas
    float* scratch1 = sm + off; off += TS * D.wmax;
    float* tin = sm + off; off += TS * (D.ns + D.na);    // [s'; a'] for the target critic
    float* tgt = sm + off; off += TS;                    // gamma (1-t) q_t
    float* acc = sm + off; off += D.n_acc;
    __shared__ double s_c[2];
    for (int q = threadIdx.x; q < D.n_acc; q += blockDim.x) acc[q] = 0.f;
    if (threadIdx.x < 2) s_c[threadIdx.x] = 0.0;
    const int nin = D.ns + D.na;
    const double igb = inv_gb(D);
    const float rbar = (float)(D.stats[ST_R] * igb);
    const int n_tiles = (D.batch + TS - 1) / TS;
    __syncthreads();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int i0 = tile * TS;
        // ---- a' = A_t(s') ------------------------------------------------------------------
        {
            float* base = scratch0;                      // target actor activations, ping-pong in scratch0/1
            for (int q = threadIdx.x; q < TS * D.ns; q += blockDim.x) {
                const int i = q / D.ns, r = q % D.ns;
                const float v = (i0 + i < D.batch) ? D.s2[(size_t)(i0 + i) * D.ns + r] : 0.f;
                base[q] = v;
                tin[i * nin + r] = v;
            }
            __syncthreads();
            float* in = scratch0; float* out = scratch1;
            for (int l = 0; l < D.At.n_layers; ++l) {
                layer_forward(D.At.params + D.At.offs[l], D.At.sizes[l], D.At.sizes[l + 1], D.At.acts[l], in, out);
                float* t = in; in = out; out = t;
            }
            for (int q = threadIdx.x; q < TS * D.na; q += blockDim.x) tin[(q / D.na) * nin + D.ns + q % D.na] = in[q];
            __syncthreads();
        }
        // ---- q_t = C_t([s'; a']) -------------------------------------------------------------
        {
            float* in = tin; float* out = scratch0; float* other = scratch1;
            for (int l = 0; l < D.Ct.n_layers; ++l) {
                layer_forward(D.Ct.params + D.Ct.offs[l], D.Ct.sizes[l], D.Ct.sizes[l + 1], D.Ct.acts[l], in, out);
                in = out; out = (out == scratch0) ? other : scratch0;
            }
            for (int i = threadIdx.x; i < TS; i += blockDim.x) {
                const bool ok = i0 + i < D.batch;
                tgt[i] = ok ? D.gamma * (1.f - (float)D.t[i0 + i]) * in[i] : 0.f;
            }
            __syncthreads();
        }
        // ---- q = C([s; a]) with kept activations --------------------------------------------
        for (int q = threadIdx.x; q < TS * nin; q += blockDim.x) {
            const int i = q / nin, r = q % nin;
            float v = 0.f;
            if (i0 + i < D.batch) v = r < D.ns ? D.s[(size_t)(i0 + i) * D.ns + r] : D.a[(size_t)(i0 + i) * D.na + (r - D.ns)];
            actsC[0][q] = v;
        }
        __syncthreads();
        net_forward(D.C, actsC);
        // ---- dLoss/dq ------------------------------------------------------------------------
        // literal (quirk Q1): loss = mean_{i,j} (r_j + T_i - q_i)^2  =>  dq_i = -(2/B)(rbar + T_i - q_i)
        // per-sample:         loss = mean_i (r_i + T_i - q_i)^2      =>  dq_i = -(2/B)(r_i  + T_i - q_i)
        const float* qv = actsC[D.C.n_layers];
        for (int i = threadIdx.x; i < TS; i += blockDim.x) {
            float dq = 0.f;
            if (i0 + i < D.batch) {
                const float c = tgt[i] - qv[i];
                const float rr = D.literal_q1 ? rbar : D.r[i0 + i];
                dq = (float)(-2.0 * igb) * (rr + c);
                const double cl = D.literal_q1 ? (double)c : (double)(D.r[i0 + i] + c);
                atomicAdd(&s_c[0], cl); atomicAdd(&s_c[1], cl * cl);
            }
            scratch0[i] = dq;
        }
        __syncthreads();
        net_backward(D.C, actsC, scratch0, scratch1, acc, false);
        __syncthreads();
    }
    float* out = D.partials + (size_t)blockIdx.x * (D.n_acc + 2);
    for (int q = threadIdx.x; q < D.n_acc; q += blockDim.x) out[q] = acc[q];
    if (threadIdx.x == 0) { out[D.n_acc] = (float)s_c[0]; out[D.n_acc + 1] = (float)s_c[1]; }
    if (D.fuse) fused_tail(D);
}

// Actor phase: gradient of -mean C([s; A(s)]) w.r.t. the actor parameters (PDEagent.jl:402-409).
__global__ void __launch_bounds__(512) ddpg_actor_kernel(const __grid_constant__ DdpgArgs D) {
    extern __shared__ __align__(16) float sm[];
    float* actsC[kMaxLayers + 1];
    float* actsA[kMaxLayers + 1];
    int off = carve_acts(D.C, sm, actsC);
    off += carve_acts(D.A, sm + off, actsA);
    float* scratch0 = sm + off; off += TS * D.wmax;
    float* scratch1 = sm + off; off += TS * D.wmax;
    float* acc = sm + off; off += D.n_acc;
    __shared__ double s_q;
    for (int q = threadIdx.x; q < D.n_acc; q += blockDim.x) acc[q] = 0.f;
    if (threadIdx.x == 0) s_q = 0.0;
    const int nin = D.ns + D.na;
    const int n_tiles = (D.batch + TS - 1) / TS;
    const double igb = inv_gb(D);
    __syncthreads();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int i0 = tile * TS;
        for (int q = threadIdx.x; q < TS * D.ns; q += blockDim.x) {
            const int i = q / D.ns, r = q % D.ns;
            const float v = (i0 + i < D.batch) ? D.s[(size_t)(i0 + i) * D.ns + r] : 0.f;
            actsA[0][q] = v;
            actsC[0][i * nin + r] = v;
        }
        __syncthreads();
        net_forward(D.A, actsA);
        const float* av = actsA[D.A.n_layers];
        for (int q = threadIdx.x; q < TS * D.na; q += blockDim.x) actsC[0][(q / D.na) * nin + D.ns + q % D.na] = av[q];
        __syncthreads();
        net_forward(D.C, actsC);
        const float* qv = actsC[D.C.n_layers];
        for (int i = threadIdx.x; i < TS; i += blockDim.x) {
            const bool ok = i0 + i < D.batch;
            scratch0[i] = ok ? (float)(-igb) : 0.f;
            if (ok) atomicAdd(&s_q, (double)qv[i]);
        }
        __syncthreads();
        float* dx = net_backward(D.C, actsC, scratch0, scratch1, nullptr, true);     // [TS][ns+na]
        float* da = (dx == scratch0) ? scratch1 : scratch0;
        for (int q = threadIdx.x; q < TS * D.na; q += blockDim.x) da[q] = dx[(q / D.na) * nin + D.ns + q % D.na];
        __syncthreads();
        net_backward(D.A, actsA, da, dx, acc, false);
        __syncthreads();
    }
    float* out = D.partials + (size_t)blockIdx.x * (D.n_acc + 2);
    for (int q = threadIdx.x; q < D.n_acc; q += blockDim.x) out[q] = acc[q];
    if (threadIdx.x == 0) { out[D.n_acc] = (float)s_q; out[D.n_acc + 1] = 0.f; }
    if (D.fuse) fused_tail(D);
}

// sum_b partials[b * stride + q] in ascending b (fixed order), with the loads issued 32 at a time ahead of the adds
// (the SM issues in order: load / add / load / add would expose one memory latency per term; with 128 CTA partials
// batches of 8 were still 16 dependent L2 round trips = 11 of the kernel's 14 us)
__device__ __forceinline__ double ordered_sum(const float* __restrict__ partials, int n_blocks, size_t stride, int q) {
    double s = 0.0;
    int b = 0;
    for (; b + 32 <= n_blocks; b += 32) {
        float v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = __ldg(partials + (size_t)(b + u) * stride + q);
#pragma unroll
        for (int u = 0; u < 32; ++u) s += (double)v[u];
    }
    for (; b + 8 <= n_blocks; b += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(partials + (size_t)(b + u) * stride + q);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; b < n_blocks; ++b) s += (double)__ldg(partials + (size_t)b * stride + q);
    return s;
}

// Same loads through L2 (the partials were written by other CTAs of the SAME launch: no non-coherent path)
__device__ __forceinline__ double ordered_sum_cg(const float* partials, int n_blocks, size_t stride, int q) {
    double s = 0.0;
    int b = 0;
    for (; b + 32 <= n_blocks; b += 32) {
        float v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = __ldcg(partials + (size_t)(b + u) * stride + q);
#pragma unroll
        for (int u = 0; u < 32; ++u) s += (double)v[u];
    }
    for (; b + 8 <= n_blocks; b += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(partials + (size_t)(b + u) * stride + q);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; b < n_blocks; ++b) s += (double)__ldcg(partials + (size_t)b * stride + q);
    return s;
}

// critic_loss / actor_loss (PDEagent.jl:393-396, 403-407) from the global sums
__device__ __forceinline__ float critic_loss_from(const double* st, int literal) {
    // literal: mean_{ij} (r_j + c_i)^2 = sum c^2/B + 2 sum c sum r / B^2 + sum r^2 / B ; per-sample: sum c'^2 / B
    const double B = st[ST_N];
    return (float)(literal ? st[ST_C2] / B + 2.0 * st[ST_C] * st[ST_R] / (B * B) + st[ST_R2] / B : st[ST_C2] / B);
}

// Tail of the fused gradient kernels: every CTA has written its partials; the last one to arrive (ticket counter,
// self-resetting) reduces them in a fixed order, sums the result over the ranks through NVLink peer memory
// (comm_allreduce_cta -- the path's one collective, executed here instead of between launches), then applies ADAM, the
// target's Polyak step and the losses.  Identical arithmetic on every rank => identical weights on every rank.
__device__ void fused_tail(const DdpgArgs& D) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(D.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const size_t stride = (size_t)(D.n_acc + 2);
    for (int q = threadIdx.x; q < D.n_acc + 2; q += blockDim.x)
        D.xbuf[q] = (float)ordered_sum_cg(D.partials, (int)gridDim.x, stride, q);
    __syncthreads();
    if (D.cm.nranks > 1) comm_allreduce_cta(D.cm, D.xbuf, D.xbuf, D.n_acc + 2);
    const double bp1 = D.betap[0], bp2 = D.betap[1];
    if (threadIdx.x == 0) {
        const double s0 = (double)D.xbuf[D.n_acc], s1 = (double)D.xbuf[D.n_acc + 1];
        D.stats_out[D.stat0] = s0;
        if (D.stat0 == ST_C) D.stats_out[ST_C2] = s1;
        if (D.losses) {                                   // actor phase: both losses are complete now
            D.losses[0] = critic_loss_from(D.stats_out, D.literal_loss);
            D.losses[1] = (float)(-s0 / D.stats_out[ST_N]);
        }
    }
    for (int q = threadIdx.x; q < D.n_acc; q += blockDim.x) {
        const float g = D.xbuf[q];
        D.grads[q] = g;
        const double gi = g;
        const float mi = (float)(D.b1 * (double)D.m[q] + (1.0 - D.b1) * gi);
        const float vi = (float)(D.b2 * (double)D.v[q] + (1.0 - D.b2) * gi * gi);
        D.m[q] = mi; D.v[q] = vi;
        const float delta = (float)((double)mi / (1.0 - bp1) / (sqrt((double)vi / (1.0 - bp2)) + D.eps) * D.eta);
        const float xn = D.x[q] - delta;
        D.x[q] = xn;
        D.target[q] = D.polyak * D.target[q] + (1.f - D.polyak) * xn;
    }
    __syncthreads();
    if (threadIdx.x == 0) { D.betap[0] = bp1 * D.b1; D.betap[1] = bp2 * D.b2; *D.ticket = 0; }
}

#include "ddpg_fast.cuh"

// grads[q] = sum over CTAs (fixed order); tail sums go to stats
__global__ void reduce_partials_kernel(int n_blocks, int n_acc, const float* __restrict__ partials, float* grads,
                                       double* stats, int stat0) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_acc + 2) return;
    const double s = ordered_sum(partials, n_blocks, (size_t)(n_acc + 2), q);
    if (q < n_acc) grads[q] = (float)s;
    else stats[stat0 + (q - n_acc)] = s;
}

// Flux ADAM on Float32 arrays with Float64 hyper-parameters.  The running beta powers live on the device; the last
// block to finish (ticket) advances them -- every block has read them by then.
__global__ void adam_kernel(int n, float* x, float* m, float* v, const float* __restrict__ g, double eta, double b1,
                            double b2, double* betap, double eps, unsigned int* ticket) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double bp1 = betap[0], bp2 = betap[1];
    if (i < n) {
        const double gi = g[i];
        const float mi = (float)(b1 * (double)m[i] + (1.0 - b1) * gi);
        const float vi = (float)(b2 * (double)v[i] + (1.0 - b2) * gi * gi);
        m[i] = mi; v[i] = vi;
        const float delta = (float)((double)mi / (1.0 - bp1) / (sqrt((double)vi / (1.0 - bp2)) + eps) * eta);
        x[i] -= delta;
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(ticket, 1u) == gridDim.x - 1) { betap[0] = bp1 * b1; betap[1] = bp2 * b2; *ticket = 0; }
}

__global__ void polyak_kernel(int n, float* dest, const float* __restrict__ src, float p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dest[i] = p * dest[i] + (1.f - p) * src[i];
}

// n_global > 0: the caller's global batch (four-phase API) replaces the sampler's count
__global__ void losses_kernel(double* stats, double n_global, int literal, int which, float* losses) {
    if (n_global > 0.0) stats[ST_N] = n_global;
    if (which == 0) losses[0] = critic_loss_from(stats, literal);
    else losses[1] = (float)(-stats[ST_Q] / stats[ST_N]);
}

void drop_graph(Agent* a) {
    if (a && a->graph) { cudaGraphExecDestroy(a->graph); a->graph = nullptr; }
}

int32_t ensure_partials(pdeb200_ctx* c, int n_blocks, int n_acc) {
    Agent* a = ag(c);
    const size_t need = (size_t)n_blocks * (n_acc + 2);
    if (need > a->partials_floats) {
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
        drop_graph(a);
        if (a->partials) cudaFree(a->partials);
        PDEB_CUDA(c, cudaMalloc(&a->partials, need * sizeof(float)));
        a->partials_floats = need;
    }
    if (n_acc + 2 > a->xbuf_cap) {
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
        drop_graph(a);
        if (a->xbuf) cudaFree(a->xbuf);
        PDEB_CUDA(c, cudaMalloc(&a->xbuf, (size_t)(n_acc + 2) * sizeof(float)));
        a->xbuf_cap = n_acc + 2;
    }
    return PDEB200_OK;
}

int32_t ensure_agent(pdeb200_ctx* c) {
    if (c->agent) return PDEB200_OK;
    Agent* a = new Agent();
    a->ns = c->obs_rows;
    a->na = c->cfg.mono ? c->cfg.n_actuators * c->a_rows : c->a_rows;
    a->ncols = (int64_t)c->cfg.n_envs * c->n_cols;
    c->agent = a;
    PDEB_CUDA(c, cudaMalloc(&a->stats, 8 * sizeof(double)));
    PDEB_CUDA(c, cudaMemset(a->stats, 0, 8 * sizeof(double)));
    PDEB_CUDA(c, cudaMalloc(&a->dev, sizeof(AgentDev)));
    PDEB_CUDA(c, cudaMemset(a->dev, 0, sizeof(AgentDev)));
    PDEB_CUDA(c, cudaMalloc(&a->tickets, 8 * sizeof(unsigned int)));      // [0] sampler [1] critic [2] actor [3] adam
    PDEB_CUDA(c, cudaMemset(a->tickets, 0, 8 * sizeof(unsigned int)));
    return PDEB200_OK;
}


int32_t ensure_batch(pdeb200_ctx* c, int batch) {
    Agent* a = ag(c);
    if (batch > a->batch_cap) {
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
        drop_graph(a);
        for (void* p : {(void*)a->bs, (void*)a->ba, (void*)a->br, (void*)a->bs2, (void*)a->bt, (void*)a->inds})
            if (p) cudaFree(p);
        PDEB_CUDA(c, cudaMalloc(&a->bs, (size_t)batch * a->ns * 4));
        PDEB_CUDA(c, cudaMalloc(&a->bs2, (size_t)batch * a->ns * 4));
        PDEB_CUDA(c, cudaMalloc(&a->ba, (size_t)batch * a->na * 4));
        PDEB_CUDA(c, cudaMalloc(&a->br, (size_t)batch * 4));
        PDEB_CUDA(c, cudaMalloc(&a->bt, (size_t)batch));
        PDEB_CUDA(c, cudaMalloc(&a->inds, (size_t)batch * 8));
        a->batch_cap = batch;
    }
    const int grid = (batch + 127) / 128;
    if (grid > a->stat_part_cap) {
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
        drop_graph(a);
        if (a->stat_part) cudaFree(a->stat_part);
        PDEB_CUDA(c, cudaMalloc(&a->stat_part, (size_t)2 * grid * sizeof(double)));
        a->stat_part_cap = grid;
    }
    a->batch = batch;
    return PDEB200_OK;
}

int32_t check_nets(pdeb200_ctx* c) {
    Agent* a = ag(c);
    for (int i = 0; i < 4; ++i)
        if (!c->nets[i].n_layers) return fail(c, PDEB200_ESTATE, "ddpg: all four networks must be set (pdeb200_net_set)");
    const HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    if (A.sizes[0] != a->ns || A.sizes[A.n_layers] != a->na) return fail(c, PDEB200_EINVAL, "ddpg: actor shape != (ns -> na)");
    if (C.sizes[0] != a->ns + a->na || C.sizes[C.n_layers] != 1) return fail(c, PDEB200_EINVAL, "ddpg: critic shape != (ns+na -> 1)");
    if (c->nets[PDEB200_NET_TARGET_ACTOR].n_params != A.n_params || c->nets[PDEB200_NET_TARGET_CRITIC].n_params != C.n_params)
        return fail(c, PDEB200_EINVAL, "ddpg: target networks must have the behavior networks' shapes");
    const int want = C.n_params + A.n_params;
    if (c->n_grads != want) {
        PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
        drop_graph(a);
        if (c->d_grads) cudaFree(c->d_grads);
        PDEB_CUDA(c, cudaMalloc(&c->d_grads, (size_t)want * sizeof(float)));
        PDEB_CUDA(c, cudaMemset(c->d_grads, 0, (size_t)want * sizeof(float)));
        c->n_grads = want;
    }
    return PDEB200_OK;
}

int net_wmax(const HostNet& n) { int w = 0; for (int l = 0; l <= n.n_layers; ++l) w = std::max(w, n.sizes[l]); return w; }
int net_act_floats(const HostNet& n) { int s = 0; for (int l = 0; l <= n.n_layers; ++l) s += TS * n.sizes[l]; return s; }

// global_batch > 0: the caller's count (four-phase API); 0: the sampler's global count on the device (stats[ST_N])
DdpgArgs make_args(pdeb200_ctx* c, double gamma, int literal, int64_t global_batch) {
    Agent* a = ag(c);
    DdpgArgs D;
    D.A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR].dev(); D.C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC].dev();
    D.At = c->nets[PDEB200_NET_TARGET_ACTOR].dev(); D.Ct = c->nets[PDEB200_NET_TARGET_CRITIC].dev();
    D.batch = a->batch; D.ns = a->ns; D.na = a->na;
    D.s = a->bs; D.a = a->ba; D.r = a->br; D.s2 = a->bs2; D.t = a->bt;
    D.gamma = (float)gamma; D.literal_q1 = literal; D.inv_global_batch = global_batch > 0 ? 1.0 / (double)global_batch : 0.0;
    D.stats = a->stats; D.partials = nullptr; D.n_acc = 0;
    D.fuse = 0; D.ticket = nullptr; D.grads = nullptr; D.stats_out = a->stats; D.stat0 = 0; D.xbuf = a->xbuf;
    D.x = D.m = D.v = D.target = nullptr; D.betap = nullptr; D.eta = D.b1 = D.b2 = D.eps = 0.0; D.polyak = 0.f;
    D.losses = nullptr; D.literal_loss = literal;
    D.wmax = std::max({net_wmax(c->nets[0]), net_wmax(c->nets[1]), a->ns + a->na});
    return D;
}

int block_threads(int wmax, int n_acc) {
    int t = std::max(64, ((std::max(wmax, std::min(n_acc, 512)) + 31) / 32) * 32);
    return std::min(t, 512);
}


// ---------------------------------------------------------------------------------------------
// Layer-wise DDPG update for networks that do not fit the shared-memory kernels above (wide critics with the
// middle layer, `drop_middle_layer = false`): activations live in global memory, every Dense layer is a GEMM --
// forward, input gradient and weight gradient on the tensor cores when the layer is dense (nn_ops.cu /
// dense_tc.cuh), CUDA-core kernels for the thin input / output layers.  Same arithmetic as update!
// (PDEagent.jl:363-418), same flat gradient layout, fixed-order reductions.
// ---------------------------------------------------------------------------------------------
inline int pad4i(int n) { return (n + 3) / 4 * 4; }

struct Acts { float* out[kMaxLayers + 1]; long long ld[kMaxLayers + 1]; };

float* arena_take(Agent* a, size_t n) {
    n = (n + 63) / 64 * 64;
    if (a->arena_used + n > a->arena_cap) return nullptr;
    float* p = a->arena + a->arena_used;
    a->arena_used += n;
    return p;
}

int32_t arena_reserve(pdeb200_ctx* c, size_t floats) {
    Agent* a = ag(c);
    a->arena_used = 0;
    if (floats <= a->arena_cap) return PDEB200_OK;
    if (a->arena) cudaFree(a->arena);
    a->arena = nullptr; a->arena_cap = 0;
    PDEB_CUDA(c, cudaMalloc(&a->arena, floats * sizeof(float)));
    PDEB_CUDA(c, cudaMemsetAsync(a->arena, 0, floats * sizeof(float), c->stream));
    a->arena_cap = floats;
    return PDEB200_OK;
}

// X[m] = [s[m]; a[m]; 0-pad]
__global__ void concat_kernel(int M, int ns, int na, int ld, const float* __restrict__ s, const float* __restrict__ a, int lda,
                              float* __restrict__ X) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= (long long)M * ld) return;
    const int m = (int)(q / ld), r = (int)(q % ld);
    X[q] = r < ns ? s[(size_t)m * ns + r] : (r < ns + na ? a[(size_t)m * lda + (r - ns)] : 0.f);
}

__global__ void copy_pad_kernel(int M, int n, int ld, const float* __restrict__ src, float* __restrict__ dst) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= (long long)M * ld) return;
    const int m = (int)(q / ld), r = (int)(q % ld);
    dst[q] = r < n ? src[(size_t)m * n + r] : 0.f;
}

// dq and the loss sums of the critic phase (single CTA: fixed summation order)
__global__ void __launch_bounds__(1024) critic_dq_kernel(int B, const float* __restrict__ q, long long ldq, const float* __restrict__ qt,
                                                         long long ldqt, const float* __restrict__ r, const uint8_t* __restrict__ t,
                                                         float gamma, int literal, double inv_gb, double* stats, float* dq,
                                                         long long lddq) {
    __shared__ double s0[1024], s1[1024];
    const float rbar = (float)(stats[ST_R] * inv_gb);
    double a = 0, b = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float cc = gamma * (1.f - (float)t[i]) * qt[(long long)i * ldqt] - q[(long long)i * ldq];
        const float rr = literal ? rbar : r[i];
        dq[(long long)i * lddq] = (float)(-2.0 * inv_gb) * (rr + cc);
        const double cl = literal ? (double)cc : (double)(r[i] + cc);
        a += cl; b += cl * cl;
    }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { stats[ST_C] = s0[0]; stats[ST_C2] = s1[0]; }
}

__global__ void __launch_bounds__(1024) actor_dq_kernel(int B, const float* __restrict__ q, long long ldq, double inv_gb, double* stats,
                                                        float* dq, long long lddq) {
    __shared__ double s0[1024];
    double a = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { a += (double)q[(long long)i * ldq]; dq[(long long)i * lddq] = (float)(-inv_gb); }
    s0[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) s0[threadIdx.x] += s0[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) stats[ST_Q] = s0[0];
}

// d[m][n] *= act'(out[m][n])
__global__ void act_grad_mul_kernel(long long M, int N, int act, const float* __restrict__ out, long long ldo, float* d, long long ldd) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= M * N) return;
    const long long m = q / N; const int n = (int)(q % N);
    d[m * ldd + n] *= act_grad(act, out[m * ldo + n]);
}

void zero_pad(pdeb200_ctx* c, long long M, int n, long long ld, float* X);

int32_t wide_forward(pdeb200_ctx* c, const HostNet& net, int M, float* x, long long ldx, Acts* A) {
    Agent* a = ag(c);
    A->out[0] = x; A->ld[0] = ldx;
    for (int l = 0; l < net.n_layers; ++l) {
        const int no = net.sizes[l + 1];
        A->ld[l + 1] = pad4i(no);
        A->out[l + 1] = arena_take(a, (size_t)M * A->ld[l + 1]);          // arena is zero-filled: padding columns stay 0
        if (!A->out[l + 1]) return fail(c, PDEB200_ECUDA, "ddpg (layer-wise path): activation arena exhausted");
        int32_t rc = dense_layer(c, M, net.sizes[l], no, A->out[l], A->ld[l], net.d_params + net.offs[l], net.acts[l],
                                 A->out[l + 1], A->ld[l + 1], a->wide_path, nullptr);
        if (rc) return rc;
        zero_pad(c, M, no, A->ld[l + 1], A->out[l + 1]);
    }
    return PDEB200_OK;
}

// d: dLoss/d(output) [M][ld of the output]; overwritten.  gflat: flat parameter gradient or nullptr.
// Returns dLoss/d(input) in *d_in (arena buffer, ld = pad4(sizes[0])) when want_input.
int32_t wide_backward(pdeb200_ctx* c, const HostNet& net, int M, const Acts& A, float* d, long long ldd, float* gflat,
                      bool want_input, float** d_in, long long* ld_in) {
    Agent* a = ag(c);
    const int L = net.n_layers;
    if (net.acts[L - 1] != 0) {
        const long long tot = (long long)M * net.sizes[L];
        act_grad_mul_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(M, net.sizes[L], net.acts[L - 1], A.out[L], A.ld[L], d, ldd);
        c->launches += 1;
    }
    for (int l = L - 1; l >= 0; --l) {
        const int K = net.sizes[l], N = net.sizes[l + 1];
        const float* W = net.d_params + net.offs[l];
        int32_t rc;
        if (gflat && (rc = dense_wgrad(c, M, K, N, d, ldd, A.out[l], A.ld[l], gflat + net.offs[l], gflat + net.offs[l] + (size_t)K * N,
                                       a->wide_path))) return rc;
        if (l > 0 || want_input) {
            const long long ldp = pad4i(K);
            float* dp = arena_take(a, (size_t)M * ldp);
            if (!dp) return fail(c, PDEB200_ECUDA, "ddpg (layer-wise path): activation arena exhausted");
            // delta of the previous layer = (d * W) .* act'(its output)
            const float* mask = l > 0 && net.acts[l - 1] != 0 ? A.out[l] : nullptr;
            if ((rc = dense_dgrad(c, M, K, N, d, ldd, W, mask, A.ld[l], l > 0 ? net.acts[l - 1] : 0, dp, ldp, a->wide_path))) return rc;
            zero_pad(c, M, K, ldp, dp);
            d = dp; ldd = ldp;
        }
    }
    if (want_input) { *d_in = d; *ld_in = ldd; }
    PDEB_CUDA(c, cudaGetLastError());
    return PDEB200_OK;
}

// activations + deltas of every network for one phase, generously (each buffer is rounded up to 64 floats)
size_t wide_arena_floats(pdeb200_ctx* c) {
    Agent* a = ag(c);
    size_t per_col = 64, n_buf = 8;
    for (int i = 0; i < 4; ++i)
        for (int l = 0; l <= c->nets[i].n_layers; ++l) { per_col += 2 * (size_t)pad4i(c->nets[i].sizes[l]); n_buf += 2; }
    return per_col * (size_t)a->batch + n_buf * 64;
}

// zero the padding columns [n, ld) of a row-major buffer (they take part in the next GEMM's contraction)
__global__ void zero_pad_kernel(long long M, int n, int ld, float* X) {
    const int w = ld - n;
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w <= 0 || q >= M * w) return;
    X[(q / w) * ld + n + (q % w)] = 0.f;
}
void zero_pad(pdeb200_ctx* c, long long M, int n, long long ld, float* X) {
    if (ld == n) return;
    const long long tot = M * (ld - n);
    zero_pad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(M, n, (int)ld, X);
    c->launches += 1;
}

int32_t wide_critic_grads(pdeb200_ctx* c, double gamma, int literal, int64_t global_batch) {
    Agent* a = ag(c);
    const int B = a->batch, ns = a->ns, na = a->na, nin = ns + na, ldin = pad4i(nin);
    const HostNet &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC], &At = c->nets[PDEB200_NET_TARGET_ACTOR], &Ct = c->nets[PDEB200_NET_TARGET_CRITIC];
    int32_t rc = arena_reserve(c, wide_arena_floats(c));
    if (rc) return rc;
    const unsigned gcat = (unsigned)(((long long)B * ldin + 255) / 256);
    // a' = A_t(s'), q_t = C_t([s'; a'])
    Acts acts_at, acts_ct, acts_c;
    const int lds = pad4i(ns);
    float* s2p = arena_take(a, (size_t)B * lds);
    float* x2 = arena_take(a, (size_t)B * ldin);
    float* x = arena_take(a, (size_t)B * ldin);
    float* dq = arena_take(a, (size_t)B * 4);
    if (dq) PDEB_CUDA(c, cudaMemsetAsync(dq, 0, (size_t)B * 4 * sizeof(float), c->stream));
    if (!s2p || !x2 || !x || !dq) return fail(c, PDEB200_ECUDA, "ddpg (layer-wise path): activation arena exhausted");
    copy_pad_kernel<<<(unsigned)(((long long)B * lds + 255) / 256), 256, 0, c->stream>>>(B, ns, lds, a->bs2, s2p);
    if ((rc = wide_forward(c, At, B, s2p, lds, &acts_at))) return rc;
    concat_kernel<<<gcat, 256, 0, c->stream>>>(B, ns, na, ldin, a->bs2, acts_at.out[At.n_layers], (int)acts_at.ld[At.n_layers], x2);
    if ((rc = wide_forward(c, Ct, B, x2, ldin, &acts_ct))) return rc;
    // q = C([s; a]) with kept activations
    concat_kernel<<<gcat, 256, 0, c->stream>>>(B, ns, na, ldin, a->bs, a->ba, na, x);
    if ((rc = wide_forward(c, C, B, x, ldin, &acts_c))) return rc;
    critic_dq_kernel<<<1, 1024, 0, c->stream>>>(B, acts_c.out[C.n_layers], acts_c.ld[C.n_layers], acts_ct.out[Ct.n_layers],
                                                 acts_ct.ld[Ct.n_layers], a->br, a->bt, (float)gamma, literal, 1.0 / (double)global_batch,
                                                 a->stats, dq, 4);
    c->launches += 4;
    if ((rc = wide_backward(c, C, B, acts_c, dq, 4, c->d_grads, false, nullptr, nullptr))) return rc;
    PDEB_CUDA(c, cudaGetLastError());
    return PDEB200_OK;
}

int32_t wide_actor_grads(pdeb200_ctx* c, int64_t global_batch) {
    Agent* a = ag(c);
    const int B = a->batch, ns = a->ns, na = a->na, nin = ns + na, ldin = pad4i(nin);
    const HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    int32_t rc = arena_reserve(c, wide_arena_floats(c));
    if (rc) return rc;
    Acts acts_a, acts_c;
    const int lds = pad4i(ns);
    float* sp = arena_take(a, (size_t)B * lds);
    float* x = arena_take(a, (size_t)B * ldin);
    float* dq = arena_take(a, (size_t)B * 4);
    float* da = arena_take(a, (size_t)B * pad4i(na));
    if (dq) PDEB_CUDA(c, cudaMemsetAsync(dq, 0, (size_t)B * 4 * sizeof(float), c->stream));
    if (!sp || !x || !dq || !da) return fail(c, PDEB200_ECUDA, "ddpg (layer-wise path): activation arena exhausted");
    copy_pad_kernel<<<(unsigned)(((long long)B * lds + 255) / 256), 256, 0, c->stream>>>(B, ns, lds, a->bs, sp);
    if ((rc = wide_forward(c, A, B, sp, lds, &acts_a))) return rc;
    concat_kernel<<<(unsigned)(((long long)B * ldin + 255) / 256), 256, 0, c->stream>>>(B, ns, na, ldin, a->bs, acts_a.out[A.n_layers],
                                                                                        (int)acts_a.ld[A.n_layers], x);
    if ((rc = wide_forward(c, C, B, x, ldin, &acts_c))) return rc;
    actor_dq_kernel<<<1, 1024, 0, c->stream>>>(B, acts_c.out[C.n_layers], acts_c.ld[C.n_layers], 1.0 / (double)global_batch, a->stats, dq, 4);
    c->launches += 3;
    float* dx = nullptr; long long lddx = 0;
    if ((rc = wide_backward(c, C, B, acts_c, dq, 4, nullptr, true, &dx, &lddx))) return rc;
    // dLoss/da = the action rows of dLoss/d[s; a]
    const int lda = pad4i(na);
    concat_kernel<<<(unsigned)(((long long)B * lda + 255) / 256), 256, 0, c->stream>>>(B, 0, na, lda, nullptr, dx + ns, (int)lddx, da);
    c->launches += 1;
    if ((rc = wide_backward(c, A, B, acts_a, da, lda, c->d_grads + C.n_params, false, nullptr, nullptr))) return rc;
    PDEB_CUDA(c, cudaGetLastError());
    return PDEB200_OK;
}

}  // namespace


double* agent_stats(pdeb200_ctx* c) { return c->agent ? ag(c)->stats : nullptr; }

void agent_invalidate_graph(pdeb200_ctx* c) {
    if (!c->agent) return;
    if (ag(c)->graph && c->stream) cudaStreamSynchronize(c->stream);
    drop_graph(ag(c));
}

void agent_free(pdeb200_ctx* c) {
    Agent* a = ag(c);
    if (!a) return;
    drop_graph(a);
    for (void* p : {(void*)a->state, (void*)a->action, (void*)a->reward, (void*)a->terminal, (void*)a->bs, (void*)a->ba,
                    (void*)a->br, (void*)a->bs2, (void*)a->bt, (void*)a->inds, (void*)a->stats, (void*)a->partials,
                    (void*)a->arena, (void*)a->stat_part, (void*)a->tickets, (void*)a->dev, (void*)a->xbuf, (void*)a->timeline})
        if (p) cudaFree(p);
    delete a;
    c->agent = nullptr;
}

}  // namespace pdeb200

using namespace pdeb200;

// ---- replay rings ---------------------------------------------------------------------------------------------------
static void ring_advance(Ring& r, int64_t n) {
    const int64_t over = std::max<int64_t>(0, r.len + n - r.cap);
    r.len = std::min(r.cap, r.len + n);
    r.start = (r.start + over) % r.cap;
}

// host rings changed without a push kernel: refresh the device mirror
static int32_t sync_rings(pdeb200_ctx* c) {
    Agent* a = ag(c);
    if (!a->ring_dirty) return PDEB200_OK;
    ring_sync_kernel<<<1, 1, 0, c->stream>>>(a->dev, a->sa, a->rt);
    PDEB_CUDA(c, cudaGetLastError());
    a->ring_dirty = false;
    c->launches += 1;
    return PDEB200_OK;
}

static int32_t push_sa(pdeb200_ctx* c, int zero_action) {
    Agent* a = ag(c);
    if (!a || !a->state) return fail(c, PDEB200_ESTATE, "trajectory not created");
    const int64_t n = a->ncols, pos0 = (a->sa.start + a->sa.len) % a->sa.cap, cap = a->sa.cap;
    const int tpb = 128; const int grid = (int)((n + tpb - 1) / tpb);
    ring_advance(a->sa, n);
    if (c->cfg.dtype == PDEB200_F64)
        push_sa_kernel<double><<<grid, tpb, 0, c->stream>>>(n, a->ns, a->na, cap, pos0, (const double*)c->state,
                                                            (const double*)c->action_in, zero_action, a->state, a->action, a->dev, a->sa);
    else
        push_sa_kernel<float><<<grid, tpb, 0, c->stream>>>(n, a->ns, a->na, cap, pos0, (const float*)c->state,
                                                           (const float*)c->action_in, zero_action, a->state, a->action, a->dev, a->sa);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

extern "C" {

int32_t pdeb200_traj_create(pdeb200_ctx* c, int64_t capacity) {
    if (!c || capacity < 1) return fail(c, PDEB200_EINVAL, "traj_create: bad argument");
    cudaSetDevice(c->device);
    int32_t rc = ensure_agent(c);
    if (rc) return rc;
    Agent* a = ag(c);
    if (capacity < 2 * a->ncols) return fail(c, PDEB200_EINVAL, "traj_create: capacity must hold at least two env steps of columns");
    agent_invalidate_graph(c);
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (void* p : {(void*)a->state, (void*)a->action, (void*)a->reward, (void*)a->terminal})
        if (p) cudaFree(p);
    a->state = a->action = a->reward = nullptr; a->terminal = nullptr;
    a->sa = Ring{capacity + 1, 0, 0};
    a->rt = Ring{capacity, 0, 0};
    a->ring_dirty = true;
    PDEB_CUDA(c, cudaMalloc(&a->state, (size_t)(capacity + 1) * a->ns * 4));
    PDEB_CUDA(c, cudaMalloc(&a->action, (size_t)(capacity + 1) * a->na * 4));
    PDEB_CUDA(c, cudaMalloc(&a->reward, (size_t)capacity * 4));
    PDEB_CUDA(c, cudaMalloc(&a->terminal, (size_t)capacity));
    return PDEB200_OK;
}

int32_t pdeb200_traj_length(const pdeb200_ctx* c, int64_t* length) {
    if (!c || !length || !c->agent) return fail(c, PDEB200_ESTATE, "traj_length: no trajectory");
    *length = static_cast<Agent*>(c->agent)->rt.len;
    return PDEB200_OK;
}

int32_t pdeb200_traj_info(const pdeb200_ctx* c, int64_t* capacity, int64_t* n_sa, int64_t* n_rt, int64_t* first_sa, int64_t* first_rt) {
    if (!c || !c->agent || !static_cast<Agent*>(c->agent)->state) return fail(c, PDEB200_ESTATE, "traj_info: no trajectory");
    const Agent* a = static_cast<Agent*>(c->agent);
    if (capacity) *capacity = a->rt.cap;
    if (n_sa) *n_sa = a->sa.len;
    if (n_rt) *n_rt = a->rt.len;
    if (first_sa) *first_sa = a->sa.start;
    if (first_rt) *first_rt = a->rt.start;
    return PDEB200_OK;
}

// ring -> logical order: two contiguous pieces [start, cap) and [0, start + len - cap)
static cudaError_t ring_read(const Ring& r, size_t row_bytes, const void* ring, void* dst, cudaStream_t st) {
    if (!dst || r.len == 0) return cudaSuccess;
    const int64_t first = std::min(r.len, r.cap - r.start);
    cudaError_t e = cudaMemcpyAsync(dst, (const char*)ring + r.start * row_bytes, first * row_bytes, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess || first == r.len) return e;
    return cudaMemcpyAsync((char*)dst + first * row_bytes, ring, (r.len - first) * row_bytes, cudaMemcpyDeviceToHost, st);
}

int32_t pdeb200_traj_get(pdeb200_ctx* c, float* state, float* action, float* reward, uint8_t* terminal) {
    if (!c) return PDEB200_EINVAL;
    Agent* a = ag(c);
    if (!a || !a->state) return fail(c, PDEB200_ESTATE, "traj_get: no trajectory");
    cudaSetDevice(c->device);
    PDEB_CUDA(c, ring_read(a->sa, (size_t)a->ns * 4, a->state, state, c->stream));
    PDEB_CUDA(c, ring_read(a->sa, (size_t)a->na * 4, a->action, action, c->stream));
    PDEB_CUDA(c, ring_read(a->rt, 4, a->reward, reward, c->stream));
    PDEB_CUDA(c, ring_read(a->rt, 1, a->terminal, terminal, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

// logical order -> ring with its oldest column at raw position `start`
static cudaError_t ring_write(const Ring& r, size_t row_bytes, void* ring, const void* src, cudaStream_t st) {
    if (r.len == 0) return cudaSuccess;
    const int64_t first = std::min(r.len, r.cap - r.start);
    cudaError_t e = cudaMemcpyAsync((char*)ring + r.start * row_bytes, src, first * row_bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess || first == r.len) return e;
    return cudaMemcpyAsync(ring, (const char*)src + first * row_bytes, (r.len - first) * row_bytes, cudaMemcpyHostToDevice, st);
}

int32_t pdeb200_traj_set(pdeb200_ctx* c, int64_t n_sa, int64_t n_rt, int64_t first_sa, int64_t first_rt, const float* state,
                         const float* action, const float* reward, const uint8_t* terminal) {
    if (!c) return PDEB200_EINVAL;
    Agent* a = ag(c);
    if (!a || !a->state) return fail(c, PDEB200_ESTATE, "traj_set: no trajectory (pdeb200_traj_create first)");
    if (n_sa < 0 || n_rt < 0 || n_sa > a->sa.cap || n_rt > a->rt.cap || (n_sa && (!state || !action)) || (n_rt && (!reward || !terminal)) ||
        first_sa < 0 || first_sa >= a->sa.cap || first_rt < 0 || first_rt >= a->rt.cap)
        return fail(c, PDEB200_EINVAL, "traj_set: bad sizes, ring positions or null arrays");
    cudaSetDevice(c->device);
    a->sa.start = first_sa; a->sa.len = n_sa;
    a->rt.start = first_rt; a->rt.len = n_rt;
    a->ring_dirty = true;
    PDEB_CUDA(c, ring_write(a->sa, (size_t)a->ns * 4, a->state, state, c->stream));
    PDEB_CUDA(c, ring_write(a->sa, (size_t)a->na * 4, a->action, action, c->stream));
    PDEB_CUDA(c, ring_write(a->rt, 4, a->reward, reward, c->stream));
    PDEB_CUDA(c, ring_write(a->rt, 1, a->terminal, terminal, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_rng_get(pdeb200_ctx* c, uint64_t* offset) {
    if (!c || !offset) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    int32_t rc = ensure_agent(c);
    if (rc) return rc;
    unsigned long long v = 0;
    PDEB_CUDA(c, cudaMemcpyAsync(&v, &ag(c)->dev->rng_offset, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    *offset = v;
    return PDEB200_OK;
}

int32_t pdeb200_rng_set(pdeb200_ctx* c, uint64_t offset) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    int32_t rc = ensure_agent(c);
    if (rc) return rc;
    rng_set_kernel<<<1, 1, 0, c->stream>>>(ag(c)->dev, (unsigned long long)offset);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

/* PreActStage: push (state[:, i], action[:, i]) for every column.  The action pushed is the one the
 * policy just produced (the staged action buffer), exactly what `agent(PRE_ACT_STAGE, env, action)` gets. */
int32_t pdeb200_traj_push_pre(pdeb200_ctx* c) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    return push_sa(c, 0);
}

int32_t pdeb200_traj_episode_end(pdeb200_ctx* c) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    return push_sa(c, 1);
}

int32_t pdeb200_traj_push_post(pdeb200_ctx* c) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    Agent* a = ag(c);
    if (!a || !a->reward) return fail(c, PDEB200_ESTATE, "trajectory not created");
    const int64_t n = a->ncols, pos0 = (a->rt.start + a->rt.len) % a->rt.cap, cap = a->rt.cap;
    const int tpb = 128; const int grid = (int)((n + tpb - 1) / tpb);
    ring_advance(a->rt, n);
    if (c->cfg.dtype == PDEB200_F64)
        push_rt_kernel<double><<<grid, tpb, 0, c->stream>>>(n, c->n_cols, cap, pos0, (const double*)c->reward, c->done,
                                                            a->reward, a->terminal, a->dev, a->rt);
    else
        push_rt_kernel<float><<<grid, tpb, 0, c->stream>>>(n, c->n_cols, cap, pos0, (const float*)c->reward, c->done,
                                                           a->reward, a->terminal, a->dev, a->rt);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

int32_t pdeb200_traj_pop_tail(pdeb200_ctx* c) {
    if (!c) return PDEB200_EINVAL;
    Agent* a = ag(c);
    if (!a || !a->state) return fail(c, PDEB200_ESTATE, "trajectory not created");
    if (a->rt.len > 0) { a->sa.len = std::max<int64_t>(0, a->sa.len - a->ncols); a->ring_dirty = true; }      // PDEagent.jl:243-251
    return PDEB200_OK;
}

}  // extern "C"

// ---- sampler ---------------------------------------------------------------------------------------------------------
// one launch: index draw (device Philox unless the host supplied indices), gather, reward statistics (+ their exchange)
static int32_t sample_launch(pdeb200_ctx* c, int batch, int draw, uint64_t seed, uint64_t offset, int dev_state) {
    Agent* a = ag(c);
    const int tpb = 128, grid = (batch + tpb - 1) / tpb;
    fetch_kernel<<<grid, tpb, 0, c->stream>>>(batch, a->ns, a->na, a->ncols, a->sa, a->rt, a->inds, draw, seed, offset, dev_state,
                                              a->dev, a->state, a->action, a->reward, a->terminal, a->bs, a->ba, a->br, a->bt,
                                              a->bs2, a->stat_part, a->tickets, a->stats, comm_dev(c));
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    if (comm_nranks(c) > 1 && comm_transport(c) == PDEB200_COMM_NCCL) return comm_allreduce_f64(c, a->stats, 3);
    return PDEB200_OK;
}

extern "C" {

int32_t pdeb200_sample(pdeb200_ctx* c, int32_t batch, const int64_t* inds_host, uint64_t seed, uint64_t offset) {
    if (!c || batch < 1) return fail(c, PDEB200_EINVAL, "sample: bad argument");
    cudaSetDevice(c->device);
    Agent* a = ag(c);
    if (!a || !a->state) return fail(c, PDEB200_ESTATE, "sample: trajectory not created");
    const int64_t range = a->rt.len - a->ncols;
    if (range < 1) return fail(c, PDEB200_ESTATE, "sample: trajectory shorter than one env step of columns");
    int32_t rc = ensure_batch(c, batch);
    if (rc) return rc;
    if (inds_host) {
        for (int i = 0; i < batch; ++i)
            if (inds_host[i] < 0 || inds_host[i] >= range) return fail(c, PDEB200_EINVAL, "sample: index out of range");
        PDEB_CUDA(c, cudaMemcpyAsync(a->inds, inds_host, (size_t)batch * 8, cudaMemcpyHostToDevice, c->stream));
    }
    if ((rc = sample_launch(c, batch, inds_host ? 0 : 1, seed, offset, 0))) return rc;
    if (inds_host) PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_set_batch(pdeb200_ctx* c, int32_t batch, const float* s, const float* a_, const float* r, const uint8_t* t,
                          const float* s2) {
    if (!c || batch < 1 || !s || !a_ || !r || !t || !s2) return fail(c, PDEB200_EINVAL, "set_batch: bad argument");
    cudaSetDevice(c->device);
    int32_t rc = ensure_agent(c);
    if (rc) return rc;
    if ((rc = ensure_batch(c, batch))) return rc;
    Agent* a = ag(c);
    PDEB_CUDA(c, cudaMemcpyAsync(a->bs, s, (size_t)batch * a->ns * 4, cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaMemcpyAsync(a->ba, a_, (size_t)batch * a->na * 4, cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaMemcpyAsync(a->br, r, (size_t)batch * 4, cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaMemcpyAsync(a->bt, t, (size_t)batch, cudaMemcpyHostToDevice, c->stream));
    PDEB_CUDA(c, cudaMemcpyAsync(a->bs2, s2, (size_t)batch * a->ns * 4, cudaMemcpyHostToDevice, c->stream));
    reward_stats_kernel<<<1, 256, 0, c->stream>>>(batch, a->br, a->stats, comm_dev(c));
    c->launches += 1;
    if (comm_nranks(c) > 1 && comm_transport(c) == PDEB200_COMM_NCCL && (rc = comm_allreduce_f64(c, a->stats, 3))) return rc;
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_get_batch(pdeb200_ctx* c, float* s, float* a_, float* r, uint8_t* t, float* s2, int64_t* inds) {
    if (!c) return PDEB200_EINVAL;
    Agent* a = ag(c);
    if (!a || !a->batch) return fail(c, PDEB200_ESTATE, "get_batch: no batch (pdeb200_sample / pdeb200_set_batch first)");
    cudaSetDevice(c->device);
    const size_t B = a->batch;
    if (s) PDEB_CUDA(c, cudaMemcpyAsync(s, a->bs, B * a->ns * 4, cudaMemcpyDeviceToHost, c->stream));
    if (a_) PDEB_CUDA(c, cudaMemcpyAsync(a_, a->ba, B * a->na * 4, cudaMemcpyDeviceToHost, c->stream));
    if (r) PDEB_CUDA(c, cudaMemcpyAsync(r, a->br, B * 4, cudaMemcpyDeviceToHost, c->stream));
    if (t) PDEB_CUDA(c, cudaMemcpyAsync(t, a->bt, B, cudaMemcpyDeviceToHost, c->stream));
    if (s2) PDEB_CUDA(c, cudaMemcpyAsync(s2, a->bs2, B * a->ns * 4, cudaMemcpyDeviceToHost, c->stream));
    if (inds) PDEB_CUDA(c, cudaMemcpyAsync(inds, a->inds, B * 8, cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

// PDEB200_DDPG_TIMELINE=1: copy out (and reset) the device timestamps of the update kernels; out: 8 words per kernel record
// {phase, t_entry, t_loop_done, t_tail_start, t_reduced, t_exchanged, t_end, -}, returns the record count in *n
int32_t pdeb200_debug_timeline(pdeb200_ctx* c, uint64_t* out, int32_t max_records, int32_t* n) {
    if (!c || !out || !n) return PDEB200_EINVAL;
    Agent* a = ag(c);
    *n = 0;
    if (!a || !a->timeline) return PDEB200_OK;
    cudaSetDevice(c->device);
    std::vector<unsigned long long> h(8192);
    PDEB_CUDA(c, cudaMemcpyAsync(h.data(), a->timeline, h.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    const int k = (int)std::min<unsigned long long>(h[0], (unsigned long long)std::min(max_records, 500));
    for (int i = 0; i < k * 8; ++i) out[i] = h[8 + i];
    *n = k;
    PDEB_CUDA(c, cudaMemsetAsync(a->timeline, 0, 8 * sizeof(unsigned long long), c->stream));
    return PDEB200_OK;
}

int32_t pdeb200_ddpg_set_path(pdeb200_ctx* c, int32_t path) {
    if (!c || path < 0 || path > 3) return fail(c, PDEB200_EINVAL, "ddpg_set_path: bad argument");
    int32_t rc = ensure_agent(c);
    if (rc) return rc;
    agent_invalidate_graph(c);
    ag(c)->force_wide = path != 0;
    ag(c)->wide_path = path == 3 ? 0 : path;
    return PDEB200_OK;
}

}  // extern "C"

// ---- DDPG update -------------------------------------------------------------------------------------------------------
namespace {

// How one update runs for the current networks and batch.  Decided ONCE for both phases: the fused shared-memory
// kernels need the critic AND the actor phase to fit; otherwise both take the layer-wise GEMM path.
struct Plan {
    bool fused = false;
    size_t smem_c = 0, smem_a = 0;
    int tpb_c = 0, tpb_a = 0, n_blocks = 0, wmax = 0;
    // register-resident kernels for the shipped two-layer shapes (ddpg_fast.cuh)
    bool fast = false;
    int f_ns = 0, f_cfg = 0, f_warps = 0, f_grid = 0, f_nx_c = 0, f_nx_a = 0;
    size_t f_smem_c = 0, f_smem_a = 0;
    int f_cluster = 0;                 // > 0: the whole update as ONE launch of a single thread-block cluster of this many CTAs
    size_t f_smem_cl = 0;
};

inline int C_params(const pdeb200_ctx* c) { return c->nets[PDEB200_NET_BEHAVIOR_CRITIC].n_params; }

template <int NS, int UPLC, int WC>
int32_t fast_configure(pdeb200_ctx* c, Plan* P) {
    using Geo = FastGeom<NS, UPLC, WC>;
    P->f_smem_c = Geo::smem_bytes(P->f_warps, P->f_nx_c);
    P->f_smem_a = Geo::smem_bytes(P->f_warps, P->f_nx_a);
    if (P->f_smem_c > 220 * 1024) { P->fast = false; return PDEB200_OK; }
    PDEB_CUDA(c, ensure_dyn_smem(ddpg_fast_critic_kernel<NS, UPLC, WC>, P->f_smem_c, c->device));
    PDEB_CUDA(c, ensure_dyn_smem(ddpg_fast_actor_kernel<NS, UPLC, WC>, P->f_smem_a, c->device));
    // one-cluster form: possible when the batch's tiles fit 16 CTAs' warps in a few rounds and the cluster can be co-scheduled
    P->f_cluster = 0;
    static const int want_cluster = [] { const char* e = getenv("PDEB200_DDPG_CLUSTER"); return e ? atoi(e) : 16; }();
    const int teams = P->f_warps / WC;
    const int n_tiles = (ag(c)->batch + 31) / 32;
    const int need = (n_tiles + teams - 1) / teams;                       // CTAs for one tile per team
    const int per_slice = 2 * ((C_params(c) + 15) / 16) + kXTail;
    const int max_cs = std::min(want_cluster, 16);
    const int per_cta = kSlicePF * P->f_warps * 32;                       // parameters one CTA's slice can hold
    const int min_cs = (C_params(c) + per_cta - 1) / per_cta;
    if (want_cluster > 0 && need <= 4 * max_cs && min_cs <= max_cs && (comm_nranks(c) == 1 || per_slice <= comm_cap(c) / 2 / kMaxSlices)) {
        int cs = 1;
        while ((cs < need || cs < min_cs) && cs < max_cs) cs *= 2;
        const size_t smem = (Geo::base_floats(P->f_warps, P->f_nx_c) + 2 * (size_t)P->f_nx_c) * sizeof(float);
        auto kern = ddpg_fast_cluster_kernel<NS, UPLC, WC>;
        if (smem <= 220 * 1024 && ensure_dyn_smem(kern, smem, c->device) == cudaSuccess &&
            (cs <= 8 || cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess)) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs); cfg.blockDim = dim3(P->f_warps * 32); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) == cudaSuccess && n_clusters >= 1) {
                P->f_cluster = cs; P->f_smem_cl = smem;
            }
        }
        cudaGetLastError();
    }
    return PDEB200_OK;
}

template <int NS, int UPLC, int WC>
void fast_launch_cluster_t(pdeb200_ctx* c, const Plan& P, const FastArgs& Fc, const FastArgs& Fa) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.f_cluster); cfg.blockDim = dim3(P.f_warps * 32); cfg.dynamicSmemBytes = P.f_smem_cl; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = P.f_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, ddpg_fast_cluster_kernel<NS, UPLC, WC>, Fc, Fa);
}

template <int NS, int UPLC, int WC>
void fast_launch_t(pdeb200_ctx* c, const Plan& P, const FastArgs& Fc, const FastArgs& Fa) {
    ddpg_fast_critic_kernel<NS, UPLC, WC><<<P.f_grid, P.f_warps * 32, P.f_smem_c, c->stream>>>(Fc);
    ddpg_fast_actor_kernel<NS, UPLC, WC><<<P.f_grid, P.f_warps * 32, P.f_smem_a, c->stream>>>(Fa);
}

#define PDEB_FAST_DISPATCH(P, CALL)                                                     \
    do {                                                                               \
        if ((P).f_cfg == 0) {                                                          \
            switch ((P).f_ns) {                                                        \
                case 1: CALL(1, 5, 1); break;                                          \
                case 3: CALL(3, 5, 1); break;                                          \
            }                                                                          \
        } else {                                                                       \
            switch ((P).f_ns) {                                                        \
                case 1: CALL(1, 3, 4); break;                                          \
                case 3: CALL(3, 3, 4); break;                                          \
                case 9: CALL(9, 3, 4); break;                                          \
                case 12: CALL(12, 3, 4); break;                                        \
            }                                                                          \
        }                                                                              \
    } while (0)

// two-layer relu networks with one action row in one of the compiled shapes?
int32_t plan_fast(pdeb200_ctx* c, Plan* P) {
    Agent* a = ag(c);
    static const bool off = [] { const char* e = getenv("PDEB200_DDPG_FAST"); return e && atoi(e) == 0; }();
    static const int env_warps = [] { const char* e = getenv("PDEB200_DDPG_WARPS"); return e ? atoi(e) : 0; }();
    const HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    const HostNet &At = c->nets[PDEB200_NET_TARGET_ACTOR], &Ct = c->nets[PDEB200_NET_TARGET_CRITIC];
    P->fast = false;
    if (off || a->force_wide || a->na != 1 || A.n_layers != 2 || C.n_layers != 2 || At.n_layers != 2 || Ct.n_layers != 2) return PDEB200_OK;
    const int ns = A.sizes[0], ha = A.sizes[1], hc = C.sizes[1];
    if (!(ns == 1 || ns == 3 || ns == 9 || ns == 12) || ha > 32 || hc > 384) return PDEB200_OK;
    if (A.acts[0] != PDEB200_ACT_RELU || C.acts[0] != PDEB200_ACT_RELU || C.acts[1] != PDEB200_ACT_IDENTITY) return PDEB200_OK;
    for (int l = 0; l <= 2; ++l)
        if (At.sizes[l] != A.sizes[l] || Ct.sizes[l] != C.sizes[l]) return PDEB200_OK;
    if (At.acts[0] != A.acts[0] || At.acts[1] != A.acts[1] || Ct.acts[0] != C.acts[0] || Ct.acts[1] != C.acts[1]) return PDEB200_OK;
    P->fast = true;
    P->f_ns = ns;
    P->f_cfg = (hc <= 160 && ns <= 3) ? 0 : 1;      // (UPL 5, 1 warp per tile) or (UPL 3, 4 warps per tile: hc <= 384, wide inputs)
    const int wc = P->f_cfg == 0 ? 1 : 4;
    P->f_warps = env_warps > 0 ? std::max(wc, env_warps / wc * wc) : 8;
    P->f_warps = std::min(P->f_warps, 8);
    const int teams = P->f_warps / wc;
    const int n_tiles = (a->batch + 31) / 32;
    P->f_grid = std::max(1, std::min((n_tiles + teams - 1) / teams, 148));
    P->f_nx_c = 2 * C.n_params + kXTail;
    P->f_nx_a = A.n_params + kXTail;
    if (comm_nranks(c) > 1 && P->f_nx_c > comm_cap(c) / 2) { P->fast = false; return PDEB200_OK; }
    int32_t rc = ensure_partials(c, P->f_grid, P->f_nx_c - 2);
    if (rc) return rc;
    static const bool want_tl = [] { const char* e = getenv("PDEB200_DDPG_TIMELINE"); return e && atoi(e) != 0; }();
    if (want_tl && !a->timeline) {                       // allocated here, never inside a stream capture
        PDEB_CUDA(c, cudaMalloc(&a->timeline, 8192 * sizeof(unsigned long long)));
        PDEB_CUDA(c, cudaMemset(a->timeline, 0, 8192 * sizeof(unsigned long long)));
    }
#define PDEB_CFG(NS_, U_, W_) rc = fast_configure<NS_, U_, W_>(c, P)
    PDEB_FAST_DISPATCH(*P, PDEB_CFG);
#undef PDEB_CFG
    return rc;
}

int32_t plan_update(pdeb200_ctx* c, Plan* P) {
    Agent* a = ag(c);
    if (!a || !a->batch) return fail(c, PDEB200_ESTATE, "ddpg: no batch (pdeb200_sample / pdeb200_set_batch first)");
    int32_t rc = check_nets(c);
    if (rc) return rc;
    const HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    P->wmax = std::max({net_wmax(c->nets[0]), net_wmax(c->nets[1]), a->ns + a->na});
    P->smem_c = ((size_t)net_act_floats(C) + 2 * (size_t)TS * P->wmax + (size_t)TS * (a->ns + a->na) + TS + C.n_params) * 4;
    P->smem_a = ((size_t)net_act_floats(C) + net_act_floats(A) + 2 * (size_t)TS * P->wmax + A.n_params) * 4;
    P->fused = !a->force_wide && P->smem_c <= 220 * 1024 && P->smem_a <= 220 * 1024;
    const int n_tiles = (a->batch + TS - 1) / TS;
    P->n_blocks = std::min(n_tiles, 2 * 148);
    P->tpb_c = block_threads(P->wmax, C.n_params);
    P->tpb_a = block_threads(P->wmax, A.n_params);
    if ((rc = ensure_partials(c, P->n_blocks, std::max(C.n_params, A.n_params)))) return rc;
    if (P->fused) {
        PDEB_CUDA(c, ensure_dyn_smem(ddpg_critic_kernel, P->smem_c, c->device));
        PDEB_CUDA(c, ensure_dyn_smem(ddpg_actor_kernel, P->smem_a, c->device));
    }
    return plan_fast(c, P);
}

struct Hyper { double gamma, polyak, lr_a, lr_c; int literal; };

// One update with the register-resident kernels: two launches.  fetch = 1: the critic kernel draws and gathers the batch
// itself from the rings (device-resident ring positions and Philox counter).
int32_t fast_update_launch(pdeb200_ctx* c, const Plan& P, const Hyper& H, int fetch, uint64_t seed) {
    Agent* a = ag(c);
    HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    HostNet &At = c->nets[PDEB200_NET_TARGET_ACTOR], &Ct = c->nets[PDEB200_NET_TARGET_CRITIC];
    FastArgs F;
    F.pA = A.d_params; F.pC = C.d_params; F.pAt = At.d_params; F.pCt = Ct.d_params;
    F.ns = A.sizes[0]; F.ha = A.sizes[1]; F.hc = C.sizes[1]; F.actA2 = A.acts[1];
    F.nA = A.n_params; F.nC = C.n_params; F.batch = a->batch;
    F.bs = a->bs; F.ba = a->ba; F.br = a->br; F.bs2 = a->bs2; F.bt = a->bt; F.inds = a->inds;
    F.fetch = fetch; F.seed = seed; F.dev = a->dev; F.ncols = a->ncols;
    F.rstate = a->state; F.raction = a->action; F.rreward = a->reward; F.rterminal = a->terminal;
    F.gamma = (float)H.gamma; F.literal = H.literal;
    F.partials = a->partials; F.xbuf = a->xbuf; F.stats = a->stats; F.losses = c->d_losses;
    F.b1 = 0.9; F.b2 = 0.999; F.eps = 1e-8; F.polyak = (float)H.polyak;
    F.cm = comm_dev(c);
    F.tl = a->timeline;
    FastArgs Fc = F, Fa = F;
    Fc.n_x = P.f_nx_c; Fc.ticket = a->tickets + 1; Fc.grads = c->d_grads;
    Fc.x = C.d_params; Fc.m = C.d_m; Fc.v = C.d_v; Fc.target = Ct.d_params; Fc.betap = C.d_betap; Fc.eta = H.lr_c;
    Fa.n_x = P.f_nx_a; Fa.ticket = a->tickets + 2; Fa.grads = c->d_grads + C.n_params; Fa.fetch = 0;
    Fa.x = A.d_params; Fa.m = A.d_m; Fa.v = A.d_v; Fa.target = At.d_params; Fa.betap = A.d_betap; Fa.eta = H.lr_a;
    if (P.f_cluster > 0) {
#define PDEB_LAUNCH(NS_, U_, W_) fast_launch_cluster_t<NS_, U_, W_>(c, P, Fc, Fa)
        PDEB_FAST_DISPATCH(P, PDEB_LAUNCH);
#undef PDEB_LAUNCH
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 1;
        return PDEB200_OK;
    }
#define PDEB_LAUNCH(NS_, U_, W_) fast_launch_t<NS_, U_, W_>(c, P, Fc, Fa)
    PDEB_FAST_DISPATCH(P, PDEB_LAUNCH);
#undef PDEB_LAUNCH
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return PDEB200_OK;
}

// Gradient kernels of the shared-memory path.  tail != nullptr: the last CTA reduces, exchanges (cm) and applies the
// optimiser; otherwise the per-CTA partials are reduced into ARR_GRADS by a second launch.  global_batch: see make_args.
int32_t critic_kernel_launch(pdeb200_ctx* c, const Plan& P, double gamma, int literal, int64_t global_batch, const Hyper* tail) {
    Agent* a = ag(c);
    HostNet& Cw = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    DdpgArgs D = make_args(c, gamma, literal, global_batch);
    D.n_acc = Cw.n_params; D.wmax = P.wmax; D.partials = a->partials; D.stat0 = ST_C;
    if (tail) {
        D.fuse = 1; D.ticket = a->tickets + 1; D.grads = c->d_grads; D.cm = comm_dev(c);
        D.x = Cw.d_params; D.m = Cw.d_m; D.v = Cw.d_v; D.target = c->nets[PDEB200_NET_TARGET_CRITIC].d_params; D.betap = Cw.d_betap;
        D.eta = tail->lr_c; D.b1 = 0.9; D.b2 = 0.999; D.eps = 1e-8; D.polyak = (float)tail->polyak;
    }
    ddpg_critic_kernel<<<P.n_blocks, P.tpb_c, P.smem_c, c->stream>>>(D);
    c->launches += 1;
    if (!tail) {
        reduce_partials_kernel<<<(D.n_acc + 2 + 127) / 128, 128, 0, c->stream>>>(P.n_blocks, D.n_acc, a->partials, c->d_grads, a->stats, ST_C);
        c->launches += 1;
    }
    PDEB_CUDA(c, cudaGetLastError());
    return PDEB200_OK;
}

int32_t actor_kernel_launch(pdeb200_ctx* c, const Plan& P, int64_t global_batch, const Hyper* tail) {
    Agent* a = ag(c);
    HostNet& Aw = c->nets[PDEB200_NET_BEHAVIOR_ACTOR];
    const int n_c = c->nets[PDEB200_NET_BEHAVIOR_CRITIC].n_params;
    DdpgArgs D = make_args(c, 0.0, tail ? tail->literal : 0, global_batch);
    D.n_acc = Aw.n_params; D.wmax = P.wmax; D.partials = a->partials; D.stat0 = ST_Q;
    if (tail) {
        D.fuse = 1; D.ticket = a->tickets + 2; D.grads = c->d_grads + n_c; D.cm = comm_dev(c);
        D.x = Aw.d_params; D.m = Aw.d_m; D.v = Aw.d_v; D.target = c->nets[PDEB200_NET_TARGET_ACTOR].d_params; D.betap = Aw.d_betap;
        D.eta = tail->lr_a; D.b1 = 0.9; D.b2 = 0.999; D.eps = 1e-8; D.polyak = (float)tail->polyak;
        D.losses = c->d_losses;
    }
    ddpg_actor_kernel<<<P.n_blocks, P.tpb_a, P.smem_a, c->stream>>>(D);
    c->launches += 1;
    if (!tail) {
        reduce_partials_kernel<<<(D.n_acc + 2 + 127) / 128, 128, 0, c->stream>>>(P.n_blocks, D.n_acc, a->partials, c->d_grads + n_c, a->stats, ST_Q);
        c->launches += 1;
    }
    PDEB_CUDA(c, cudaGetLastError());
    return PDEB200_OK;
}

int32_t adam_apply(pdeb200_ctx* c, HostNet& n, const float* g, double lr) {
    adam_kernel<<<(n.n_params + 127) / 128, 128, 0, c->stream>>>(n.n_params, n.d_params, n.d_m, n.d_v, g, lr, 0.9, 0.999, n.d_betap, 1e-8,
                                                                  ag(c)->tickets + 3);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

int32_t losses_launch(pdeb200_ctx* c, int64_t global_batch, int literal, int which) {
    losses_kernel<<<1, 1, 0, c->stream>>>(ag(c)->stats, (double)global_batch, literal, which, c->d_losses);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

// gradient phases into ARR_GRADS (reduced over this rank's CTAs, not over ranks), by whichever path the plan chose
int32_t critic_phase(pdeb200_ctx* c, const Plan& P, double gamma, int literal, int64_t global_batch) {
    return P.fused ? critic_kernel_launch(c, P, gamma, literal, global_batch, nullptr) : wide_critic_grads(c, gamma, literal, global_batch);
}
int32_t actor_phase(pdeb200_ctx* c, const Plan& P, int64_t global_batch) {
    return P.fused ? actor_kernel_launch(c, P, global_batch, nullptr) : wide_actor_grads(c, global_batch);
}

int32_t polyak_both(pdeb200_ctx* c, double polyak) {
    HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    HostNet &At = c->nets[PDEB200_NET_TARGET_ACTOR], &Ct = c->nets[PDEB200_NET_TARGET_CRITIC];
    polyak_kernel<<<(A.n_params + 127) / 128, 128, 0, c->stream>>>(A.n_params, At.d_params, A.d_params, (float)polyak);
    polyak_kernel<<<(C.n_params + 127) / 128, 128, 0, c->stream>>>(C.n_params, Ct.d_params, C.d_params, (float)polyak);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return PDEB200_OK;
}

// can the optimiser run in the gradient kernels' tails (two launches per update)?
bool tail_fusable(pdeb200_ctx* c, const Plan& P) {
    if (!P.fused) return false;
    const int nr = comm_nranks(c);
    if (nr == 1) return true;
    const int n_max = std::max(c->nets[PDEB200_NET_BEHAVIOR_CRITIC].n_params, c->nets[PDEB200_NET_BEHAVIOR_ACTOR].n_params);
    return comm_transport(c) == PDEB200_COMM_PEER && n_max + 2 <= comm_cap(c) / 2;
}

// One whole update on the staged batch (PDEagent.jl:363-418).
int32_t enqueue_update(pdeb200_ctx* c, const Plan& P, const Hyper& H) {
    Agent* a = ag(c);
    int32_t rc;
    if (P.fast && tail_fusable(c, P)) return fast_update_launch(c, P, H, 0, 0);
    if (tail_fusable(c, P)) {
        // two launches: {critic gradients; last CTA: reduce + exchange + ADAM + Polyak} {actor gradients; last CTA: same +
        // losses}.  The target critic is not read after the critic phase and the behavior critic is not written by the
        // actor phase, so its Polyak step can run right after its ADAM step (reference order: both at the end,
        // PDEagent.jl:411-417).  The global batch count comes from the sampler's statistics on the device.
        if ((rc = critic_kernel_launch(c, P, H.gamma, H.literal, 0, &H))) return rc;
        return actor_kernel_launch(c, P, 0, &H);
    }
    // layer-wise (wide network) path, or NCCL transport: allreduce between the gradient and the optimiser kernels.
    // Shared-memory kernels read the global batch count from the sampler's statistics (gb = 0); the layer-wise path
    // takes it by value = batch x nranks (equal local batches required there).
    const int nr = comm_nranks(c);
    const int64_t gb = P.fused ? 0 : (int64_t)a->batch * nr;
    HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    if ((rc = critic_phase(c, P, H.gamma, H.literal, gb))) return rc;
    if (nr > 1 && ((rc = comm_allreduce_f32(c, c->d_grads, C.n_params)) || (rc = comm_allreduce_f64(c, a->stats + ST_C, 2)))) return rc;
    if ((rc = losses_launch(c, gb, H.literal, 0))) return rc;
    if ((rc = adam_apply(c, C, c->d_grads, H.lr_c))) return rc;
    if ((rc = actor_phase(c, P, gb))) return rc;
    if (nr > 1 && ((rc = comm_allreduce_f32(c, c->d_grads + C.n_params, A.n_params)) || (rc = comm_allreduce_f64(c, a->stats + ST_Q, 1)))) return rc;
    if ((rc = losses_launch(c, gb, H.literal, 1))) return rc;
    if ((rc = adam_apply(c, A, c->d_grads + C.n_params, H.lr_a))) return rc;
    return polyak_both(c, H.polyak);
}

}  // namespace

extern "C" {

int32_t pdeb200_ddpg_critic_grads(pdeb200_ctx* c, double gamma, int32_t literal_q1, int64_t global_batch) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    Plan P;
    int32_t rc = plan_update(c, &P);
    if (rc) return rc;
    if (global_batch < ag(c)->batch) return fail(c, PDEB200_EINVAL, "ddpg: global_batch < local batch");
    if ((rc = critic_phase(c, P, gamma, literal_q1, global_batch))) return rc;
    return losses_launch(c, global_batch, literal_q1, 0);
}

int32_t pdeb200_ddpg_critic_apply(pdeb200_ctx* c, double lr) {
    if (!c || !c->d_grads) return fail(c, PDEB200_ESTATE, "ddpg: no gradients");
    cudaSetDevice(c->device);
    return adam_apply(c, c->nets[PDEB200_NET_BEHAVIOR_CRITIC], c->d_grads, lr);
}

int32_t pdeb200_ddpg_actor_grads(pdeb200_ctx* c, int64_t global_batch) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    Plan P;
    int32_t rc = plan_update(c, &P);
    if (rc) return rc;
    if (global_batch < ag(c)->batch) return fail(c, PDEB200_EINVAL, "ddpg: global_batch < local batch");
    if ((rc = actor_phase(c, P, global_batch))) return rc;
    return losses_launch(c, global_batch, 0, 1);
}

int32_t pdeb200_ddpg_actor_apply(pdeb200_ctx* c, double lr, double polyak) {
    if (!c || !c->d_grads) return fail(c, PDEB200_ESTATE, "ddpg: no gradients");
    cudaSetDevice(c->device);
    HostNet &A = c->nets[PDEB200_NET_BEHAVIOR_ACTOR], &C = c->nets[PDEB200_NET_BEHAVIOR_CRITIC];
    int32_t rc = adam_apply(c, A, c->d_grads + C.n_params, lr);
    if (rc) return rc;
    return polyak_both(c, polyak);
}

int32_t pdeb200_ddpg_update(pdeb200_ctx* c, double gamma, double polyak, double lr_actor, double lr_critic, int32_t literal_q1) {
    if (!c || !c->agent) return fail(c, PDEB200_ESTATE, "ddpg_update: no batch");
    cudaSetDevice(c->device);
    Plan P;
    int32_t rc = plan_update(c, &P);
    if (rc) return rc;
    const Hyper H{gamma, polyak, lr_actor, lr_critic, literal_q1};
    return enqueue_update(c, P, H);
}

int32_t pdeb200_train_updates(pdeb200_ctx* c, int32_t n_updates, int32_t batch, double gamma, double polyak, double lr_actor,
                              double lr_critic, int32_t literal_q1, uint64_t seed) {
    if (!c || n_updates < 1 || batch < 1) return fail(c, PDEB200_EINVAL, "train_updates: bad argument");
    cudaSetDevice(c->device);
    Agent* a = ag(c);
    if (!a || !a->state) return fail(c, PDEB200_ESTATE, "train_updates: trajectory not created");
    if (a->rt.len - a->ncols < 1) return fail(c, PDEB200_ESTATE, "train_updates: trajectory shorter than one env step of columns");
    int32_t rc = ensure_batch(c, batch);
    if (rc) return rc;
    Plan P;
    if ((rc = plan_update(c, &P))) return rc;
    if ((rc = sync_rings(c))) return rc;
    const Hyper H{gamma, polyak, lr_actor, lr_critic, literal_q1};
    static const bool no_graph = [] { const char* e = getenv("PDEB200_NO_GRAPH"); return e && atoi(e) != 0; }();
    const bool fast = P.fast && tail_fusable(c, P);            // sampler fused into the critic kernel: two launches per update
    auto one_update = [&]() -> int32_t {
        if (fast) return fast_update_launch(c, P, H, 1, seed);
        int32_t r = sample_launch(c, batch, 1, seed, 0, 1);
        return r ? r : enqueue_update(c, P, H);
    };
    if (no_graph || !tail_fusable(c, P)) {
        for (int k = 0; k < n_updates; ++k)
            if ((rc = one_update())) return rc;
        return PDEB200_OK;
    }
    // update_loops x {sample, critic, actor} as one graph: everything that changes between launches (ring positions,
    // Philox counter, beta powers, exchange epochs) lives in device memory
    Agent::GraphKey k;
    k.n = n_updates; k.batch = batch; k.literal = literal_q1; k.gamma = gamma; k.polyak = polyak; k.lr_a = lr_actor; k.lr_c = lr_critic; k.seed = seed;
    const Agent::GraphKey& o = a->gkey;
    const bool same = a->graph && o.n == k.n && o.batch == k.batch && o.literal == k.literal && o.gamma == k.gamma && o.polyak == k.polyak &&
                      o.lr_a == k.lr_a && o.lr_c == k.lr_c && o.seed == k.seed;
    if (!same) {
        drop_graph(a);
        const int64_t l0 = c->launches;
        PDEB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < n_updates && !rc; ++i) rc = one_update();
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        c->launches = l0;
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(c, PDEB200_ECUDA, std::string("train_updates: graph capture failed: ") + cudaGetErrorString(e));
        const cudaError_t e2 = cudaGraphInstantiate(&a->graph, g, 0);
        cudaGraphDestroy(g);
        if (e2 != cudaSuccess) { a->graph = nullptr; return fail(c, PDEB200_ECUDA, std::string("train_updates: cudaGraphInstantiate: ") + cudaGetErrorString(e2)); }
        a->gkey = k;
    }
    PDEB_CUDA(c, cudaGraphLaunch(a->graph, c->stream));
    c->launches += (int64_t)(fast ? (P.f_cluster > 0 ? 1 : 2) : 3) * n_updates;
    return PDEB200_OK;
}

}  // extern "C"
