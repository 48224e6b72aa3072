// Host-side context shared by the translation units of libpdeb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pdeb200.h"
#include "common.cuh"

namespace pdeb200 {

struct HostNet {
    int n_layers = 0;
    int sizes[kMaxLayers + 1] = {0};
    int acts[kMaxLayers] = {0};
    int offs[kMaxLayers] = {0};
    int n_params = 0;
    float* d_params = nullptr;      // weights
    float* d_m = nullptr;           // ADAM first moment
    float* d_v = nullptr;           // ADAM second moment
    double* d_betap = nullptr;      // running beta powers (Flux ADAM state `βp`), 2 doubles ON THE DEVICE: advanced by the
                                    // optimiser kernels themselves so that a captured CUDA graph needs no host-side state
    NetDev dev() const {
        NetDev n;
        n.params = d_params; n.n_layers = n_layers;
        for (int i = 0; i <= kMaxLayers; ++i) n.sizes[i] = sizes[i];
        for (int i = 0; i < kMaxLayers; ++i) { n.acts[i] = acts[i]; n.offs[i] = offs[i]; }
        return n;
    }
};

struct EllHost {
    int* d_idx = nullptr;
    void* d_w = nullptr;
    int nnz_max = 0, n_rows = 0;
};

}  // namespace pdeb200

struct pdeb200_ctx {
    pdeb200_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    mutable std::string err;
    size_t esz = 8;                 // sizeof(T)
    int npts = 0;                   // grid points per field
    int y_elems = 0;                // scalars of T per environment in env.y
    int p_elems = 0;                // scalars of T per environment in env.p
    int fields = 1;
    int obs_rows = 0, n_cols = 0, a_rows = 1, n_rew = 0, wrows = 1;
    int64_t launches = 0;
    bool bases_set = false, y0_set = false;

    // environment arrays
    void *y = nullptr, *y0 = nullptr, *p = nullptr, *state = nullptr, *action = nullptr, *action_in = nullptr,
         *delta_action = nullptr, *reward = nullptr, *sensors = nullptr;
    uint8_t* done = nullptr; double* time = nullptr; int* steps = nullptr;
    void* result_block = nullptr;   // [reward | done | state] in one allocation (reward / done / state point into it)
    size_t res_reward_off = 0, res_state_off = 0, res_done_off = 0, res_bytes = 0;
    size_t res_copy_bytes = 0;      // what result_packed receives: res_bytes, or the [reward | done] prefix (pdeb200_result_select)
    int* d_nsub = nullptr;           // adaptive mode: {accepted, rejected} substeps per environment
    void* d_hlast = nullptr;         // adaptive mode: last accepted step size per environment (warm start of the controller)
    uint8_t* d_mask = nullptr; int* d_list = nullptr; int* d_counts = nullptr; double* d_rsum = nullptr; void* d_noise = nullptr; void* vmax = nullptr;
    // pdeb200_noise_prefetch: the NEXT policy call's host noise, uploaded on its own stream while the current step runs
    void* d_noise_q[2] = {nullptr, nullptr}; cudaStream_t copy_stream = nullptr; cudaEvent_t noise_ready[2] = {nullptr, nullptr};
    unsigned long long noise_put = 0, noise_got = 0;     // prefetches issued / consumed (at most two outstanding)

    // bases
    pdeb200::EllHost sens, actT;
    int* d_a2s = nullptr; void* d_sens_sum = nullptr;

    // KS spectral constants
    int N1 = 0, N2 = 0;
    void *tw12 = nullptr, *tw21 = nullptr, *c1 = nullptr, *cN = nullptr, *ainvh = nullptr, *hm = nullptr;
    // KS sensor gather: layout permutation of the natural-order state in shared memory and the sensor table's
    // indices under it (ks.cu::sensor_layout); rebuilt after pdeb200_set_bases
    int *ks_perm = nullptr, *ks_sens_idx = nullptr; bool ks_layout_dirty = true, ks_layout_on = false;
    // KS sensors as ONE spectral multiply + inverse transform (shift-invariant, equally spaced sensor bases; ks.cu::ks_bases_changed)
    void* ks_sens_hat = nullptr; int ks_sens_sp = 0;

    // problem-specific opaque state (KSeg / NS translation units)
    void* prob = nullptr;
    void* prob_p_phys = nullptr;    // NS: physical-space actuation sum before its FFT (owned by ns.cu)

    // networks + replay + ddpg
    pdeb200::HostNet nets[4];
    float* d_grads = nullptr; int n_grads = 0; float* d_losses = nullptr;
    void* agent = nullptr;
    void* comm = nullptr;           // pdeb200::Comm (comm.cu) after pdeb200_comm_init

    // timing
    const char* core_kernel = "";  // name of the core kernel the most recent step launched (pdeb200_last_core_kernel)
    bool timing = false; cudaEvent_t ev0 = nullptr, ev1 = nullptr, evc0 = nullptr, evc1 = nullptr; bool timed = false;
};

namespace pdeb200 {

extern thread_local std::string g_last_error;

inline int32_t fail(const pdeb200_ctx* c, int32_t code, const std::string& msg) {
    if (c) c->err = msg;
    g_last_error = msg;
    return code;
}

#define PDEB_CUDA(ctx, expr)                                                                         \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return pdeb200::fail(ctx, PDEB200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is ONE value per (kernel, device) for the whole process.  Contexts driven
// from different host threads (bench.py's e2e shards, one-thread-per-GPU hosts) share it, so it is only ever RAISED, under
// a lock: a per-thread cache (round 1) let a thread with a smaller batch lower it under another thread's larger launch.
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, size_t bytes, int device) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> have;
    if (bytes <= 48 * 1024) return cudaSuccess;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = have[{(const void*)kernel, device}];
    if (bytes <= cur) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

template <typename T> struct DT;
template <> struct DT<float>  { static constexpr int id = PDEB200_F32; };
template <> struct DT<double> { static constexpr int id = PDEB200_F64; };

template <typename T>
inline ObsRewardParams<T> make_obs_params(const pdeb200_ctx* c) {
    ObsRewardParams<T> P;
    const pdeb200_config& g = c->cfg;
    P.n_sensors = g.n_sensors; P.n_act = g.n_actuators; P.fields = c->fields;
    P.window = g.window_size; P.temporal = g.temporal_steps; P.memory = g.memory_size; P.a_rows = c->a_rows;
    P.obs_rows = c->obs_rows; P.mono = g.mono;
    P.spa = (g.problem == PDEB200_NS2D || g.problem == PDEB200_KSEG2D) ? g.sensors_per_axis : 0;
    P.check_max = g.check_max_value;
    P.obs_scale = (T)g.obs_scale; P.r_gain = (T)g.reward_gain; P.r_pow = (T)g.reward_pow; P.r_div = (T)g.reward_div;
    P.r_offset = (T)g.reward_offset; P.a_pun = (T)g.action_punish; P.da_pun = (T)g.delta_action_punish;
    P.max_value = (T)g.max_value; P.dt = g.dt; P.te = g.te;
    P.a2s = c->d_a2s; P.sens_sum = (const T*)c->d_sens_sum;
    return P;
}

// problem back-ends (one translation unit each).  *_core advances env.y by one env step from (y, p),
// and writes the raw sensor dots and max|y|; *_sensors computes sensor dots from env.y (reset path).
int32_t ks_setup(pdeb200_ctx* c);
int32_t ks_core(pdeb200_ctx* c);
int32_t ks_cost(const pdeb200_ctx* c, double* bytes, double* flops);
int32_t ks_bases_changed(pdeb200_ctx* c, const double* sensor_basis);   // dense [n_sensors][nx] float64, as given to set_bases
void ks_free(pdeb200_ctx* c);

int32_t kseg_setup(pdeb200_ctx* c);
int32_t kseg_core(pdeb200_ctx* c);
int32_t kseg_cost(const pdeb200_ctx* c, double* bytes, double* flops);
void kseg_free(pdeb200_ctx* c);

int32_t kseg2d_setup(pdeb200_ctx* c);
int32_t kseg2d_core(pdeb200_ctx* c);
int32_t kseg2d_cost(const pdeb200_ctx* c, double* bytes, double* flops);

int32_t ns_setup(pdeb200_ctx* c);
int32_t ns_core(pdeb200_ctx* c);
int32_t ns_sensors(pdeb200_ctx* c, const uint8_t* d_mask);
int32_t ns_cost(const pdeb200_ctx* c, double* bytes, double* flops);
void ns_free(pdeb200_ctx* c);

// layer-wise network operators (nn_ops.cu): tensor cores for dense layers, CUDA cores for thin ones
int32_t dense_layer(pdeb200_ctx* c, int M, int K, int N, const float* X, long long ldx, const float* W, int act, float* Y,
                    long long ldy, int path, int* used_tc);
int32_t dense_dgrad(pdeb200_ctx* c, int M, int K, int N, const float* dY, long long lddy, const float* W, const float* mask,
                    long long ldm, int mask_act, float* dX, long long lddx, int path);
int32_t dense_wgrad(pdeb200_ctx* c, int M, int K, int N, const float* dY, long long lddy, const float* X, long long ldx,
                    float* gW, float* gb, int path);

void agent_free(pdeb200_ctx* c);
double* agent_stats(pdeb200_ctx* c);   // 8 doubles, or nullptr before the first batch
void agent_invalidate_graph(pdeb200_ctx* c);   // networks / rings / communicator changed: drop the captured update graph

// multi-GPU (comm.cu)
struct CommDev;
CommDev comm_dev(const pdeb200_ctx* c);                 // nranks = 1 unless the peer-memory transport is up
int comm_nranks(const pdeb200_ctx* c);
int comm_transport(const pdeb200_ctx* c);
int comm_cap(const pdeb200_ctx* c);
int32_t comm_allreduce_f32(pdeb200_ctx* c, float* dev, size_t n);      // NCCL on the context's stream
int32_t comm_allreduce_f64(pdeb200_ctx* c, double* dev, size_t n);
int32_t comm_check(pdeb200_ctx* c);                    // PDEB200_ECOMM if a peer wait timed out
void comm_free(pdeb200_ctx* c);

}  // namespace pdeb200
