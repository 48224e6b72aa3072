// Shared device-side helpers for the pdeb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdeb200 {

template <typename T> struct V2;
template <> struct V2<float>  { using type = float2;  static __host__ __device__ float2  make(float a, float b)   { return make_float2(a, b); } };
template <> struct V2<double> { using type = double2; static __host__ __device__ double2 make(double a, double b) { return make_double2(a, b); } };

constexpr int kMaxLayers = 4;       // Dense layers per network (reference uses 2 or 3)
constexpr int kFusedActorMaxWidth = 64;

// Sparse (ELL) gather table: row i = sum_j w[j][i] * x[idx[j][i]], j < nnz_max.
// j-major so that threads owning consecutive rows read consecutive addresses.
// Padding entries have w = 0 and idx = a valid index.
template <typename T>
struct EllTable {
    const int* idx;
    const T* w;
    int nnz_max;
    int n_rows;
};

// A Flux Chain of Dense layers, float32, device resident.
// params: per layer W (column-major (out,in), i.e. W[o + out*i]) then b (out).
struct NetDev {
    const float* params;
    int n_layers;
    int sizes[kMaxLayers + 1];
    int acts[kMaxLayers];
    int offs[kMaxLayers];           // offset of layer l's W in params
};

__device__ __forceinline__ float act_apply(int kind, float v) {
    if (kind == 1) return v > 0.f ? v : 0.f;         // relu
    if (kind == 2) return tanhf(v);                  // tanh
    return v;
}

// NIN-NH-1 actor evaluated in registers.  w: the flat parameter vector W0 (NH x NIN, column-major) | b0 | W1 (1 x NH) |
// b1; same accumulation order as the runtime-shaped loops (ascending input index from 0, bias added last).
// Restates Flux Dense: y = act.(W*x .+ b)   (src/PDEagent.jl:18-30).
template <int NIN, int NH>
__device__ __forceinline__ float actor_two_layer(const float* w, const float* x, int act0, int act1) {
    float h[NH];
#pragma unroll
    for (int u = 0; u < NH; ++u) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NIN; ++i) acc = fmaf(w[u + NH * i], x[i], acc);
        h[u] = acc + w[NIN * NH + u];
    }
#pragma unroll
    for (int u = 0; u < NH; ++u) h[u] = act_apply(act0, h[u]);
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < NH; ++u) acc = fmaf(w[NIN * NH + NH + u], h[u], acc);
    return act_apply(act1, acc + w[NIN * NH + NH + NH]);
}

template <typename T> __device__ __forceinline__ T pow_t(T a, T b);
template <> __device__ __forceinline__ float  pow_t<float>(float a, float b)    { return powf(a, b); }
// a >= 0, b > 0 (reward terms |x|^1.3, |x|^1.1): exp(b log a) with CUDA's double log / exp (<= 1 ulp each) is within
// ~|b ln a| * 2^-53 <= 3e-15 relative of the correctly rounded power -- far inside the 1e-12 parity bound -- and avoids
// the out-of-line call (stack frame, ~3x the instructions) of the general pow().
template <> __device__ __forceinline__ double pow_t<double>(double a, double b) { return a > 0.0 ? exp(b * log(a)) : 0.0; }

template <typename T> __device__ __forceinline__ T clamp_t(T v, T lim) { return v < -lim ? -lim : (v > lim ? lim : v); }

// Parameters shared by every environment kernel (KS / KSeg / NS): observation and
// reward assembly restating featurize / reward_function of the setup files.
template <typename T>
struct ObsRewardParams {
    int n_sensors, n_act, fields;         // fields: 1 (KS, NS) or 2 (KSeg u,v)
    int window, temporal, memory, a_rows; // a_rows = 1 + memory
    int obs_rows;                         // ns
    int mono;
    int spa;                              // NS: sensors per axis (2-D window), 0 for 1-D
    int check_max;
    T obs_scale, r_gain, r_pow, r_div, r_offset, a_pun, da_pun, max_value;
    double dt, te;
    const int* a2s;                       // [n_act], 0-based
    const T* sens_sum;                    // [n_sensors] sum of each sensor basis (for r_offset)
};

}  // namespace pdeb200
