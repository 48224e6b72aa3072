// Keller-Segel 1-D chemotaxis back-end: fixed-step RK4 on the finite-difference rhs.
//
// Restates /root/reference/scripts/Keller-Segel/setup/KellerSegelSetup.jl
//   f       :213-232  second-order central differences with the zero-flux edge copies
//                     U[1,1]=U[1,2], U[end,3]=U[end,2] (quirk Q5)
//   do_step :234-239  the reference drives f with OrdinaryDiffEq's adaptive RK4() (rtol=atol=1e-8,
//                     third-party step controller); here the same classical RK4 tableau runs
//                     `oversampling` fixed substeps per env step (default 40: error vs the exact ODE
//                     solution ~5e-9, below the reference's own tolerance; tests/test_kseg_*.py).
// One CTA per environment, one thread per grid point; (u, v) live in registers, the four RK4 stage
// states go through a double-buffered shared-memory line for the neighbour reads; sensor dots and
// max|y| are produced from the on-chip state like in the KS core kernel.
#include <algorithm>
#include <cmath>

#include "ctx.hpp"

namespace pdeb200 {
namespace {

template <typename T>
struct KsegArgs {
    int nx, S, n_sensors;
    T h, c1, c2;                 // substep, 0.5/dx, 1/dx^2
    EllTable<T> sens;
    T* y;                        // [B][nx][2] (Julia (2,nx) column-major)
    const T* p;                  // [B][nx]
    T* sensors_out;              // [B][2][n_sensors]
    T* vmax_out;                 // [B]
};

template <typename T>
__device__ __forceinline__ void rhs(const typename V2<T>::type* line, int i, int nx, T p, T c1, T c2, T& du, T& dv) {
    using C = typename V2<T>::type;
    const C c = line[i];
    const C l = line[i == 0 ? 0 : i - 1];          // edge copy: left neighbour of the first cell is itself
    const C r = line[i == nx - 1 ? i : i + 1];     // right neighbour of the last cell is itself
    const T u1 = (-c1) * l.x + T(0) * c.x + c1 * r.x;
    const T u2 = c2 * l.x + (T(-2) * c2) * c.x + c2 * r.x;
    const T v1 = (-c1) * l.y + T(0) * c.y + c1 * r.y;
    const T v2 = c2 * l.y + (T(-2) * c2) * c.y + c2 * r.y;
    dv = v2 - c.y + c.x + p;
    du = u2 + c.x - T(5.6) * u1 * v1 - T(5.6) * c.x * v2 - c.x * c.x;
}

template <typename T>
__global__ void __launch_bounds__(1024) kseg_step_kernel(const __grid_constant__ KsegArgs<T> A) {
    using C = typename V2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* line0 = reinterpret_cast<C*>(smem_raw);
    C* line1 = line0 + A.nx;
    __shared__ T s_red[32];
    const int env = blockIdx.x, i = threadIdx.x, nx = A.nx;
    const bool on = i < nx;
    C* yg = reinterpret_cast<C*>(A.y) + (size_t)env * nx;
    C y = on ? yg[i] : V2<T>::make(T(0), T(0));
    const T p = on ? A.p[(size_t)env * nx + i] : T(0);
    const T h = A.h, h2 = T(0.5) * A.h, h6 = A.h / T(6);
    for (int s = 0; s < A.S; ++s) {
        T k1u, k1v, k2u, k2v, k3u, k3v, k4u, k4v;
        if (on) line0[i] = y;
        __syncthreads();
        if (on) { rhs<T>(line0, i, nx, p, A.c1, A.c2, k1u, k1v); line1[i] = V2<T>::make(y.x + h2 * k1u, y.y + h2 * k1v); }
        __syncthreads();
        if (on) { rhs<T>(line1, i, nx, p, A.c1, A.c2, k2u, k2v); line0[i] = V2<T>::make(y.x + h2 * k2u, y.y + h2 * k2v); }
        __syncthreads();
        if (on) { rhs<T>(line0, i, nx, p, A.c1, A.c2, k3u, k3v); line1[i] = V2<T>::make(y.x + h * k3u, y.y + h * k3v); }
        __syncthreads();
        if (on) {
            rhs<T>(line1, i, nx, p, A.c1, A.c2, k4u, k4v);
            y.x = y.x + h6 * (k1u + T(2) * (k2u + k3u) + k4u);
            y.y = y.y + h6 * (k1v + T(2) * (k2v + k3v) + k4v);
        }
    }
    if (on) { yg[i] = y; line0[i] = y; }
    // max |y| over both fields
    T m = on ? fmax(fabs(y.x), fabs(y.y)) : T(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((i & 31) == 0) s_red[i >> 5] = m;
    __syncthreads();
    if (i == 0) {
        T mm = T(0);
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) mm = fmax(mm, s_red[w]);
        A.vmax_out[env] = mm;
    }
    const int ns = A.n_sensors;
    for (int q = i; q < 2 * ns; q += blockDim.x) {
        const int f = q / ns, k = q % ns;
        T acc = T(0);
        for (int j = 0; j < A.sens.nnz_max; ++j) {
            const C v = line0[A.sens.idx[j * ns + k]];
            acc += (f ? v.y : v.x) * A.sens.w[j * ns + k];
        }
        A.sensors_out[(size_t)env * 2 * ns + q] = acc;
    }
}

template <typename T>
int32_t launch(pdeb200_ctx* c) {
    const pdeb200_config& g = c->cfg;
    KsegArgs<T> A;
    const double dx = g.Lx / g.nx;
    A.nx = g.nx; A.S = g.oversampling; A.n_sensors = g.n_sensors;
    A.h = (T)(g.dt / g.oversampling); A.c1 = (T)(0.5 / dx); A.c2 = (T)(1.0 / (dx * dx));
    A.sens = EllTable<T>{c->sens.d_idx, (const T*)c->sens.d_w, c->sens.nnz_max, c->sens.n_rows};
    A.y = (T*)c->y; A.p = (const T*)c->p; A.sensors_out = (T*)c->sensors; A.vmax_out = (T*)c->vmax;
    const int tpb = ((g.nx + 31) / 32) * 32;
    const size_t smem = (size_t)2 * g.nx * 2 * sizeof(T);
    kseg_step_kernel<T><<<g.n_envs, tpb, smem, c->stream>>>(A);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

}  // namespace

int32_t kseg_setup(pdeb200_ctx* c) {
    if (c->cfg.nx < 3 || c->cfg.nx > 1024) return fail(c, PDEB200_EUNSUPPORTED, "KSeg: 3 <= nx <= 1024 (one thread per grid point)");
    if (c->cfg.oversampling < 1) return fail(c, PDEB200_EINVAL, "KSeg: oversampling (RK4 substeps) must be >= 1");
    return PDEB200_OK;
}

int32_t kseg_core(pdeb200_ctx* c) { return c->cfg.dtype == PDEB200_F64 ? launch<double>(c) : launch<float>(c); }

// bytes: (u,v) in + out, p in, action in, obs + reward out; flops: 4 stages x ~36 flops x 2 fields per point
int32_t kseg_cost(const pdeb200_ctx* c, double* bytes, double* flops) {
    const pdeb200_config& g = c->cfg;
    const double w = (double)c->esz;
    if (bytes) *bytes = 2 * 2 * g.nx * w + g.n_actuators * (w * c->a_rows + w * c->obs_rows + w) + 1;
    if (flops) *flops = (double)g.oversampling * g.nx * (4 * 36 + 16) + 2.0 * 2 * c->sens.nnz_max * g.n_sensors;
    return PDEB200_OK;
}

void kseg_free(pdeb200_ctx*) {}

}  // namespace pdeb200
