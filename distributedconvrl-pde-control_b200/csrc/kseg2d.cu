// Keller-Segel 2-D chemotaxis back-end (BASELINE config 3: 128 x 128, two fields, distributed actuators).
//
// The reference ships only the 1-D model (scripts/Keller-Segel/setup/KellerSegelSetup.jl:213-239); this is
// its direct 2-D generalisation: the same reaction / chemotaxis terms and constants,
//   v_t = lap v - v + u + p
//   u_t = lap u + u - 5.6 grad u . grad v - 5.6 u lap v - u^2
// second-order central differences per axis with the SAME zero-flux edge rule per axis (the neighbour
// outside the domain is the cell itself, KellerSegelSetup.jl:220-223, quirk Q5), classical RK4 with
// `oversampling` fixed substeps.  Data that does not depend on y reproduces the 1-D model bit for bit
// operation order (the y-differences are exact zeros), which is how it is pinned to the reference's
// golden rows (tests/test_kseg2d.py).
//
// Mapping: one 4-CTA thread-block cluster per environment.  Each CTA owns ny/4 rows; a thread owns a
// vertical strip of <= 8 points and keeps (y_n, RK accumulator, p) for them in registers.  Only the
// current stage state lives in shared memory (double buffered, interleaved (u,v)); after every stage each
// CTA pushes its two boundary rows into the neighbours' halo rows through distributed shared memory and
// the cluster synchronises once.  HBM sees the state once in and once out per env step.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

#include "ctx.hpp"
#include "glue.cuh"

namespace cg = cooperative_groups;

namespace pdeb200 {
namespace {

constexpr int kCS = 4;          // CTAs per cluster (= per environment)
constexpr int kTX = 128;        // threads along x (nx <= 128)
constexpr int kStrips = 4;      // strips of rows per CTA
constexpr int kR = 8;           // rows per strip (max)

template <typename T>
struct Kseg2dArgs {
    int nx, ny, rows, S;         // rows = ny / kCS
    T h, c1x, c2x, c1y, c2y;
    T* y;                        // [B][ny][nx][2]
    const T* p;                  // [B][ny][nx]
    T* vmax_out;                 // [B], zeroed before the launch
};

template <typename T> struct BitsOf;
template <> struct BitsOf<float>  { using type = int; };
template <> struct BitsOf<double> { using type = long long; };

template <typename T>
__device__ __forceinline__ void atomic_max_nonneg(T* addr, T v) {
    using I = typename BitsOf<T>::type;
    I bits;
    memcpy(&bits, &v, sizeof(T));
    atomicMax(reinterpret_cast<I*>(addr), bits);        // order of non-negative IEEE values == order of their bits
}

// FULL: the BASELINE shape (nx = 128, ny = 128 -> 32 rows per CTA, 8 per strip).  Every point is live, so the
// per-point predicates disappear, the four RK stages are unrolled (no stage branches inside the row loop) and all
// shared-memory addresses are "row pointer + immediate"; ncu on the runtime-shaped version showed 40 % integer /
// 28 % FP64 instructions (profiles/r1_kseg2d.md).
template <typename T, bool FULL>
__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kTX * kStrips, 1)
kseg2d_step_kernel(const __grid_constant__ Kseg2dArgs<T> A) {
    using C = typename V2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int env = blockIdx.x / kCS;
    const int nx = FULL ? kTX : A.nx, rows = FULL ? kStrips * kR : A.rows;
    const int x = threadIdx.x % kTX, strip = threadIdx.x / kTX;
    const int R = FULL ? kR : (rows + kStrips - 1) / kStrips;
    const int r0 = strip * R;
    const int slab = (rows + 2) * nx;                     // one buffer: halo row, `rows` rows, halo row
    C* const buf0 = reinterpret_cast<C*>(smem_raw);
    C* const up0 = rank > 0 ? cluster.map_shared_rank(buf0, rank - 1) : nullptr;           // CTA owning the rows above
    C* const dn0 = rank < kCS - 1 ? cluster.map_shared_rank(buf0, rank + 1) : nullptr;     // ... below
    const bool xon = FULL || x < nx;
    const int grow0 = rank * rows;                        // global index of local row 0

    C y0[kR], acc[kR];
    T pp[kR];
    bool on[kR];
    const C* yg = reinterpret_cast<const C*>(A.y) + (size_t)env * A.ny * nx;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
        const int lr = r0 + r;
        on[r] = FULL || (xon && r < R && lr < rows);
        y0[r] = on[r] ? yg[(size_t)(grow0 + lr) * nx + x] : V2<T>::make(T(0), T(0));
        pp[r] = on[r] ? A.p[((size_t)env * A.ny + grow0 + lr) * nx + x] : T(0);
        acc[r] = V2<T>::make(T(0), T(0));
    }
    // which of this thread's rows touch a neighbour CTA / the domain edge (row index within the strip)
    const int r_first = 0, r_last = FULL ? kR - 1 : min(R, rows - r0) - 1;
    const bool push_up = up0 != nullptr && r0 == 0;                       // my first row = their lower halo
    const bool push_dn = dn0 != nullptr && r0 + r_last == rows - 1;       // my last row = their upper halo
    const bool top_edge = grow0 + r0 == 0;                                // edge copy along y (top)
    const bool bot_edge = grow0 + r0 + r_last == A.ny - 1;                // edge copy along y (bottom)
    // publish a stage state: own rows into buf[b], boundary rows into the neighbours' halo rows
    auto publish = [&](int b, const C* st) {
        C* row = buf0 + b * slab + (r0 + 1) * nx + x;
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            if (!on[r]) continue;
            row[r * nx] = st[r];
            if (r == r_first && push_up) up0[b * slab + (rows + 1) * nx + x] = st[r];
            if (r == r_last && push_dn) dn0[b * slab + x] = st[r];
        }
    };
    cluster.sync();              // every CTA of the cluster is running before anyone writes into its shared memory
    publish(0, y0);
    cluster.sync();
    const T h = A.h, h2 = T(0.5) * A.h, h6 = A.h / T(6);
    const int xl = x == 0 ? 0 : x - 1, xr = x >= nx - 1 ? (xon ? x : 0) : x + 1;
    const T c1x = A.c1x, c2x = A.c2x, c1y = A.c1y, c2y = A.c2y, m2x = T(-2) * A.c2x, m2y = T(-2) * A.c2y;
    int cur = 0;
    // one RK stage over the strip; the centre row slides down the strip so every stage state is read three times
    // per point (left, right, below) instead of five
    auto run_stage = [&](const int stage) {
        const C* Br = buf0 + cur * slab + (r0 + 1) * nx;
        // the new stage state goes straight into the other buffer (nobody reads it before the cluster barrier below)
        const int b = cur ^ 1;
        C* row = buf0 + b * slab + (r0 + 1) * nx + x;
        C c = on[0] ? Br[x] : V2<T>::make(T(0), T(0));
        C u = (top_edge || !on[0]) ? c : Br[x - nx];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            if (!on[r]) continue;
            const C l = Br[r * nx + xl];
            const C rt = Br[r * nx + xr];
            const C d = (r == r_last && bot_edge) ? c : Br[(r + 1) * nx + x];
            // same operation order as the 1-D f (KellerSegelSetup.jl:225-229) per axis; the reference's literal
            // "+ 0 * centre" term of the first differences is dropped: it adds an exact zero for finite data, and
            // the compiler may not remove it itself (NaN / inf semantics)
            const T u1x = (-c1x) * l.x + c1x * rt.x;
            const T u2x = c2x * l.x + m2x * c.x + c2x * rt.x;
            const T v1x = (-c1x) * l.y + c1x * rt.y;
            const T v2x = c2x * l.y + m2x * c.y + c2x * rt.y;
            const T u1y = (-c1y) * u.x + c1y * d.x;
            const T u2y = c2y * u.x + m2y * c.x + c2y * d.x;
            const T v1y = (-c1y) * u.y + c1y * d.y;
            const T v2y = c2y * u.y + m2y * c.y + c2y * d.y;
            const T lapu = u2x + u2y, lapv = v2x + v2y;
            const T kv = lapv - c.y + c.x + pp[r];
            const T ku = lapu + c.x - (T(5.6) * u1x * v1x + T(5.6) * u1y * v1y) - T(5.6) * c.x * lapv - c.x * c.x;
            C st;
            if (stage == 1) { acc[r] = V2<T>::make(ku, kv); st = V2<T>::make(y0[r].x + h2 * ku, y0[r].y + h2 * kv); }
            else if (stage == 2) { acc[r].x += T(2) * ku; acc[r].y += T(2) * kv; st = V2<T>::make(y0[r].x + h2 * ku, y0[r].y + h2 * kv); }
            else if (stage == 3) { acc[r].x += T(2) * ku; acc[r].y += T(2) * kv; st = V2<T>::make(y0[r].x + h * ku, y0[r].y + h * kv); }
            else { y0[r].x += h6 * (acc[r].x + ku); y0[r].y += h6 * (acc[r].y + kv); st = y0[r]; }
            row[r * nx] = st;
            if (r == r_first && push_up) up0[b * slab + (rows + 1) * nx + x] = st;
            if (r == r_last && push_dn) dn0[b * slab + x] = st;
            u = c; c = d;
        }
        cluster.sync();
        cur ^= 1;
    };
    for (int s = 0; s < A.S; ++s) {
        if (FULL) { run_stage(1); run_stage(2); run_stage(3); run_stage(4); }
        else {
#pragma unroll 1
            for (int stage = 1; stage <= 4; ++stage) run_stage(stage);
        }
    }
    C* yo = reinterpret_cast<C*>(A.y) + (size_t)env * A.ny * nx;
    T m = T(0);
#pragma unroll
    for (int r = 0; r < kR; ++r)
        if (on[r]) { yo[(size_t)(grow0 + r0 + r) * nx + x] = y0[r]; m = fmax(m, fmax(fabs(y0[r].x), fabs(y0[r].y))); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg<T>(A.vmax_out + env, m);
}

template <typename T>
int32_t launch(pdeb200_ctx* c) {
    const pdeb200_config& g = c->cfg;
    Kseg2dArgs<T> A;
    const double dx = g.Lx / g.nx, dy = g.Ly / g.ny;
    A.nx = g.nx; A.ny = g.ny; A.rows = g.ny / kCS; A.S = g.oversampling;
    A.h = (T)(g.dt / g.oversampling);
    A.c1x = (T)(0.5 / dx); A.c2x = (T)(1.0 / (dx * dx)); A.c1y = (T)(0.5 / dy); A.c2y = (T)(1.0 / (dy * dy));
    A.y = (T*)c->y; A.p = (const T*)c->p; A.vmax_out = (T*)c->vmax;
    const size_t smem = (size_t)2 * (A.rows + 2) * g.nx * 2 * sizeof(T);
    const bool full = g.nx == kTX && g.ny == kCS * kStrips * kR;
    auto kern = full ? kseg2d_step_kernel<T, true> : kseg2d_step_kernel<T, false>;
    PDEB_CUDA(c, ensure_dyn_smem(kern, smem, c->device));
    PDEB_CUDA(c, cudaMemsetAsync(c->vmax, 0, (size_t)g.n_envs * sizeof(T), c->stream));
    kern<<<g.n_envs * kCS, kTX * kStrips, smem, c->stream>>>(A);
    PDEB_CUDA(c, cudaGetLastError());
    // sensor dots from the new state (5 x 5 boxes: a few dozen taps per sensor)
    sensors_phys_kernel<T><<<g.n_envs, 128, 0, c->stream>>>(
        2, c->npts, g.n_sensors, EllTable<T>{c->sens.d_idx, (const T*)c->sens.d_w, c->sens.nnz_max, c->sens.n_rows}, nullptr,
        (const T*)c->y, 1, (T*)c->sensors, nullptr);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return PDEB200_OK;
}

}  // namespace

int32_t kseg2d_setup(pdeb200_ctx* c) {
    const pdeb200_config& g = c->cfg;
    if (g.nx < 3 || g.nx > kTX) return fail(c, PDEB200_EUNSUPPORTED, "KSeg2D: 3 <= nx <= 128");
    if (g.ny < kCS || g.ny % kCS || g.ny / kCS > kStrips * kR)
        return fail(c, PDEB200_EUNSUPPORTED, "KSeg2D: ny must be a multiple of 4, 4 <= ny <= 128");
    if (g.oversampling < 1) return fail(c, PDEB200_EINVAL, "KSeg2D: oversampling (RK4 substeps) must be >= 1");
    if (g.sensors_per_axis < 1 || g.sensors_per_axis * g.sensors_per_axis != g.n_sensors)
        return fail(c, PDEB200_EINVAL, "KSeg2D: n_sensors must equal sensors_per_axis^2 (2-D observation windows)");
    return PDEB200_OK;
}

int32_t kseg2d_core(pdeb200_ctx* c) { return c->cfg.dtype == PDEB200_F64 ? launch<double>(c) : launch<float>(c); }

// bytes: (u,v) in + out, p in, action in, obs + reward out; flops: 4 stages x ~70 flops per point per substep
int32_t kseg2d_cost(const pdeb200_ctx* c, double* bytes, double* flops) {
    const pdeb200_config& g = c->cfg;
    const double w = (double)c->esz, np = (double)g.nx * g.ny;
    if (bytes) *bytes = 2 * 2 * np * w + np * w + g.n_actuators * (w * c->a_rows + w * c->obs_rows + w) + 1;
    if (flops) *flops = (double)g.oversampling * np * (4 * 70 + 16) + 2.0 * 2 * c->sens.nnz_max * g.n_sensors;
    return PDEB200_OK;
}

}  // namespace pdeb200
