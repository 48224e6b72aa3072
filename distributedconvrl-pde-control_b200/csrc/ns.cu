// Navier-Stokes back-end: 2-D incompressible flow in vorticity form, pseudo-spectral, fixed-step RK4.
//
// Restates (batched over environments; y = omega_hat, the spectral vorticity, like the reference):
//   rk4 / rhs / advection / pad / chop   src/fluid_rk4.jl:122-229
//   do_step (RK4 x oversampling)          scripts/Fluid/setup/FluidSetup.jl:163-172
//   prepare_action's fft(p)               FluidSetup.jl:247-261   (the physical sum is actuate_kernel's)
//   real(ifft(env.y)) + sensor dots       FluidSetup.jl:189-197, 205-217
//
// One rhs evaluation = three kernels; the 3/2-rule work arrays never exist as full 2-D complex arrays:
//   A  ns_ypass_inv : per (env, kx >= 0 column): the four spectra  u_hat = i ky psi_hat, v_hat = -i kx psi_hat,
//                     i kx omega_hat, i ky omega_hat  are formed on the fly from omega_hat, zero-padded along ky and
//                     inverse-transformed along y.  Only the kx >= 0 half is needed because every field is REAL in
//                     physical space: real(ifft2(X)) == ifft2(X_h), X_h[k] = (P[k] + conj(P[-k]))/2 with P = pad(X).
//                     X_h is evaluated literally (both k and -k are read), so arbitrary complex input -- including
//                     the non-Hermitian Nyquist row/column that pad()/chop() create -- is treated like the reference.
//   B  ns_xpass     : per pair of y-lines: Hermitian extension along kx, two complex inverse FFTs give (u + i v) and
//                     (omega_x + i omega_y), the product -u omega_x - v omega_y of two lines is packed into ONE
//                     complex forward FFT along x and split again; only kx in [0, N/2] is kept (chop).
//   C  ns_ypass_fwd : per (env, kx column): forward FFT along y, chop along ky, scale 1.5*1.5, and the RK4 stage
//                     update for column kx and (by conjugate symmetry of the FFT of a real field) column -kx:
//                     k = -nu k^2 f + nonlin + p_hat;  stage combinations of rk4().
// Arrays are env-major, Julia column-major inside an environment: y[env][i (kx)][j (ky)] complex.
// A and B exist in two forms: the round-1 kernels keep one line per warp in registers (fft_pass.cuh; PDEB200_NS_LEGACY=1),
// the default ones (ns_ypass_inv4_kernel, ns_xpass4_kernel) transform four lines per warp in place in shared memory
// (fft_batch.cuh).  adaptive = 1 adds per-environment step control on top of the same kernels (ns_adapt_*).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "ctx.hpp"
#include "fft_batch.cuh"
#include "fft_pass.cuh"
#include "glue.cuh"

namespace pdeb200 {

namespace {

template <int P1, int P2> struct NsGeom {
    static constexpr int NP = P1 * P2;
    static constexpr int RMAX = P1 > P2 ? P1 : P2;
    static constexpr int SA = P2 * PassStride<P2, P1>::value;      // inverse pass <P2, P1>
    static constexpr int SB = P1 * PassStride<P1, P2>::value;      // forward pass <P1, P2>
    static constexpr int XB0 = (SA > SB ? SA : SB) > NP ? (SA > SB ? SA : SB) : NP;
    static constexpr int XB = ((XB0 + 7) / 8) * 8 + 1;             // == 1 (mod 8): conflict-free 8-column tiles
};


struct NsProb {
    int N = 0, NP = 0, NH = 0, NHP = 0, chunk = 0, n_sm = 148;
    bool legacy = false;
    int n1 = 0, n2 = 0, p1 = 0, p2 = 0;
    void *tw_inv = nullptr, *tw_fwd = nullptr, *tw_n = nullptr, *tw_b = nullptr;
    void *kx = nullptr, *ky = nullptr;
    void *fst = nullptr, *acc = nullptr, *W = nullptr, *Q = nullptr;
    void *p_phys = nullptr, *om_phys = nullptr, *tmpc = nullptr;
    void *y2 = nullptr, *ad_state = nullptr; int* ad_active = nullptr; int* h_active = nullptr;   // adaptive mode
};

template <typename T>
struct NsArgs {
    using C = typename V2<T>::type;
    int N, NH, NHP, stage;
    const C* tw_inv; const C* tw_fwd;
    const C* tw_b;             // batched kernels: table of fft_pass<BP2, BP1> for THEIR factorisation (tw_b[n*BP2 + k] = W^(n k))
    const T* kx; const T* ky;
    const C* fin;              // stage input (y at stage 1, fst afterwards)
    C* y; C* fst; C* acc;      // y: the step's base state (read by stages 2-4)
    C* yout;                   // stage 4 writes the step's result here (== y for the in-place fixed-step path)
    const T* dt_env;           // adaptive mode: per-environment step size (nullptr: dt)
    const C* phat;
    C* W;                      // [env][4][NP][NHP]
    C* Q;                      // [env][NP][NHP]
    T nu, dt, scale;
};

// padded index -> index in the unpadded array, or -1 if pad() leaves that entry zero (fluid_rk4.jl:205-208)
__device__ __forceinline__ int unpad_idx(int kp, int NP, int N) {
    const int s = kp <= NP / 2 ? kp : kp - NP;
    if (s > N / 2 || s <= -(N / 2)) return -1;
    return s < 0 ? s + N : s;
}

template <typename T> struct NsMinCtas { static constexpr int value = sizeof(T) == 8 ? 1 : 2; };
// Round-1 kernels (one line per warp; PDEB200_NS_LEGACY=1): kernel A ran best as three 4-warp CTAs per SM in fp64.
template <typename T> struct NsColsA { static constexpr int value = 4; };
template <typename T> struct NsColsC { static constexpr int value = 8; };
// Kernel B: a ninth warp fits 227 KB of shared memory in fp64 (24.2 KB per warp + 12 KB of twiddles) but measured
// slower (568 -> 554 env-steps/s: 192 line pairs do not divide by 9 and B is the FP64-pipe-bound kernel of the three)
template <typename T> struct NsColsB { static constexpr int value = 4; };

// ---- A: inverse transform along y of the four padded, Hermitian-symmetrised spectra ---------------------
// Input staging is branch-free and uses all 32 lanes: a CTA-wide table maps each padded row to the unpadded
// rows of its two contributions (k and -k), the column pair (kx, -kx) sits in shared memory, and the two
// psi-based fields reuse the columns after an in-place division by k^2 (one division per entry, like the
// reference's `psihat = omghat ./ kx2ky2`).
template <typename T, int P1, int P2, int NN, int COLS>
__global__ void __launch_bounds__(COLS * 32, (sizeof(T) == 8 ? 1 : 2) * (COLS > 6 ? 1 : (COLS > 4 ? 2 : (COLS > 3 ? 3 : 4))))
ns_ypass_inv_kernel(const __grid_constant__ NsArgs<T> A) {
    using G = NsGeom<P1, P2>;
    using C = typename V2<T>::type;
    constexpr int NP = G::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = NN;                                         // grid size, NH, NHP known at compile time
    constexpr int NH = NN / 2 + 1, NHP = (NH + 3) / 4 * 4;
    C* s_tw = reinterpret_cast<C*>(smem_raw);
    C* s_xb0 = s_tw + NP;
    C* s_col0 = s_xb0 + COLS * G::XB;
    T* s_ky = reinterpret_cast<T*>(s_col0 + COLS * 2 * N);
    short2* s_tab = reinterpret_cast<short2*>(s_ky + N);
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int env = blockIdx.y, a0 = blockIdx.x * COLS, a = a0 + w;
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        s_tw[i] = A.tw_inv[i];
        s_tab[i] = make_short2((short)unpad_idx(i, NP, N), (short)unpad_idx((NP - i) % NP, NP, N));
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_ky[i] = A.ky[i];
    C* xb = s_xb0 + w * G::XB;
    C* colA = s_col0 + (size_t)w * 2 * N;
    C* colB = colA + N;
    const bool live = a < NH;
    const int ib = (N - a) % N;                                   // column of -kx in the unpadded array
    const bool xpartner = unpad_idx((NP - a) % NP, NP, N) >= 0;   // pad() keeps (-kx) ?
    T kxa = T(0), kxb = T(0);
    if (live) {
        const C* src = A.fin + (size_t)env * N * N;
        for (int j = t; j < N; j += 32) { colA[j] = src[(size_t)a * N + j]; colB[j] = src[(size_t)ib * N + j]; }
        kxa = A.kx[a]; kxb = A.kx[ib];
    }
    __syncthreads();
#pragma unroll 1
    for (int fi = 0; fi < 4; ++fi) {
        const int f = (fi + 2) & 3;              // omega-based fields (2, 3) first, then psi-based (0, 1)
        if (live) {
            if (fi == 2) {
                // psi_hat = omega_hat ./ kx2ky2, psi_hat[1,1] = 0   (fluid_rk4.jl:152-153)
                for (int j = t; j < N; j += 32) {
                    const T kyv = s_ky[j];
                    const T ka = kyv * kyv + kxa * kxa, kb = kyv * kyv + kxb * kxb;
                    const C ca = colA[j], cb = colB[j];
                    colA[j] = (j == 0 && a == 0) ? V2<T>::make(T(0), T(0)) : V2<T>::make(ca.x / ka, ca.y / ka);
                    colB[j] = (j == 0 && ib == 0) ? V2<T>::make(T(0), T(0)) : V2<T>::make(cb.x / kb, cb.y / kb);
                }
                __syncwarp();
            }
            // multiplier i*m: u = i ky psi, v = -i kx psi, w_x = i kx omega, w_y = i ky omega
            const bool kytype = (f == 0 || f == 3);
            const T mxa = f == 1 ? -kxa : kxa, mxb = f == 1 ? -kxb : kxb;
#pragma unroll
            for (int kyp0 = 0; kyp0 < NP; kyp0 += 32) {
                const int kyp = kyp0 + t;
                if (NP % 32 != 0 && kyp >= NP) break;
                // 32-row blocks inside N/2 < ky' < NP - N/2 are zero in pad(X) for both k and -k: no loads
                if (kyp0 > N / 2 && kyp0 + 31 < NP - N / 2) { xb[kyp] = V2<T>::make(T(0), T(0)); continue; }
                const short2 jj = s_tab[kyp];
                const int j1 = jj.x, j2 = xpartner ? (int)jj.y : -1;
                const int i1 = j1 < 0 ? 0 : j1, i2 = j2 < 0 ? 0 : j2;
                const C c1 = colA[i1], c2 = colB[i2];
                T m1 = kytype ? s_ky[i1] : mxa, m2 = kytype ? s_ky[i2] : mxb;
                m1 = j1 < 0 ? T(0) : m1; m2 = j2 < 0 ? T(0) : m2;
                // X_h = (X(k) + conj(X(-k))) / 2 with X = i m c
                xb[kyp] = V2<T>::make(T(0.5) * (-m1 * c1.y - m2 * c2.y), T(0.5) * (m1 * c1.x - m2 * c2.x));
            }
            __syncwarp();
            T zr[G::RMAX], zi[G::RMAX];
            if (t < P1) {
#pragma unroll
                for (int r = 0; r < P2; ++r) { const C v = xb[t + P1 * r]; zr[r] = v.x; zi[r] = v.y; }
            }
            __syncwarp();
            fft_pass<T, P2, P1, +1>(zr, zi, xb, s_tw, t);
            if (t < P2) {
#pragma unroll
                for (int r = 0; r < P1; ++r) xb[t + P2 * r] = V2<T>::make(zr[r], zi[r]);
            }
        }
        __syncthreads();
        {
            // 8-column tile store: thread -> (column c, row y), rows advance by 32 per iteration
            const int c = threadIdx.x % COLS, y0 = threadIdx.x / COLS;
            C* dst = A.W + ((size_t)(env * 4 + f) * NP + y0) * NHP + a0 + c;
            const C* srow = s_xb0 + c * G::XB + y0;
            if (a0 + c < NH) {
#pragma unroll 4
                for (int y = y0; y < NP; y += 32) { *dst = *srow; dst += (size_t)32 * NHP; srow += 32; }
            }
        }
        __syncthreads();
    }
}

// ---- mbarrier + bulk-copy (TMA, 1-D) helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}

// ---- B: x transforms, the quadratic term, forward x transform (two y-lines per warp) ---------------------
// The four half-spectrum lines a warp consumes per y-line are contiguous rows of W: they are staged into
// shared memory by 1-D bulk copies (cp.async.bulk, completion on a per-warp mbarrier) two phases ahead of
// their use, so the HBM/L2 latency overlaps the FFT arithmetic of the previous phase.
template <typename T, int P1, int P2, int NN, int COLS>
__global__ void __launch_bounds__(COLS * 32, NsMinCtas<T>::value * (8 / COLS))
ns_xpass_kernel(const __grid_constant__ NsArgs<T> A) {
    using G = NsGeom<P1, P2>;
    using C = typename V2<T>::type;
    constexpr int NP = G::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = NN, NH = NN / 2 + 1, NHP = (NH + 3) / 4 * 4;
    C* s_twi = reinterpret_cast<C*>(smem_raw);
    C* s_twf = s_twi + NP;
    C* s_xb0 = s_twf + NP;
    C* s_uv0 = s_xb0 + COLS * G::XB;
    C* s_in0 = s_uv0 + COLS * NP;                              // [warp][slot 2][line 2][NHP]
    T* s_q0 = reinterpret_cast<T*>(s_in0 + (size_t)COLS * 4 * NHP);
    uint64_t* s_bar0 = reinterpret_cast<uint64_t*>(s_q0 + (size_t)COLS * NP);
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int env = blockIdx.y;
    const int y0 = (blockIdx.x * COLS + w) * 2;
    for (int i = threadIdx.x; i < NP; i += blockDim.x) { s_twi[i] = A.tw_inv[i]; s_twf[i] = A.tw_fwd[i]; }
    uint64_t* bar = s_bar0 + w * 2;
    if (t == 0) {
        mbar_init(bar, 1); mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (y0 >= NP) return;
    C* xb = s_xb0 + w * G::XB;
    C* uv = s_uv0 + (size_t)w * NP;
    C* sin = s_in0 + (size_t)w * 4 * NHP;
    T* q1 = s_q0 + (size_t)w * NP;
    const uint32_t line_bytes = (uint32_t)(NHP * sizeof(C));
    // phase ph = 2*L + h: y-line y0 + L, fields (2h, 2h + 1); slot = ph & 1
    auto issue = [&](int ph) {
        const int L = ph >> 1, h = ph & 1, slot = ph & 1;
        const C* Wa = A.W + ((size_t)(env * 4 + 2 * h) * NP + (y0 + L)) * NHP;
        mbar_expect_tx(bar + slot, 2 * line_bytes);
        bulk_g2s(sin + (size_t)(slot * 2) * NHP, Wa, line_bytes, bar + slot);
        bulk_g2s(sin + (size_t)(slot * 2 + 1) * NHP, Wa + (size_t)NP * NHP, line_bytes, bar + slot);
    };
    if (t == 0) { issue(0); issue(1); }
    T zr[G::RMAX], zi[G::RMAX];
#pragma unroll 1
    for (int ph = 0; ph < 4; ++ph) {
        const int L = ph >> 1, h = ph & 1, slot = ph & 1;
        mbar_wait(bar + slot, (uint32_t)(ph >> 1) & 1);
        const C* Wa = sin + (size_t)(slot * 2) * NHP;
        const C* Wb = Wa + NHP;
        if (t < P1) {
#pragma unroll
            for (int r = 0; r < P2; ++r) {
                const int kxp = t + P1 * r;
                const int s = kxp <= NP / 2 ? kxp : kxp - NP;
                const int a = s < 0 ? -s : s;
                T re = T(0), im = T(0);
                if (a <= N / 2) {
                    C U = Wa[a], V = Wb[a];
                    if (s < 0) { U.y = -U.y; V.y = -V.y; }              // G(y, -kx) = conj(G(y, kx))
                    re = U.x - V.y; im = U.y + V.x;                      // U + i V
                }
                zr[r] = re; zi[r] = im;
            }
        }
        __syncwarp();                                                    // every lane is done with the slot
        if (t == 0 && ph + 2 < 4) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(ph + 2);
        }
        fft_pass<T, P2, P1, +1>(zr, zi, xb, s_twi, t);
        if (t < P2) {
            if (h == 0) {
#pragma unroll
                for (int r = 0; r < P1; ++r) uv[t + P2 * r] = V2<T>::make(zr[r], zi[r]);
            } else {
#pragma unroll
                for (int r = 0; r < P1; ++r) {
                    const int x = t + P2 * r;
                    const C g = uv[x];
                    const T q = -(g.x * zr[r] + g.y * zi[r]) * A.scale;   // -u w_x - v w_y  (fluid_rk4.jl:175)
                    if (L == 0) q1[x] = q;
                    else { zr[r] = q1[x]; zi[r] = q; }
                }
            }
        }
    }
    fft_pass<T, P1, P2, -1>(zr, zi, xb, s_twf, t);
    if (t < P1) {
#pragma unroll
        for (int r = 0; r < P2; ++r) xb[t + P1 * r] = V2<T>::make(zr[r], zi[r]);
    }
    __syncwarp();
    C* Qa = A.Q + ((size_t)env * NP + y0) * NHP;
    C* Qb = Qa + NHP;
    for (int a = t; a < NH; a += 32) {
        const C za = xb[a], zb = xb[(NP - a) % NP];
        Qa[a] = V2<T>::make(T(0.5) * (za.x + zb.x), T(0.5) * (za.y - zb.y));
        Qb[a] = V2<T>::make(T(0.5) * (za.y + zb.y), T(-0.5) * (za.x - zb.x));
    }
}

// 1 / q for q > 0, normal range: hardware seed + two Newton steps (4 DFMA; the IEEE division is ~25 instructions and was
// 10 % of kernel A4's samples).  Within 1-2 ulp of the quotient -- the psi_hat = omega_hat ./ kx2ky2 of fluid_rk4.jl:152
// is then a multiplication; far inside the 1e-12 parity bound.
__device__ __forceinline__ double fast_rcp(double q) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(q));
    double e = fma(-q, r, 1.0);
    r = fma(r, e, r);
    e = fma(-q, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ float fast_rcp(float q) { return 1.0f / q; }

// ---- A and B, batched form (fft_batch.cuh): a warp owns FOUR lines in shared memory ------------------------
// A4: one warp = one kx >= 0 column, its four fields are the four lines.  The column pair (kx, -kx) is read
//     straight from global memory into registers (16 independent 16-byte loads per lane, no staging table), each
//     entry is divided by k^2 once (as a reciprocal) and the four symmetrised spectra are written in natural
//     layout; after the batched inverse transform the CTA stores all four fields as COLS-column tiles.
//     Measured and dropped: twiddles in registers (230 registers: 175 -> 193 us per 148 environments) and a
//     persistent CTA that issues the next job's column loads before its tile store (222-255 registers: 198-216 us).
//     FPW = fields per warp: 4 (one warp per column) or 2 (two warps per column: psi-based u, v and omega-based
//     w_x, w_y; half the shared memory per warp, i.e. twice the resident warps, for a row pass with 48 tasks on
//     2 x 32 lanes).
template <typename T, int P1, int P2, int NN, int COLS, int FPW>
__global__ void __launch_bounds__(COLS * (4 / FPW) * 32, FPW == 2 ? 2 : 1)
ns_ypass_inv4_kernel(const __grid_constant__ NsArgs<T> A) {
    using G = BatchLayout<P1, P2>;
    using C = typename V2<T>::type;
    constexpr int NP = G::N, N = NN, NH = NN / 2 + 1, NHP = (NH + 3) / 4 * 4;
    constexpr int WS = FPW * G::LS + 2;                          // per-warp stride: == 2 (mod 8) entries, see tile store
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* s_tw = reinterpret_cast<C*>(smem_raw);
    C* s_xb0 = s_tw + NP;
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int wc = w % COLS, half = w / COLS;                    // FPW == 2: half 0 = psi-based fields 0, 1; half 1 = fields 2, 3
    const int env = blockIdx.y, a0 = blockIdx.x * COLS, a = a0 + wc;
    // twiddle table: the global loads are issued here, the shared-memory stores after the column loads below, so that
    // the two latencies overlap (a load-store loop up front was 8 % of the samples, all of them waiting)
    constexpr int NTW = (NP + COLS * (4 / FPW) * 32 - 1) / (COLS * (4 / FPW) * 32);
    C twv[NTW];
#pragma unroll
    for (int i = 0; i < NTW; ++i) {
        const int q = i * (int)blockDim.x + threadIdx.x;
        twv[i] = q < NP ? A.tw_b[q] : V2<T>::make(T(0), T(0));
    }
    auto stage_tw = [&]() {
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            const int q = i * (int)blockDim.x + threadIdx.x;
            if (q < NP) s_tw[q] = twv[i];
        }
    };
    C* xb = s_xb0 + w * WS;
    if (a < NH) {
        const int ib = (N - a) % N;                               // column of -kx in the unpadded array
        const bool xpartner = unpad_idx((NP - a) % NP, NP, N) >= 0;   // pad() keeps (-kx) ?
        const T kxa = A.kx[a], kxb = A.kx[ib];
        const C* srcA = A.fin + (size_t)env * N * N + (size_t)a * N;
        const C* srcB = A.fin + (size_t)env * N * N + (size_t)ib * N;
        constexpr int NB = (NP + 31) / 32;
        C c1[NB], c2[NB];
        T k1[NB], k2[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int kyp = i * 32 + t;
            c1[i] = c2[i] = V2<T>::make(T(0), T(0)); k1[i] = k2[i] = T(0);
            if (i * 32 > N / 2 && i * 32 + 31 < NP - N / 2) continue;     // all-zero block of pad()
            if (NP % 32 != 0 && kyp >= NP) continue;
            const int j1 = unpad_idx(kyp, NP, N);
            const int j2 = xpartner ? unpad_idx((NP - kyp) % NP, NP, N) : -1;
            if (j1 >= 0) { c1[i] = srcA[j1]; k1[i] = A.ky[j1]; }
            if (j2 >= 0) { c2[i] = srcB[j2]; k2[i] = A.ky[j2]; }
        }
        stage_tw();
        const bool psi = FPW == 4 || half == 0, omg = FPW == 4 || half == 1;
        C* opsi = xb;
        C* oomg = xb + (FPW == 4 ? 2 * G::LS : 0);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int kyp = i * 32 + t;
            if (NP % 32 != 0 && kyp >= NP) continue;
            const int o = G::nat(kyp);
            if (i * 32 > N / 2 && i * 32 + 31 < NP - N / 2) {
                const C z = V2<T>::make(T(0), T(0));
#pragma unroll
                for (int f = 0; f < FPW; ++f) xb[f * G::LS + o] = z;
                continue;
            }
            // X_h = (X(k) + conj(X(-k))) / 2 with X = i m c:   u = i ky psi, v = -i kx psi, w_x = i kx omega, w_y = i ky omega
            if (psi) {
                // psi_hat = omega_hat ./ kx2ky2, psi_hat[1,1] = 0   (fluid_rk4.jl:152-153); absent entries are 0
                const T qa = k1[i] * k1[i] + kxa * kxa, qb = k2[i] * k2[i] + kxb * kxb;
                const T ra = qa > T(0) ? fast_rcp(qa) : T(0), rb = qb > T(0) ? fast_rcp(qb) : T(0);
                const C p1 = V2<T>::make(c1[i].x * ra, c1[i].y * ra), p2 = V2<T>::make(c2[i].x * rb, c2[i].y * rb);
                opsi[o]         = V2<T>::make(T(0.5) * (-k1[i] * p1.y - k2[i] * p2.y), T(0.5) * (k1[i] * p1.x - k2[i] * p2.x));
                opsi[G::LS + o] = V2<T>::make(T(0.5) * (kxa * p1.y + kxb * p2.y),      T(0.5) * (-kxa * p1.x + kxb * p2.x));
            }
            if (omg) {
                oomg[o]         = V2<T>::make(T(0.5) * (-kxa * c1[i].y - kxb * c2[i].y), T(0.5) * (kxa * c1[i].x - kxb * c2[i].x));
                oomg[G::LS + o] = V2<T>::make(T(0.5) * (-k1[i] * c1[i].y - k2[i] * c2[i].y), T(0.5) * (k1[i] * c1[i].x - k2[i] * c2[i].x));
            }
        }
    }
    if (a >= NH) stage_tw();
    __syncthreads();                                             // twiddles staged; xb written by its own warp
    if (a < NH)
        fft_batch_nt<T, P1, P2, FPW, +1, false, BatchNoSync>(
            xb, s_tw, BatchTwiddles<T, P1, P2>(), t, [&](int line, int col, int r) { return xb[line * G::LS + r * G::S + col]; },
            [](int) {});
    __syncthreads();
    {
        // COLS-column tile store: thread -> (column c, row y); consecutive lanes = COLS columns x 32/COLS rows.
        // Row y of the transposed layout sits at (y % P2) * S + y / P2; the per-warp stride shifts column c by
        // 2c entries, so the 8 lanes of a quarter warp (COLS = 4: 4 columns x 2 rows) hit 32 distinct banks.
        constexpr int TPF = COLS * 32;                            // threads per tile-store group
        const int grp = threadIdx.x / TPF, tid = threadIdx.x % TPF;
        const int c = tid % COLS, y0 = tid / COLS;
        if (a0 + c < NH) {
#pragma unroll 1
            for (int f = grp * FPW; f < grp * FPW + FPW; ++f) {
                const C* src = s_xb0 + ((f / FPW) * COLS + c) * WS + (f % FPW) * G::LS;
                C* dst = A.W + ((size_t)(env * 4 + f) * NP + y0) * NHP + a0 + c;
                int ym = y0 % P2, yd = y0 / P2;
#pragma unroll 4
                for (int y = y0; y < NP; y += 32) {
                    *dst = src[ym * G::S + yd];
                    dst += (size_t)32 * NHP;
                    ym += 32 % P2; yd += 32 / P2;
                    if (ym >= P2) { ym -= P2; ++yd; }
                }
            }
        }
    }
}

// B4: one warp = two y-lines per job, persistent over jobs.  The four inverse transforms along x (u + i v and
//     w_x + i w_y of both lines) are ONE batch; the two real products are packed into one forward transform.
//     The raw half-spectrum rows of W are staged by 1-D bulk copies (cp.async.bulk, one mbarrier per line) INTO the
//     line buffer they will be expanded in (U at entry 0, V at entry NHP), and the copies of the next job are issued
//     as soon as a line buffer is dead (lines 1-3 after the product, line 0 after the split), so that they run under
//     the forward transform and the stores.
template <typename T, int P1, int P2, int NN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
ns_xpass4_kernel(const __grid_constant__ NsArgs<T> A, int n_jobs) {
    using G = BatchLayout<P1, P2>;
    using C = typename V2<T>::type;
    constexpr int NP = G::N, N = NN, NH = NN / 2 + 1, NHP = (NH + 3) / 4 * 4;
    static_assert(2 * NHP <= G::LS, "raw rows must fit the line buffer");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool TWR = true;                                    // twiddles in registers (shared-memory tables measured slower: 195 vs 180 us)
    static_assert(G::TWREG, "kernel B4 needs a factorisation with P1 | 32");
    C* s_twi = nullptr;
    C* s_twf = nullptr;
    C* s_xb0 = reinterpret_cast<C*>(smem_raw);
    uint64_t* s_bar0 = reinterpret_cast<uint64_t*>(s_xb0 + (size_t)WARPS * 4 * G::LS);
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    C* xb = s_xb0 + (size_t)w * 4 * G::LS;
    uint64_t* bar = s_bar0 + w * 4;
    constexpr uint32_t row_bytes = (uint32_t)(NHP * sizeof(C));
    constexpr int JPE = NP / 2;                                   // jobs per environment
    const int stride = gridDim.x * WARPS;
    int job = blockIdx.x * WARPS + w;
    // line l of a job: y-line y0 + (l >> 1), fields (2h, 2h + 1) with h = l & 1
    auto issue = [&](int jb, int l) {
        const int env = jb / JPE, y = (jb % JPE) * 2 + (l >> 1);
        const C* Wu = A.W + ((size_t)(env * 4 + 2 * (l & 1)) * NP + y) * NHP;
        mbar_expect_tx(bar + l, 2 * row_bytes);
        bulk_g2s(xb + l * G::LS, Wu, row_bytes, bar + l);
        bulk_g2s(xb + l * G::LS + NHP, Wu + (size_t)NP * NHP, row_bytes, bar + l);
    };
    if (t == 0) {
        for (int l = 0; l < 4; ++l) mbar_init(bar + l, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (job < n_jobs) for (int l = 0; l < 4; ++l) issue(job, l);
    }
    BatchTwiddles<T, P1, P2> tw;
    tw.load(A.tw_b, t);
    __syncthreads();
    constexpr int NA = (NH + 31) / 32;
    constexpr int LPR = 32 / P1;                                  // lines per column-pass round
    uint32_t ph = 0;
#pragma unroll 1
    for (; job < n_jobs; job += stride, ph ^= 1) {
        const int env = job / JPE, y0 = (job % JPE) * 2;
        // inverse batch; its column pass expands the raw rows while loading: Z = U + i V on kx >= 0,
        // conj(U) + i conj(V) on kx < 0 (G(y, -kx) = conj(G(y, kx)) for a real field), pad()'s zero band between
        fft_batch_nt<T, P1, P2, 4, +1, TWR, BatchSync>(
            xb, s_twi, tw, t,
            [&](int line, int col, int r) {
                const C* ln = xb + line * G::LS;
                const int e = col + P1 * r, lo = P1 * r, hi = P1 * r + P1 - 1;     // lo, hi fold after unrolling
                bool pos, neg;
                if (hi <= N / 2) { pos = true; neg = false; }
                else if (lo >= NP - N / 2 && lo > N / 2) { pos = false; neg = true; }
                else if (lo > N / 2 && hi < NP - N / 2) { pos = false; neg = false; }
                else { pos = e <= N / 2; neg = !pos && e >= NP - N / 2; }
                if (pos) { const C U = ln[e], V = ln[NHP + e]; return V2<T>::make(U.x - V.y, U.y + V.x); }
                if (neg) { const C U = ln[NP - e], V = ln[NHP + NP - e]; return V2<T>::make(U.x + V.y, V.x - U.y); }
                return V2<T>::make(T(0), T(0));
            },
            [&](int round) {
                for (int l = round * LPR; l < round * LPR + LPR; ++l) mbar_wait(bar + l, ph);
            });
        const bool more = job + stride < n_jobs;
        // forward transform of the two real products packed into one complex line; its row pass forms
        // q = -u w_x - v w_y  (fluid_rk4.jl:175) from the four lines at the same (transposed) positions
        fft_batch_tn<T, P1, P2, 1, -1, TWR>(
            xb, s_twf, tw, t,
            [&](int, int row, int r) {
                const int pos = row * G::S + r;
                const C g0 = xb[pos], z0 = xb[G::LS + pos], g1 = xb[2 * G::LS + pos], z1 = xb[3 * G::LS + pos];
                return V2<T>::make(-(g0.x * z0.x + g0.y * z0.y) * A.scale, -(g1.x * z1.x + g1.y * z1.y) * A.scale);
            },
            [&]() {
                if (t == 0 && more) {                              // lines 1-3 are dead: next job's rows
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    for (int l = 1; l < 4; ++l) issue(job + stride, l);
                }
            });
        C za[NA], zb[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int a = i * 32 + t;
            if (a < NH) { za[i] = xb[G::nat(a)]; zb[i] = xb[G::nat((NP - a) % NP)]; }
        }
        __syncwarp();
        if (t == 0 && more) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(job + stride, 0);
        }
        C* Qa = A.Q + ((size_t)env * NP + y0) * NHP;
        C* Qb = Qa + NHP;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int a = i * 32 + t;
            if (a < NH) {
                Qa[a] = V2<T>::make(T(0.5) * (za[i].x + zb[i].x), T(0.5) * (za[i].y - zb[i].y));
                Qb[a] = V2<T>::make(T(0.5) * (za[i].y + zb[i].y), T(-0.5) * (za[i].x - zb[i].x));
            }
        }
    }
}

// ---- C: forward transform along y, chop, RK4 stage update ---------------------------------------------------
// RK4 stage update of U entries at once: all loads are issued before the first store (the arrays may alias
// from the compiler's point of view, so it cannot do this reordering itself).
template <typename T, int U>
__device__ __forceinline__ void rk_update(const NsArgs<T>& A, const T dt, const size_t* idx, const T* k2, const typename V2<T>::type* nl) {
    using C = typename V2<T>::type;
    C fs[U], ph[U], f0[U], ac[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { fs[u] = A.fin[idx[u]]; ph[u] = A.phat[idx[u]]; }
    if (A.stage != 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) { f0[u] = A.y[idx[u]]; ac[u] = A.acc[idx[u]]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        // rhs = lin + advection + p,  lin = -nu * (kx2ky2 .* omghat)     (fluid_rk4.jl:138-142)
        const T kr = (-A.nu * (k2[u] * fs[u].x) + nl[u].x) + ph[u].x;
        const T ki = (-A.nu * (k2[u] * fs[u].y) + nl[u].y) + ph[u].y;
        if (A.stage == 1) {
            A.acc[idx[u]] = V2<T>::make(kr, ki);
            A.fst[idx[u]] = V2<T>::make(fs[u].x + (T(0.5) * dt) * kr, fs[u].y + (T(0.5) * dt) * ki);
        } else if (A.stage == 4) {
            A.yout[idx[u]] = V2<T>::make(f0[u].x + (dt / T(6)) * (ac[u].x + kr), f0[u].y + (dt / T(6)) * (ac[u].y + ki));
        } else {
            A.acc[idx[u]] = V2<T>::make(ac[u].x + T(2) * kr, ac[u].y + T(2) * ki);
            const T c = A.stage == 2 ? T(0.5) * dt : dt;
            A.fst[idx[u]] = V2<T>::make(f0[u].x + c * kr, f0[u].y + c * ki);
        }
    }
}

template <typename T, int P1, int P2, int NN, int COLS>
__global__ void __launch_bounds__(COLS * 32, (sizeof(T) == 8 ? 2 : 4) * (8 / COLS))     // load-latency bound: 16+ warps/SM
ns_ypass_fwd_kernel(const __grid_constant__ NsArgs<T> A) {
    using G = NsGeom<P1, P2>;
    using C = typename V2<T>::type;
    constexpr int NP = G::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = NN, NH = NN / 2 + 1, NHP = (NH + 3) / 4 * 4;
    C* s_tw = reinterpret_cast<C*>(smem_raw);
    C* s_xb0 = s_tw + NP;
    T* s_ky = reinterpret_cast<T*>(s_xb0 + COLS * G::XB);
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int env = blockIdx.y, a0 = blockIdx.x * COLS, a = a0 + w;
    for (int i = threadIdx.x; i < NP; i += blockDim.x) s_tw[i] = A.tw_fwd[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_ky[i] = A.ky[i];
    {
        const int c = threadIdx.x % COLS, y0 = threadIdx.x / COLS;
        const C* src = A.Q + ((size_t)env * NP + y0) * NHP + a0 + c;
        C* srow = s_xb0 + c * G::XB + y0;
        if (a0 + c < NH) {
#pragma unroll 4
            for (int y = y0; y < NP; y += 32) { *srow = *src; src += (size_t)32 * NHP; srow += 32; }
        }
    }
    __syncthreads();
    if (a >= NH) return;
    C* xb = s_xb0 + w * G::XB;
    T zr[G::RMAX], zi[G::RMAX];
    if (t < P2) {
#pragma unroll
        for (int r = 0; r < P1; ++r) { const C v = xb[t + P2 * r]; zr[r] = v.x; zi[r] = v.y; }
    }
    __syncwarp();
    fft_pass<T, P1, P2, -1>(zr, zi, xb, s_tw, t);
    if (t < P1) {
#pragma unroll
        for (int r = 0; r < P2; ++r) xb[t + P1 * r] = V2<T>::make(zr[r], zi[r]);
    }
    __syncwarp();
    const int ib = (N - a) % N;
    const bool two = ib != a;
    const T kxa = A.kx[a], kxb = A.kx[ib];
    const T dt = A.dt_env ? A.dt_env[env] : A.dt;
    const size_t base = (size_t)env * N * N;
    constexpr int U = 2;                                          // rows per lane in flight (N / 32 is even)
    for (int j0 = t; j0 < N; j0 += 32 * U) {
        size_t idx[2 * U]; T k2[2 * U]; C nl[2 * U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + 32 * u;
            const int s = j <= N / 2 ? j : j - N;                // chop: rows (-N/2, N/2]   (fluid_rk4.jl:224-227)
            const int kyp = s < 0 ? s + NP : s;
            const T kyv = s_ky[j];
            idx[u] = base + (size_t)a * N + j; k2[u] = kyv * kyv + kxa * kxa; nl[u] = xb[kyp];
            // F[ky, -kx] = conj(F[-ky, kx]) for the FFT of a real field (includes the +N/2 Nyquist row,
            // whose partner is the -N/2 row of the padded transform)
            C v = xb[(NP - kyp) % NP];
            v.y = -v.y;
            idx[U + u] = base + (size_t)ib * N + j; k2[U + u] = kyv * kyv + kxb * kxb; nl[U + u] = v;
        }
        if (two) rk_update<T, 2 * U>(A, dt, idx, k2, nl);
        else rk_update<T, U>(A, dt, idx, k2, nl);
    }
}

// ---- generic batched line FFT for the N x N transforms of featurize / prepare_action ----------------------
// One warp per line.  Element e of line l of environment b sits at  b*env_stride + l*line_stride + e*elem_stride.
template <typename T, int N1, int N2, int SIGN>
__global__ void __launch_bounds__(256) ns_lines_kernel(const void* __restrict__ in, void* __restrict__ out,
                                                       const typename V2<T>::type* __restrict__ tw, int in_real,
                                                       int out_real, T out_scale, int n_lines, long long line_stride,
                                                       long long elem_stride, long long env_stride) {
    using C = typename V2<T>::type;
    constexpr int N = N1 * N2;
    constexpr int XB = N1 * PassStride<N1, N2>::value;
    __shared__ C s_tw[N];
    __shared__ C s_xb[8][XB];
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_tw[i] = tw[i];
    __syncthreads();
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const int line = blockIdx.x * 8 + w;
    if (line >= n_lines) return;
    const long long base = (long long)blockIdx.y * env_stride + (long long)line * line_stride;
    constexpr int RMAX = N1 > N2 ? N1 : N2;
    T zr[RMAX], zi[RMAX];
    if (t < N2) {
#pragma unroll
        for (int r = 0; r < N1; ++r) {
            const long long o = base + (long long)(t + N2 * r) * elem_stride;
            if (in_real) { zr[r] = reinterpret_cast<const T*>(in)[o]; zi[r] = T(0); }
            else { const C v = reinterpret_cast<const C*>(in)[o]; zr[r] = v.x; zi[r] = v.y; }
        }
    }
    fft_pass<T, N1, N2, SIGN>(zr, zi, s_xb[w], s_tw, t);
    if (t < N1) {
#pragma unroll
        for (int r = 0; r < N2; ++r) {
            const long long o = base + (long long)(t + N1 * r) * elem_stride;
            if (out_real) reinterpret_cast<T*>(out)[o] = zr[r] * out_scale;
            else reinterpret_cast<C*>(out)[o] = V2<T>::make(zr[r] * out_scale, zi[r] * out_scale);
        }
    }
}

// max |omega_hat| per environment (PDEenv.jl:227 with check_max_value = "y"; quirk Q6)
template <typename T>
__global__ void __launch_bounds__(256) ns_vmax_kernel(const typename V2<T>::type* __restrict__ y, int n, T* vmax) {
    __shared__ T s[8];
    const typename V2<T>::type* ye = y + (size_t)blockIdx.x * n;
    T m = T(0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, hypot(ye[i].x, ye[i].y));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < 8; ++i) m = fmax(m, s[i]); vmax[blockIdx.x] = m; }
}

// ---- adaptive-step parity mode (FluidSetup.jl:178-186) ---------------------------------------------
// The shipped Fluid scripts wire do_step2 = solve(ODEProblem(f, y, (t, t + dt), p), RK4(), reltol = tol, abstol = tol).
// OrdinaryDiffEq's controller is third-party and unpinned, so -- like the Keller-Segel adaptive mode (kseg.cu) -- this is
// an error-controlled integrator of the SAME tableau: classical RK4 with step doubling per ENVIRONMENT (one step of h
// against two of h/2, e = (y2 - y1)/15, the extrapolated y2 + e kept, the 16x larger error of the single full step held
// below the tolerance; RMS norm over the N^2 complex entries of |16 e| / (atol + rtol max(|y|, |y_new|)); factor
// 0.9 err^(-1/5) in [0.2, 5]; first trial step dt / oversampling or the previous env step's last accepted size).
// The environments advance in lock step through ATTEMPTS (3 RK4 steps = 12 rhs launches for all of them); each has its
// own t, h and accept / reject decision, a finished one rides along with step size 0.
constexpr int kNsMaxAttempts = 2000;

template <typename T>
struct NsAdapt {
    T* t; T* h; T* hs; T* hh;       // [B] time inside the env step, natural step, this attempt's step and its half
    T* hlast; int* nsub;            // context arrays: warm start, {accepted, rejected}
    int* n_active;                  // environments that need another attempt
    T dt, h0, rtol, atol;
};

template <typename T>
__global__ void ns_adapt_begin_kernel(NsAdapt<T> D, int B) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    const T h = D.hlast[e] > T(0) ? D.hlast[e] : D.h0;
    const T hs = h >= D.dt ? D.dt : h;
    D.t[e] = T(0); D.h[e] = h; D.hs[e] = hs; D.hh[e] = T(0.5) * hs;
    D.nsub[2 * e] = 0; D.nsub[2 * e + 1] = 0;
}

// one CTA per environment: error norm of the attempt, accept / reject, controller, next attempt's step
template <typename T>
__global__ void __launch_bounds__(1024) ns_adapt_finish_kernel(NsAdapt<T> D, typename V2<T>::type* __restrict__ y,
                                                               const typename V2<T>::type* __restrict__ y1,
                                                               const typename V2<T>::type* __restrict__ y2, int n) {
    using C = typename V2<T>::type;
    const int e = blockIdx.x;
    const T hs = D.hs[e];
    if (!(hs > T(0))) return;                                    // finished earlier
    C* ye = y + (size_t)e * n;
    const C* y1e = y1 + (size_t)e * n;
    const C* y2e = y2 + (size_t)e * n;
    double e2 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const C a = ye[i], b1 = y1e[i], b2 = y2e[i];
        const double ex = ((double)b2.x - (double)b1.x) / 15.0, ey = ((double)b2.y - (double)b1.y) / 15.0;
        const double nx = (double)b2.x + ex, ny = (double)b2.y + ey;
        const double sc = (double)D.atol + (double)D.rtol * fmax(hypot((double)a.x, (double)a.y), hypot(nx, ny));
        e2 += 256.0 * (ex * ex + ey * ey) / (sc * sc);
    }
    __shared__ double s_red[32];
    __shared__ int s_ok;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = e2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) tot += s_red[w];
        const double err = sqrt(tot / n);
        const bool ok = err <= 1.0;                              // NaN compares false: rejected, step shrinks
        const T h = D.h[e];
        T t = D.t[e];
        const bool last = t + h >= D.dt;
        if (ok) { t = last ? D.dt : t + hs; ++D.nsub[2 * e]; }
        else ++D.nsub[2 * e + 1];
        double fac = err == err ? (err > 0.0 ? 0.9 * pow(err, -0.2) : 5.0) : 0.2;
        fac = fmin(5.0, fmax(0.2, fac));
        // an accepted step that was truncated to land on t + dt says nothing about the natural step size: keep h
        const T hn = (ok && hs < h) ? h : (T)((double)hs * fac);
        // give up (state left at the last accepted value, warm start forgotten) when the step underflows -- a state that
        // is already non-finite rejects every attempt -- or after kNsMaxAttempts attempts
        const bool stuck = !(hn > D.dt * T(1e-10)) || D.nsub[2 * e] + D.nsub[2 * e + 1] >= kNsMaxAttempts;
        const bool done = !(t < D.dt) || stuck;
        const T hsn = done ? T(0) : (t + hn >= D.dt ? D.dt - t : hn);
        D.t[e] = t; D.h[e] = hn; D.hs[e] = hsn; D.hh[e] = T(0.5) * hsn;
        if (done) D.hlast[e] = stuck ? T(0) : hn; else atomicAdd(D.n_active, 1);
        s_ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (s_ok) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const C b1 = y1e[i], b2 = y2e[i];
            ye[i] = V2<T>::make((T)((double)b2.x + ((double)b2.x - (double)b1.x) / 15.0),
                                (T)((double)b2.y + ((double)b2.y - (double)b1.y) / 15.0));
        }
    }
}

struct Fact { int n, n1, n2; };
const Fact kFacts[] = {{64, 8, 8}, {96, 8, 12}, {128, 8, 16}, {192, 12, 16}, {256, 16, 16}, {384, 16, 24}};

inline NsProb* prob(const pdeb200_ctx* c) { return static_cast<NsProb*>(c->prob); }

template <typename T>
int32_t upload(pdeb200_ctx* c, void** dst, const std::vector<T>& v) {
    PDEB_CUDA(c, cudaMalloc(dst, v.size() * sizeof(T)));
    PDEB_CUDA(c, cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PDEB200_OK;
}

// table for fft_pass<T, P, Q, .>: tw[n*P + t] = exp(-2 pi i n t / (P Q)), n < Q, t < P
template <typename T>
std::vector<typename V2<T>::type> twiddles(int P, int Q) {
    const long double two_pi = 6.283185307179586476925286766559L;
    const int N = P * Q;
    std::vector<typename V2<T>::type> tw((size_t)N);
    for (int n = 0; n < Q; ++n)
        for (int t = 0; t < P; ++t) {
            const long double a = -two_pi * (long double)((long long)n * t % N) / N;
            tw[(size_t)n * P + t] = V2<T>::make((T)cosl(a), (T)sinl(a));
        }
    return tw;
}

template <typename T>
int32_t setup_t(pdeb200_ctx* c) {
    using C = typename V2<T>::type;
    NsProb* P = prob(c);
    const pdeb200_config& g = c->cfg;
    const int N = P->N;
    int32_t rc;
    if ((rc = upload<C>(c, &P->tw_inv, twiddles<T>(P->p2, P->p1)))) return rc;    // fft_pass<P2, P1>
    if ((rc = upload<C>(c, &P->tw_fwd, twiddles<T>(P->p1, P->p2)))) return rc;    // fft_pass<P1, P2>
    if ((rc = upload<C>(c, &P->tw_n, twiddles<T>(P->n1, P->n2)))) return rc;      // fft_pass<N1, N2>
    {
        const int bp1 = 32 % P->p1 == 0 ? P->p1 : 16, bp2 = P->NP / bp1;            // the batched kernels' factorisation
        if ((rc = upload<C>(c, &P->tw_b, twiddles<T>(bp2, bp1)))) return rc;        // fft_pass<BP2, BP1>
    }
    // kx = [0:(nx/2); (-nx/2+1):(-1)] / Lx * 2pi  (Nyquist kept, positive side; FluidSetup.jl:106-107)
    std::vector<T> kx(N), ky(N);
    for (int i = 0; i < N; ++i) {
        const double s = i <= N / 2 ? i : i - N;
        kx[i] = (T)(s / g.Lx * 2 * M_PI);
        ky[i] = (T)(s / g.Ly * 2 * M_PI);
    }
    if ((rc = upload<T>(c, &P->kx, kx))) return rc;
    if ((rc = upload<T>(c, &P->ky, ky))) return rc;
    const size_t B = g.n_envs, nn = (size_t)N * N;
    PDEB_CUDA(c, cudaMalloc(&P->fst, B * nn * sizeof(C)));
    PDEB_CUDA(c, cudaMalloc(&P->acc, B * nn * sizeof(C)));
    PDEB_CUDA(c, cudaMalloc(&P->tmpc, B * nn * sizeof(C)));
    PDEB_CUDA(c, cudaMalloc(&P->p_phys, B * nn * sizeof(T)));
    PDEB_CUDA(c, cudaMalloc(&P->om_phys, B * nn * sizeof(T)));
    PDEB_CUDA(c, cudaMemset(P->p_phys, 0, B * nn * sizeof(T)));
    PDEB_CUDA(c, cudaMalloc(&P->W, (size_t)P->chunk * 4 * P->NP * P->NHP * sizeof(C)));
    PDEB_CUDA(c, cudaMalloc(&P->Q, (size_t)P->chunk * P->NP * P->NHP * sizeof(C)));
    c->prob_p_phys = P->p_phys;
    if (g.adaptive) {
        PDEB_CUDA(c, cudaMalloc(&P->y2, B * nn * sizeof(C)));
        PDEB_CUDA(c, cudaMalloc(&P->ad_state, B * 4 * sizeof(T)));
        PDEB_CUDA(c, cudaMalloc((void**)&P->ad_active, sizeof(int)));
        PDEB_CUDA(c, cudaMallocHost((void**)&P->h_active, sizeof(int)));
    }
    return PDEB200_OK;
}

template <typename T, int P1, int P2>
size_t smem_a(int N) {
    using G = NsGeom<P1, P2>; using C = typename V2<T>::type;
    constexpr int COLS = NsColsA<T>::value;
    return ((size_t)G::NP + COLS * G::XB + COLS * 2 * N) * sizeof(C) + (size_t)N * sizeof(T) +
           (size_t)G::NP * sizeof(short2);
}
template <typename T, int P1, int P2>
size_t smem_b(int NHP) {
    using G = NsGeom<P1, P2>; using C = typename V2<T>::type;
    constexpr int COLS = NsColsB<T>::value;
    return ((size_t)2 * G::NP + COLS * G::XB + COLS * G::NP + (size_t)COLS * 4 * NHP) * sizeof(C) +
           (size_t)COLS * G::NP * sizeof(T) + (size_t)COLS * 2 * sizeof(uint64_t);
}
template <typename T, int P1, int P2>
size_t smem_c(int N) {
    using G = NsGeom<P1, P2>; using C = typename V2<T>::type;
    return ((size_t)G::NP + NsColsC<T>::value * G::XB) * sizeof(C) + (size_t)N * sizeof(T);
}

template <typename K>
int32_t set_smem(pdeb200_ctx* c, K kern, size_t bytes) {
    if (bytes > 227 * 1024) return fail(c, PDEB200_EUNSUPPORTED, "NS: shared-memory budget exceeded");
    PDEB_CUDA(c, ensure_dyn_smem(kern, bytes, c->device));
    return PDEB200_OK;
}

// RK4 x oversampling for all environments (FluidSetup.jl:163-172), in chunks of `chunk` environments so that
// the work arrays W and Q of a chunk can stay in L2 between the three kernels of a stage.
template <typename T, int P1, int P2, int NN>
int32_t rk4_t(pdeb200_ctx* c) {
    using C = typename V2<T>::type;
    NsProb* P = prob(c);
    const pdeb200_config& g = c->cfg;
    const int N = P->N, NP = P->NP;
    const size_t nn = (size_t)N * N;
    constexpr int COLS_A = NsColsA<T>::value;
    auto kA = ns_ypass_inv_kernel<T, P1, P2, NN, COLS_A>;
    constexpr int COLS_B = NsColsB<T>::value;
    auto kB = ns_xpass_kernel<T, P1, P2, NN, COLS_B>;
    constexpr int COLS_C = NsColsC<T>::value;
    auto kC = ns_ypass_fwd_kernel<T, P1, P2, NN, COLS_C>;
    const size_t sa = smem_a<T, P1, P2>(N), sb = smem_b<T, P1, P2>(P->NHP), sc = smem_c<T, P1, P2>(N);
    int32_t rc;
    if ((rc = set_smem(c, kA, sa)) || (rc = set_smem(c, kB, sb)) || (rc = set_smem(c, kC, sc))) return rc;
    // batched form of A and B (fft_batch.cuh); PDEB200_NS_LEGACY=1 keeps the one-line-per-warp kernels
    constexpr int COLS_A4 = 4, WARPS_B4 = 8;
    using BL = BatchLayout<P1, P2>;
    // the batched kernels need P1 | 32: 192 points are 16 x 12 there (12 x 16 in the one-line-per-warp kernels)
    constexpr int BP1 = BL::TWREG ? P1 : 16, BP2 = BL::TWREG ? P2 : BL::N / 16;
    using BLB = BatchLayout<BP1, BP2>;
    constexpr int FPW_A4 = 4;      // 2 (two warps per column, 16 resident warps) measured slower: 205 vs 173 us per 148 environments
    auto kA4 = ns_ypass_inv4_kernel<T, BP1, BP2, NN, COLS_A4, FPW_A4>;
    auto kB4 = ns_xpass4_kernel<T, BP1, BP2, NN, WARPS_B4>;
    const size_t sa4 = ((size_t)BLB::N + COLS_A4 * (4 / FPW_A4) * (FPW_A4 * BLB::LS + 2)) * sizeof(C);
    const size_t sb4 = ((size_t)WARPS_B4 * 4 * BLB::LS) * sizeof(C) + (size_t)WARPS_B4 * 4 * sizeof(uint64_t);
    const bool batched = !P->legacy;
    if (batched && ((rc = set_smem(c, kA4, sa4)) || (rc = set_smem(c, kB4, sb4)))) return rc;
    int b4_ctas = 1;                                              // resident CTAs per SM of the persistent kernel B
    if (batched) {
        PDEB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b4_ctas, kB4, WARPS_B4 * 32, sb4));
        b4_ctas = std::max(1, b4_ctas);
    }
    NsArgs<T> A;
    A.N = N; A.NH = P->NH; A.NHP = P->NHP;
    A.tw_inv = (const C*)P->tw_inv; A.tw_fwd = (const C*)P->tw_fwd; A.tw_b = (const C*)P->tw_b;
    A.kx = (const T*)P->kx; A.ky = (const T*)P->ky;
    A.W = (C*)P->W; A.Q = (C*)P->Q;
    A.nu = (T)g.nu; A.dt = (T)(g.dt / g.oversampling);
    // ifft normalisation of the two factors (1/NP^2 each) and the 1.5*1.5 of fluid_rk4.jl:178
    const double np2 = (double)NP * NP;
    A.scale = (T)((g.ifpad ? 2.25 : 1.0) / (np2 * np2));
    A.dt_env = nullptr;
    // one classical RK4 step (4 rhs evaluations) of `ne` environments: A.y -> A.yout
    auto rk4_step = [&](int ne) {
        for (int stage = 1; stage <= 4; ++stage) {
            A.stage = stage;
            A.fin = stage == 1 ? A.y : A.fst;
            if (batched) {
                kA4<<<dim3((P->NH + COLS_A4 - 1) / COLS_A4, ne), COLS_A4 * (4 / FPW_A4) * 32, sa4, c->stream>>>(A);
                kB4<<<std::min(P->n_sm * b4_ctas, (ne * (NP / 2) + WARPS_B4 - 1) / WARPS_B4), WARPS_B4 * 32, sb4, c->stream>>>(A, ne * (NP / 2));
            } else {
                kA<<<dim3((P->NH + COLS_A - 1) / COLS_A, ne), COLS_A * 32, sa, c->stream>>>(A);
                kB<<<dim3((NP / 2 + COLS_B - 1) / COLS_B, ne), COLS_B * 32, sb, c->stream>>>(A);
            }
            kC<<<dim3((P->NH + COLS_C - 1) / COLS_C, ne), COLS_C * 32, sc, c->stream>>>(A);
            c->launches += 3;
        }
    };
    if (g.adaptive) {
        const int B = g.n_envs;
        NsAdapt<T> D;
        T* st = (T*)P->ad_state;
        D.t = st; D.h = st + B; D.hs = st + 2 * (size_t)B; D.hh = st + 3 * (size_t)B;
        D.hlast = (T*)c->d_hlast; D.nsub = c->d_nsub; D.n_active = P->ad_active;
        D.dt = (T)g.dt; D.h0 = (T)(g.dt / g.oversampling); D.rtol = (T)g.rtol; D.atol = (T)g.atol;
        A.fst = (C*)P->fst; A.acc = (C*)P->acc; A.phat = (const C*)c->p;
        C* y = (C*)c->y; C* y1 = (C*)P->tmpc; C* y2 = (C*)P->y2;
        ns_adapt_begin_kernel<T><<<(B + 255) / 256, 256, 0, c->stream>>>(D, B);
        c->launches += 1;
        for (int attempt = 0; attempt <= kNsMaxAttempts; ++attempt) {
            A.y = y;  A.yout = y1; A.dt_env = D.hs; rk4_step(B);          // one step of h
            A.y = y;  A.yout = y2; A.dt_env = D.hh; rk4_step(B);          // two steps of h/2
            A.y = y2; A.yout = y2; A.dt_env = D.hh; rk4_step(B);
            PDEB_CUDA(c, cudaMemsetAsync(P->ad_active, 0, sizeof(int), c->stream));
            ns_adapt_finish_kernel<T><<<B, 1024, 0, c->stream>>>(D, y, y1, y2, (int)nn);
            c->launches += 1;
            PDEB_CUDA(c, cudaMemcpyAsync(P->h_active, P->ad_active, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            PDEB_CUDA(c, cudaStreamSynchronize(c->stream));          // the host decides whether another attempt is needed
            if (*P->h_active == 0) break;
        }
        PDEB_CUDA(c, cudaGetLastError());
        return PDEB200_OK;
    }
    for (int e0 = 0; e0 < g.n_envs; e0 += P->chunk) {
        const int ne = std::min(P->chunk, g.n_envs - e0);
        A.y = (C*)c->y + (size_t)e0 * nn; A.yout = A.y; A.fst = (C*)P->fst + (size_t)e0 * nn; A.acc = (C*)P->acc + (size_t)e0 * nn;
        A.phat = (const C*)c->p + (size_t)e0 * nn;
        for (int s = 0; s < g.oversampling; ++s) rk4_step(ne);
    }
    PDEB_CUDA(c, cudaGetLastError());
    return PDEB200_OK;
}

// N x N transform of all environments: pass along j (contiguous), then along i (stride N).
template <typename T, int N1, int N2, int SIGN>
int32_t fft2_t(pdeb200_ctx* c, const void* in, int in_real, void* out, int out_real, double scale) {
    using C = typename V2<T>::type;
    NsProb* P = prob(c);
    const int N = P->N;
    const long long nn = (long long)N * N;
    const dim3 grid((N + 7) / 8, c->cfg.n_envs);
    auto k = ns_lines_kernel<T, N1, N2, SIGN>;
    k<<<grid, 256, 0, c->stream>>>(in, P->tmpc, (const C*)P->tw_n, in_real, 0, T(1), N, (long long)N, 1LL, nn);
    k<<<grid, 256, 0, c->stream>>>(P->tmpc, out, (const C*)P->tw_n, 0, out_real, (T)scale, N, 1LL, (long long)N, nn);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return PDEB200_OK;
}

template <typename T, int N1, int N2>
int32_t sensors_t(pdeb200_ctx* c, const uint8_t* d_mask) {
    NsProb* P = prob(c);
    const int N = P->N;
    // y = real(ifft(env.y))   (FluidSetup.jl:189, 206-208)
    int32_t rc = fft2_t<T, N1, N2, +1>(c, c->y, 0, P->om_phys, 1, 1.0 / ((double)N * N));
    if (rc) return rc;
    sensors_phys_kernel<T><<<c->cfg.n_envs, 128, 0, c->stream>>>(
        1, c->npts, c->cfg.n_sensors, EllTable<T>{c->sens.d_idx, (const T*)c->sens.d_w, c->sens.nnz_max, c->sens.n_rows},
        d_mask, (const T*)P->om_phys, 0, (T*)c->sensors, (T*)c->vmax);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    if (c->cfg.check_max_value == PDEB200_CHECK_Y) {
        ns_vmax_kernel<T><<<c->cfg.n_envs, 256, 0, c->stream>>>((const typename V2<T>::type*)c->y, N * N, (T*)c->vmax);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 1;
    }
    return PDEB200_OK;
}

template <typename T, int N1, int N2, int P1, int P2>
int32_t core_t(pdeb200_ctx* c) {
    NsProb* P = prob(c);
    // p_hat = fft(p)   (FluidSetup.jl:260); the physical sum was written by actuate_kernel
    int32_t rc = fft2_t<T, N1, N2, -1>(c, P->p_phys, 1, c->p, 0, 1.0);
    if (rc) return rc;
    if ((rc = rk4_t<T, P1, P2, N1 * N2>(c))) return rc;
    return sensors_t<T, N1, N2>(c, nullptr);
}

template <typename T>
int32_t core_dispatch(pdeb200_ctx* c) {
    NsProb* P = prob(c);
    const int key = P->N * 1000 + P->NP;
    switch (key) {
        case 64 * 1000 + 96:   return core_t<T, 8, 8, 8, 12>(c);
        case 64 * 1000 + 64:   return core_t<T, 8, 8, 8, 8>(c);
        case 128 * 1000 + 192: return core_t<T, 8, 16, 12, 16>(c);
        case 128 * 1000 + 128: return core_t<T, 8, 16, 8, 16>(c);
        case 256 * 1000 + 384: return core_t<T, 16, 16, 16, 24>(c);
        case 256 * 1000 + 256: return core_t<T, 16, 16, 16, 16>(c);
    }
    return fail(c, PDEB200_EUNSUPPORTED, "NS: unsupported grid");
}

template <typename T>
int32_t sensors_dispatch(pdeb200_ctx* c, const uint8_t* d_mask) {
    switch (prob(c)->N) {
        case 64:  return sensors_t<T, 8, 8>(c, d_mask);
        case 128: return sensors_t<T, 8, 16>(c, d_mask);
        case 256: return sensors_t<T, 16, 16>(c, d_mask);
    }
    return fail(c, PDEB200_EUNSUPPORTED, "NS: unsupported grid");
}

}  // namespace

int32_t ns_setup(pdeb200_ctx* c) {
    const pdeb200_config& g = c->cfg;
    if (g.nx != g.ny) return fail(c, PDEB200_EUNSUPPORTED, "NS: nx must equal ny (the reference builds kx2ky2 for square grids)");
    if (g.nx != 64 && g.nx != 128 && g.nx != 256) return fail(c, PDEB200_EUNSUPPORTED, "NS: nx must be 64, 128 or 256");
    if (g.oversampling < 1) return fail(c, PDEB200_EINVAL, "NS: oversampling must be >= 1");
    if (g.sensors_per_axis < 1 || g.sensors_per_axis * g.sensors_per_axis != g.n_sensors)
        return fail(c, PDEB200_EINVAL, "NS: n_sensors must equal sensors_per_axis^2");
    NsProb* P = new NsProb();
    c->prob = P;
    P->N = g.nx;
    P->NP = g.ifpad ? g.nx * 3 / 2 : g.nx;
    P->NH = P->N / 2 + 1;
    P->NHP = (P->NH + 3) / 4 * 4;
    for (const Fact& f : kFacts) {
        if (f.n == P->N) { P->n1 = f.n1; P->n2 = f.n2; }
        if (f.n == P->NP) { P->p1 = f.n1; P->p2 = f.n2; }
    }
    const char* e = getenv("PDEB200_NS_CHUNK");
    P->chunk = e ? std::max(1, atoi(e)) : g.n_envs;
    P->chunk = std::min(P->chunk, g.n_envs);
    cudaDeviceGetAttribute(&P->n_sm, cudaDevAttrMultiProcessorCount, c->device);
    e = getenv("PDEB200_NS_LEGACY");
    P->legacy = e && atoi(e) != 0;
    return g.dtype == PDEB200_F64 ? setup_t<double>(c) : setup_t<float>(c);
}

int32_t ns_core(pdeb200_ctx* c) {
    return c->cfg.dtype == PDEB200_F64 ? core_dispatch<double>(c) : core_dispatch<float>(c);
}

int32_t ns_sensors(pdeb200_ctx* c, const uint8_t* d_mask) {
    return c->cfg.dtype == PDEB200_F64 ? sensors_dispatch<double>(c, d_mask) : sensors_dispatch<float>(c, d_mask);
}

// Algorithmic cost of one env step:
//   bytes: omega_hat in + out (complex), action in, obs + reward out, done
//   flops: per rhs 4 inverse + 1 forward real-data 2-D FFTs of the padded size at 2.5 n log2 n per real point,
//          pruned to the non-zero lines, + pointwise work; 4 rhs per RK4 substep
int32_t ns_cost(const pdeb200_ctx* c, double* bytes, double* flops) {
    const pdeb200_config& g = c->cfg;
    const NsProb* P = prob(c);
    const double N = g.nx, NP = P->NP, w = (double)c->esz;
    if (bytes) *bytes = 2 * (2 * N * N * w) + g.n_actuators * (w * c->a_rows + w * c->obs_rows + w) + 1;
    if (flops) {
        const double cfft = 5 * NP * std::log2(NP);                      // one complex line
        const double nh = N / 2 + 1;
        const double per_rhs = 4 * nh * cfft + (NP / 2) * 5 * cfft + nh * cfft + 40 * N * N + 6 * NP * NP;
        *flops = 4.0 * g.oversampling * per_rhs + 2 * 2 * N * cfft * 2;
    }
    return PDEB200_OK;
}

void ns_free(pdeb200_ctx* c) {
    NsProb* P = prob(c);
    if (!P) return;
    for (void* p : {P->tw_inv, P->tw_fwd, P->tw_n, P->tw_b, P->kx, P->ky, P->fst, P->acc, P->W, P->Q, P->p_phys, P->om_phys, P->tmpc})
        if (p) cudaFree(p);
    for (void* p : {P->y2, P->ad_state, (void*)P->ad_active}) if (p) cudaFree(p);
    if (P->h_active) cudaFreeHost(P->h_active);
    delete P;
    c->prob = nullptr;
    c->prob_p_phys = nullptr;
}

}  // namespace pdeb200
