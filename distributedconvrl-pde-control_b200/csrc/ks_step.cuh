// Fused Kuramoto-Sivashinsky environment step for thousands of environments.
//
// Core kernel of the env step: per environment
//   do_step          scripts/KS/setup/KSSetup.jl:130-160   (CNAB2 pseudo-spectral, `S` substeps)
//   sensor dots      KSSetup.jl:168-170, 200-202           (<y, g_i>, shared by reward and featurize)
//   max |y|          src/PDEenv.jl:227                     (divergence guard)
// with the PDE state resident in registers/shared memory across all substeps.  The column-level
// work on either side (actor, prepare_action sum, reward, observation windows, clock) runs in the
// full-occupancy kernels of glue.cuh; profiling showed that keeping it inside this register-heavy
// kernel made it instruction-cache and latency bound (profiles/r1_ks_step_f64.md, r1a-r1c).
//
// Mapping (B200-first, not a translation of the reference's FFTW calls):
//   * Two environments are packed into ONE complex sequence z = u_a + i*u_b.  Every
//     spectral operation of CNAB2 is complex-linear with Hermitian-symmetric
//     coefficients, and the only nonlinearity, u -> u^2, acts on Re and Im
//     separately in physical space, so the pair is advanced with exactly the
//     complex FFTs the reference performs on one zero-imaginary array -- no
//     real-FFT pre/post passes.
//   * A length-N FFT (N = N1*N2) is a four-step FFT: DFT_N1 in registers ->
//     twiddle -> transpose through shared memory -> DFT_N2 in registers.  TP =
//     max(N1,N2) threads (a half warp for N<=256) own one pair; only __syncwarp()
//     is needed.  Alternating the (N1,N2) and (N2,N1) factorizations for forward
//     and inverse transforms makes each transform's output layout the next one's
//     input layout, so there is never a reorder pass.
//   * u_hat lives in registers across substeps; N^{n-1} and the forcing term live
//     in shared memory; coefficients/twiddles are shared by all pairs of the CTA.
#pragma once
#include "common.cuh"
#include "dft_gen.cuh"
#include "fft_pass.cuh"

namespace pdeb200 {

template <typename T>
struct KsArgs {
    using C = typename V2<T>::type;
    int n_envs, S, n_sensors;
    // spectral constants, spectral natural order k = 0..N-1
    const C* tw12;                     // [N1][N2]  exp(-2*pi*i*a*b/N)
    const C* tw21;                     // [N2][N1]
    const T* c1;                       // A_inv * B
    const T* cN;                       // A_inv * (-alpha/2) / N^2 * 3h/2
    const T* ainvh;                    // A_inv * h (global, read once per env step)
    const C* hm;                       // h * fft(mu*cos(...)) or nullptr (mu == 0)
    T third, inv_n;                    // (h/2)/(3h/2); 1/N
    EllTable<T> sens;                  // rows: sensors, gather over grid points (indices already under `perm`)
    const int* perm;                   // position of grid point n in the natural-order shared line, or nullptr (identity)
    const C* sens_hat;                 // conj(fft(g_0)) when the sensor bases are equally spaced shifted copies of g_0, else nullptr
    int sens_sp;                       // their spacing in grid points: <y, g_i> = ifft(u_hat . sens_hat)[sens_sp * i]
    T* y;                              // [B][N] in/out
    const T* p;                        // [B][N] actuation field (physical)
    T* sensors_out;                    // [B][n_sensors] raw <y, g_i>
    T* vmax_out;                       // [B] max |y|
};

template <int N1, int N2> struct KsGeom {
    static constexpr int N = N1 * N2;
    static constexpr int RMAX = N1 > N2 ? N1 : N2;
    static constexpr int TP = RMAX <= 8 ? 8 : (RMAX <= 16 ? 16 : 32);        // threads per env pair
    static constexpr int S12 = N1 * PassStride<N1, N2>::value;
    static constexpr int S21 = N2 * PassStride<N2, N1>::value;
    static constexpr int XB = (S12 > S21 ? S12 : S21) > N ? (S12 > S21 ? S12 : S21) : N;
};

// Dynamic shared memory per CTA:
//   [tw12 | tw21 | c1 | cN] shared by all pairs, then CTA-level arrays
//   XB[PAIRS][xb] (exchange), PREV[PAIRS][N], F[PAIRS][N].
// PREV and F are dead after the job loop; the sensor gather table is staged there.
// with_prev = false (one substep: N^{n-1} is never read, quirk Q2 seeds it from N^n): the PREV array is not allocated
template <typename T, int N1, int N2>
__host__ __device__ inline size_t ks_smem_bytes(int pairs_per_cta, bool with_prev = true) {
    using G = KsGeom<N1, N2>;
    using C = typename V2<T>::type;
    size_t shared = (size_t)(N1 == N2 ? 1 : 2) * G::N * sizeof(C) + 2 * (size_t)G::N * sizeof(T);
    size_t per_pair = ((size_t)G::XB + (with_prev ? 2 : 1) * G::N) * sizeof(C);
    return shared + per_pair * pairs_per_cta;
}

// Register budget: fp64 keeps z and u_hat (4*RMAX doubles) in registers, so it is compiled for 8 resident
// warps/SM (255 registers); fp32 for 16 warps/SM (128 registers).  The CTA size is a RUNTIME choice
// (blockDim.x = 32*w, w <= kMaxWarps): the host picks w and the CTAs/SM so that the grid fills whole waves.
// LOWREG (fp64 only): F stays in shared memory like in fp32, the kernel is compiled for 12 resident warps/SM (168 registers):
// the variant for few substeps (oversampling <= 4), where the loads / stores / sensor gather around the transforms weigh as
// much as the transforms and latency hiding matters more than the 64 shared-memory wavefronts F costs per substep.
template <typename T, bool LOWREG = false> struct KsMaxWarps { static constexpr int value = sizeof(T) == 8 ? (LOWREG ? 12 : 8) : 16; };

// SPEC: the sensor dots come from one more inverse transform after the job loop (u_hat . conj(fft(g_0)); ks.cu::ks_bases_changed)
// instead of the gather epilogue.  A compile-time switch: keeping u_hat alive past the loop costs the many-substep kernel 24 more
// spilled bytes in its hot loop (183 -> 195 us at oversampling 30), so only the few-substep launches (oversampling <= 4) use it
// (414 -> 392 us at oversampling 1, 131 072 environments).  Doing it as an extra job INSIDE the loop was worse on both (217 / 458 us).
template <typename T, int N1, int N2, bool LOWREG = false, bool SPEC = false>
__global__ void __launch_bounds__(KsMaxWarps<T, LOWREG>::value * 32, 1)
ks_step_kernel(const __grid_constant__ KsArgs<T> A) {
    using G = KsGeom<N1, N2>;
    using C = typename V2<T>::type;
    constexpr int N = G::N, TP = G::TP, RMAX = G::RMAX;
    const int NT = blockDim.x;
    const int PAIRS = NT / TP;
    constexpr int CU = 4;                            // sensors processed together per thread (ILP)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* s_tw12 = reinterpret_cast<C*>(smem_raw);
    C* s_tw21 = (N1 == N2) ? s_tw12 : s_tw12 + N;
    T* s_c1 = reinterpret_cast<T*>(s_tw21 + N);
    T* s_cN = s_c1 + N;
    C* s_xb0 = reinterpret_cast<C*>(s_cN + N);

    const int n_s = A.n_sensors;
    const int t = threadIdx.x % TP;
    const int pic = threadIdx.x / TP;
    const int pair = blockIdx.x * PAIRS + pic;
    C* xb = s_xb0 + (size_t)pic * G::XB;             // exchange buffer; natural-order y at the end
    const bool has_prev = A.S > 1;                   // one substep: no N^{n-1} array (host sizes the shared memory accordingly)
    C* s_prev0 = s_xb0 + (size_t)PAIRS * G::XB;
    C* s_prev = s_prev0 + (size_t)pic * N;           // N^{n-1} (raw fft of (N u)^2)
    C* s_F = s_prev0 + (has_prev ? (size_t)PAIRS * N : 0) + (size_t)pic * N;   // A_inv*h*p_hat + h*m_hat
    unsigned char* dead = reinterpret_cast<unsigned char*>(s_prev0);
    const size_t dead_bytes = (size_t)(has_prev ? 2 : 1) * PAIRS * N * sizeof(C);


    const int ea = 2 * pair, eb = 2 * pair + 1;
    const bool va = ea < A.n_envs, vb = eb < A.n_envs;
    // Pairs beyond the batch keep running on zeros (they share barriers with live pairs).

    T zr[RMAX], zi[RMAX], ur[RMAX], ui[RMAX];
    // F = A_inv*h*p_hat + h*m_hat of this thread's modes: in registers where the register file allows it (fp64 is
    // compiled for 255 registers anyway; fp32 keeps it in shared memory to stay at 128, N > 256 would only spill)
    constexpr bool F_REGS = sizeof(T) == 8 && RMAX <= 16 && !LOWREG;
    T fr[F_REGS ? RMAX : 1], fi[F_REGS ? RMAX : 1];

    // ---- load p into z and y into u (physical layout: thread t < N2 holds n = t + N2*r) -----
#pragma unroll
    for (int r = 0; r < N1; ++r) {
        const int n = t + N2 * r;
        const bool in = t < N2;
        zr[r] = (va && in) ? __ldg(A.p + (size_t)ea * N + n) : T(0);
        zi[r] = (vb && in) ? __ldg(A.p + (size_t)eb * N + n) : T(0);
        ur[r] = (va && in) ? A.y[(size_t)ea * N + n] : T(0);
        ui[r] = (vb && in) ? A.y[(size_t)eb * N + n] : T(0);
    }
    // constant tables after the state loads are in flight (the staging stores wait for their own loads)
    for (int i = threadIdx.x; i < N; i += NT) {
        s_tw12[i] = A.tw12[i];
        if (N1 != N2) s_tw21[i] = A.tw21[i];
        s_c1[i] = A.c1[i]; s_cN[i] = A.cN[i];
    }
    __syncthreads();

    // ---- unified job loop: two forward-only jobs, then S CNAB2 substeps, then the final inverse
    //   job -2: F = A_inv*h*fft(p) + h*m_hat        job -1: u_hat = fft(y)
    //   job n>=0 (KSSetup.jl:144-156): z = ifft(u_hat); z = z^2; z = fft(z); CNAB2 update
    //   job S: only the inverse transform (KSSetup.jl:158)
    for (int job = -2; job <= A.S; ++job) {
        if (job >= 0) {
            // z = u_hat here, at the top, so that z is not live across the loop's back edge (a copy at the end of the
            // CNAB2 update kept 4*N2 extra registers loop-carried and cost 64 register moves per substep)
#pragma unroll
            for (int r = 0; r < N2; ++r) { zr[r] = ur[r]; zi[r] = ui[r]; }
            fft_pass<T, N2, N1, +1>(zr, zi, xb, s_tw12, t);          // z = N * u  (physical)
            if (job == A.S) break;
#pragma unroll
            for (int r = 0; r < N1; ++r) { zr[r] = zr[r] * zr[r]; zi[r] = zi[r] * zi[r]; }
        } else if (job == -1) {
#pragma unroll
            for (int r = 0; r < N1; ++r) { zr[r] = ur[r]; zi[r] = ui[r]; }
        }
        fft_pass<T, N1, N2, -1>(zr, zi, xb, s_tw21, t);
        if (t < N1) {
            if (job == -2) {
#pragma unroll
                for (int r = 0; r < N2; ++r) {
                    const int k = t + N1 * r;
                    const T ah = __ldg(A.ainvh + k);
                    T f_r = ah * zr[r], f_i = ah * zi[r];
                    if (A.hm) { const C m = A.hm[k]; f_r += m.x - m.y; f_i += m.x + m.y; }   // m_hat*(1+i)
                    if (F_REGS) { fr[r] = f_r; fi[r] = f_i; }
                    else s_F[k] = V2<T>::make(f_r, f_i);
                }
            } else if (job == -1) {
#pragma unroll
                for (int r = 0; r < N2; ++r) { ur[r] = zr[r]; ui[r] = zi[r]; }
            } else {
#pragma unroll
                for (int r = 0; r < N2; ++r) {
                    const int k = t + N1 * r;
                    // N^{n-1} := N^n on the first substep (KSSetup.jl:140-141,145; quirk Q2).  The
                    // reference gets N^n there from fft(y^2) and here from fft(ifft(fft(y))^2): equal
                    // to round-off, one transform fewer.
                    const C z1 = (job == 0) ? V2<T>::make(zr[r], zi[r]) : s_prev[k];
                    T f_r, f_i;
                    if (F_REGS) { f_r = fr[r]; f_i = fi[r]; }
                    else { const C f = s_F[k]; f_r = f.x; f_i = f.y; }
                    if (has_prev) s_prev[k] = V2<T>::make(zr[r], zi[r]);
                    // u = c1*u + i*cn*(3h/2 N^n - h/2 N^{n-1}) + F  with cn32 = cn*3h/2:  u = c1*u + F + i*cn32*(N^n - N^{n-1}/3)
                    const T tr = fma(-A.third, z1.x, zr[r]);
                    const T ti = fma(-A.third, z1.y, zi[r]);
                    const T c1 = s_c1[k], cn = s_cN[k];
                    ur[r] = fma(-cn, ti, fma(c1, ur[r], f_r));
                    ui[r] = fma(cn, tr, fma(c1, ui[r], f_i));
                }
            }
        }
    }
    // ---- y = real/imag(ifft(u_hat)); store; keep a copy in xb for the sensors.  With equally spaced sensors the
    // copy is laid out in `spacing` planes (perm) so that, for a given tap, the lanes of a pair read CONSECUTIVE
    // 16-byte entries: the plain natural order costs 8 wavefronts per half warp and tap (lanes 4 points = 64 B
    // apart), which ncu showed as ~10 us of shared-memory time per launch at 8192 environments ----
    // the sensor table's entries are requested now (u_hat / F registers are dead) and stored into the dead PREV / F
    // region after the barrier below: their L2 latency hides behind the epilogue instead of heading the sensor phase
    if (SPEC) {
        // ---- sensors in spectral space: y first, then ONE more inverse transform of u_hat . conj(fft(g_0)) ------------------
        T vmax_a = T(0), vmax_b = T(0);
        if (t < N2) {
#pragma unroll
            for (int r = 0; r < N1; ++r) {
                const int n = t + N2 * r;
                const T ya = zr[r] * A.inv_n, yb = zi[r] * A.inv_n;
                if (va) A.y[(size_t)ea * N + n] = ya;
                if (vb) A.y[(size_t)eb * N + n] = yb;
                vmax_a = fmax(vmax_a, fabs(ya)); vmax_b = fmax(vmax_b, fabs(yb));
            }
        }
#pragma unroll
        for (int o = TP / 2; o > 0; o >>= 1) {
            vmax_a = fmax(vmax_a, __shfl_xor_sync(0xffffffffu, vmax_a, o, TP));
            vmax_b = fmax(vmax_b, __shfl_xor_sync(0xffffffffu, vmax_b, o, TP));
        }
        if (t == 0) {
            if (va) A.vmax_out[ea] = vmax_a;
            if (vb) A.vmax_out[eb] = vmax_b;
        }
        if (t < N1) {
#pragma unroll
            for (int r = 0; r < N2; ++r) {
                const C h = __ldg(A.sens_hat + t + N1 * r);
                zr[r] = ur[r] * h.x - ui[r] * h.y;
                zi[r] = ur[r] * h.y + ui[r] * h.x;
            }
        }
        fft_pass<T, N2, N1, +1>(zr, zi, xb, s_tw12, t);
        if (t < N2) {
#pragma unroll
            for (int r = 0; r < N1; ++r) {
                const int n = t + N2 * r;
                const int i = n / A.sens_sp;
                if (i * A.sens_sp == n && i < n_s) {
                    if (va) A.sensors_out[(size_t)ea * n_s + i] = zr[r] * A.inv_n;
                    if (vb) A.sensors_out[(size_t)eb * n_s + i] = zi[r] * A.inv_n;
                }
            }
        }
        return;
    }
    const size_t n_tab = (size_t)A.sens.nnz_max * n_s;
    const bool staged = n_tab * (sizeof(T) + sizeof(int)) <= dead_bytes;       // CTA-uniform
    constexpr int PF = 8;
    T pf_w[PF]; int pf_i[PF];
    if (staged) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const size_t i = threadIdx.x + (size_t)k * NT;
            pf_w[k] = i < n_tab ? __ldg(A.sens.w + i) : T(0);
            pf_i[k] = i < n_tab ? __ldg(A.sens.idx + i) : 0;
        }
    }
    T vmax_a = T(0), vmax_b = T(0);
    if (t < N2) {
#pragma unroll
        for (int r = 0; r < N1; ++r) {
            const int n = t + N2 * r;
            const T ya = zr[r] * A.inv_n, yb = zi[r] * A.inv_n;
            xb[A.perm ? __ldg(A.perm + n) : n] = V2<T>::make(ya, yb);
            if (va) A.y[(size_t)ea * N + n] = ya;
            if (vb) A.y[(size_t)eb * N + n] = yb;
            vmax_a = fmax(vmax_a, fabs(ya)); vmax_b = fmax(vmax_b, fabs(yb));
        }
    }
#pragma unroll
    for (int o = TP / 2; o > 0; o >>= 1) {
        vmax_a = fmax(vmax_a, __shfl_xor_sync(0xffffffffu, vmax_a, o, TP));
        vmax_b = fmax(vmax_b, __shfl_xor_sync(0xffffffffu, vmax_b, o, TP));
    }
    if (t == 0) {
        if (va) A.vmax_out[ea] = vmax_a;
        if (vb) A.vmax_out[eb] = vmax_b;
    }
    __syncthreads();                     // PREV / F dead from here

    // ---- sensors: raw dots <y, g_i> for both envs (CU sensors per thread in flight) ---------
    T* s_w = reinterpret_cast<T*>(dead);
    int* s_i = reinterpret_cast<int*>(s_w + n_tab);
    if (staged) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const size_t i = threadIdx.x + (size_t)k * NT;
            if (i < n_tab) { s_w[i] = pf_w[k]; s_i[i] = pf_i[k]; }
        }
        for (size_t i = threadIdx.x + (size_t)PF * NT; i < n_tab; i += NT) { s_w[i] = __ldg(A.sens.w + i); s_i[i] = __ldg(A.sens.idx + i); }
    }
    __syncthreads();
    // called once with shared-memory and once with global pointers so that the staged case compiles to LDS
    auto dots = [&](const int* __restrict__ tidx, const T* __restrict__ tw) {
        for (int i0 = t; i0 < n_s; i0 += CU * TP) {
            T sa[CU], sb[CU];
#pragma unroll
            for (int m = 0; m < CU; ++m) { sa[m] = T(0); sb[m] = T(0); }
#pragma unroll 3
            for (int j = 0; j < A.sens.nnz_max; ++j) {
                int idx[CU]; T w[CU];
#pragma unroll
                for (int m = 0; m < CU; ++m) {
                    const int i = i0 + m * TP;
                    idx[m] = i < n_s ? tidx[j * n_s + i] : 0;
                    w[m] = i < n_s ? tw[j * n_s + i] : T(0);
                }
#pragma unroll
                for (int m = 0; m < CU; ++m) {
                    const C v = xb[idx[m]];
                    sa[m] += v.x * w[m]; sb[m] += v.y * w[m];
                }
            }
#pragma unroll
            for (int m = 0; m < CU; ++m) {
                const int i = i0 + m * TP;
                if (i < n_s) {
                    if (va) A.sensors_out[(size_t)ea * n_s + i] = sa[m];
                    if (vb) A.sensors_out[(size_t)eb * n_s + i] = sb[m];
                }
            }
        }
    };
    if (staged) dots(s_i, s_w);
    else dots(A.sens.idx, A.sens.w);
}

}  // namespace pdeb200
