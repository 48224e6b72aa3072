// Fused Kuramoto-Sivashinsky environment step for thousands of environments.
//
// One launch performs, per environment and for `n_steps` consecutive env steps,
//   [actor forward]  src/PDEagent.jl:189,204 (optional, fused policy inference)
//   prepare_action   scripts/KS/setup/KSSetup.jl:231-245
//   do_step          KSSetup.jl:130-160   (CNAB2 pseudo-spectral, `S` substeps)
//   reward_function  KSSetup.jl:162-184
//   featurize        KSSetup.jl:190-229
//   clock / done     src/PDEenv.jl:224-240
// with the PDE state resident in registers/shared memory for the whole launch.
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
#include "obs_reward.cuh"

namespace pdeb200 {

template <typename T>
struct KsArgs {
    using C = typename V2<T>::type;
    int n_envs, S, n_steps;
    int use_actor;                     // 1: actions from the fused actor; 0: from actions_in
    int write_p;                       // materialise env.p
    ObsRewardParams<T> P;
    // spectral constants, spectral natural order k = 0..N-1
    const C* tw12;                     // [N1][N2]  exp(-2*pi*i*k1*t/N)
    const C* tw21;                     // [N2][N1]
    const T* c1;                       // A_inv * B
    const T* cN;                       // A_inv * (-alpha/2) / N^2
    const T* ainvh;                    // A_inv * h
    const C* hm;                       // h * fft(mu*cos(...)) or nullptr (mu == 0)
    T dt32, dt2, inv_n, n_scale, power;
    EllTable<T> sens;                  // rows: sensors, gather over grid points
    EllTable<T> actT;                  // rows: grid points, gather over actuators
    // environment arrays (device)
    T* y; T* p; T* state; T* action; T* delta_action; T* reward; T* sensors_out;
    uint8_t* done; double* time; int* steps;
    const T* actions_in;
    NetDev actor; T act_limit;
    double* reward_sum;                // optional [B]
};

template <int P_, int Q_> struct PassStride { static constexpr int value = (Q_ % 2 == 0) ? Q_ + 1 : Q_; };

// in : threads t < Q hold elements (t + Q*r), r < P        (registers zr/zi[0..P))
// out: threads t < P hold elements (t + P*r), r < Q
// transform: X[k] = sum_n x[n] exp(SIGN*2*pi*i*n*k/(P*Q)); tw = [P][Q] forward twiddles.
template <typename T, int P, int Q, int SIGN>
__device__ __forceinline__ void fft_pass(T* __restrict__ zr, T* __restrict__ zi, typename V2<T>::type* xb,
                                         const typename V2<T>::type* __restrict__ tw, int t) {
    using C = typename V2<T>::type;
    constexpr int STRIDE = PassStride<P, Q>::value;
    if (t < Q) {
        dft_r<P, T, SIGN>(zr, zi);
        xb[t] = V2<T>::make(zr[0], zi[0]);
#pragma unroll
        for (int k = 1; k < P; ++k) {
            const C w = tw[k * Q + t];
            const T wi = (SIGN < 0) ? w.y : -w.y;
            xb[k * STRIDE + t] = V2<T>::make(zr[k] * w.x - zi[k] * wi, zr[k] * wi + zi[k] * w.x);
        }
    }
    __syncwarp();
    if (t < P) {
#pragma unroll
        for (int n = 0; n < Q; ++n) {
            const C v = xb[t * STRIDE + n];
            zr[n] = v.x; zi[n] = v.y;
        }
        dft_r<Q, T, SIGN>(zr, zi);
    }
    __syncwarp();
}

template <int N1, int N2> struct KsGeom {
    static constexpr int N = N1 * N2;
    static constexpr int RMAX = N1 > N2 ? N1 : N2;
    static constexpr int TP = RMAX <= 8 ? 8 : (RMAX <= 16 ? 16 : 32);        // threads per env pair
    static constexpr int S12 = N1 * PassStride<N1, N2>::value;
    static constexpr int S21 = N2 * PassStride<N2, N1>::value;
    static constexpr int XB = (S12 > S21 ? S12 : S21) > N ? (S12 > S21 ? S12 : S21) : N;
};

// dynamic shared memory per CTA, in units of C (complex) unless noted
template <typename T, int N1, int N2>
__host__ __device__ inline size_t ks_smem_bytes(int pairs_per_cta, int n_sensors, int n_act) {
    using G = KsGeom<N1, N2>;
    using C = typename V2<T>::type;
    size_t shared = (size_t)(N1 == N2 ? 1 : 2) * G::N * sizeof(C) + 3 * (size_t)G::N * sizeof(T);
    size_t per_pair = ((size_t)G::XB + 2 * G::N + n_sensors + 2 * n_act) * sizeof(C);
    return shared + per_pair * pairs_per_cta;
}

template <typename T, int N1, int N2, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ks_step_kernel(const __grid_constant__ KsArgs<T> A) {
    using G = KsGeom<N1, N2>;
    using C = typename V2<T>::type;
    constexpr int N = G::N, TP = G::TP, RMAX = G::RMAX;
    constexpr int PAIRS = WARPS * 32 / TP;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* s_tw12 = reinterpret_cast<C*>(smem_raw);
    C* s_tw21 = (N1 == N2) ? s_tw12 : s_tw12 + N;
    T* s_c1 = reinterpret_cast<T*>(s_tw21 + N);
    T* s_cN = s_c1 + N;
    T* s_ah = s_cN + N;
    C* s_pair0 = reinterpret_cast<C*>(s_ah + N);

    const ObsRewardParams<T>& P = A.P;
    const int n_s = P.n_sensors, n_a = P.n_act;
    const int per_pair = G::XB + 2 * N + n_s + 2 * n_a;
    const int t = threadIdx.x % TP;
    const int pic = threadIdx.x / TP;
    const int pair = blockIdx.x * PAIRS + pic;
    C* xb = s_pair0 + (size_t)pic * per_pair;        // exchange / natural-order field buffer
    C* s_prev = xb + G::XB;                          // N^{n-1} (raw fft of (N u)^2)
    C* s_F = s_prev + N;                             // A_inv*h*p_hat + h*m_hat
    C* s_sens = s_F + N;                             // raw sensor dots (env a, env b)
    C* s_act = s_sens + n_s;                         // action row 0 (env a, env b)
    C* s_dact = s_act + n_a;                         // delta_action row 0

    for (int i = threadIdx.x; i < N; i += WARPS * 32) {
        s_tw12[i] = A.tw12[i];
        if (N1 != N2) s_tw21[i] = A.tw21[i];
        s_c1[i] = A.c1[i]; s_cN[i] = A.cN[i]; s_ah[i] = A.ainvh[i];
    }
    __syncthreads();

    const int ea = 2 * pair, eb = 2 * pair + 1;
    const bool va = ea < A.n_envs, vb = eb < A.n_envs;
    // Pairs beyond the batch keep running on zeros (they share __syncwarp with live pairs).

    T zr[RMAX], zi[RMAX], ur[RMAX], ui[RMAX];

    // ---- load y (physical layout: thread t < N2 holds n = t + N2*r) -----------------------
#pragma unroll
    for (int r = 0; r < N1; ++r) {
        const int n = t + N2 * r;
        zr[r] = (va && t < N2) ? A.y[(size_t)ea * N + n] : T(0);
        zi[r] = (vb && t < N2) ? A.y[(size_t)eb * N + n] : T(0);
    }
    for (int j = t; j < n_a; j += TP) {
        s_act[j] = V2<T>::make(va ? A.action[((size_t)ea * n_a + j) * P.a_rows] : T(0),
                               vb ? A.action[((size_t)eb * n_a + j) * P.a_rows] : T(0));
    }
    double time_a = va ? A.time[ea] : 0.0, time_b = vb ? A.time[eb] : 0.0;
    double rsum_a = 0.0, rsum_b = 0.0;
    __syncwarp();

    for (int step = 0; step < A.n_steps; ++step) {
        // ---- actions: fused actor or supplied -------------------------------------------
        for (int j = t; j < n_a; j += TP) {
            T na[2] = {T(0), T(0)};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int env = e ? eb : ea;
                if (!(e ? vb : va)) continue;
                T* acol = A.action + ((size_t)env * n_a + j) * P.a_rows;
                T* dcol = A.delta_action + ((size_t)env * n_a + j) * P.a_rows;
                if (A.use_actor && !P.mono) {
                    float x[kFusedActorMaxWidth], h[kFusedActorMaxWidth];
                    const T* scol = A.state + ((size_t)env * n_a + j) * P.obs_rows;
                    for (int r = 0; r < P.obs_rows; ++r) x[r] = (float)scol[r];
                    mlp_forward_small(A.actor, x, h);
                    for (int r = 0; r < P.a_rows; ++r) {
                        const T v = clamp_t<T>((T)x[r], A.act_limit);
                        dcol[r] = v - acol[r]; acol[r] = v;
                        if (r == 0) na[e] = v;
                    }
                } else if (!A.use_actor) {
                    const T* icol = A.actions_in + ((size_t)env * n_a + j) * P.a_rows;
                    for (int r = 0; r < P.a_rows; ++r) {
                        const T v = icol[r];
                        dcol[r] = v - acol[r]; acol[r] = v;
                        if (r == 0) na[e] = v;
                    }
                }
            }
            if (!(A.use_actor && P.mono)) {
                const C old = s_act[j];
                s_dact[j] = V2<T>::make(na[0] - old.x, na[1] - old.y);
                s_act[j] = V2<T>::make(na[0], na[1]);
            }
        }
        if (A.use_actor && P.mono) {
            // global agent: one column per env, n_a outputs (KSglobalSetup.jl); thread 0/1 of the pair
            if (t < 2 && (t ? vb : va)) {
                const int env = t ? eb : ea;
                float x[kFusedActorMaxWidth], h[kFusedActorMaxWidth];
                const T* scol = A.state + (size_t)env * P.obs_rows;
                for (int r = 0; r < P.obs_rows; ++r) x[r] = (float)scol[r];
                mlp_forward_small(A.actor, x, h);
                for (int j = 0; j < n_a; ++j) {
                    const T v = clamp_t<T>((T)x[j], A.act_limit);
                    T* acol = A.action + (size_t)env * n_a + j;
                    A.delta_action[(size_t)env * n_a + j] = v - *acol; *acol = v;
                    T* sa = reinterpret_cast<T*>(s_act + j) + t;
                    T* sd = reinterpret_cast<T*>(s_dact + j) + t;
                    *sd = v - *sa; *sa = v;
                }
            }
        }
        __syncwarp();

        // keep y in u while z is used for the side transforms
#pragma unroll
        for (int r = 0; r < N1; ++r) { ur[r] = zr[r]; ui[r] = zi[r]; }

        // ---- p = sum_i power*a_i*g_i (gather over actuators covering each point) ----------
        if (t < N2) {
#pragma unroll
            for (int r = 0; r < N1; ++r) {
                const int n = t + N2 * r;
                T pa = T(0), pb = T(0);
                for (int j = 0; j < A.actT.nnz_max; ++j) {
                    const int idx = A.actT.idx[j * N + n];
                    const T w = A.actT.w[j * N + n];
                    const C a = s_act[idx];
                    pa += (A.power * a.x) * w;
                    pb += (A.power * a.y) * w;
                }
                zr[r] = pa; zi[r] = pb;
                if (A.write_p) {
                    if (va) A.p[(size_t)ea * N + n] = pa;
                    if (vb) A.p[(size_t)eb * N + n] = pb;
                }
            }
        }
        fft_pass<T, N1, N2, -1>(zr, zi, xb, s_tw12, t);
        if (t < N1) {
#pragma unroll
            for (int r = 0; r < N2; ++r) {
                const int k = t + N1 * r;
                const T ah = s_ah[k];
                T fr = ah * zr[r], fi = ah * zi[r];
                if (A.hm) { const C m = A.hm[k]; fr += m.x - m.y; fi += m.x + m.y; }   // m_hat*(1+i)
                s_F[k] = V2<T>::make(fr, fi);
            }
        }
        // ---- N^0 = fft((N*y)^2) -> s_prev ; u_hat = fft(y) ---------------------------------
#pragma unroll
        for (int r = 0; r < N1; ++r) {
            const T a = ur[r] * A.n_scale, b = ui[r] * A.n_scale;
            zr[r] = a * a; zi[r] = b * b;
        }
        fft_pass<T, N1, N2, -1>(zr, zi, xb, s_tw12, t);
        if (t < N1) {
#pragma unroll
            for (int r = 0; r < N2; ++r) s_prev[t + N1 * r] = V2<T>::make(zr[r], zi[r]);
        }
#pragma unroll
        for (int r = 0; r < N1; ++r) { zr[r] = ur[r]; zi[r] = ui[r]; }
        fft_pass<T, N1, N2, -1>(zr, zi, xb, s_tw12, t);
#pragma unroll
        for (int r = 0; r < N2; ++r) { ur[r] = zr[r]; ui[r] = zi[r]; }

        // ---- CNAB2 substeps (KSSetup.jl:144-156) -----------------------------------------
        for (int n = 0; n < A.S; ++n) {
            fft_pass<T, N2, N1, +1>(zr, zi, xb, s_tw21, t);          // z = N * u  (physical)
#pragma unroll
            for (int r = 0; r < N1; ++r) { zr[r] = zr[r] * zr[r]; zi[r] = zi[r] * zi[r]; }
            fft_pass<T, N1, N2, -1>(zr, zi, xb, s_tw12, t);          // z = N^2 * fft(u^2)
            if (t < N1) {
#pragma unroll
                for (int r = 0; r < N2; ++r) {
                    const int k = t + N1 * r;
                    const C z1 = s_prev[k];
                    const C f = s_F[k];
                    s_prev[k] = V2<T>::make(zr[r], zi[r]);
                    const T dr = A.dt32 * zr[r] - A.dt2 * z1.x;
                    const T di = A.dt32 * zi[r] - A.dt2 * z1.y;
                    const T c1 = s_c1[k], cn = s_cN[k];
                    // u = c1*u + i*cn*d + F
                    ur[r] = c1 * ur[r] - cn * di + f.x;
                    ui[r] = c1 * ui[r] + cn * dr + f.y;
                    zr[r] = ur[r]; zi[r] = ui[r];
                }
            }
        }
        fft_pass<T, N2, N1, +1>(zr, zi, xb, s_tw21, t);
        T vmax_a = T(0), vmax_b = T(0);
        if (t < N2) {
#pragma unroll
            for (int r = 0; r < N1; ++r) {
                zr[r] *= A.inv_n; zi[r] *= A.inv_n;
                xb[t + N2 * r] = V2<T>::make(zr[r], zi[r]);
                vmax_a = fmax(vmax_a, fabs(zr[r])); vmax_b = fmax(vmax_b, fabs(zi[r]));
            }
        }
        __syncwarp();

        // ---- sensors: raw dots <y, g_i> for both envs ------------------------------------
        for (int i = t; i < n_s; i += TP) {
            T sa = T(0), sb = T(0);
            for (int j = 0; j < A.sens.nnz_max; ++j) {
                const int idx = A.sens.idx[j * n_s + i];
                const T w = A.sens.w[j * n_s + i];
                const C v = xb[idx];
                sa += v.x * w; sb += v.y * w;
            }
            s_sens[i] = V2<T>::make(sa, sb);
            if (A.sensors_out) {
                if (va) A.sensors_out[(size_t)ea * n_s + i] = sa;
                if (vb) A.sensors_out[(size_t)eb * n_s + i] = sb;
            }
        }
        __syncwarp();

        // ---- reward + observation columns -------------------------------------------------
        T racc_a = T(0), racc_b = T(0), rmax_a = T(0), rmax_b = T(0);
        for (int j = t; j < n_a; j += TP) {
            const C a = s_act[j], d = s_dact[j];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int env = e ? eb : ea;
                if (!(e ? vb : va)) continue;
                auto sens = [&](int, int i) { return e ? s_sens[i].y : s_sens[i].x; };
                T rj;
                if (!P.mono) {
                    const size_t col = (size_t)env * n_a + j;
                    rj = assemble_column<T>(P, sens, j, e ? a.y : a.x, e ? d.y : d.x,
                                            A.action + col * P.a_rows, A.state + col * P.obs_rows, false);
                    A.reward[col] = rj;
                } else {
                    const int m = P.a2s[j];
                    const T raw = sens(0, m) - P.r_offset * P.sens_sum[m];
                    const T s = pow_t<T>(fabs(P.r_gain * raw), P.r_pow) / P.r_div;
                    const T a0 = e ? a.y : a.x, d0 = e ? d.y : d.x;
                    rj = -fabs(s) - P.a_pun * a0 * a0 - P.da_pun * d0 * d0;
                }
                if (e) { racc_b += rj; rmax_b = fmax(rmax_b, fabs(rj)); }
                else   { racc_a += rj; rmax_a = fmax(rmax_a, fabs(rj)); }
            }
        }
        if (P.mono) {
            // state = reshape(sensors, (n_sensors, 1)) [+ temporal stacking], KSglobalSetup.jl:222-238
            for (int e = 0; e < 2; ++e) {
                if (!(e ? vb : va)) continue;
                T* scol = A.state + (size_t)(e ? eb : ea) * P.obs_rows;
                if (P.temporal > 1) {
                    __syncwarp();
                    for (int r = P.obs_rows - P.memory - 1 - t; r >= n_s; r -= TP) {
                        // shift down by n_s rows; done back-to-front in TP-wide sweeps
                        scol[r] = scol[r - n_s];
                    }
                    __syncwarp();
                }
                for (int i = t; i < n_s; i += TP) scol[i] = (e ? s_sens[i].y : s_sens[i].x) * P.obs_scale;
            }
        }
#pragma unroll
        for (int o = TP / 2; o > 0; o >>= 1) {
            racc_a += __shfl_xor_sync(0xffffffffu, racc_a, o, TP);
            racc_b += __shfl_xor_sync(0xffffffffu, racc_b, o, TP);
            rmax_a = fmax(rmax_a, __shfl_xor_sync(0xffffffffu, rmax_a, o, TP));
            rmax_b = fmax(rmax_b, __shfl_xor_sync(0xffffffffu, rmax_b, o, TP));
            vmax_a = fmax(vmax_a, __shfl_xor_sync(0xffffffffu, vmax_a, o, TP));
            vmax_b = fmax(vmax_b, __shfl_xor_sync(0xffffffffu, vmax_b, o, TP));
        }
        const T rmean_a = racc_a / T(n_a), rmean_b = racc_b / T(n_a);
        if (P.mono) { rmax_a = fabs(rmean_a); rmax_b = fabs(rmean_b); }
        rsum_a += (double)rmean_a; rsum_b += (double)rmean_b;
        time_a += P.dt; time_b += P.dt;
        if (t == 0) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (!(e ? vb : va)) continue;
                const int env = e ? eb : ea;
                if (P.mono) A.reward[env] = e ? rmean_b : rmean_a;
                const double tm = e ? time_b : time_a;
                bool dn = tm >= P.te;
                if (P.check_max == 1) dn = dn || ((e ? vmax_b : vmax_a) > P.max_value);
                else if (P.check_max == 2) dn = dn || ((e ? rmax_b : rmax_a) > P.max_value);
                A.done[env] = dn ? 1 : 0;
                A.time[env] = tm;
                A.steps[env] += 1;
            }
        }
        __syncwarp();
    }

    // ---- store y ------------------------------------------------------------------------
    if (t < N2) {
#pragma unroll
        for (int r = 0; r < N1; ++r) {
            const int n = t + N2 * r;
            if (va) A.y[(size_t)ea * N + n] = zr[r];
            if (vb) A.y[(size_t)eb * N + n] = zi[r];
        }
    }
    if (A.reward_sum && t == 0) {
        if (va) A.reward_sum[ea] += rsum_a;
        if (vb) A.reward_sum[eb] += rsum_b;
    }
}

}  // namespace pdeb200
